"""Out of scope for the B200 hot path: executes the reference's models/model_def.py unchanged (see _reference.py)."""
from ._reference import reference_module as _ref

_m = _ref("model_def")
globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
