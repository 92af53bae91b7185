"""Out of scope for the B200 hot path: executes the reference's models/alexnet.py unchanged (see _reference.py)."""
from ._reference import reference_module as _ref

_m = _ref("alexnet")
globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
