"""Import alias: the real package directory carries the (hyphenated) name the build contract
prescribes, which cannot be written in an ``import`` statement.  ``import npp_b200`` loads it."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
PACKAGE_NAME = "learning-continuous-implicit-representation-for-near-periodic-patterns_b200"
_pkg = importlib.import_module(PACKAGE_NAME)
sys.modules[__name__] = _pkg


class _AliasFinder:
    """``import npp_b200.x`` must give the SAME module object as the hyphenated package's ``x`` (a second copy would
    carry its own library handle, caches and classes: isinstance checks across the two would fail silently)."""

    @staticmethod
    def find_spec(name, path=None, target=None):
        if not name.startswith(__name__ + "."):
            return None
        real = PACKAGE_NAME + name[len(__name__):]
        mod = importlib.import_module(real)
        sys.modules[name] = mod
        return importlib.util.spec_from_loader(name, loader=_AliasLoader(mod))


class _AliasLoader:
    def __init__(self, mod):
        self.mod = mod

    def create_module(self, spec):
        return self.mod

    def exec_module(self, module):
        pass


import importlib.util  # noqa: E402
sys.meta_path.insert(0, _AliasFinder)
for _k, _v in list(sys.modules.items()):
    if _k.startswith(PACKAGE_NAME + "."):
        sys.modules[__name__ + _k[len(PACKAGE_NAME):]] = _v
