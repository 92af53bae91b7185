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
