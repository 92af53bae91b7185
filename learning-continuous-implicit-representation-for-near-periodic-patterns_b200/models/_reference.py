"""Loader for reference modules that stay reference PyTorch (periodicity search, patch losses).

Nothing is copied: the module is executed from the reference checkout found on sys.path (the reference scripts
append their root there) or named by NPP_REFERENCE_ROOT."""
import importlib.util
import os
import sys

_cache = {}


def reference_root():
    cand = [os.environ.get("NPP_REFERENCE_ROOT")] + list(sys.path)
    for c in cand:
        if c and os.path.isfile(os.path.join(c, "models", "networks.py")) and \
                os.path.isdir(os.path.join(c, "NPP_proposal")):
            return c
    raise ImportError("the reference checkout is not on sys.path (set NPP_REFERENCE_ROOT); it is needed only for the "
                      "parts that stay reference PyTorch (periodicity search, style/LPIPS/contextual losses)")


def reference_module(name):
    if name not in _cache:
        path = os.path.join(reference_root(), "models", name + ".py")
        spec = importlib.util.spec_from_file_location(f"models._ref_{name}", path,
                                                      submodule_search_locations=None)
        mod = importlib.util.module_from_spec(spec)
        mod.__package__ = "models"
        sys.modules[spec.name] = mod
        spec.loader.exec_module(mod)
        _cache[name] = mod
    return _cache[name]
