"""Bridge from the top-level ``models`` package (how the reference scripts import it) to the core package,
whose hyphenated directory name cannot appear in an import statement."""
import importlib
import os
import sys

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ROOT = os.path.dirname(_PKG_DIR)
if _ROOT not in sys.path:
    sys.path.append(_ROOT)
_NAME = os.path.basename(_PKG_DIR)
core = importlib.import_module(_NAME)
plan = importlib.import_module(_NAME + ".plan")
native = importlib.import_module(_NAME + "._native")
EncoderSpec = plan.EncoderSpec
Plan = plan.Plan
NppAdaptiveLoss = importlib.import_module(_NAME + ".robust_loss").NppAdaptiveLoss
fused_l2_img2mse = importlib.import_module(_NAME + ".robust_loss").fused_l2_img2mse
