"""Runner (test infrastructure): executes an UNCHANGED reference train script for a few iterations, either on the
reference's own ``models`` package or with this repo's drop-in ``models`` package shadowing it, and prints the
per-iteration losses as one JSON line.

    python tests/run_reference_script.py --impl dropin|reference [--script NPP_completion/train.py] [--iters 20]
                                         [--loss_type l2|robust_loss_adaptive] [--workdir DIR]

The reference checkout is baseline/_ref (copied by __graft_entry__.build(); /root/reference in the build container).
What is shimmed is environment, not algorithm: modules this image lacks (configargparse,
matplotlib, imageio, skimage, kornia, torch_dct: stubbed), numpy aliases removed since the reference was written,
torchvision weight downloads (no network: random-initialised VGG/AlexNet, as BASELINE.json's north_star allows), and
a synthesised ``config.odgt`` (the file the periodicity search would write, NPP_proposal/search.py:228-239).
"""
import argparse
import json
import os
import re
import runpy
import shutil
import sys
import tempfile
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "learning-continuous-implicit-representation-for-near-periodic-patterns_b200")
IMAGE = "20150911134910-5dcfbc24"          # bundled 211 x 325 example (BASELINE.json configs[0])


def reference_root():
    for cand in (os.environ.get("NPP_REFERENCE_ROOT"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "NPP_completion", "train.py")):
            return cand
    return None


def install_shims():
    import numpy as np
    for k, v in (("float", float), ("int", int), ("bool", bool)):
        if not hasattr(np, k):
            setattr(np, k, v)

    def stub(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        parent, _, child = name.rpartition(".")
        if parent:
            setattr(stub(parent), child, m)
        return m

    class _ConfigArgParser(argparse.ArgumentParser):
        def add_argument(self, *a, **k):
            k.pop("is_config_file", None)
            return super().add_argument(*a, **k)

    stub("configargparse", ArgumentParser=_ConfigArgParser)
    stub("torch_dct")
    stub("imageio", imread=lambda *a, **k: None, imwrite=lambda *a, **k: None)
    stub("kornia")
    stub("gco")
    nop = lambda *a, **k: None  # noqa: E731
    plt = stub("matplotlib.pyplot", imsave=nop, imshow=nop, show=nop, figure=nop, savefig=nop, plot=nop, close=nop,
               subplot=nop, title=nop, axis=nop, scatter=nop, colorbar=nop, clf=nop)
    stub("matplotlib", pyplot=plt, use=nop)
    stub("matplotlib.cm")
    stub("skimage")
    stub("skimage.feature", peak_local_max=nop, canny=nop)
    stub("skimage.filters", gaussian=nop, sobel=nop)
    stub("skimage.morphology", disk=nop, dilation=nop, erosion=nop, binary_dilation=nop, binary_erosion=nop)
    # no network: the "pretrained" torchvision backbones of LPIPS / contextual / style losses are random-initialised
    import torchvision.models as tvm

    def no_download(fn):
        def make(*a, **k):
            k.pop("pretrained", None)
            k["weights"] = None
            return fn(**k)
        return make
    import torchvision.models.alexnet
    import torchvision.models.squeezenet
    import torchvision.models.vgg
    for mod in (tvm, tvm.vgg, tvm.alexnet, tvm.squeezenet):
        for name in ("vgg16", "vgg19", "alexnet", "squeezenet1_1"):
            fn = getattr(mod, name, None)
            if fn is not None and not getattr(fn, "_npp_no_download", False):
                wrapped = no_download(fn)
                wrapped._npp_no_download = True
                setattr(mod, name, wrapped)


def make_datadir(ref, workdir, topk=3):
    """Copy of the bundled example with the config.odgt the periodicity search would have written."""
    src = os.path.join(ref, "data", "completion", "input", IMAGE)
    dst = os.path.join(workdir, "data", IMAGE)
    shutil.copytree(src, dst)
    p = 36.0
    info = {
        "fpath_masked_img": "masked_img.png", "fpath_valid_mask": "valid_mask.png", "fpath_mask": "unknown_mask.png",
        "fpath_gt_img": "gt_img.png",
        "selected_angles": [[90.0, 180.0]] * 3,
        "selected_periods": [[p, 0.9 * p], [p / 2, 0.45 * p], [2 * p, 1.8 * p]],
        "selected_shifts": [[[0.0, p], [0.9 * p, 0.0]], [[0.0, p / 2], [0.45 * p, 0.0]], [[0.0, 2 * p], [1.8 * p, 0.0]]],
    }
    with open(os.path.join(dst, "config.odgt"), "w") as fh:
        fh.write(json.dumps(info) + "\n")
    return dst


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", required=True, choices=["dropin", "reference"])
    ap.add_argument("--script", default="NPP_completion/train.py")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--loss_type", default="l2")
    ap.add_argument("--workdir", default=None)
    ap.add_argument("--no_patch_losses", action="store_true",
                    help="pixel loss only (the script's --use_contextual_loss / --use_perceptual_loss switches, which turn them OFF)")
    a = ap.parse_args()
    ref = reference_root()
    if ref is None:
        print(json.dumps({"unavailable": "no reference checkout (baseline/_ref)"}))
        return
    import torch
    if not torch.cuda.is_available():
        print(json.dumps({"unavailable": "needs a CUDA device (the script sets the CUDA default tensor type)"}))
        return
    workdir = a.workdir or tempfile.mkdtemp(prefix="npp_script_")
    install_shims()
    datadir = make_datadir(ref, workdir)
    if a.impl == "dropin":
        sys.path.insert(0, PKG)            # shadows `models`; the script appends the reference root itself
        sys.path.insert(1, ROOT)
        os.environ.setdefault("NPP_REFERENCE_ROOT", ref)
    script = os.path.join(ref, a.script)
    argv = [script, "--datadir", datadir, "--basedir", os.path.join(workdir, "results"), "--N_iters", str(a.iters + 1),
            "--i_print", "1", "--i_testset", str(a.iters), "--loss_type", a.loss_type]
    if a.no_patch_losses:
        argv += ["--use_contextual_loss", "--use_perceptual_loss"]
    sys.argv = argv
    lines = []
    from tqdm import tqdm
    orig_write = tqdm.write

    def capture(s, *x, **k):
        lines.append(str(s))
        return orig_write(s, file=sys.stderr)
    tqdm.write = staticmethod(capture)
    import contextlib
    t0 = __import__("time").time()
    with contextlib.redirect_stdout(sys.stderr):
        cwd = os.getcwd()
        os.chdir(ref)
        try:
            runpy.run_path(script, run_name="__main__")
        finally:
            os.chdir(cwd)
    losses = []
    for s in lines:
        m = re.search(r"Iter: (\d+) Loss: ([-+0-9.eE]+)", s)
        if m:
            losses.append(float(m.group(2)))
    import models
    print(json.dumps({"impl": a.impl, "script": a.script, "loss_type": a.loss_type, "iters": a.iters, "losses": losses,
                      "models_package": os.path.dirname(models.__file__), "seconds": __import__("time").time() - t0}))
    shutil.rmtree(workdir, ignore_errors=True)


if __name__ == "__main__":
    main()
