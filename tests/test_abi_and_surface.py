"""CPU-only checks: the C-ABI library loads and exports every symbol include/npp_b200.h declares, the Python
surface keeps the reference's names and argument lists, and the product path fails loudly without a GPU."""
import inspect
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "learning-continuous-implicit-representation-for-near-periodic-patterns_b200")


def _models():
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    import models.embedder
    import models.helpers
    import models.mse_calculator
    import models.networks
    import models.sampler
    return sys.modules["models"]


def test_header_symbols_exported(native):
    header = open(os.path.join(ROOT, "include", "npp_b200.h")).read()
    declared = set(re.findall(r"\b(npp_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(native.SIGNATURES), declared ^ set(native.SIGNATURES)
    lib = native.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.npp_abi_version() == 1


def test_library_is_sm100a_tcgen05():
    """The shipped .so really contains the Blackwell instructions (UTCHMMA = tcgen05.mma, UTMALDG/UTMASTG = TMA,
    LDTM = tcgen05.ld) and no legacy HMMA path."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    so = os.path.join(PKG, "csrc", "libnpp_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM"):
        assert mnemonic in sass, mnemonic
    assert " HMMA" not in sass


def test_reference_signatures():
    _models()
    from models.embedder import get_embedder
    from models.helpers import batchify, create_npp_net, render, run_network
    from models.mse_calculator import img2mse
    from models.networks import NPP_Net, NPP_Net_top1
    from models.sampler import GridPatchSampler

    def names(f):
        return list(inspect.signature(f).parameters)

    assert names(get_embedder) == ["multires", "i", "res", "selected_angles", "selected_periods", "freq_scales",
                                   "freq_offsets", "angle_offsets", "is_search"]
    assert names(NPP_Net.__init__)[1:] == ["input_ch_periodic", "input_ch_periodic_aux", "freq_scales", "freq_offsets",
                                           "angle_offsets", "D", "W", "freq_nerf", "output_ch", "skips", "activation"]
    assert names(NPP_Net_top1.__init__)[1:] == ["input_ch_periodic", "freq_scales", "freq_offsets", "angle_offsets",
                                                "D", "W", "freq_nerf", "output_ch", "skips", "activation"]
    assert names(NPP_Net.forward)[1:] == ["x", "x_periodic"]
    from models.networks import NPP_Net_light
    assert names(NPP_Net_light.__init__)[1:] == ["input_ch_periodic", "freq_scales", "freq_offsets", "angle_offsets",
                                                 "D", "W", "input_ch", "output_ch", "skips", "activation"]
    assert names(NPP_Net_light.forward)[1:] == ["x", "x_periodic"]
    assert names(create_npp_net) == ["args", "selected_angles", "selected_periods", "res", "percep_net", "is_search",
                                     "style_net"]
    assert names(render) == ["select_coords_emb", "select_coords_emb_periodic", "args", "network_query_fn",
                             "network_fn"]
    assert names(img2mse) == ["x", "y", "loss_type", "adaptive", "mask"]
    assert names(run_network) == ["inputs", "inputs_periodic", "fn", "netchunk"]
    assert names(batchify) == ["fn", "chunk"]
    assert names(GridPatchSampler.__init__)[1:] == ["img", "mask", "N_samples", "patch_size", "height", "width",
                                                    "pool_train", "pool_val", "selected_shifts", "no_reg_sampling"]
    assert names(GridPatchSampler.sample_patches)[1:] == ["topk", "invalid_ratio"]
    # names the reference train scripts pick up through their star imports
    import models.helpers as H
    import models.sampler as S
    for n in ("torch", "np", "nn", "F", "os", "device", "adaptive_pix", "create_npp_net", "render", "img2mse",
              "weights_init_normal", "batchify", "run_network"):
        assert hasattr(H, n), n
    for n in ("np", "F", "torch", "GridPatchSampler", "extract_glimpse"):
        assert hasattr(S, n), n


def test_embedder_coords_mode_shapes():
    import torch
    _models()
    from models.embedder import get_embedder
    torch.manual_seed(0)
    emb, d = get_embedder(10, 0, (64, 48))
    assert d == 21
    per = [get_embedder(10, 0, (64, 48), selected_angles=torch.tensor([90.0, 180.0]),
                        selected_periods=torch.tensor([12.0, 11.0]), freq_scales=[1],
                        freq_offsets=[0, -1, 1, 0.5, -0.5], angle_offsets=[0]) for _ in range(3)]
    assert [p[1] for p in per] == [22, 22, 22]
    coords = torch.tensor([[0.0, 0.0], [5.0, 7.0]])
    outs = [emb.embed(p[0].embed(coords.clone())) for p in per]
    full = torch.cat(outs, 1)
    assert full.shape == (2, 2) and torch.equal(full, coords)      # coordinates pass through untouched
    assert get_embedder(10, -1)[1] == 3
    # RNG parity: the Fourier frequencies are the reference's draw, torch.normal(0,1,(10,1))*10 under the same seed
    torch.manual_seed(0)
    ref = (torch.normal(mean=0.0, std=1.0, size=(10, 1)) * 10).reshape(-1).numpy()
    torch.manual_seed(0)
    emb2, _ = get_embedder(10, 0, (64, 48))
    assert (emb2.freqs == ref).all()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    _models()
    from models.embedder import get_embedder
    from models.networks import NPP_Net_top1
    import npp_b200
    get_embedder(10, 0, (64, 48))
    get_embedder(10, 0, (64, 48), selected_angles=torch.tensor([90.0, 180.0]), selected_periods=torch.tensor([12.0, 11.0]),
                 freq_scales=[1], freq_offsets=[0, -1, 1, 0.5, -0.5], angle_offsets=[0])
    with pytest.raises(npp_b200._native.NppError):
        NPP_Net_top1(22, [1], [0, -1, 1, 0.5, -0.5], [0], D=8, W=512, freq_nerf=21, activation='snake')


def test_search_mode_embedders_coords_mode_shapes():
    """get_embedder(..., is_search=True) (models/embedder.py:76-95): 42 / 20 columns, coordinates passed through in the
    default 'coords' embed mode (NPP_Net_light encodes inside the kernels)."""
    import torch
    _models()
    from models.embedder import get_embedder
    torch.manual_seed(0)
    emb, d = get_embedder(10, 0, (64, 48), is_search=True)
    per, dp = get_embedder(10, 0, (64, 48), selected_angles=torch.tensor([90.0, 180.0]),
                           selected_periods=torch.tensor([12.0, 11.0]), freq_scales=[1],
                           freq_offsets=[0, -1, 1, 0.5, -0.5], angle_offsets=[0], is_search=True)
    assert (d, dp) == (42, 20) and emb.is_search and not per.include_input
    coords = torch.tensor([[0.0, 0.0], [5.0, 7.0]])
    assert torch.equal(emb.embed(coords.clone()), coords) and torch.equal(per.embed(coords), coords)
    # same default-RNG consumption as the reference: one torch.normal(size=(multires, 1)) per positional embedder
    torch.manual_seed(0)
    ref_draw = torch.normal(mean=0.0, std=1.0, size=(10, 1)) * 10
    assert torch.equal(emb.freq_bands, ref_draw)


def test_weights_init_normal_touches_conv_and_batchnorm_only():
    import torch
    _models()
    from models.helpers import weights_init_normal
    torch.manual_seed(0)
    lin, conv, bn = torch.nn.Linear(4, 4), torch.nn.Conv2d(2, 2, 3), torch.nn.BatchNorm2d(3)
    before = lin.weight.detach().clone()
    torch.nn.Sequential(lin, conv, bn).apply(weights_init_normal)
    assert torch.equal(lin.weight, before)                       # Linear keeps PyTorch's default init
    assert conv.weight.abs().max() < 0.2 and abs(bn.weight.mean().item() - 1.0) < 0.1 and bn.bias.abs().max() == 0


def test_alias_package_shares_module_objects():
    """`import npp_b200.x` and the hyphenated package's `x` are ONE module (one library handle, one set of classes:
    isinstance checks across the two import paths hold)."""
    import importlib
    import npp_b200  # noqa: F401
    real = "learning-continuous-implicit-representation-for-near-periodic-patterns_b200"
    for sub in ("_native", "plan", "robust_loss", "dp", "search_fits"):
        a = importlib.import_module("npp_b200." + sub)
        b = importlib.import_module(real + "." + sub)
        assert a is b, sub


def test_run_fits_refuses_the_adaptive_loss():
    """The concurrent-fit path trains with 'l2' only; the reference's default --loss_type must not be silently replaced."""
    import pytest
    from npp_b200 import search_fits
    with pytest.raises(NotImplementedError):
        search_fits.run_fits([], None, None, loss_type="robust_loss_adaptive")
