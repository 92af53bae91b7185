"""Adaptive robust pixel loss (Barron, CVPR 2019) -- the reference's default ``--loss_type robust_loss_adaptive``.

Mirrors the surface of ``externel_lib/robust_loss_pytorch/adaptive.py::AdaptiveLossFunction`` that the NPP-Net scripts use
(``models/helpers.py:8-9`` builds ``adaptive_pix = AdaptiveLossFunction(num_dims=3, float_dtype=np.float32, device=0)``,
``models/mse_calculator.py:24-25`` calls ``adaptive.lossfun(diff)``, ``helpers.py:144`` hands its parameters to Adam):
``latent_alpha`` / ``latent_scale`` parameters of shape [1, num_dims], ``alpha()``, ``scale()``, ``lossfun(x)``.

On CUDA, ``fused_img2mse`` computes the masked mean loss and all its gradients in one pass through libnpp_b200
(``npp_robust_adaptive_fwd_bwd``); ``lossfun`` is the same arithmetic written with torch ops (any device, any shape
[N, num_dims]).  log Z(alpha) comes from this package's own table (``data/robust_logz_table.npz``, exact integral by
quadrature, ``tools/make_robust_logz_table.py``), interpolated with cubic Hermite polynomials.
"""
import os

import numpy as np
import torch
import torch.nn as nn

from . import _native as nat

_SHIFT = float(np.log(np.expm1(1.0)))    # inv_softplus(1)           (util.py:51-53)
_EPS = float(np.finfo(np.float32).eps)
_TABLE = None


def _table():
    global _TABLE
    if _TABLE is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "robust_logz_table.npz")
        with np.load(path) as f:
            _TABLE = (float(f["alpha_max"]), f["values"].astype(np.float32), f["derivs"].astype(np.float32))
    return _TABLE


def log_partition(alpha: torch.Tensor) -> torch.Tensor:
    """log Z(alpha) for 0 <= alpha <= 2, differentiable (derivative of the interpolant)."""
    amax, val, der = _table()
    n = val.shape[0]
    h = amax / (n - 1)
    v = torch.as_tensor(val, device=alpha.device, dtype=alpha.dtype)
    d = torch.as_tensor(der, device=alpha.device, dtype=alpha.dtype) * h
    pos = torch.clamp(alpha / h, 0.0, n - 1 - 1e-3)
    i = pos.detach().floor().long()
    u = pos - i.to(pos.dtype)
    u2, u3 = u * u, u * u * u
    return ((2 * u3 - 3 * u2 + 1) * v[i] + (u3 - 2 * u2 + u) * d[i] + (-2 * u3 + 3 * u2) * v[i + 1] + (u3 - u2) * d[i + 1])


class _FusedAdaptiveMSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, mask, latent_alpha, latent_scale, owner):
        n = x.shape[0]

        def dense(t):   # no copy in the common case (contiguous fp32)
            t = t.detach()
            return t if t.dtype == torch.float32 and t.is_contiguous() else t.contiguous().float()

        xc, yc = dense(x), dense(y)
        mc = None if mask is None else dense(mask)
        la, ls = dense(latent_alpha), dense(latent_scale)
        zv, zd, amax = owner._device_table(x.device)
        gx = torch.empty_like(xc)
        buf = torch.empty(16, device=x.device, dtype=torch.float32)      # [0:9] reduction scratch, [9:16] results
        nat.check(nat.lib().npp_robust_adaptive_fwd_bwd(
            xc.data_ptr(), yc.data_ptr(), nat.ptr(mc), n, la.data_ptr(), ls.data_ptr(), owner._cfg_arr.ctypes.data,
            zv.data_ptr(), zd.data_ptr(), zv.shape[0], amax, buf.data_ptr(), buf.data_ptr() + 36, gx.data_ptr(),
            nat.current_stream()))
        ctx.save_for_backward(gx, buf)
        ctx.shapes = (latent_alpha.shape, latent_scale.shape)
        return buf[9].clone()

    @staticmethod
    def backward(ctx, g):
        gx, buf = ctx.saved_tensors
        sa, ss = ctx.shapes
        gl = g * buf[10:16]
        return g * gx, None, None, gl[0:3].reshape(sa), gl[3:6].reshape(ss), None


class _FusedL2(torch.autograd.Function):
    """img2mse(x, y, 'l2', None, mask) on a CUDA [N,3] prediction: loss and dL/dx in one kernel."""

    @staticmethod
    def forward(ctx, x, y, mask):
        def dense(t):
            t = t.detach()
            return t if t.dtype == torch.float32 and t.is_contiguous() else t.contiguous().float()

        xc, yc = dense(x), dense(y)
        mc = None if mask is None else dense(mask)
        gx = torch.empty_like(xc)
        loss = torch.empty((), device=x.device, dtype=torch.float32)
        nat.check(nat.lib().npp_l2_fwd_bwd(xc.data_ptr(), yc.data_ptr(), nat.ptr(mc), x.shape[0], loss.data_ptr(),
                                           gx.data_ptr(), nat.current_stream()))
        ctx.save_for_backward(gx)
        return loss

    @staticmethod
    def backward(ctx, g):
        (gx,) = ctx.saved_tensors
        return g * gx, None, None


def fused_l2_img2mse(x, y, mask=None):
    return _FusedL2.apply(x, y, mask)


class NppAdaptiveLoss(nn.Module):
    def __init__(self, num_dims, float_dtype=np.float32, device="cuda", alpha_lo=0.001, alpha_hi=1.999, alpha_init=None,
                 scale_lo=1e-5, scale_init=1.0):
        super().__init__()
        if not (0 < alpha_lo < alpha_hi < 2):
            raise ValueError("NppAdaptiveLoss needs 0 < alpha_lo < alpha_hi < 2 (the reference default is 0.001 .. 1.999)")
        if not (0 < scale_lo < scale_init):
            raise ValueError("NppAdaptiveLoss needs 0 < scale_lo < scale_init")
        if float_dtype not in (np.float32, torch.float32):
            raise ValueError("NppAdaptiveLoss is float32 only")
        if alpha_init is None:
            alpha_init = (alpha_lo + alpha_hi) / 2.0
        if not (alpha_lo < alpha_init < alpha_hi):
            raise ValueError("`alpha_init` must be in (`alpha_lo`, `alpha_hi`)")
        self.num_dims = num_dims
        self.float_dtype = torch.float32
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self._cfg = (float(alpha_lo), float(alpha_hi), float(scale_lo), float(scale_init))
        self._cfg_arr = np.asarray(self._cfg, np.float32)
        p = (alpha_init - alpha_lo) / (alpha_hi - alpha_lo)
        latent0 = float(-np.log(1.0 / p - 1.0))                       # inv_affine_sigmoid (util.py:77-85)
        self.latent_alpha = nn.Parameter(torch.full((1, num_dims), latent0, dtype=torch.float32, device=self.device))
        self.latent_scale = nn.Parameter(torch.zeros((1, num_dims), dtype=torch.float32, device=self.device))
        self._dev_table = {}

    def alpha(self):
        lo, hi, _, _ = self._cfg
        return torch.sigmoid(self.latent_alpha) * (hi - lo) + lo       # affine_sigmoid (util.py:64-74)

    def scale(self):
        _, _, lo, ref = self._cfg
        return (ref - lo) * torch.nn.functional.softplus(self.latent_scale + _SHIFT) + lo   # affine_softplus (util.py:88-95)

    def lossfun(self, x, **kwargs):
        """Per-element negative log-likelihood, shape of x ([N, num_dims]) -- torch ops, differentiable."""
        x = torch.as_tensor(x)
        assert x.dim() == 2 and x.shape[1] == self.num_dims and x.dtype == torch.float32
        alpha, scale = self.alpha().to(x.device), self.scale().to(x.device)
        b = torch.clamp(torch.abs(alpha - 2.0), min=_EPS)
        a = torch.clamp(torch.abs(alpha), min=_EPS)
        loss = (b / a) * (torch.pow((x / scale) ** 2 / b + 1.0, 0.5 * alpha) - 1.0)      # general.py:84-118, 0 < alpha < 2
        return loss + torch.log(scale) + log_partition(alpha)                            # distribution.py:204-209

    def _device_table(self, device):
        key = str(device)
        if key not in self._dev_table:
            amax, val, der = _table()
            self._dev_table[key] = (torch.from_numpy(val).to(device), torch.from_numpy(der).to(device), amax)
        return self._dev_table[key]

    def fused_img2mse(self, x, y, mask=None):
        """mean over [N,3] of lossfun((x - y) * (mask + 0.3 (1 - mask))) in one CUDA pass (mse_calculator.py:13-27)."""
        if not (x.is_cuda and x.dim() == 2 and x.shape[1] == 3 and self.num_dims == 3):
            raise ValueError("fused_img2mse needs a CUDA [N,3] prediction and num_dims == 3")
        return _FusedAdaptiveMSE.apply(x, y, mask, self.latent_alpha, self.latent_scale, self)
