"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): numpy restatement of the reference's adaptive robust pixel loss.

Follows, in float64:
  * models/mse_calculator.py:13-27               img2mse(..., 'robust_loss_adaptive', adaptive, mask)
  * externel_lib/robust_loss_pytorch/adaptive.py:138-198   alpha = affine_sigmoid(latent), scale = affine_softplus(latent)
  * externel_lib/robust_loss_pytorch/util.py:64-95          affine_sigmoid / affine_softplus / inv_softplus
  * externel_lib/robust_loss_pytorch/general.py:84-118      lossfun, the "otherwise" branch (0 < alpha < 2)
  * externel_lib/robust_loss_pytorch/distribution.py:143-210  nllfun = lossfun + log(scale) + log Z(alpha)
log Z(alpha) is the reference's spline approximation of an integral; the oracle evaluates the integral itself by adaptive
quadrature (scipy), independently of the product's interpolation table.  Pinned against the live reference (values and
autograd gradients) by tests/golden/golden_robust.npz.
"""
import numpy as np

EPS = float(np.finfo(np.float32).eps)
DEFAULT_CFG = (0.001, 1.999, 1e-5, 1.0)   # alpha_lo, alpha_hi, scale_lo, scale_init   (adaptive.py:52-58 defaults)
SHIFT = float(np.log(np.expm1(1.0)))       # inv_softplus(1)                           (util.py:51-53,88)


def _sigmoid(v):
    return 1.0 / (1.0 + np.exp(-v))


def adaptive_params(latent_alpha, latent_scale, cfg=DEFAULT_CFG):
    """-> alpha, scale, d alpha / d latent, d scale / d latent (arrays of shape [3])."""
    alpha_lo, alpha_hi, scale_lo, scale_ref = cfg
    la = np.asarray(latent_alpha, np.float64).reshape(-1)
    ls = np.asarray(latent_scale, np.float64).reshape(-1)
    sg = _sigmoid(la)
    alpha = alpha_lo + (alpha_hi - alpha_lo) * sg
    t = ls + SHIFT
    scale = (scale_ref - scale_lo) * np.log1p(np.exp(t)) + scale_lo
    return alpha, scale, (alpha_hi - alpha_lo) * sg * (1 - sg), (scale_ref - scale_lo) * _sigmoid(t)


def rho(d, alpha, scale):
    b = np.maximum(np.abs(alpha - 2.0), EPS)
    a = np.maximum(np.abs(alpha), EPS)
    return (b / a) * (((d / scale) ** 2 / b + 1.0) ** (0.5 * alpha) - 1.0)


def log_partition(alpha):
    """log Z(alpha) and its derivative for scalar 0 < alpha < 2 by quadrature."""
    from scipy import integrate
    a = float(alpha)
    f = lambda x: np.exp(-rho(x, a, 1.0))
    z = 2.0 * integrate.quad(f, 0, np.inf, limit=800, epsabs=1e-13, epsrel=1e-13)[0]

    def drho(x):
        b = 2.0 - a
        u = x * x / b + 1.0
        p = u ** (0.5 * a)
        return (-2.0 / (a * a)) * (p - 1.0) + (b / a) * p * (0.5 * np.log(u) + 0.5 * a * (x * x / (b * b)) / u)

    dz = -2.0 * integrate.quad(lambda x: drho(x) * f(x), 0, np.inf, limit=800, epsabs=1e-12, epsrel=1e-12)[0]
    return np.log(z), dz / z


def adaptive_img2mse(x, y, mask, latent_alpha, latent_scale, cfg=DEFAULT_CFG):
    """-> loss, dL/dx [N,3], dL/dlatent_alpha [3], dL/dlatent_scale [3]."""
    x = np.asarray(x, np.float64)
    y = np.asarray(y, np.float64)
    n = x.shape[0]
    w = np.ones((n, 1)) if mask is None else np.asarray(mask, np.float64).reshape(n, 1)
    w = w + (1.0 - w) * 0.3
    d = (x - y) * w
    alpha, scale, dal, dsl = adaptive_params(latent_alpha, latent_scale, cfg)
    lz = np.array([log_partition(a) for a in alpha])
    b = np.maximum(np.abs(alpha - 2.0), EPS)
    q = (d / scale) ** 2
    u = q / b + 1.0
    p = u ** (0.5 * alpha)
    r = (b / alpha) * (p - 1.0)
    inv = 1.0 / (3.0 * n)
    loss = inv * r.sum() + (np.log(scale) + lz[:, 0]).sum() / 3.0
    gx = (d / scale ** 2) * (p / u) * w * inv
    dr_da = (-2.0 / alpha ** 2) * (p - 1.0) + (b / alpha) * p * (0.5 * np.log(u) + 0.5 * alpha * (q / b ** 2) / u)
    dr_ds = -(q / scale) * (p / u)
    g_alpha = (inv * dr_da.sum(0) + lz[:, 1] / 3.0) * dal
    g_scale = (inv * dr_ds.sum(0) + (1.0 / 3.0) / scale) * dsl
    return loss, gx, g_alpha, g_scale
