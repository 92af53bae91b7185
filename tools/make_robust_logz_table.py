#!/usr/bin/env python
"""Tabulates log Z(alpha) and d log Z / d alpha of Barron's general robust distribution on alpha in [0, 2].

    Z(alpha) = integral over x of exp(-rho(x, alpha, 1)),   rho = (|alpha-2| / alpha) ((x^2 / |alpha-2| + 1)^(alpha/2) - 1)

(reference: externel_lib/robust_loss_pytorch/distribution.py:143-171 approximates the same function with a cubic
spline fitted to a Meijer-G evaluation; general.py:84-118 is rho).  Here both the value and the derivative come from
adaptive quadrature in float64, so the table is this repo's own data, not a copy of the reference's resource file; the
CPU tests compare it with the reference spline (agreement ~1e-8).

    python tools/make_robust_logz_table.py     # writes <package>/data/robust_logz_table.npz
"""
import os

import numpy as np
from scipy import integrate

N = 2049  # knots on [0, 2]; cubic Hermite between them
EPS = float(np.finfo(np.float32).eps)


def rho(x, a):
    b = max(abs(a - 2.0), EPS)
    aa = max(abs(a), EPS)
    return (b / aa) * ((x * x / b + 1.0) ** (0.5 * a) - 1.0)


def drho_dalpha(x, a):
    """d rho / d alpha at scale 1 for 0 < alpha < 2 (b = 2 - alpha, db/dalpha = -1)."""
    b = 2.0 - a
    u = x * x / b + 1.0
    p = u ** (0.5 * a)
    dlead = -1.0 / a - b / (a * a)                       # d(b/a)/dalpha
    du = x * x / (b * b)                                  # du/dalpha
    dp = p * (0.5 * np.log(u) + 0.5 * a * du / u)
    return dlead * (p - 1.0) + (b / a) * dp


def main():
    alphas = np.linspace(0.0, 2.0, N)
    val = np.zeros(N)
    der = np.zeros(N)
    for i, a in enumerate(alphas):
        if a == 0.0:
            z = np.pi * np.sqrt(2.0)                      # Cauchy
        elif a == 2.0:
            z = np.sqrt(2.0 * np.pi)                      # normal
        else:
            z = 2.0 * integrate.quad(lambda x: np.exp(-rho(x, a)), 0, np.inf, limit=800, epsabs=1e-13, epsrel=1e-13)[0]
        val[i] = np.log(z)
        if 0.0 < a < 2.0:
            dz = -2.0 * integrate.quad(lambda x: drho_dalpha(x, a) * np.exp(-rho(x, a)), 0, np.inf, limit=800,
                                       epsabs=1e-12, epsrel=1e-12)[0]
            der[i] = dz / z
    # one-sided finite differences of the (smooth) value table at the two ends
    h = alphas[1] - alphas[0]
    der[0] = (-3 * val[0] + 4 * val[1] - val[2]) / (2 * h)
    der[-1] = (3 * val[-1] - 4 * val[-2] + val[-3]) / (2 * h)
    here = os.path.dirname(os.path.abspath(__file__))
    pkg = os.path.join(os.path.dirname(here), "learning-continuous-implicit-representation-for-near-periodic-patterns_b200")
    out = os.path.join(pkg, "data", "robust_logz_table.npz")
    np.savez(out, alpha_max=np.float64(2.0), values=val.astype(np.float64), derivs=der.astype(np.float64))
    # self-check: derivative table vs central differences of the value table
    cd = (val[2:] - val[:-2]) / (2 * h)
    print("wrote", out, "max |der - central diff| =", np.abs(der[1:-1] - cd).max())


if __name__ == "__main__":
    main()
