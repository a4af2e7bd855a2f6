#!/usr/bin/env python
"""Fits the FP32 approximations used by the production probit site pass (csrc/samplers.cuh: phi_pair_fast,
ndtri_fast) and prints their coefficients and error bounds.  Run once on the host; the numbers are pasted into the
kernel source and re-checked on the device by tests/test_gpu_primitives.py.

  erfc(z)  = t exp(-z^2 + P(t)),  t = 1 / (1 + z / 2),  z >= 0          (form of Numerical Recipes' erfcc)
  -ndtri(q), q <= 0.075:  rational in r = sqrt(-ln q)
  ndtri(p), |p - 1/2| <= 0.425:  (p - 1/2) A(s) / B(s),  s = 0.180625 - (p - 1/2)^2   (form of Wichura's AS241 PPND7)
"""
import numpy as np
from scipy.special import erfcx, ndtri
from numpy.polynomial import chebyshev as C, polynomial as P


def cheb_nodes(a, b, m):
    k = np.arange(m)
    return 0.5 * (a + b) + 0.5 * (b - a) * np.cos(np.pi * (k + 0.5) / m)


def fit_rational(x, y, dn, dd, iters=30):
    """y ~ N(x)/D(x), D(0)=1, relative-error weighted, iteratively reweighted (Loeb) least squares."""
    w = 1.0 / np.abs(y)
    d = np.ones_like(x)
    for _ in range(iters):
        A = np.concatenate([np.vander(x, dn + 1, increasing=True), -(y[:, None]) * np.vander(x, dd + 1, increasing=True)[:, 1:]], axis=1)
        sol, *_ = np.linalg.lstsq(A * (w / d)[:, None], y * (w / d), rcond=None)
        num, den = sol[:dn + 1], np.concatenate([[1.0], sol[dn + 1:]])
        d = np.abs(P.polyval(x, den))
    return num, den


def f32_eval_rational(x, num, den):
    x = x.astype(np.float32)
    n = np.float32(num[-1]) * np.ones_like(x)
    for c in num[-2::-1]:
        n = n * x + np.float32(c)
    d = np.float32(den[-1]) * np.ones_like(x)
    for c in den[-2::-1]:
        d = d * x + np.float32(c)
    return n / d


def main():
    # ---- erfc ----
    z = np.concatenate([np.linspace(0, 6, 20001), np.linspace(6, 40, 4001)])
    t = 1.0 / (1.0 + 0.5 * z)
    target = np.log(erfcx(z) / t)
    for deg in (7, 8, 9, 10):
        coef = P.polyfit(t, target, deg)
        err = np.abs(P.polyval(t, coef) - target).max()
        print(f'erfc: degree {deg}: max |log error| = {err:.3e}')
    deg = 9
    coef = P.polyfit(t, target, deg)
    print('erfc P(t) coefficients (low -> high):', ', '.join(f'{c:.9e}f' for c in coef))
    # float32 evaluation of the whole formula
    z32 = z.astype(np.float32)
    t32 = np.float32(1) / (np.float32(1) + np.float32(0.5) * z32)
    p32 = np.float32(coef[-1]) * np.ones_like(t32)
    for c in coef[-2::-1]:
        p32 = p32 * t32 + np.float32(c)
    from scipy.special import erfc
    val = t32 * np.exp((p32 - z32 * z32).astype(np.float32))
    m = z <= 6.0
    print('erfc FP32 evaluation: max relative error (z <= 6):', np.abs(val[m] / erfc(z32[m].astype(np.float64)) - 1).max())
    # ---- ndtri tail ----
    q = np.exp(-np.linspace(np.sqrt(-np.log(0.075)), 6.6, 40001) ** 2)
    r = np.sqrt(-np.log(q))
    y = -ndtri(q)
    for dn, dd in ((3, 2), (3, 3), (4, 3), (4, 4)):
        num, den = fit_rational(r, y, dn, dd)
        e64 = np.abs(P.polyval(r, num) / P.polyval(r, den) / y - 1).max()
        e32 = np.abs(f32_eval_rational(r, num, den) / y - 1).max()
        print(f'ndtri tail ({dn},{dd}): max rel err fp64 {e64:.3e}  fp32 {e32:.3e}')
    num, den = fit_rational(r, y, 3, 3)
    print('tail num:', ', '.join(f'{c:.9e}f' for c in num))
    print('tail den:', ', '.join(f'{c:.9e}f' for c in den))
    # ---- ndtri central ----
    d = np.linspace(1e-9, 0.425, 40001)
    s = 0.180625 - d * d
    yc = ndtri(0.5 + d) / d
    for dn, dd in ((2, 2), (3, 2), (3, 3)):
        num, den = fit_rational(s, yc, dn, dd)
        e64 = np.abs(P.polyval(s, num) / P.polyval(s, den) / yc - 1).max()
        e32 = np.abs(f32_eval_rational(s, num, den) / yc - 1).max()
        print(f'ndtri central ({dn},{dd}): max rel err fp64 {e64:.3e}  fp32 {e32:.3e}')
    num, den = fit_rational(s, yc, 3, 3)
    print('central num:', ', '.join(f'{c:.9e}f' for c in num))
    print('central den:', ', '.join(f'{c:.9e}f' for c in den))


if __name__ == '__main__':
    main()
