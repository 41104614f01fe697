"""Bivariate normal CDF (oracle, numpy, float64).  TEST INFRASTRUCTURE ONLY.

Restates the published algorithm of A. Genz, "Numerical computation of rectangular bivariate and
trivariate normal and t probabilities", Statistics and Computing 14 (2004), routine BVND — the
algorithm behind R ``pbivnorm`` which the reference's external op ``wesselb/bvn-cdf`` wraps
(``doc/paper.tex:341``; loaded at ``src/core/tf_util.py:9-13``; the single call site is
``src/core/exponentiated_quadratic.py:552``: ``bvn_cdf(x1, x2, rho)`` on three flat FP64 vectors).
``wesselb/bvn-cdf`` is not vendored and not pinned by the reference (no commit, no version).
"""
import numpy as np
from scipy.special import erfc

_TWO_PI = 2.0 * np.pi


def _gauss_legendre(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return x, w


_GL = {6: _gauss_legendre(6), 12: _gauss_legendre(12), 20: _gauss_legendre(20)}


def phid(z):
    """Standard normal CDF."""
    return 0.5 * erfc(-np.asarray(z, dtype=np.float64) / np.sqrt(2.0))


def gl_order(r):
    """Genz's choice of Gauss-Legendre order from the correlation."""
    r = abs(float(r))
    return 6 if r < 0.3 else (12 if r < 0.75 else 20)


def _bvnd_scalar_rho(dh, dk, r):
    """P(X > dh, Y > dk) for arrays ``dh, dk`` and ONE correlation ``r`` (Genz BVND)."""
    h = np.asarray(dh, dtype=np.float64).copy()
    k = np.asarray(dk, dtype=np.float64).copy()
    x, w = _GL[gl_order(r)]
    hk = h * k
    if abs(r) < 0.925:
        bvn = np.zeros_like(h)
        if abs(r) > 0:
            hs = 0.5 * (h * h + k * k)
            asr = np.arcsin(r)
            for xj, wj in zip(x, w):
                sn = np.sin(asr * (xj + 1.0) / 2.0)
                bvn += wj * np.exp((sn * hk - hs) / (1.0 - sn * sn))
            bvn *= asr / (2.0 * _TWO_PI)
        return bvn + phid(-h) * phid(-k)

    if r < 0:
        k = -k
        hk = -hk
    bvn = np.zeros_like(h)
    if abs(r) < 1:
        a_s = (1.0 - r) * (1.0 + r)
        a = np.sqrt(a_s)
        bs = (h - k) ** 2
        c = (4.0 - hk) / 8.0
        d = (12.0 - hk) / 16.0
        asr = -(bs / a_s + hk) / 2.0
        with np.errstate(over='ignore', under='ignore', invalid='ignore'):
            t = a * np.exp(asr) * (1.0 - c * (bs - a_s) * (1.0 - d * bs / 5.0) / 3.0
                                   + c * d * a_s * a_s / 5.0)
        bvn = np.where(asr > -100.0, t, 0.0)
        b = np.sqrt(bs)
        with np.errstate(over='ignore', under='ignore', invalid='ignore'):
            t = np.exp(-hk / 2.0) * np.sqrt(_TWO_PI) * phid(-b / a) * b \
                * (1.0 - c * bs * (1.0 - d * bs / 5.0) / 3.0)
        bvn = bvn - np.where(-hk < 100.0, t, 0.0)
        a = a / 2.0
        for xj, wj in zip(x, w):
            xs = (a * (xj + 1.0)) ** 2
            rs = np.sqrt(1.0 - xs)
            asr = -(bs / xs + hk) / 2.0
            with np.errstate(over='ignore', under='ignore', invalid='ignore'):
                t = a * wj * np.exp(asr) * (np.exp(-hk * xs / (2.0 * (1.0 + rs) ** 2)) / rs
                                            - (1.0 + c * xs * (1.0 + d * xs)))
            bvn = bvn + np.where(asr > -100.0, t, 0.0)
        bvn = -bvn / _TWO_PI
    if r > 0:
        return bvn + phid(-np.maximum(h, k))
    bvn = -bvn
    extra = np.where(h < 0, phid(k) - phid(h), phid(-h) - phid(-k))
    return bvn + np.where(k > h, extra, 0.0)


def bvnd(dh, dk, r):
    """P(X > dh, Y > dk) with correlation ``r`` (array-valued ``r`` is grouped by value)."""
    dh, dk, r = np.broadcast_arrays(np.asarray(dh, np.float64), np.asarray(dk, np.float64),
                                    np.asarray(r, np.float64))
    out = np.empty(dh.shape, dtype=np.float64)
    flat_r = r.ravel()
    oh, ok, oo = dh.ravel(), dk.ravel(), out.ravel()
    for rv in np.unique(flat_r):
        sel = flat_r == rv
        oo[sel] = _bvnd_scalar_rho(oh[sel], ok[sel], float(rv))
    return oo.reshape(dh.shape)


def bvn_cdf(x1, x2, rho):
    """Phi_2(x1, x2; rho) = P(X <= x1, Y <= x2): same contract as the reference's ``bvn_cdf`` op
    (``src/core/exponentiated_quadratic.py:547-552``)."""
    return bvnd(-np.asarray(x1, np.float64), -np.asarray(x2, np.float64), rho)


def bvn_cdf_partials(x1, x2, rho):
    """(dPhi2/dx1, dPhi2/dx2, dPhi2/drho) — the gradient the external op registers."""
    x1 = np.asarray(x1, np.float64)
    x2 = np.asarray(x2, np.float64)
    rho = np.asarray(rho, np.float64)
    om = 1.0 - rho * rho
    s = np.sqrt(om)
    pdf = lambda z: np.exp(-0.5 * z * z) / np.sqrt(_TWO_PI)
    d1 = pdf(x1) * phid((x2 - rho * x1) / s)
    d2 = pdf(x2) * phid((x1 - rho * x2) / s)
    dr = np.exp(-(x1 * x1 - 2.0 * rho * x1 * x2 + x2 * x2) / (2.0 * om)) / (_TWO_PI * s)
    return d1, d2, dr
