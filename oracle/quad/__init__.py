"""ctypes front-end of oracle/quad/libelbo_quad.so: the VCGPCM ELBO, its 7 terms and directional derivatives in IEEE
binary128 (libquadmath).  TEST INFRASTRUCTURE ONLY: the arbiter between FP64 implementations at ill-conditioned
(trained) points; see elbo_quad.c."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, 'libelbo_quad.so')
_lib = None


def build(force=False):
    src = os.path.join(_HERE, 'elbo_quad.c')
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(['bash', os.path.join(_HERE, 'build.sh')], stdout=subprocess.DEVNULL)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB)
        dp, i32, dbl = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
        L.elbo_quad.argtypes = [i32, dp, dp, i32, dp, i32, dp, dp, dbl, i32, dp, dp]
        L.elbo_quad_dderiv.argtypes = [i32, dp, dp, i32, dp, i32, dp, dp, dp, dbl, dbl, i32, i32, dp]
        L.bvn_quad.argtypes = [dp, dp, dp, ctypes.c_long, dp]
        L.bvn_quad.restype = None
        _lib = L
    return _lib


def _c(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def elbo(params, t, y, th, tx, reg, causal=True):
    """(elbo, terms[7]) rounded to double from binary128, plus the low parts: ``(e_hi, e_lo, t_hi[7], t_lo[7])``."""
    t, y, th, tx, params = _c(t), _c(y), _c(th), _c(tx), _c(params)
    e, tm = np.zeros(2), np.zeros(14)
    rc = lib().elbo_quad(len(t), t.ctypes.data, y.ctypes.data, len(th), th.ctypes.data, len(tx), tx.ctypes.data,
                         params.ctypes.data, float(reg), int(bool(causal)), e.ctypes.data, tm.ctypes.data)
    if rc:
        raise RuntimeError('elbo_quad: Cholesky failed (code %d)' % rc)
    return e[0], e[1], tm[0::2].copy(), tm[1::2].copy()


def dderiv(params, direction, t, y, th, tx, reg, causal=True, h=1e-9, richardson=False):
    """Directional derivative of the ELBO along ``direction`` by central differences in binary128 (the function is
    smooth at this scale: D(1e-8) and D(1e-10) agree to 17 digits, h < 1e-14 only adds round-off):
    ``(value, D(h), D(2h))``; with ``richardson`` the value is ``(4 D(h) - D(2h)) / 3`` (4 evaluations instead of 2)."""
    t, y, th, tx, params, direction = _c(t), _c(y), _c(th), _c(tx), _c(params), _c(direction)
    out = np.zeros(3)
    rc = lib().elbo_quad_dderiv(len(t), t.ctypes.data, y.ctypes.data, len(th), th.ctypes.data, len(tx),
                                tx.ctypes.data, params.ctypes.data, direction.ctypes.data, float(h), float(reg),
                                int(bool(causal)), int(bool(richardson)), out.ctypes.data)
    if rc:
        raise RuntimeError('elbo_quad_dderiv: Cholesky failed (code %d)' % rc)
    return out[0], out[1], out[2]


def bvn_cdf(x, y, rho):
    x, y, rho = _c(x).ravel(), _c(y).ravel(), _c(rho).ravel()
    out = np.empty_like(x)
    lib().bvn_quad(x.ctypes.data, y.ctypes.data, rho.ctypes.data, len(x), out.ctypes.data)
    return out
