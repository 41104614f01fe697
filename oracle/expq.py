"""Gaussian integrals of exponentiated quadratic forms (oracle).  TEST INFRASTRUCTURE ONLY.

Restates the *algorithm* of the reference's ``src/core/exponentiated_quadratic.py`` — collect the
quadratic / linear / constant coefficients of the integration variables, then apply the closed forms

* whole line   (``EQ._integrate``,        ``exponentiated_quadratic.py:490-498``),
* half line    (``EQ._integrate_half1``,  ``:500-509``)  ->  ``exp * (1 - erf)``,
* quadrant     (``EQ._integrate_half2``,  ``:511-559``)  ->  ``exp * 2 pi / sqrt(det) * Phi_2``,
* boxes by shifting every finite limit to 0 and inclusion-exclusion (``EQ.integrate_box``,
  ``:430-469``; ``translate_var`` ``:471-481``)

on a data structure of its own: a polynomial is a ``dict`` that maps a monomial (a sorted tuple of
``(variable, power)`` pairs) to a coefficient.  Coefficients and the values bound to free variables
may be python floats, numpy arrays or torch tensors (float64); torch values make every integral
differentiable, with the bivariate normal CDF supplied by :class:`oracle.bvn_torch.BvnCdf`.

Pinned by the reference's own known-answer tests (``exponentiated_quadratic_test.py:24-42``) in
``tests/test_oracle_golden.py``.
"""
import math

import numpy as np

try:  # torch is optional for the numpy-only uses of this module
    import torch
except Exception:  # pragma: no cover
    torch = None

from . import bvn as _bvn

inf = math.inf


# ----------------------------------------------------------------------------- backend shim
def _is_torch(*xs):
    return torch is not None and any(isinstance(x, torch.Tensor) for x in xs)


def _as_torch(x):
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(x, dtype=torch.float64)


def _exp(x):
    return torch.exp(x) if _is_torch(x) else np.exp(x)


def _sqrt(x):
    return torch.sqrt(x) if _is_torch(x) else np.sqrt(x)


def _erf(x):
    if _is_torch(x):
        return torch.erf(x)
    from scipy.special import erf
    return erf(x)


def _squeeze(x):
    """``tf.squeeze`` of the reference (``exponentiated_quadratic.py:398,507,559``)."""
    if _is_torch(x):
        return torch.squeeze(x)
    return np.squeeze(np.asarray(x))


def _bvn_cdf(x1, x2, rho):
    if _is_torch(x1, x2, rho):
        from .bvn_torch import bvn_cdf as tb
        x1, x2, rho = _as_torch(x1), _as_torch(x2), _as_torch(rho)
        x1, x2, rho = torch.broadcast_tensors(x1, x2, rho)
        return tb(x1, x2, rho)
    return _bvn.bvn_cdf(x1, x2, rho)


# ----------------------------------------------------------------------------- polynomials
def _mono_mul(m1, m2):
    powers = dict(m1)
    for v, p in m2:
        powers[v] = powers.get(v, 0) + p
    return tuple(sorted((v, p) for v, p in powers.items() if p != 0))


class Poly(object):
    """Sparse multivariate polynomial: ``{monomial: coefficient}``."""

    __slots__ = ('terms',)

    def __init__(self, terms=None):
        self.terms = dict(terms) if terms else {}

    # construction helpers
    @staticmethod
    def const(c):
        return Poly({(): c})

    @staticmethod
    def var(name, power=1):
        return Poly({((name, power),): 1.0})

    # algebra
    def __add__(self, other):
        other = _poly(other)
        out = dict(self.terms)
        for m, c in other.terms.items():
            out[m] = out[m] + c if m in out else c
        return Poly(out)

    __radd__ = __add__

    def __neg__(self):
        return Poly({m: -c for m, c in self.terms.items()})

    def __sub__(self, other):
        return self + (-_poly(other))

    def __rsub__(self, other):
        return _poly(other) + (-self)

    def __mul__(self, other):
        other = _poly(other)
        out = {}
        for m1, c1 in self.terms.items():
            for m2, c2 in other.terms.items():
                m = _mono_mul(m1, m2)
                out[m] = out[m] + c1 * c2 if m in out else c1 * c2
        return Poly(out)

    __rmul__ = __mul__

    def __pow__(self, power):
        if not isinstance(power, int) or power < 0:
            raise RuntimeError('can only raise to nonnegative integers')
        out = Poly.const(1.0)
        for _ in range(power):
            out = out * self
        return out

    # structure
    def substitute(self, name, poly):
        """Replace variable ``name`` by the polynomial ``poly``."""
        poly = _poly(poly)
        out = Poly()
        for m, c in self.terms.items():
            rest = tuple((v, p) for v, p in m if v != name)
            power = sum(p for v, p in m if v == name)
            out = out + Poly({rest: c}) * poly ** power
        return out

    def coefficient(self, **powers):
        """Polynomial multiplying ``prod_v v**powers[v]`` *exactly* (the variables named in
        ``powers`` must occur with exactly that power; power 0 means "does not occur")."""
        out = {}
        for m, c in self.terms.items():
            d = dict(m)
            if all(d.get(v, 0) == p for v, p in powers.items()):
                rest = tuple((v, p) for v, p in m if v not in powers)
                out[rest] = out[rest] + c if rest in out else c
        return Poly(out)

    def is_constant(self):
        return all(m == () for m in self.terms)

    def eval(self, **var_map):
        total = 0.0
        for m, c in self.terms.items():
            val = c
            for v, p in m:
                val = val * var_map[v] ** p
            total = total + val
        return total


def _poly(x):
    return x if isinstance(x, Poly) else Poly.const(x)


const = Poly.const
var = Poly.var


def _is_inf(x):
    return isinstance(x, (int, float)) and math.isinf(x)


# ----------------------------------------------------------------------------- exp(quadratic)
class EQ(object):
    """``const * exp(poly)``."""

    def __init__(self, poly, const=1.0):
        self.poly = _poly(poly)
        self.const = const

    def __mul__(self, other):
        return EQ(self.poly + other.poly, self.const * other.const)

    def __neg__(self):
        return EQ(self.poly, -self.const)

    def substitute(self, name, poly):
        return EQ(self.poly.substitute(name, poly), self.const)

    def translate(self, name, shift):
        """Substitute ``name -> name + shift`` (moves the integration limit ``shift`` to 0)."""
        return EQ(self.poly.substitute(name, var(name) + _poly(shift)), self.const)

    def eval(self, **var_map):
        return _squeeze(self.const * _exp(self.poly.eval(**var_map)))

    # -- whole line
    def integrate_line(self, name):
        a = self.poly.coefficient(**{name: 2})
        b = self.poly.coefficient(**{name: 1})
        c = self.poly.coefficient(**{name: 0})
        if not a.is_constant():
            raise ValueError('quadratic coefficient must be constant')
        a = a.eval()
        return EQ(const(-.25 / a) * b ** 2 + c, self.const * (-math.pi / a) ** .5)

    # -- (-inf, 0]
    def integrate_half(self, names, **var_map):
        if len(names) == 1:
            return self._half1(names[0], **var_map)
        if len(names) == 2:
            return self._half2(names[0], names[1], **var_map)
        raise NotImplementedError()

    def _half1(self, name, **var_map):
        a = self.poly.coefficient(**{name: 2})
        b = self.poly.coefficient(**{name: 1})
        c = self.poly.coefficient(**{name: 0})
        if not a.is_constant():
            raise ValueError('quadratic coefficient must be constant')
        a, b, c = a.eval(**var_map), b.eval(**var_map), c.eval(**var_map)
        return _squeeze(.5 * self.const * (-math.pi / a) ** .5
                        * _exp(-.25 * b ** 2 / a + c)
                        * (1 - _erf(.5 * b / (-a) ** .5)))

    def _half2(self, n1, n2, **var_map):
        p = self.poly
        a11 = p.coefficient(**{n1: 2, n2: 0})
        a22 = p.coefficient(**{n2: 2, n1: 0})
        a12 = p.coefficient(**{n1: 1, n2: 1})
        b1 = p.coefficient(**{n1: 1, n2: 0})
        b2 = p.coefficient(**{n2: 1, n1: 0})
        c = p.coefficient(**{n1: 0, n2: 0})
        if not (a11.is_constant() and a22.is_constant() and a12.is_constant()):
            raise ValueError('quadratic coefficients must be constant')
        # exponent = -1/2 tau^T A tau + b^T tau + c
        a11, a22, a12 = -2 * a11.eval(), -2 * a22.eval(), -1 * a12.eval()
        b1, b2, c = b1.eval(**var_map), b2.eval(**var_map), c.eval(**var_map)
        det = a11 * a22 - a12 ** 2
        s11, s12, s22 = a22 / det, -a12 / det, a11 / det          # Sigma = A^{-1}
        mu1 = s11 * b1 + s12 * b2
        mu2 = s12 * b1 + s22 * b2
        x1 = -mu1 / s11 ** .5
        x2 = -mu2 / s22 ** .5
        rho = s12 / (s11 * s22) ** .5
        cdf = _bvn_cdf(x1, x2, rho)
        quad = .5 * (s11 * b1 ** 2 + s22 * b2 ** 2 + 2 * s12 * b1 * b2)
        return _squeeze(self.const * cdf * _exp(quad + c) * (2 * math.pi / det ** .5))

    # -- boxes
    def integrate_box(self, *vars_and_lims, **var_map):
        """``vars_and_lims``: triples ``(name, lower, upper)``; an infinite lower limit means
        ``-inf`` and an infinite upper limit ``+inf`` (``exponentiated_quadratic.py:430-441``)."""
        expq = self
        finite = []
        for name, lower, upper in vars_and_lims:
            if _is_inf(lower) and _is_inf(upper):
                expq = expq.integrate_line(name)
            else:
                finite.append((name, lower, upper))
        if not finite:
            return expq.eval(**var_map)
        parts = [expq]
        for name, lower, upper in finite:
            new = []
            for part in parts:
                if not _is_inf(upper):
                    new.append(part.translate(name, upper))
                if not _is_inf(lower):
                    new.append(-part.translate(name, lower))
            parts = new
        names = [name for name, _, _ in finite]
        total = 0.0
        for part in parts:
            total = total + part.integrate_half(names, **var_map)
        return total


# kernels of the model (exponentiated_quadratic.py:43-54, 71-80)
def kh(alpha, gamma, x, y):
    return EQ(-const(alpha) * (x ** 2 + y ** 2) - const(gamma) * (x - y) ** 2)


def kxs(omega, x, y):
    return EQ(-const(omega) * (x - y) ** 2)
