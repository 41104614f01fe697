"""The reference's only known-answer tests (src/core/exponentiated_quadratic_test.py:24-42) restated
against oracle.expq: they pin integrate_box / integrate_half (incl. the BVN path) to 5-7 decimals."""
import numpy as np

from oracle import expq
from oracle.expq import EQ, const, var, inf


def _exp1():
    t1, t2, t3 = var('t1'), var('t2'), var('t3')
    return EQ(-const(1) * t1 ** 2 - const(2) * t2 ** 2 - const(.5) * t1 * t2 - const(2) * t1 * t3
              + const(3) * t2 + const(4))


def test1_quadrant():
    ref = np.array([[55.81808295, 11.76773162], [11.76773162, 55.81808295]])
    res = _exp1().integrate_box(('t1', -inf, 0), ('t2', -inf, 0), t3=np.eye(2))
    np.testing.assert_almost_equal(res, ref, decimal=6)


def test2_box():
    ref = np.array([[217.3921457, 318.3540954], [318.3540954, 217.3921457]])
    res = _exp1().integrate_box(('t1', const(-1), const(2)), ('t2', var('t3'), const(3)), t3=np.eye(2))
    np.testing.assert_almost_equal(res, ref, decimal=5)


def test3_half_line():
    t1 = var('t1')
    exp2 = EQ(const(-1) * t1 ** 2 + const(-.5) * t1 + const(4))
    res = exp2.integrate_half(['t1'])
    np.testing.assert_almost_equal(res, 65.73974603)


def test_golden_against_quadrature():
    """The same three integrals by direct numerical quadrature (the goldens are only ~1e-7 accurate)."""
    from scipy import integrate
    f = lambda t2, t1, t3: np.exp(-t1 ** 2 - 2 * t2 ** 2 - .5 * t1 * t2 - 2 * t1 * t3 + 3 * t2 + 4)
    for t3, want in [(1., 55.81808295), (0., 11.76773162)]:
        val, _ = integrate.dblquad(f, -12, 0, -12, 0, args=(t3,), epsabs=1e-11, epsrel=1e-11)
        got = _exp1().integrate_box(('t1', -inf, 0), ('t2', -inf, 0), t3=np.array(t3))
        assert abs(val - want) < 5e-6
        assert abs(got - val) < 1e-8 * val
    val, _ = integrate.quad(lambda t: np.exp(-t ** 2 - .5 * t + 4), -15, 0, epsabs=1e-12, epsrel=1e-13)
    t1 = var('t1')
    got = EQ(const(-1) * t1 ** 2 + const(-.5) * t1 + const(4)).integrate_half(['t1'])
    assert abs(got - val) < 1e-11 * val
