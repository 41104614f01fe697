"""The oracle's closed-form Psi statistics against (i) the reference's own
integrands pushed through the restated integrate_box on the reference's 6-D layout and (ii) direct
quadrature; the oracle's gradient against central finite differences."""
import numpy as np
import pytest
import torch
from scipy import integrate

from oracle import model as om
from tests.cases import make_case


@pytest.mark.parametrize('name', ['toy_small', 'toy_acausal_model', 'sweep'])
def test_closed_forms_equal_generic_integrals(name):
    c = make_case(name, n=7)
    hyp = [torch.tensor(v, dtype=torch.float64) for v in c['hyp']]
    a0, Ahh0, Axx0, Ahx0 = om.psi_generic(c['t'], c['th'], c['tx'], *hyp, causal=c['causal'])
    a1, Ahh1, Axx1, Ahx1 = om.psi_closed(c['t'], c['th'], c['tx'], *hyp, causal=c['causal'])
    assert abs(float(a0) - float(a1)) < 1e-14
    np.testing.assert_allclose(Ahh0.numpy(), Ahh1.numpy(), atol=1e-14, rtol=1e-12)
    np.testing.assert_allclose(Axx0.numpy(), Axx1.numpy(), atol=1e-14, rtol=1e-10)
    np.testing.assert_allclose(Ahx0.numpy(), Ahx1.numpy(), atol=1e-14, rtol=1e-11)


def test_closed_forms_equal_quadrature():
    c = make_case('toy_small', n=5)
    al, ga, om_ = c['hyp']
    kh = lambda x, y: np.exp(-al * (x * x + y * y) - ga * (x - y) ** 2)
    kxs = lambda x, y: np.exp(-om_ * (x - y) ** 2)
    a, Ahh, Axx, Ahx = [np.asarray(v) for v in om.psi_closed(c['t'], c['th'], c['tx'], *c['hyp'])]
    t = c['t'][3]
    i, j, k, l = 4, 6, 9, 11
    thi, thj, txk, txl = c['th'][i], c['th'][j], c['tx'][k], c['tx'][l]
    lo = t - 2.0
    v, _ = integrate.quad(lambda s: kh(t - s, t - s), lo, t, epsabs=1e-15, epsrel=1e-13, limit=400)
    assert abs(v - float(a)) < 1e-12
    v, _ = integrate.quad(lambda s: kh(t - s, thi) * kh(thj, t - s), lo, t, epsabs=1e-16, epsrel=1e-13, limit=400)
    assert abs(v - Ahh[i, j]) < 1e-13
    v, _ = integrate.quad(lambda s: kh(t - s, thi) * kxs(s, txk), lo, t, epsabs=1e-16, epsrel=1e-13, limit=400)
    assert abs(v - Ahx[3, i, k]) < 1e-13
    v, _ = integrate.dblquad(lambda s2, s1: kh(t - s1, t - s2) * kxs(s1, txk) * kxs(txl, s2), lo, t, lo, t,
                             epsabs=1e-14, epsrel=1e-11)
    assert abs(v - Axx[3, k, l]) < 1e-11


@pytest.mark.parametrize('name', ['toy_small', 'sweep'])
def test_gradient_against_finite_differences(name):
    c = make_case(name, n=20)
    e0, terms, g = om.elbo_and_grad(c['params'], c['t'], c['y'], c['th'], c['tx'], c['reg'], c['causal'])
    assert abs(terms.sum() - e0) < 1e-9 * abs(e0)
    rng = np.random.default_rng(1)
    idx = list(range(5)) + list(rng.choice(np.arange(5, len(g)), size=6, replace=False))
    for i in idx:
        h = 1e-6
        p1, p2 = c['params'].copy(), c['params'].copy()
        p1[i] += h
        p2[i] -= h
        f1 = om.elbo_and_grad(p1, c['t'], c['y'], c['th'], c['tx'], c['reg'], c['causal'])[0]
        f2 = om.elbo_and_grad(p2, c['t'], c['y'], c['th'], c['tx'], c['reg'], c['causal'])[0]
        fd = (f1 - f2) / (2 * h)
        assert abs(fd - g[i]) < 2e-5 * max(1.0, abs(g[i])), (i, fd, g[i])


def test_frozen_regime_matches_full_at_freeze_point():
    c = make_case('toy_small')
    full = om.elbo_and_grad(c['params'], c['t'], c['y'], c['th'], c['tx'], c['reg'])
    fr = om.precompute(c['params'], c['t'], c['y'], c['th'], c['tx'], c['reg'])
    froz = om.elbo_and_grad(c['params'], c['t'], c['y'], c['th'], c['tx'], c['reg'], frozen=fr)
    assert abs(full[0] - froz[0]) < 1e-10 * abs(full[0])
    np.testing.assert_allclose(full[2][:2], froz[2][:2], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(full[2][5:], froz[2][5:], rtol=1e-9, atol=1e-8)
    assert np.all(froz[2][2:5] == 0)


def test_fpi_is_coordinate_ascent():
    """oracle.fpi (src/core/cgpcm.py:479-516): every round of the fixed-point iteration raises the saturated ELBO,
    and convert's q(z) is the optimal q(z) of the final q(u) (its mean solves reg(P) mz = lam)."""
    c = make_case('toy_small')
    args = (c['t'], c['y'], c['th'], c['tx'], c['reg'])
    e = [om.elbo_and_grad(c['params'], *args)[0]]
    for num in (1, 2, 5):
        mu, var, mz, vz = om.fpi(c['params'], *args, num)
        p = c['params'].copy()
        p[5:5 + c['nh']] = mu
        p[5 + c['nh']:] = var
        e.append(om.elbo_and_grad(p, *args)[0])
    assert e[0] < e[1] < e[2] < e[3]
    # num = 0: q(u) is returned unchanged (up to the Cholesky round trip of reg(L L^T))
    mu0, var0, _, _ = om.fpi(c['params'], *args, 0)
    np.testing.assert_allclose(mu0, c['params'][5:5 + c['nh']], rtol=0, atol=1e-15)
    L = om.vec_to_tril(om.T(c['params'][5 + c['nh']:])).numpy()
    L0 = om.vec_to_tril(om.T(var0)).numpy()
    np.testing.assert_allclose(L0 @ L0.T, L @ L.T + c['reg'] * np.eye(c['nh']), rtol=1e-12, atol=1e-14)


def test_predict_f_oracle_is_a_smoother():
    """oracle.predict_f (src/core/cgpcm.py:781-846): after a few fixed-point rounds the predictive mean follows the
    observations, the variance is positive, and the SMF variant with the mean of q(u) as the only sample equals
    the plug-in formulas evaluated by hand."""
    c = make_case('toy_small')
    args = (c['t'], c['y'], c['th'], c['tx'], c['reg'])
    mu, var, _, _ = om.fpi(c['params'], *args, 8)
    p = c['params'].copy()
    p[5:5 + c['nh']] = mu
    p[5 + c['nh']:] = var
    m, v = om.predict_f(p, *args, c['t'], [mu], smf=False)
    assert np.all(v > 0) and np.mean((m - c['y']) ** 2) < 0.6
    m2, v2 = om.predict_f(p, *args, c['t'][:5], [mu, mu], smf=True)
    m1, v1 = om.predict_f(p, *args, c['t'][:5], [mu], smf=True)
    np.testing.assert_allclose(m2, m1, rtol=1e-13)
    np.testing.assert_allclose(v2, v1, rtol=1e-12)


@pytest.mark.parametrize('causal', [True, False])
def test_center_statistics_closed_forms(causal):
    """oracle.psi_center_closed against the reference's integrands pushed through the restated integrate_box
    (``_a_center`` / ``_Ahh_center``, src/core/cgpcm.py:164-166,190-192), and against the diagonal forms at t = 0."""
    t = np.array([-.3, -.05, 0., .02, .11, .4])
    th = np.linspace(-.02, .2, 7)
    alpha, gamma = 39.27, 294.5
    a1, A1 = om.psi_center_generic(t, th, alpha, gamma, causal)
    a2, A2 = om.psi_center_closed(t, th, alpha, gamma, causal)
    assert float((a1 - a2).abs().max()) < 1e-16 and float((A1 - A2).abs().max()) < 1e-15
    assert float(a2[2]) == pytest.approx(float(om.psi_a(om.T(alpha), causal)), rel=1e-15)
    np.testing.assert_allclose(A2[2].numpy(), om.psi_Ahh(th, om.T(alpha), om.T(gamma), causal).numpy(), rtol=1e-13, atol=1e-17)


@pytest.mark.parametrize('causal', [True, False])
def test_pair_statistics_depend_on_the_lag_only(causal):
    """The AKM's pair statistics ``_a(t)`` / ``_Ahh(t)`` (src/core/cgpcm.py:156-158,182-184; upper limit min(t1, t2)),
    integrated from the reference's integrands by the restated integrate_box, equal the centre statistics
    (:164-166,190-192) at the lag ``t1 - t2``: what ``cgpcm_akm_sample`` relies on."""
    t = np.array([0., .1, .35, .5, .72, .73])
    th = np.linspace(-.05, .2, 7)
    a, Ahh = om.psi_pairs_generic(t, th, 5., 30., causal)
    lags = (t[:, None] - t[None, :]).ravel()
    ac, Ahhc = om.psi_center_closed(lags, th, 5., 30., causal)
    assert float(abs(a.reshape(-1) - ac).max()) <= 1e-14
    assert float(abs(Ahh.reshape(-1, 7, 7) - Ahhc).max()) <= 1e-14
    # and the covariance of AKM.f is symmetric positive definite
    p = np.zeros(5 + 7 + 28)
    p[:5] = np.log([.1, 1.3, 5., 30., 1.])
    rng = np.random.default_rng(0)
    f, K = om.akm_f(p, th, 1e-6, t, 3 * rng.standard_normal(7), rng.standard_normal(6), causal=causal)
    # symmetric up to the rounding of tr(iKh Ahh), iKh ~ 1/reg
    assert np.abs(K - K.T).max() <= 1e-8 * np.abs(K).max() and np.linalg.eigvalsh(.5 * (K + K.T)).min() > 0
    assert f.shape == (6,) and np.all(np.isfinite(f))
