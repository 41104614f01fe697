"""SURVEY.md §8f rank 4, third part: the Approximate Kernel Model sampler (src/core/cgpcm.py:295-422) — the generator of
the toy experiment's series (data.load_akm, src/core/data.py:594-641) — on the GPU against the oracle, which integrates
the reference's pair integrands _a / _Ahh (cgpcm.py:156-158,182-184) with the restated integrate_box."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')
import cgpcm_b200
from cgpcm_b200 import AKM, Data, Session, config
from cgpcm_b200.data import load_akm
from oracle import model as om
from tests.cases import make_case


@pytest.mark.parametrize('name', ['toy_small', 'toy_acausal_model'])
def test_akm_sample_matches_oracle(name):
    c = make_case(name)
    nh = c['nh']
    rng = np.random.default_rng(4)
    t = np.sort(rng.uniform(0., 1., 19))                       # uneven inputs: all 19^2 lags are distinct
    h = c['params'][5:5 + nh] + .3 * rng.standard_normal(nh)
    e = rng.standard_normal(19)
    om.PW_DISTS_EXACT = True
    try:
        f0, K0 = om.akm_f(c['params'], c['th'], c['reg'], t, h, e, causal=c['causal'])
    finally:
        om.PW_DISTS_EXACT = False
    eng = cgpcm_b200.Engine(nh, c['nx'], causal=c['causal'])
    eng.set_data(c['t'], c['y'], c['th'], c['tx'])
    f1, K1 = eng.akm_sample(c['params'], t, h, e, reg=c['reg'], want_cov=True)
    # K = a + h^T Ahh h - tr(iKh Ahh): the trace sums nh^2 products with iKh ~ 1/reg = 1e6 times larger than the
    # result, so two correct FP64 evaluations differ by ~eps nh^2 / reg relative (measured 1.3e-7 at nh = 21)
    tol = 1e-6 * np.abs(K0).max()
    assert np.abs(K1 - K0).max() <= tol
    assert np.abs(K1 - K1.T).max() <= tol
    # f = sqrt(s2_f) chol(K) e inherits K's error through the Cholesky factor: compare through the GPU's own K too
    f_from_k1 = np.sqrt(np.exp(c['params'][1])) * (np.linalg.cholesky(K1) @ e)
    assert np.abs(f1 - f_from_k1).max() <= 1e-10 * np.abs(f_from_k1).max()
    # ... and the Cholesky factor amplifies K's difference by cond(K)
    assert np.abs(f1 - f0).max() <= 1e-6 * np.linalg.cond(K0) * np.abs(f0).max()
    assert eng.akm_sample(c['params'], t, h, e, reg=c['reg']).shape == (19,)
    with pytest.raises(ValueError):
        eng.akm_sample(c['params'], t, h[:-1], e, reg=c['reg'])
    with pytest.raises((ValueError, RuntimeError)):
        eng.akm_sample(c['params'], t, h, np.full(19, np.nan), reg=c['reg'])


@pytest.mark.parametrize('causal', [True, False])
def test_load_akm_api(causal):
    """data.load_akm(sess, causal, n, nh, tau_w, tau_f, resample): the toy experiment's generator, normalised like the
    reference (zero mean / unit std series, unit-energy filter, unit-maximum kernel)."""
    config.reg = 1e-6
    np.random.seed(1005 if causal else 1030)
    try:
        f, k, h = load_akm(Session(), causal=causal, n=80, nh=21, tau_w=.05, tau_f=.025, resample=1)  # toy.py x data_scale
    finally:
        config.reg = 1e-8
    assert f.x.shape == (80,) and np.all(np.isfinite(f.y))
    assert abs(f.mean) < 1e-12 and f.std == pytest.approx(1.0)
    assert k.x.shape == (301,) and k.max == pytest.approx(1.0)
    assert k.y[150] == pytest.approx(1.0, abs=1e-6)                      # a kernel peaks at lag 0 ...
    assert np.abs(k.y - k.y[::-1]).max() <= 1e-6                         # ... and is even
    assert h.energy == pytest.approx(1.0)
    assert (h.x.min() >= 0) if causal else (h.x.min() < 0)


def test_akm_methods_and_k_prior():
    config.reg = 1e-6
    np.random.seed(3)
    try:
        akm = AKM.from_recipe(sess=Session(), e=Data(np.linspace(0, 1, 30)), nx=0, nh=15, tau_w=.1, tau_f=.05, causal=True)
        assert akm.nh == 15 and np.size(akm.tx) == 0
        hd = np.linspace(-1, 1, 15)
        akm.sample(np.linspace(0, 1, 30), h=hd)
        np.testing.assert_array_equal(akm.h_draw.ravel(), hd)
        f1 = akm.f()
        f2 = akm.f()
        np.testing.assert_array_equal(f1.y, f2.y)                        # same draws -> same function
        akm.sample_f(np.linspace(0, 1, 30))
        assert np.abs(akm.f().y - f1.y).max() > 0                        # new e -> new function
        mu, lo, up = akm.k_prior(np.linspace(-.4, .4, 41), iters=40, granularity=10)
        assert mu.x.shape == (41,) and len(lo) == len(up) == 3
        assert all(np.all(l.y <= u.y + 1e-15) for l, u in zip(lo, up))
        with pytest.raises(ValueError):
            cgpcm_b200.VCGPCM.from_recipe(Session(), Data(np.linspace(0, 1, 30)), nx=0, nh=15, tau_w=.1, tau_f=.05,
                                          causal=True)
    finally:
        config.reg = 1e-8
