"""VCGPCM.predict_f (src/core/cgpcm.py:781-846) on the GPU against the oracle's literal
restatement, through the C-ABI (cgpcm_predict_f) and the reference-facing Python API."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')
import cgpcm_b200
from cgpcm_b200 import VCGPCM, Data, Session, config
from oracle import model as om
from tests.cases import make_case, ulp_noise


def _trained(c, rounds=6):
    om.PW_DISTS_EXACT = True
    try:
        mu, var, _, _ = om.fpi(c['params'], c['t'], c['y'], c['th'], c['tx'], c['reg'], rounds)
    finally:
        om.PW_DISTS_EXACT = False
    p = c['params'].copy()
    p[5:5 + c['nh']] = mu
    p[5 + c['nh']:] = var
    return p


@pytest.mark.parametrize('name,smf', [('toy_small', False), ('toy_small', True), ('sweep_hi', False), ('crude', True)])
def test_predict_f_matches_oracle(name, smf):
    c = make_case(name)
    p = _trained(c)
    rng = np.random.default_rng(2)
    nh = c['nh']
    L = om.vec_to_tril(om.T(p[5 + nh:])).numpy()
    Lc = np.linalg.cholesky(L @ L.T + c['reg'] * np.eye(nh))
    samples = np.stack([p[5:5 + nh] + Lc @ rng.standard_normal(nh) for _ in range(4)])
    lo, hi = c['t'].min(), c['t'].max()
    t_star = np.concatenate([np.linspace(lo, hi, 37), [lo - .05 * (hi - lo), hi + .05 * (hi - lo)]])   # incl. extrapolation
    om.PW_DISTS_EXACT = True
    try:
        (m0, v0), (mn, vn) = ulp_noise(
            lambda pp, th: om.predict_f(pp, c['t'], c['y'], th, c['tx'], c['reg'], t_star, samples, smf=smf), p, c['th'],
            trials=2)
    finally:
        om.PW_DISTS_EXACT = False
    outs = []
    for opts in (dict(), dict(chunk=32, store=0)):                     # test points in one chunk / in 32-point chunks
        eng = cgpcm_b200.Engine(c['nh'], c['nx'])
        for k, v in opts.items():
            eng.set_option(k, v)
        eng.set_data(c['t'], c['y'], c['th'], c['tx'])
        with pytest.raises(ValueError):
            eng.predict_f(p, t_star, samples, smf=smf, reg=c['reg'])   # needs the frozen statistics
        eng.precompute(*c['hyp'], reg=c['reg'])
        m1, v1 = eng.predict_f(p, t_star, samples, smf=smf, reg=c['reg'])
        assert np.abs(m1 - m0).max() <= 1e-8 * np.abs(m0).max() + 3 * mn
        assert np.abs(v1 - v0).max() <= 1e-8 * max(np.abs(v0).max(), np.abs(m0).max() ** 2) + 3 * vn
        outs.append((m1, v1))
    assert np.abs(outs[0][0] - outs[1][0]).max() <= 1e-8 * np.abs(m0).max() + 3 * mn
    e = cgpcm_b200.Engine(c['nh'], c['nx'])
    e.set_data(c['t'], c['y'], c['th'], c['tx'])
    e.precompute(*c['hyp'], reg=c['reg'])
    m_empty, v_empty = e.predict_f(p, np.zeros(0), samples, reg=c['reg'])
    assert m_empty.shape == (0,) and v_empty.shape == (0,)


def test_predict_f_api():
    c = make_case('toy_small')
    config.reg = c['reg']
    np.random.seed(4)
    sess = Session()
    mod = VCGPCM.from_recipe(sess, Data(c['t'], c['y']), nx=c['nx'], nh=c['nh'], tau_w=.1, tau_f=.05, causal=True,
                             noise_init=1e-2)
    mod.precompute()
    mod.fpi(10)
    pred = mod.predict_f(c['t'], samples_h=8)
    assert pred.mean.x.shape == c['t'].shape and np.all(pred.std.y >= 0)
    assert np.all(pred.lower.y <= pred.mean.y) and np.all(pred.mean.y <= pred.upper.y)
    # after the fixed-point iteration the posterior mean explains the (unit-variance) observations to a large part
    assert np.mean((pred.mean.y - c['y']) ** 2) < 0.6
    mod.undo_precompute()
    pred2 = mod.predict_f(Data(c['t'][:7]), samples_h=mod.sample(iters=3, burn=2))     # SMF path, not precomputed
    assert pred2.mean.y.shape == (7,) and np.all(np.isfinite(pred2.std.y))
    config.reg = 1e-8


@pytest.mark.parametrize('name', ['toy_small', 'hrir', 'toy_acausal_model'])
def test_kernel_samples_match_oracle(name):
    """predict_k's Monte-Carlo kernel samples (src/core/cgpcm.py:610-634; centre statistics :164-166,190-192)."""
    c = make_case(name)
    rng = np.random.default_rng(9)
    nh = c['nh']
    samples = c['params'][5:5 + nh] + .3 * rng.standard_normal((11, nh))
    span = c['th'].max() - c['th'].min()
    t = np.concatenate([np.linspace(-1.5 * span, 1.5 * span, 45), [0.0]])
    om.PW_DISTS_EXACT = True
    try:
        want = om.kernel_samples(c['params'], c['th'], c['reg'], t, samples, causal=c['causal'])
    finally:
        om.PW_DISTS_EXACT = False
    eng = cgpcm_b200.Engine(c['nh'], c['nx'], causal=c['causal'])
    eng.set_data(c['t'], c['y'], c['th'], c['tx'])
    got = eng.kernel_samples(c['params'], t, samples, reg=c['reg'])
    assert got.shape == (46, 11)
    # the trace against iKh (entries ~ 1/reg) cancels most of h^T Ahh h: relative to the terms that are summed
    scale = np.abs(want).max() + 1.0 / c['reg'] * 1e-9
    assert np.abs(got - want).max() <= 1e-9 * scale + 1e-7 * np.abs(want).max()
    assert eng.kernel_samples(c['params'], np.zeros(0), samples, reg=c['reg']).shape == (0, 11)


def test_predict_k_api():
    c = make_case('toy_small')
    config.reg = c['reg']
    np.random.seed(6)
    mod = VCGPCM.from_recipe(Session(), Data(c['t'], c['y']), nx=c['nx'], nh=c['nh'], tau_w=.1, tau_f=.05,
                             causal=True, noise_init=1e-2)
    t = np.linspace(-.3, .3, 61)
    k = mod.predict_k(t, samples_h=16)
    assert k.mean.x.shape == (61,) and np.all(k.lower.y <= k.upper.y + 1e-15)
    assert k.mean.y[30] == pytest.approx(k.mean.y.max())            # a (normalised) kernel peaks at lag 0
    assert np.abs(k.mean.y - k.mean.y[::-1]).max() < 0.25           # and is roughly even
    p = mod.predict_k(t, samples_h=[mod.sample_q() for _ in range(5)], psd=True)
    assert p.mean.x.shape == (61 + 4000,) and np.all(p.mean.y >= 0)
    with pytest.raises(AssertionError):
        mod.predict_k(np.array([0., .1, .3]), samples_h=2, psd=True)
    config.reg = 1e-8


@pytest.mark.parametrize('name,n', [('toy_small', 37), ('sweep_hi', 64), ('hrir', 50)])
def test_filter_samples_match_oracle(name, n):
    """predict_h / predict_psd's posterior draws of the filter (src/core/cgpcm.py:663-779)."""
    c = make_case(name)
    rng = np.random.default_rng(12)
    nh = c['nh']
    samples = c['params'][5:5 + nh] + .3 * rng.standard_normal((6, nh))
    lo, hi = c['th'].min(), c['th'].max()
    t = np.linspace(lo - .3 * (hi - lo), hi + .3 * (hi - lo), n)
    noise = rng.standard_normal((n, 6))
    om.PW_DISTS_EXACT = True
    try:
        want = om.filter_samples(c['params'], c['th'], c['reg'], t, samples, noise)
        mean_only = om.filter_samples(c['params'], c['th'], c['reg'], t, samples, np.zeros_like(noise))
    finally:
        om.PW_DISTS_EXACT = False
    eng = cgpcm_b200.Engine(c['nh'], c['nx'], causal=c['causal'])
    eng.set_data(c['t'], c['y'], c['th'], c['tx'])
    got0 = eng.filter_samples(c['params'], t, samples, np.zeros_like(noise), reg=c['reg'])
    got = eng.filter_samples(c['params'], t, samples, noise, reg=c['reg'])
    scale = np.abs(want).max()
    # conditional mean Kuh^T h: plain kernel evaluations and one GEMM
    assert np.abs(got0 - mean_only).max() <= 1e-12 * scale
    # the noise term goes through chol(reg(Ktt - A^T A)), the Nystrom residual of a kernel matrix with cond ~ 1/reg
    # (measured agreement: 3e-11 relative at reg = 1e-6)
    resid = np.abs((got - got0) - (want - mean_only)).max()
    assert resid <= 1e-8 * scale, resid


def test_predict_h_and_psd_api():
    c = make_case('toy_small')
    config.reg = c['reg']
    np.random.seed(8)
    mod = VCGPCM.from_recipe(Session(), Data(c['t'], c['y']), nx=c['nx'], nh=c['nh'], tau_w=.1, tau_f=.05,
                             causal=True, noise_init=1e-2)
    t = np.linspace(-.1, .3, 81)
    h = mod.predict_h(t, samples_h=12)
    assert np.all(h.mean.x >= 0) and h.mean.x.shape == (61,)          # causal model: the positive part of t
    assert np.all(np.isfinite(h.mean.y)) and np.all(h.lower.y <= h.upper.y + 1e-15)
    h_raw = mod.predict_h(t, samples_h=[mod.sample_q() for _ in range(4)], normalise=False, phase_transform=None)
    assert h_raw.mean.y.shape == (61,)
    p = mod.predict_psd(t, samples_h=6)
    assert p.mean.x.shape == (2 * 81 - 1 + 4000,) and np.all(p.mean.y >= 0)
    hz = mod.predict_h(t, samples_h=5, phase_transform='zero_phase')          # experiment.predict's third transform
    assert hz.mean.x.shape == (61,) and hz.mean.x[30] == 0.0 and hz.mean.x[0] == pytest.approx(-30 * .005)
    assert np.argmax(hz.mean.y) == 30                                         # a zero-phase signal peaks at its centre
    with pytest.raises(NotImplementedError):
        mod.predict_h(t, samples_h=2, phase_transform='linear_phase')
    config.reg = 1e-8
