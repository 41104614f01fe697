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
