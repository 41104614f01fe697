"""the stochastic SMF bound elbo(smf=True, sample) / elbo_smf and the slice sampler's target
(src/core/cgpcm.py:527-531,594-608,848-872) on the GPU against the oracle, through the C-ABI (cgpcm_elbo_smf) and the
reference-facing Python API."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')
import cgpcm_b200
from cgpcm_b200 import VCGPCM, Data, Session, config, MODE_FROZEN, MODE_FULL
from oracle import model as om
from tests.cases import make_case, ulp_noise


@pytest.mark.parametrize('name', ['toy_small', 'toy_test', 'sweep_hi', 'crude'])
def test_elbo_smf_and_loglik_match_oracle(name):
    c = make_case(name)
    rng = np.random.default_rng(5)
    nh = c['nh']
    eng = cgpcm_b200.Engine(c['nh'], c['nx'])
    eng.set_data(c['t'], c['y'], c['th'], c['tx'])
    for trial in range(2):
        s = c['params'][5:5 + nh] * (1 + .2 * rng.standard_normal(nh)) + .01 * rng.standard_normal(nh)
        om.PW_DISTS_EXACT = True
        try:
            # oracle value and its conditioning noise (sample() factorises P *without* jitter: at the toy shape the
            # oracle's own log_lik moves by 4e-6 when its inputs move by 2 ulp)
            (e0, t0, ll0), (en, tn, lln) = ulp_noise(
                lambda p, th: om.elbo_smf(p, c['t'], c['y'], th, c['tx'], c['reg'], s), c['params'], c['th'])
        finally:
            om.PW_DISTS_EXACT = False
        e1, t1, ll1 = eng.elbo_smf(c['params'], s, mode=MODE_FULL, reg=c['reg'])
        scale = max(abs(e0), np.abs(t0).max())
        assert abs(e1 - e0) <= 1e-9 * scale + 3 * en and np.abs(t1 - t0).max() <= 1e-9 * scale + 3 * tn
        assert abs(ll1 - ll0) <= 1e-9 * max(abs(ll0), np.abs(t0[2:4]).max()) + 3 * lln
        # the precomputed regime gives the same numbers (up to the reformulation noise of the two regimes)
        eng.precompute(*c['hyp'], reg=c['reg'])
        e2, t2, ll2 = eng.elbo_smf(c['params'], s, mode=MODE_FROZEN, reg=c['reg'])
        assert abs(e2 - e1) <= 1e-7 * scale and abs(ll2 - ll1) <= 1e-7 * max(abs(ll1), 1.0)
    with pytest.raises(ValueError):
        eng.elbo_smf(c['params'], np.zeros(nh + 1), reg=c['reg'])
    with pytest.raises(ValueError):
        eng.elbo_smf(c['params'], np.full(nh, np.nan), reg=c['reg'])


def test_smf_api_and_sampler():
    c = make_case('toy_small')
    config.reg = c['reg']
    np.random.seed(11)
    sess = Session()
    mod = VCGPCM.from_recipe(sess, Data(c['t'], c['y']), nx=c['nx'], nh=c['nh'], tau_w=.1, tau_f=.05, causal=True,
                             noise_init=1e-2)
    mod.precompute()
    mod.fpi(5)
    # the SMF bound at the mean of q(u) equals the oracle's
    mean = mod.vars['mu_u'].value.ravel().copy()
    elbo, terms = mod.elbo(smf=True, sample=mean)
    om.PW_DISTS_EXACT = True
    try:
        want = om.elbo_smf(mod._pack(), c['t'], c['y'], mod.th, mod.tx, config.reg, mean)
    finally:
        om.PW_DISTS_EXACT = False
    scale = np.abs(want[1]).max()
    assert abs(sess.run(elbo) - want[0]) <= 1e-7 * scale
    assert abs(sum(sess.run([tm['tensor'] for tm in terms])) - want[0]) <= 1e-7 * scale
    assert sess.run(-elbo) == pytest.approx(-sess.run(elbo), rel=1e-14)
    # elbo(smf=True) without a sample draws from q(u): finite, different from call to call
    e_rand, _ = mod.elbo(smf=True)
    a, b = sess.run(e_rand), sess.run(e_rand)
    assert np.isfinite(a) and np.isfinite(b) and a != b
    # posterior samples by elliptical slice sampling; Monte-Carlo SMF estimate
    samples = mod.sample(iters=12, burn=4)
    assert len(samples) == 12 and samples[0].shape == (c['nh'], 1)
    assert all(np.all(np.isfinite(s)) for s in samples)
    est, se = mod.elbo_smf(samples)
    assert np.isfinite(est) and se >= 0
    config.reg = 1e-8
