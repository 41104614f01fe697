"""VCGPCM.fpi / convert (src/core/cgpcm.py:479-516,577-592) on the GPU against the oracle's
restatement, through the C-ABI (cgpcm_fpi) and through the reference-facing Python API."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')
import cgpcm_b200
from cgpcm_b200 import VCGPCM, Data, Session, config
from oracle import model as om
from tests.cases import make_case


@pytest.mark.parametrize('name,num,high_reg', [('toy_small', 1, False), ('toy_small', 4, True), ('toy_test', 2, False),
                                               ('sweep_hi', 3, False), ('toy_small', 0, False)])
def test_fpi_matches_oracle(name, num, high_reg):
    c = make_case(name)
    om.PW_DISTS_EXACT = True
    try:
        want = om.fpi(c['params'], c['t'], c['y'], c['th'], c['tx'], c['reg'], num, high_reg=high_reg)
        # conditioning noise of the iteration (two inversions of matrices with cond ~ 1/reg per round): how far the
        # oracle's own result moves when its inputs move by <= 2 ulp (cf. tests/cases.py: oracle_noise_floor)
        rng = np.random.default_rng(0)
        noise = [0.0] * 4
        for _ in range(3):
            p2 = c['params'].copy()
            p2[:5] *= 1 + rng.integers(-2, 3, 5) * 1.1e-16
            th2 = c['th'] * (1 + rng.integers(-2, 3, c['nh']) * 1.1e-16)
            w2 = om.fpi(p2, c['t'], c['y'], th2, c['tx'], c['reg'], num, high_reg=high_reg)
            noise = [max(a, float(np.abs(x - y).max())) for a, x, y in zip(noise, w2, want)]
    finally:
        om.PW_DISTS_EXACT = False
    eng = cgpcm_b200.Engine(c['nh'], c['nx'])
    eng.set_data(c['t'], c['y'], c['th'], c['tx'])
    with pytest.raises(ValueError):
        eng.fpi(c['params'], num, reg=c['reg'])                     # needs the frozen statistics
    eng.precompute(*c['hyp'], reg=c['reg'])
    got = eng.fpi(c['params'], num, high_reg=high_reg, reg=c['reg'])
    for g, w, nz, what in zip(got, want, noise, ['mu_u', 'var_u', 'mu_z', 'var_z']):
        assert np.abs(g - w).max() <= 1e-8 * np.abs(w).max() + 3 * nz, (what, np.abs(g - w).max(), np.abs(w).max(), nz)
    # same answer without the resident blocks, with small chunks, and on a second call
    eng2 = cgpcm_b200.Engine(c['nh'], c['nx'])
    eng2.set_option('store', 0)
    eng2.set_option('chunk', 32)
    eng2.set_data(c['t'], c['y'], c['th'], c['tx'])
    eng2.precompute(*c['hyp'], reg=c['reg'])
    again = eng2.fpi(c['params'], num, high_reg=high_reg, reg=c['reg'])
    for g, a, nz in zip(got, again, noise):                         # another summation order through the inversions
        assert np.abs(g - a).max() <= 1e-8 * np.abs(g).max() + 3 * nz
    twice = eng.fpi(c['params'], num, high_reg=high_reg, reg=c['reg'])
    for g, a in zip(got, twice):
        np.testing.assert_array_equal(g, a)


def test_fpi_api_raises_the_elbo_and_convert_assigns_qz():
    c = make_case('toy_test')
    config.reg = c['reg']
    np.random.seed(3)
    sess = Session()
    mod = VCGPCM.from_recipe(sess, Data(c['t'], c['y']), nx=c['nx'], nh=c['nh'], tau_w=.1, tau_f=.05, causal=True,
                             noise_init=1e-2)
    elbo, _ = mod.elbo()
    e0 = sess.run(elbo)
    mod.fpi(3)                                   # not precomputed: statistics at the current hyper-parameters
    e1 = sess.run(elbo)
    mod.precompute()
    mod.fpi(3)
    e2 = sess.run(elbo)
    assert e0 < e1 < e2 + 1e-9 * abs(e2)
    mod.convert()
    assert mod.vars['mu_z'].value.shape == (c['nx'], 1)
    assert mod.vars['var_z'].value.shape == (c['nx'] * (c['nx'] + 1) // 2,)
    # the z = False variants (src/core/cgpcm.py:479-516,577-592): the iteration on q(z) raises the bound saturated for
    # q(u); convert(z=False) then assigns the optimal q(u), which the z = True bound scores at least as high as before
    ez, terms_z = mod.elbo(z=False)
    assert [tm['name'] for tm in terms_z][1:4] == ['p(u) complexity', 'q*(u) complexity', 'q*(u) fit']
    z0 = sess.run(ez)
    assert sum(sess.run([tm['tensor'] for tm in terms_z])) == pytest.approx(z0, rel=1e-12)
    mod.fpi(2, z=False)
    z1 = sess.run(ez)
    assert z1 >= z0 - 1e-9 * abs(z0)
    mod.convert(z=False)
    assert sess.run(elbo) >= e2 - 1e-6 * abs(e2)
    with pytest.raises(NotImplementedError):
        (-ez).value_and_grad([mod.vars['mu_z']])
    # one more round from the fixed point of many rounds changes (almost) nothing
    mod.fpi(40)
    a = mod.vars['mu_u'].value.copy()
    mod.fpi(1)
    # (the map contracts slowly along the directions the data do not determine: 1.1e-3 after 40 rounds)
    assert np.abs(mod.vars['mu_u'].value - a).max() <= 5e-3 * np.abs(a).max()
    config.reg = 1e-8
