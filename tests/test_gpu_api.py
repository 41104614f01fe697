"""The reference-facing Python API on the GPU: the training schedule of src/core/experiment.py:209-250
(precompute -> L-BFGS on q(u) -> + noise -> undo_precompute -> + hyper-parameters) runs unchanged, and
the ELBO it reaches is the oracle's ELBO at the same variables."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')
import cgpcm_b200
from cgpcm_b200 import VCGPCM, Data, Session, config, learn
from oracle import model as om
from tests.cases import oracle_noise_floor


def test_train_schedule_toy():
    config.reg = 1e-6
    np.random.seed(1005)
    rng = np.random.default_rng(1005)
    n = 120
    t = np.linspace(0, 1, n)
    w = np.exp(-300 * np.linspace(-.2, .2, 41) ** 2)
    y = np.convolve(rng.standard_normal(n + 40), w, mode='valid')
    y = (y - y.mean()) / y.std()
    sess = Session()
    mod = VCGPCM.from_recipe(sess, Data(t, y), nx=40, nh=21, tau_w=.1, tau_f=.05, causal=True, noise_init=1e-2)
    mod.precompute()
    elbo, terms = mod.elbo()
    e_start = sess.run(elbo)
    fetches = [{'name': 'ELBO', 'tensor': elbo, 'modifier': '.2e'},
               {'name': 's2', 'tensor': mod.s2, 'modifier': '.2e'}]
    learn.minimise_lbfgs(sess, -elbo, vars=[mod.vars['mu_u'], mod.vars['var_u']], iters=30,
                         fetches_config=fetches + terms, name='pretraining using L-BFGS', quiet=True)
    e_pre = sess.run(elbo)
    learn.minimise_lbfgs(sess, -elbo, vars=[mod.vars['mu_u'], mod.vars['var_u'], mod.vars['s2_f'], mod.vars['s2']],
                         iters=40, fetches_config=fetches + terms, name='training using L-BFGS', quiet=True)
    e_main = sess.run(elbo)
    mod.undo_precompute()
    elbo, terms = mod.elbo()
    # frozen == full at the freeze point, up to the reformulation noise of the two regimes (sum_Bxx is
    # pre-summed in one, H = m2 - iKh contracted in the other; cond(Kh) ~ 1/reg amplifies the difference)
    assert sess.run(elbo) == pytest.approx(e_main, rel=1e-7)
    learn.minimise_lbfgs(sess, -elbo, vars=[mod.vars[k] for k in ['mu_u', 'var_u', 's2_f', 's2', 'gamma', 'omega',
                                                                   'alpha']],
                         iters=15, fetches_config=fetches + terms, name='posttraining using L-BFGS', quiet=True)
    e_post = sess.run(elbo)
    assert e_start < e_pre <= e_main + 1e-9 and e_main <= e_post + 1e-9
    assert sum(sess.run([tm['tensor'] for tm in terms])) == pytest.approx(e_post, rel=1e-12)
    # the oracle agrees at the trained variables
    p = mod._pack()
    om.PW_DISTS_EXACT = True
    try:
        want = om.elbo_and_grad(p, t, y, mod.th, mod.tx, config.reg)
    finally:
        om.PW_DISTS_EXACT = False
    # the trained ELBO (~ -8) is a sum of terms of magnitude ~1e2..1e3 that cancel: parity is 1e-9 relative
    # to the largest term, term by term
    got_terms = np.array(sess.run([tm['tensor'] for tm in terms]))
    scale = np.abs(want[1]).max()
    assert np.abs(got_terms - want[1]).max() <= 1e-9 * scale, (got_terms, want[1])
    assert abs(want[0] - e_post) <= 1e-9 * scale
    gsel = mod._evaluate(True, ['mu_u', 'var_u', 's2_f', 's2', 'gamma', 'omega', 'alpha'])[2]
    # at the trained point (s2 ~ 4e-3, cond(Kh) ~ 1/reg) the gradient is a cancellation of terms ~1e6 times
    # larger than itself: the oracle's own gradient moves by ~1e-5 when its inputs move by 1 ulp.  The bar is
    # 1e-9 relative or that measured conditioning noise, whichever is larger.
    om.PW_DISTS_EXACT = True
    try:
        _, gnoise = oracle_noise_floor(p, t, y, mod.th, mod.tx, config.reg)
    finally:
        om.PW_DISTS_EXACT = False
    assert np.abs(gsel - want[2]).max() <= 1e-9 * np.abs(want[2]).max() + 3 * gnoise, gnoise
    mats = mod.mats
    assert mats['sum_Axx'].shape == (40, 40) and mats['Ahh'].shape == (21, 21)
    config.reg = 1e-8
