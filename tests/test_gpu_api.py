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
from tests.cases import oracle_noise_floor  # noqa: F401 (kept for interactive use)


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
    # the oracle agrees at the trained variables.  At the trained point (s2 ~ 4e-3, cond(Kh) ~ 1/reg) the ELBO (~ -8) is a
    # sum of terms of magnitude ~1e2..1e3 that cancel and the gradient a cancellation of terms ~1e6 times larger than
    # itself: the oracle's own terms / gradient move by ~1e-5 when its inputs move by 2 ulp (measured here).  The bar is
    # 1e-9 relative to the largest term, term by term, plus 3 x that measured conditioning noise.
    from tests.cases import ulp_noise
    p = mod._pack()
    om.PW_DISTS_EXACT = True
    try:
        want, (enoise, tnoise, gnoise) = ulp_noise(
            lambda pp, th: om.elbo_and_grad(pp, t, y, th, mod.tx, config.reg), p, mod.th, trials=4)
    finally:
        om.PW_DISTS_EXACT = False
    got_terms = np.array(sess.run([tm['tensor'] for tm in terms]))
    scale = np.abs(want[1]).max()
    assert np.abs(got_terms - want[1]).max() <= 1e-9 * scale + 3 * tnoise, (got_terms, want[1], tnoise)
    # the ELBO is the sum of the 7 terms: its bar is the sum of theirs (the sampled noise of the sum alone is a 4-trial
    # estimate and has been seen a few per cent below the actual difference); tests/test_gpu_quad.py compares both
    # sides with a quad-precision evaluation at trained points
    assert abs(want[0] - e_post) <= 1e-9 * scale + 3 * max(enoise, 7 * tnoise)
    gsel = mod._evaluate(True, ['mu_u', 'var_u', 's2_f', 's2', 'gamma', 'omega', 'alpha'])[2]
    # the gradient at a trained point is the residual of an optimisation: terms ~scale whose derivatives cancel.  The
    # binary128 arbiter (tests/test_gpu_quad.py, profiles/r02_quad_truth.json) shows that both FP64 evaluations carry
    # ~1e-9 of THAT scale, so the bar for their difference is 2e-9 of max(|g|_max, largest term) + the sampled noise
    assert np.abs(gsel - want[2]).max() <= 2e-9 * max(np.abs(want[2]).max(), scale) + 3 * gnoise, gnoise
    mats = mod.mats
    assert mats['sum_Axx'].shape == (40, 40) and mats['Ahh'].shape == (21, 21)
    # the derived sums of src/core/cgpcm.py:255-267 against the FP64 restatement at the trained hyper-parameters
    om.PW_DISTS_EXACT = True
    try:
        m_or, _ = om.precompute(p, t, y, mod.th, mod.tx, config.reg)
    finally:
        om.PW_DISTS_EXACT = False
    for key in ('sum_Bxx', 'sum_Bhh', 'sum_Ahx_y', 'sum_Axx'):
        want_m = m_or[key].numpy()
        assert np.abs(mats[key] - want_m).max() <= 1e-7 * max(1.0, np.abs(want_m).max()), key
    assert mats['sum_b'] == pytest.approx(float(m_or['sum_b']), rel=1e-6, abs=1e-6)
    assert mats['sum_a'] == pytest.approx(len(t) * mats['a'])
    config.reg = 1e-8
