"""The oracle restatement against the REFERENCE'S OWN CODE.

`tests/golden/ref/*.npz` are outputs of `oracle/_ref` (py3-patched copies of /root/reference/src running on the
TensorFlow stand-in `oracle/tfshim`, built by `oracle/build_ref.py`; generator `tools/make_ref_golden.py`): what
`VCGPCM.from_recipe`, `mod.mats`, `mod._optimal_q`, `mod.elbo()` + `tf.gradients`, `mod.precompute()`, `mod.fpi()`,
`mod.convert()` and `mod.predict_f()` of src/core/cgpcm.py return on the seeded inputs of tests/cases.py.

Here: (1) the reference's own unit tests run on the stand-in; (2) a live run of the reference reproduces a committed
fixture (so the fixtures are what the reference computes, not a copy of the oracle); (3) the oracle agrees with the
fixtures within BASELINE.json's tolerances -- Psi matrices 1e-10 absolute, ELBO / gradient 1e-9 relative -- in the
reference's own arithmetic (`pw_dists2` as |x|^2 - 2xy + |y|^2, `1 - erf`).

The GPU side of the same comparison is tests/test_gpu_ref.py.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import build_ref, model as om, ref
from tests.cases import make_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, 'tests', 'golden', 'ref')
REF_CASES = ['toy_small', 'toy_small_cid', 'toy_test', 'toy_acausal_model', 'ou', 'hrir', 'crude', 'crude_shifted', 'sweep', 'sweep_hi']

PSI_ATOL = 1e-10
REL = 1e-9
REF_NOISE_GAIN = 100.     # the exponent noise of ref_noise() passes through iKx / iKh into the ELBO and its gradient

needs_ref = pytest.mark.skipif(not ref.available(), reason='needs oracle/_ref or /root/reference')


def ref_noise(c):
    """Rounding noise of the REFERENCE'S arithmetic, relative: its integrals expand the exponent into polynomials in
    absolute time (coefficients ~ omega t^2, exponentiated_quadratic.py:490-559) and `pw_dists2` forms
    |x|^2 - 2xy + |y|^2 (tf_util.py:24-31), so an offset in t costs eps * max|t|^2 * max(omega, alpha + gamma) in every
    exponent: 1e-6 at the crude-oil time stamps (t ~ 2010), < 1e-12 for every other named shape.  The oracle and the
    CUDA path work with differences t - tx and do not have this term (case 'crude_shifted' shows both agree with the
    reference to the plain tolerances once the origin is moved)."""
    a, g, o = c['hyp']
    return 1.1e-16 * float(np.abs(c['t']).max()) ** 2 * max(o, a + g)


def load(name):
    path = os.path.join(REF_DIR, name + '.npz')
    if not os.path.exists(path):
        pytest.skip('no reference fixture for ' + name)
    with np.load(path) as z:
        return {k: (z[k][()] if z[k].ndim == 0 else z[k]) for k in z.files}


@needs_ref
def test_reference_unit_tests_pass_on_the_stand_in():
    """src/core/exponentiated_quadratic_test.py (the reference's only known-answer tests) through oracle/_ref."""
    ref.ensure_built()
    paths, env = build_ref.paths()
    code = ('import sys, unittest, warnings; warnings.simplefilter("ignore"); sys.path[:0] = %r; '
            'import exponentiated_quadratic_test as t; '
            'r = unittest.TextTestRunner().run(unittest.defaultTestLoader.loadTestsFromModule(t)); '
            'sys.exit(0 if r.wasSuccessful() and r.testsRun == 3 else 1)' % (paths,))
    p = subprocess.run([sys.executable, '-c', code], env=dict(os.environ, **env), stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]


@needs_ref
def test_live_reference_reproduces_the_fixture():
    """The committed fixture really is the reference's output: rerun it.  (Term order inside the reference's
    polynomial sums follows Python's per-process hash seeds, so two runs differ in the last bits.)"""
    f = load('toy_small')
    c = make_case('toy_small')
    rec = c['recipe']
    r = ref.call(t=c['t'], y=c['y'], nx=rec['nx'], nh=rec['nh'], tau_w=rec['tau_w'], tau_f=rec['tau_f'],
                 causal=c['causal'], causal_id=False, reg=c['reg'], params=c['params'], want_per_n=True)
    assert np.array_equal(r['th'], c['th']) and np.array_equal(r['tx'], c['tx'])
    scale = np.abs(f['terms']).max()
    assert abs(r['elbo'] - f['elbo']) <= 1e-11 * scale
    assert np.abs(r['grad'] - f['grad']).max() <= 1e-10 * np.abs(f['grad']).max()
    assert np.abs(r['mat_Axx'] - f['mat_Axx']).max() <= 1e-14
    assert np.abs(r['mat_Ahx'] - f['mat_Ahx']).max() <= 1e-14


@pytest.mark.parametrize('name', REF_CASES)
def test_recipe(name):
    """`from_recipe` (src/core/cgpcm.py:32-109): inducing inputs and float32-rounded initial values."""
    f, c = load(name), make_case(name)
    assert np.array_equal(f['th'], c['th']) and np.array_equal(f['tx'], c['tx'])
    rec = om.recipe(c['t'], **c['recipe'])
    want = np.log([rec['s2'], rec['s2_f'], rec['alpha'], rec['gamma'], rec['omega']])
    np.testing.assert_allclose(f['recipe_vars'], want, rtol=0, atol=1e-15)


@pytest.mark.parametrize('name', REF_CASES)
def test_psi_and_model_matrices(name):
    """`_construct_model_matrices` (src/core/cgpcm.py:231-268): closed forms of the oracle vs the reference's
    symbolic integrals, 1e-10 absolute on every Psi entry; derived sums relative to their scale."""
    f, c = load(name), make_case(name)
    a, g, o = c['hyp']
    m, k = om.precompute(c['params'], c['t'], c['y'], c['th'], c['tx'], c['reg'], c['causal'], causal_id=c['causal_id'])
    assert abs(float(m['a']) - f['mat_a']) <= PSI_ATOL
    assert np.abs(m['Ahh'].numpy() - f['mat_Ahh']).max() <= PSI_ATOL
    if 'mat_Axx' in f:
        assert np.abs(m['Axx'].numpy() - f['mat_Axx']).max() <= PSI_ATOL
        assert np.abs(m['Ahx'].numpy() - f['mat_Ahx']).max() <= PSI_ATOL
    n = len(c['t'])
    rn = ref_noise(c)
    # sums over n: n entries of <= 1e-10 each (+ the reference's own rounding noise where t carries an offset)
    assert np.abs(m['sum_Axx'].numpy() - f['mat_sum_Axx']).max() <= (PSI_ATOL + rn) * n
    assert np.abs(m['sum_Ahx_y'].numpy() - f['mat_sum_Ahx_y']).max() <= (PSI_ATOL + rn) * n * np.abs(c['y']).max()
    assert np.abs(k['Kh'].numpy() - f['mat_Kh']).max() <= 1e-14
    assert np.abs(k['Kx'].numpy() - f['mat_Kx']).max() <= (1e-14 + rn) * np.abs(f['mat_Kx']).max()


@pytest.mark.parametrize('name', REF_CASES)
def test_elbo_terms_and_gradient_full_regime(name):
    """`VCGPCM.elbo()` and `tf.gradients` (src/core/cgpcm.py:518-575) vs the oracle: 1e-9 relative (ELBO: to the
    largest term; gradient: to its max-norm), plus the measured conditioning noise of the path at trained-like
    points (tests/cases.py: oracle_noise_floor)."""
    from tests.cases import oracle_noise_floor
    f, c = load(name), make_case(name)
    e, terms, g = om.elbo_and_grad(c['params'], c['t'], c['y'], c['th'], c['tx'], c['reg'], c['causal'], causal_id=c['causal_id'])
    en, gn = oracle_noise_floor(c['params'], c['t'], c['y'], c['th'], c['tx'], c['reg'], c['causal'], trials=2)
    scale = np.abs(f['terms']).max()
    rel = REL + REF_NOISE_GAIN * ref_noise(c)         # plain 1e-9 for every case but 'crude' (see ref_noise)
    assert abs(e - f['elbo']) <= rel * scale + 3 * en, (abs(e - f['elbo']) / scale, en / scale)
    assert np.abs(terms - f['terms']).max() <= rel * scale + 3 * en
    gs = np.abs(f['grad']).max()
    assert np.abs(g - f['grad']).max() <= rel * gs + 3 * gn, (np.abs(g - f['grad']).max() / gs, gn / gs)


@pytest.mark.parametrize('name', REF_CASES)
def test_precomputed_regime(name):
    """`precompute()` then new variable values (src/core/cgpcm.py:270-292): value, terms and the gradient w.r.t.
    (log s2, log s2_f, mu_u, var_u) -- and w.r.t. (alpha, gamma, omega), which in the reference is NOT zero: only
    `mats` are frozen, the prior kernels Kx, Kh and the prior of q(u) stay symbolic."""
    f, c = load(name), make_case(name)
    e, terms, g = om.elbo_and_grad(f['params_frozen'], c['t'], c['y'], c['th'], c['tx'], c['reg'], c['causal'],
                                   frozen=om.precompute(c['params'], c['t'], c['y'], c['th'], c['tx'], c['reg'],
                                                        c['causal'], causal_id=c['causal_id']),
                                   frozen_kernels='symbolic')
    scale = np.abs(f['terms_frozen']).max()
    rel = REL + REF_NOISE_GAIN * ref_noise(c)
    assert abs(e - f['elbo_frozen']) <= rel * scale
    gs = np.abs(f['grad_frozen']).max()
    assert np.abs(g - f['grad_frozen']).max() <= rel * gs, np.abs(g - f['grad_frozen']).max() / gs


@pytest.mark.parametrize('name', REF_CASES)
def test_optimal_q(name):
    """`_optimal_q(z=True)` (src/core/cgpcm.py:458-477)."""
    f, c = load(name), make_case(name)
    import torch
    m, k = om.precompute(c['params'], c['t'], c['y'], c['th'], c['tx'], c['reg'], c['causal'], causal_id=c['causal_id'])
    s2, s2_f, alpha, gamma, omega, mu_u, var_u = om.unpack(c['params'], c['nh'])
    Lq = om.vec_to_tril(var_u)
    var = om.reg(Lq @ Lq.T, c['reg'])
    lam, P = om.optimal_q(m, k, s2, s2_f, mu_u, var + mu_u @ mu_u.T, True)
    rel = REL + REF_NOISE_GAIN * ref_noise(c)
    assert np.abs(P.numpy() - f['optq_P']).max() <= rel * np.abs(f['optq_P']).max()
    assert np.abs(lam.numpy() - f['optq_lam']).max() <= rel * max(1e-300, np.abs(f['optq_lam']).max())


@pytest.mark.parametrize('name', REF_CASES)
def test_fpi_convert_and_predict_f(name):
    """`fpi(num=3)` + `convert()` (src/core/cgpcm.py:479-516,577-592, `Normal.from_natural` distribution.py:20-33)
    and `predict_f` (cgpcm.py:781-846, the non-SMF branch on the reference's own draws) in the precomputed regime.
    The fixed-point map is ill-conditioned (1 / reg), so q(u) is compared through what it is used for: the ELBO at
    the result, and the predictive mean / standard deviation."""
    f, c = load(name), make_case(name)
    if 'fpi_error' in f:
        pytest.skip('the reference\'s own fpi fails on this case: ' + str(f['fpi_error']))
    nh = c['nh']
    rel = 1e-7 + REF_NOISE_GAIN * ref_noise(c)
    mu_u, var_u, mu_z, var_z = om.fpi(f['params_frozen'], c['t'], c['y'], c['th'], c['tx'], c['reg'], 3,
                                      causal=c['causal'], causal_id=c['causal_id'])
    p_ref = np.concatenate([f['params_frozen'][:5], f['fpi_mu_u'], f['fpi_var_u']])
    p_own = np.concatenate([f['params_frozen'][:5], mu_u, var_u])
    e_ref = om.elbo_and_grad(p_ref, c['t'], c['y'], c['th'], c['tx'], c['reg'], c['causal'], causal_id=c['causal_id'])[0]
    e_own = om.elbo_and_grad(p_own, c['t'], c['y'], c['th'], c['tx'], c['reg'], c['causal'], causal_id=c['causal_id'])[0]
    assert abs(e_ref - f['fpi_elbo']) <= rel * abs(f['fpi_elbo'])      # the ELBO of the reference's q(u), two ways
    assert abs(e_own - f['fpi_elbo']) <= 1e-5 * abs(f['fpi_elbo'])     # ... and of the oracle's own iteration
    mu, var = om.predict_f(p_ref, c['t'], c['y'], c['th'], c['tx'], c['reg'], f['t_star'], f['pred_samples'],
                           smf=False, causal=c['causal'], causal_id=c['causal_id'])
    sc = max(np.abs(f['pred_mean']).max(), np.abs(f['pred_std']).max())
    ptol = 1e-6 + 10 * REF_NOISE_GAIN * ref_noise(c)
    assert np.abs(mu - f['pred_mean']).max() <= ptol * sc
    assert np.abs(np.sqrt(var) - f['pred_std']).max() <= ptol * sc


@pytest.mark.parametrize('name', REF_CASES)
def test_z_false_variants(name):
    """`elbo(z=False)`, `fpi(2, z=False)` + `convert(z=False)` (src/core/cgpcm.py:472-476,499,537-540,584-592) from an
    explicit q(z), precomputed regime."""
    f, c = load(name), make_case(name)
    if 'qz_error' in f:
        pytest.skip('the reference\'s own elbo(z=False) fails on this case: ' + str(f['qz_error']))
    rel = 10 * REL + REF_NOISE_GAIN * ref_noise(c)
    e, terms = om.elbo_qz(f['params_frozen'], f['qz_mu'], f['qz_var'], c['t'], c['y'], c['th'], c['tx'], c['reg'],
                          c['causal'], causal_id=c['causal_id'])
    scale = np.abs(f['qz_terms']).max()
    assert abs(e - f['qz_elbo']) <= rel * scale
    assert np.abs(terms - f['qz_terms']).max() <= rel * scale
    if 'qz_fpi_error' in f:
        return
    mu_u, var_u, mu_z, var_z = om.fpi_qz(f['params_frozen'], f['qz_mu'], f['qz_var'], c['t'], c['y'], c['th'], c['tx'],
                                         c['reg'], 2, c['causal'], causal_id=c['causal_id'])
    # compared through what the result is used for: the bound at the iterated q(z), and the z = True bound at the
    # converted q(u)
    e_own = om.elbo_qz(f['params_frozen'], mu_z, var_z, c['t'], c['y'], c['th'], c['tx'], c['reg'], c['causal'],
                       causal_id=c['causal_id'])[0]
    e_ref = om.elbo_qz(f['params_frozen'], f['qz_fpi_mu_z'], f['qz_fpi_var_z'], c['t'], c['y'], c['th'], c['tx'],
                       c['reg'], c['causal'], causal_id=c['causal_id'])[0]
    assert abs(e_own - e_ref) <= 1e-5 * abs(e_ref)
    p_own = np.concatenate([f['params_frozen'][:5], mu_u, var_u])
    p_ref = np.concatenate([f['params_frozen'][:5], f['qz_fpi_mu_u'], f['qz_fpi_var_u']])
    eu_own = om.elbo_and_grad(p_own, c['t'], c['y'], c['th'], c['tx'], c['reg'], c['causal'], causal_id=c['causal_id'])[0]
    eu_ref = om.elbo_and_grad(p_ref, c['t'], c['y'], c['th'], c['tx'], c['reg'], c['causal'], causal_id=c['causal_id'])[0]
    assert abs(eu_own - eu_ref) <= 1e-5 * abs(eu_ref)
