"""The CUDA path (through the C-ABI) against the REFERENCE'S OWN CODE: tests/golden/ref/*.npz are outputs of
`oracle/_ref` (py3-patched copies of /root/reference/src on the TensorFlow stand-in; tools/make_ref_golden.py) for the
seeded inputs of tests/cases.py.  BASELINE.json's bars: Psi matrices 1e-10 absolute, ELBO and gradient 1e-9 relative.

The product runs with option pw_dists = 1 here: the prior kernels are formed with the reference's
|x|^2 - 2xy + |y|^2 (`pw_dists2`, src/core/tf_util.py:24-31) so that the comparison is arithmetic-for-arithmetic.
The one named shape where the reference's own arithmetic is noisy (crude-oil time stamps, t ~ 2010: its integrals
expand polynomials in absolute time) is bounded by that noise (tests/test_ref_parity.py: ref_noise) and repeated
with the origin of time moved (`crude_shifted`), where the plain bars hold.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')
import cgpcm_b200
from cgpcm_b200 import GRAD_ALL, MODE_FROZEN, MODE_FULL
from tests.cases import make_case, oracle_noise_floor
from tests.test_ref_parity import REF_CASES, REF_NOISE_GAIN, load, ref_noise

PSI_ATOL = 1e-10
REL = 1e-9


def _engine(c, **opts):
    eng = cgpcm_b200.Engine(c['nh'], c['nx'], causal=c['causal'], causal_id=c['causal_id'])
    eng.set_option('pw_dists', 1)
    for k, v in opts.items():
        eng.set_option(k, v)
    eng.set_data(c['t'], c['y'], c['th'], c['tx'])
    return eng


@pytest.mark.parametrize('name', REF_CASES)
def test_psi_matrices(name):
    f, c = load(name), make_case(name)
    got = _engine(c, cull=0.0).psi(*c['hyp'], per_observation='mat_Axx' in f)
    rn = ref_noise(c)
    n = len(c['t'])
    assert abs(got['a'] - f['mat_a']) <= PSI_ATOL
    assert np.abs(got['Ahh'] - f['mat_Ahh']).max() <= PSI_ATOL
    if 'mat_Axx' in f:
        assert np.abs(got['Axx'] - f['mat_Axx']).max() <= PSI_ATOL
        assert np.abs(got['Ahx'] - f['mat_Ahx']).max() <= PSI_ATOL
    assert np.abs(got['sum_Axx'] - f['mat_sum_Axx']).max() <= (PSI_ATOL + rn) * n
    assert np.abs(got['sum_Ahx_y'] - f['mat_sum_Ahx_y']).max() <= (PSI_ATOL + rn) * n * np.abs(c['y']).max()


@pytest.mark.parametrize('name', REF_CASES)
@pytest.mark.parametrize('cull', [0.0, 746.0])
def test_elbo_terms_and_gradient(name, cull):
    f, c = load(name), make_case(name)
    e, terms, g = _engine(c, cull=cull).elbo_grad(c['params'], mode=MODE_FULL, grad_mask=GRAD_ALL, reg=c['reg'])
    en, gn = oracle_noise_floor(c['params'], c['t'], c['y'], c['th'], c['tx'], c['reg'], c['causal'], trials=2)
    rel = REL + REF_NOISE_GAIN * ref_noise(c)
    scale = np.abs(f['terms']).max()
    assert abs(e - f['elbo']) <= rel * scale + 3 * en, (abs(e - f['elbo']) / scale, en / scale)
    assert np.abs(terms - f['terms']).max() <= rel * scale + 3 * en
    gs = np.abs(f['grad']).max()
    assert np.abs(g - f['grad']).max() <= rel * gs + 3 * gn, (np.abs(g - f['grad']).max() / gs, gn / gs)


@pytest.mark.parametrize('name', REF_CASES)
def test_precomputed_regime(name):
    """precompute() at `params`, then value and FULL gradient at `params_frozen`: in the reference only `mats` are
    frozen -- the prior kernels stay functions of (alpha, gamma, omega), so those gradient entries are not zero
    (src/core/cgpcm.py:270-292, 214-229)."""
    f, c = load(name), make_case(name)
    eng = _engine(c)
    eng.precompute(*c['hyp'], reg=c['reg'])
    e, terms, g = eng.elbo_grad(f['params_frozen'], mode=MODE_FROZEN, grad_mask=GRAD_ALL, reg=c['reg'])
    rel = 10 * REL + REF_NOISE_GAIN * ref_noise(c)
    scale = np.abs(f['terms_frozen']).max()
    assert abs(e - f['elbo_frozen']) <= rel * scale
    assert np.abs(terms - f['terms_frozen']).max() <= rel * scale
    gs = np.abs(f['grad_frozen']).max()
    assert np.abs(g - f['grad_frozen']).max() <= rel * gs, np.abs(g - f['grad_frozen']).max() / gs


@pytest.mark.parametrize('name', REF_CASES)
def test_fpi_and_predict_f(name):
    """fpi(3) + convert() and predict_f on the reference's own draws (non-SMF branch), precomputed regime."""
    f, c = load(name), make_case(name)
    eng = _engine(c)
    eng.precompute(*c['hyp'], reg=c['reg'])
    if 'fpi_error' in f:
        # the reference's own iteration leaves the positive-definite cone here (tf.cholesky raises): so must ours
        with pytest.raises(cgpcm_b200.CgpcmError):
            eng.fpi(f['params_frozen'], 3, reg=c['reg'])
        return
    mu_u, var_u, mu_z, var_z = eng.fpi(f['params_frozen'], 3, reg=c['reg'])
    p_own = np.concatenate([f['params_frozen'][:5], mu_u, var_u])
    p_ref = np.concatenate([f['params_frozen'][:5], f['fpi_mu_u'], f['fpi_var_u']])
    e_own = eng.elbo_grad(p_own, mode=MODE_FROZEN, reg=c['reg'], want_grad=False)[0]
    e_ref = eng.elbo_grad(p_ref, mode=MODE_FROZEN, reg=c['reg'], want_grad=False)[0]
    rel = 1e-7 + REF_NOISE_GAIN * ref_noise(c)
    assert abs(e_ref - f['fpi_elbo']) <= rel * abs(f['fpi_elbo'])
    assert abs(e_own - f['fpi_elbo']) <= 1e-5 * abs(f['fpi_elbo'])
    mu, var = eng.predict_f(p_ref, f['t_star'], f['pred_samples'], smf=False, reg=c['reg'])
    sc = max(np.abs(f['pred_mean']).max(), np.abs(f['pred_std']).max())
    ptol = 1e-6 + 10 * REF_NOISE_GAIN * ref_noise(c)
    assert np.abs(mu - f['pred_mean']).max() <= ptol * sc
    assert np.abs(np.sqrt(var) - f['pred_std']).max() <= ptol * sc


@pytest.mark.parametrize('name', REF_CASES)
def test_z_false_variants(name):
    """elbo(z=False), fpi(2, z=False) + convert(z=False) from an explicit q(z) against the reference's own outputs."""
    f, c = load(name), make_case(name)
    eng = _engine(c)
    eng.precompute(*c['hyp'], reg=c['reg'])
    if 'qz_error' in f:
        with pytest.raises(cgpcm_b200.CgpcmError):       # the reference's tf.cholesky raises here: so must ours
            eng.elbo_qz(f['params_frozen'], f['qz_mu'], f['qz_var'], reg=c['reg'])
        return
    rel = 10 * REL + REF_NOISE_GAIN * ref_noise(c)
    e, terms = eng.elbo_qz(f['params_frozen'], f['qz_mu'], f['qz_var'], reg=c['reg'])
    scale = np.abs(f['qz_terms']).max()
    assert abs(e - f['qz_elbo']) <= rel * scale
    assert np.abs(terms - f['qz_terms']).max() <= rel * scale
    if 'qz_fpi_error' in f:
        return
    mu_u, var_u, mu_z, var_z = eng.fpi_qz(f['params_frozen'], f['qz_mu'], f['qz_var'], 2, reg=c['reg'])
    e_own = eng.elbo_qz(f['params_frozen'], mu_z, var_z, reg=c['reg'])[0]
    e_ref = eng.elbo_qz(f['params_frozen'], f['qz_fpi_mu_z'], f['qz_fpi_var_z'], reg=c['reg'])[0]
    assert abs(e_own - e_ref) <= 1e-5 * abs(e_ref)
    p_own = np.concatenate([f['params_frozen'][:5], mu_u, var_u])
    p_ref = np.concatenate([f['params_frozen'][:5], f['qz_fpi_mu_u'], f['qz_fpi_var_u']])
    eu_own = eng.elbo_grad(p_own, mode=MODE_FROZEN, reg=c['reg'], want_grad=False)[0]
    eu_ref = eng.elbo_grad(p_ref, mode=MODE_FROZEN, reg=c['reg'], want_grad=False)[0]
    assert abs(eu_own - eu_ref) <= 1e-5 * abs(eu_ref)
