"""Committed golden vectors (tools/make_golden.py): the oracle must keep reproducing them on the CPU, and the
CUDA path must reproduce them on the GPU through the C-ABI."""
import glob
import os

import numpy as np
import pytest

from oracle import model as om

# (bench_m200.npz -- the bench shape's own fixture, M = 200 and N = 1e4 -- has its own schema and test: test_gpu_quad.py)
FILES = [f for f in sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden', '*.npz')))
         if os.path.basename(f) != 'bench_m200.npz']


def _load(f):
    d = np.load(f)
    return {k: d[k] for k in d.files}


def test_fixtures_exist():
    assert len(FILES) >= 6


@pytest.mark.parametrize('f', FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracle_reproduces_fixture(f):
    d = _load(f)
    om.PW_DISTS_EXACT = True
    try:
        e, terms, g = om.elbo_and_grad(d['params'], d['t'], d['y'], d['th'], d['tx'], float(d['reg']), bool(d['causal']))
    finally:
        om.PW_DISTS_EXACT = False
    assert abs(e - float(d['elbo'])) <= 1e-11 * abs(float(d['elbo']))
    np.testing.assert_allclose(g, d['grad'], rtol=0, atol=1e-10 * np.abs(d['grad']).max())


@pytest.mark.gpu
@pytest.mark.parametrize('f', FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_gpu_reproduces_fixture(f):
    import cgpcm_b200
    d = _load(f)
    eng = cgpcm_b200.Engine(len(d['th']), len(d['tx']), causal=bool(d['causal']))
    eng.set_data(d['t'], d['y'], d['th'], d['tx'])
    psi = eng.psi(*d['hyp'])
    assert np.abs(psi['sum_Axx'] - d['sum_Axx']).max() < 1e-10           # BASELINE.json: Psi within 1e-10 absolute
    assert np.abs(psi['sum_Ahx_y'] - d['sum_Ahx_y']).max() < 1e-10
    assert np.abs(psi['Ahh'] - d['Ahh']).max() < 1e-10 and abs(psi['a'] - float(d['a'])) < 1e-10
    reg = float(d['reg'])
    e, terms, g = eng.elbo_grad(d['params'], reg=reg)
    scale = max(abs(float(d['elbo'])), np.abs(d['terms']).max())
    assert abs(e - float(d['elbo'])) <= 1e-9 * scale                     # ELBO within 1e-9 relative
    assert np.abs(terms - d['terms']).max() <= 1e-9 * scale
    assert np.abs(g - d['grad']).max() <= 1e-9 * np.abs(d['grad']).max()  # gradient within 1e-9 of its scale
    eng.precompute(*d['hyp'], reg=reg)
    e, terms, g = eng.elbo_grad(d['params_frozen'], mode=cgpcm_b200.MODE_FROZEN, reg=reg)
    scale = max(abs(float(d['elbo_frozen'])), np.abs(d['terms_frozen']).max())
    assert abs(e - float(d['elbo_frozen'])) <= 1e-9 * scale
    assert np.abs(g - d['grad_frozen']).max() <= 1e-9 * np.abs(d['grad_frozen']).max()


AKM_FILES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden', 'akm', '*.npz')))


@pytest.mark.parametrize('f', AKM_FILES, ids=[os.path.basename(f)[:-4] for f in AKM_FILES])
def test_oracle_reproduces_akm_fixture(f):
    """AKM.f() (src/core/cgpcm.py:382-392) from the reference's pair integrands: covariance and draw."""
    d = _load(f)
    om.PW_DISTS_EXACT = True
    try:
        f1, K1 = om.akm_f(d['params'], d['th'], float(d['reg']), d['t'], d['h'], d['e'], causal=bool(d['causal']))
    finally:
        om.PW_DISTS_EXACT = False
    assert len(AKM_FILES) == 2
    np.testing.assert_allclose(K1, d['K'], rtol=0, atol=1e-9 * np.abs(d['K']).max())
    np.testing.assert_allclose(f1, d['f'], rtol=0, atol=1e-6 * np.abs(d['f']).max())


@pytest.mark.gpu
@pytest.mark.parametrize('f', AKM_FILES, ids=[os.path.basename(f)[:-4] for f in AKM_FILES])
def test_gpu_reproduces_akm_fixture(f):
    import cgpcm_b200
    d = _load(f)
    eng = cgpcm_b200.Engine(len(d['th']), len(d['tx']), causal=bool(d['causal']))
    eng.set_data(np.zeros(1), np.zeros(1), d['th'], d['tx'])
    f1, K1 = eng.akm_sample(d['params'], d['t'], d['h'], d['e'], reg=float(d['reg']), want_cov=True)
    # tr(iKh Ahh) sums nh^2 products 1/reg times larger than the result (tests/test_gpu_akm.py)
    assert np.abs(K1 - d['K']).max() <= 1e-6 * np.abs(d['K']).max()
    assert np.abs(f1 - d['f']).max() <= 1e-6 * np.linalg.cond(d['K']) * np.abs(d['f']).max()
