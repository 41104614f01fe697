"""The product's formulation (reduced contractions + hand-written adjoint chain, prototyped in numpy
in tests/adjoint_proto.py exactly as csrc/cgpcm.cu evaluates it) against autograd of the literal
oracle.  This is the CPU-side check of the mathematics the CUDA library implements."""
import numpy as np
import pytest

from oracle import model as om
from tests import adjoint_proto as ap
from tests.cases import make_case


@pytest.fixture
def exact_dists():
    """crude's time stamps are ~2010: the reference's |x|^2 - 2xy + |y|^2 distance formula leaves ~1e-7
    relative rounding noise in Kx; compare against the oracle evaluated with exact differences."""
    om.PW_DISTS_EXACT = True
    yield
    om.PW_DISTS_EXACT = False


@pytest.mark.parametrize('name', ['toy_small', 'sweep', 'crude'])
def test_full_regime(name, exact_dists):
    c = make_case(name, n=30)
    e0, t0, g0 = om.elbo_and_grad(c['params'], c['t'], c['y'], c['th'], c['tx'], c['reg'])
    e1, t1, g1 = ap.elbo_grad(c['params'], c['t'], c['y'], c['th'], c['tx'], c['reg'], mode=1)
    assert abs(e0 - e1) < 1e-9 * abs(e0)
    np.testing.assert_allclose(t1, t0, rtol=1e-9, atol=1e-9 * abs(e0))
    scale = np.max(np.abs(g0))
    np.testing.assert_allclose(g1, g0, rtol=1e-7, atol=1e-9 * scale)


def test_reference_distance_formula_noise_is_bounded():
    """Against the *faithful* oracle (reference distance formula) the crude shape agrees to 1e-6 only —
    the reference's own rounding noise, not a property of the reformulation."""
    c = make_case('crude', n=30)
    e0, _, _ = om.elbo_and_grad(c['params'], c['t'], c['y'], c['th'], c['tx'], c['reg'])
    e1, _, _ = ap.elbo_grad(c['params'], c['t'], c['y'], c['th'], c['tx'], c['reg'], mode=1)
    assert abs(e0 - e1) < 1e-6 * abs(e0)


@pytest.mark.parametrize('name', ['toy_small', 'ou'])
def test_frozen_regime(name):
    c = make_case(name, n=30)
    fr = om.precompute(c['params'], c['t'], c['y'], c['th'], c['tx'], c['reg'])
    e0, t0, g0 = om.elbo_and_grad(c['params'], c['t'], c['y'], c['th'], c['tx'], c['reg'], frozen=fr)
    m, k = fr
    frozen = dict(Kx=k['Kx'].numpy(), iKh=k['iKh'].numpy(), iKx=k['iKx'].numpy(),
                  logdetKx=float(om.log_det(k['Lx'])), a=float(m['a']), Ahh=m['Ahh'].numpy(),
                  sum_Axx=m['sum_Axx'].numpy(), A=m['Ahx'].numpy(),
                  Q=(m['sum_Ahh'] - m['sum_Bhh']).numpy(), Y=m['sum_Ahx_y'].numpy())
    e1, t1, g1 = ap.elbo_grad(c['params'], c['t'], c['y'], c['th'], c['tx'], c['reg'], mode=0, frozen=frozen)
    assert abs(e0 - e1) < 1e-9 * abs(e0)
    scale = np.max(np.abs(g0))
    np.testing.assert_allclose(g1, g0, rtol=1e-7, atol=1e-9 * scale)


def test_triangular_window_route_matches_full_blocks():
    """Option `tri` of the CUDA path, restated on the host (tests/adjoint_proto.windowed_gram): Q = sum_n A_n iKx A_n^T and
    Hbar = sum_n A_n C1bar A_n^T through Cholesky factors of the window blocks -- iKx[w] and -C1bar[w] are positive
    definite on every window -- give the ELBO and gradient of the full-block products, on a sweep-like series whose
    windows of exact zeros are narrower than nx and in chunks that do not divide the series."""
    from tests.workload import sweep_workload
    wl = sweep_workload(260, 128)
    th = np.ascontiguousarray(wl['th'][::10])                         # 13 filter inducing points keep the host loops short
    a, g, o = wl['hyp']
    mu_u, var_u = om.init_q(th, a, g, wl['reg'], np.random.default_rng(5))
    p = om.pack(0.1, float(np.exp(wl['params'][1])), a, g, o, mu_u, var_u)
    args = (p, wl['t'], wl['y'], th, wl['tx'], wl['reg'])
    e0, t0, g0 = ap.elbo_grad(*args)
    e1, t1, g1 = ap.elbo_grad(*args, tri_chunk=48)
    A, _ = ap.ahx_with_tangents(wl['t'], th, wl['tx'], a, g, o)
    widest = ap.windowed_gram(A[:96], np.eye(128), 48)[1]
    assert 8 < widest < 128, widest                                   # the windows are exercised
    sc = np.abs(t0).max()
    assert abs(e1 - e0) <= 1e-9 * sc and np.abs(t1 - t0).max() <= 1e-9 * sc
    assert np.abs(g1 - g0).max() <= 1e-9 * np.abs(g0).max()
