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
