"""ELBO and gradient from the CUDA path (through the C-ABI) against the CPU oracle (torch autograd over
the literal restatement of the reference) on the named shapes, both regimes.

Tolerances (BASELINE.json): ELBO within 1e-9 relative; gradient within 1e-9 of the gradient's scale
(max-norm), element-wise."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')
import cgpcm_b200
from cgpcm_b200 import GRAD_ALL, GRAD_MU_U, GRAD_S2, GRAD_S2F, GRAD_VAR_U, MODE_FROZEN, MODE_FULL
from oracle import model as om
from tests.cases import CASES, make_case

ELBO_RTOL = 1e-9
GRAD_RTOL = 1e-9


@pytest.fixture(autouse=True)
def exact_dists():
    """The product forms (x - y)^2 directly; the reference's |x|^2 - 2xy + |y|^2 carries rounding noise of
    its own (up to 1e-7 relative for crude's time stamps ~2010, see tests/test_adjoint_chain.py).  The
    parity target is the oracle with exact differences; test_faithful_distance_formula bounds the rest."""
    om.PW_DISTS_EXACT = True
    yield
    om.PW_DISTS_EXACT = False


def _engine(c, **opts):
    eng = cgpcm_b200.Engine(c['nh'], c['nx'], causal=c['causal'])
    for k, v in opts.items():
        eng.set_option(k, v)
    eng.set_data(c['t'], c['y'], c['th'], c['tx'])
    return eng


def _check(got, want, what=''):
    e1, t1, g1 = got
    e0, t0, g0 = want
    assert abs(e1 - e0) <= ELBO_RTOL * abs(e0), (what, e1, e0)
    np.testing.assert_allclose(t1, t0, rtol=0, atol=ELBO_RTOL * max(abs(e0), np.abs(t0).max()))
    if g1 is not None:
        scale = np.abs(g0).max()
        err = np.abs(g1 - g0).max()
        assert err <= GRAD_RTOL * scale, (what, err / scale, int(np.argmax(np.abs(g1 - g0))))


@pytest.mark.parametrize('name', CASES)
def test_full_regime(name):
    c = make_case(name)
    want = om.elbo_and_grad(c['params'], c['t'], c['y'], c['th'], c['tx'], c['reg'], c['causal'])
    got = _engine(c).elbo_grad(c['params'], mode=MODE_FULL, grad_mask=GRAD_ALL, reg=c['reg'])
    _check(got, want, name)


@pytest.mark.parametrize('name', ['toy_test', 'ou', 'hrir', 'sweep'])
def test_frozen_regime(name):
    c = make_case(name)
    fr = om.precompute(c['params'], c['t'], c['y'], c['th'], c['tx'], c['reg'], c['causal'])
    eng = _engine(c)
    eng.precompute(*c['hyp'], reg=c['reg'])
    # move q(u) and the noise away from the freeze point
    p = c['params'].copy()
    rng = np.random.default_rng(7)
    p[0] += .3
    p[1] -= .2
    p[5:] *= 1 + .05 * rng.standard_normal(p.shape[0] - 5)
    want = om.elbo_and_grad(p, c['t'], c['y'], c['th'], c['tx'], c['reg'], c['causal'], frozen=fr)
    qmask = GRAD_S2 | GRAD_S2F | GRAD_MU_U | GRAD_VAR_U          # what the precomputed training phases optimise
    got = eng.elbo_grad(p, mode=MODE_FROZEN, grad_mask=qmask, reg=c['reg'])
    _check(got, want, name)
    assert np.all(got[2][2:5] == 0)
    # all entries: in the reference only `mats` are frozen (src/core/cgpcm.py:270-284); the prior kernels stay
    # functions of (alpha, gamma, omega), so tf.gradients through Kx, Lx and the prior of q(u) is not zero -- and the
    # value follows the hyper-parameters when they move away from the freeze point
    p[2:5] += np.array([.05, -.03, .04])
    want = om.elbo_and_grad(p, c['t'], c['y'], c['th'], c['tx'], c['reg'], c['causal'], frozen=fr,
                            frozen_kernels='symbolic')
    got = eng.elbo_grad(p, mode=MODE_FROZEN, grad_mask=GRAD_ALL, reg=c['reg'])
    _check(got, want, name)
    assert np.all(got[2][2:5] != 0)


def test_grad_mask_and_value_only():
    c = make_case('toy_small')
    eng = _engine(c)
    full = eng.elbo_grad(c['params'], reg=c['reg'])
    e, t, g = eng.elbo_grad(c['params'], reg=c['reg'], want_grad=False)
    assert g is None and e == full[0]
    e, t, g = eng.elbo_grad(c['params'], grad_mask=GRAD_MU_U | GRAD_S2, reg=c['reg'])
    nh = c['nh']
    np.testing.assert_allclose(g[5:5 + nh], full[2][5:5 + nh], rtol=1e-12, atol=1e-12)
    assert g[0] == pytest.approx(full[2][0], rel=1e-12)
    assert np.all(g[1:5] == 0) and np.all(g[5 + nh:] == 0)


@pytest.mark.parametrize('name', ['toy_small', 'sweep_wide'])
def test_chunking_and_culling_do_not_change_the_result(name):
    c = make_case(name)
    ref = _engine(c, cull=0.0, chunk=4096).elbo_grad(c['params'], reg=c['reg'])
    # chunk = 0: the planner's choice (512 / 1024 / 2048 columns x nx, window-snapped or not, by its cost model)
    for opts in [dict(cull=0.0, chunk=32), dict(cull=80.0, chunk=64), dict(cull=80.0, chunk=1024), dict(cull=80.0, chunk=0),
                 dict(cull=746.0, chunk=0)]:
        got = _engine(c, **opts).elbo_grad(c['params'], reg=c['reg'])
        assert abs(got[0] - ref[0]) <= 1e-11 * abs(ref[0]), opts
        assert np.abs(got[2] - ref[2]).max() <= 1e-10 * np.abs(ref[2]).max(), opts


@pytest.mark.parametrize('name', ['toy_small', 'sweep_wide'])
def test_sweep_stores_do_not_change_the_result(name):
    """Option "store" (Ahx blocks and H*Ahx kept resident for the backward sweep; the frozen regime's blocks kept
    between evaluations) against regenerate / recompute: the same bits."""
    c = make_case(name)
    outs = []
    for store in (1, 0):
        eng = _engine(c, chunk=64)
        eng.set_option('store', store)
        full = eng.elbo_grad(c['params'], reg=c['reg'])
        eng.precompute(*c['hyp'], reg=c['reg'])
        p = c['params'].copy()
        p[5:] *= 1.02
        fr1 = eng.elbo_grad(p, mode=MODE_FROZEN, reg=c['reg'])
        fr2 = eng.elbo_grad(p, mode=MODE_FROZEN, reg=c['reg'])          # second call reuses the resident blocks
        again = eng.elbo_grad(c['params'], reg=c['reg'])                 # full regime overwrites them
        fr3 = eng.elbo_grad(p, mode=MODE_FROZEN, reg=c['reg'])          # ... and the frozen regime regenerates
        assert fr1[0] == fr2[0] == fr3[0] and np.array_equal(fr1[2], fr2[2]) and np.array_equal(fr1[2], fr3[2])
        assert again[0] == full[0] and np.array_equal(again[2], full[2])
        outs.append((full, fr1))
    for a, b in zip(outs[0], outs[1]):
        assert a[0] == b[0]
        np.testing.assert_array_equal(a[1], b[1])
        np.testing.assert_array_equal(a[2], b[2])


@pytest.mark.parametrize('name', ['toy_small', 'sweep_wide', 'sweep_hi'])
def test_psi_tensor_of_the_frozen_regime(name):
    """Option "gram": the precomputed regime contracts the resident fourth-order tensor G = sum_n Ahx_n (x) Ahx_n
    instead of sweeping over the observations (gram_kernels.cuh).  Forced on (2) against off (0): the frozen ELBO and
    gradient, a fixed-point round and the SMF bound agree to the rounding of a different summation order."""
    c = make_case(name)
    p = c['params'].copy()
    p[0] += .2
    p[5:] *= 1.03
    rng = np.random.default_rng(1)
    smp = p[5:5 + c['nh']] + .01 * rng.standard_normal(c['nh'])
    outs = []
    for gram in (2, 0):
        eng = _engine(c, chunk=64, gram=gram)
        eng.precompute(*c['hyp'], reg=c['reg'])
        fr = eng.elbo_grad(p, mode=MODE_FROZEN, reg=c['reg'])
        launches = eng.last_timing()['launches']
        fpi = eng.fpi(p, 1, reg=c['reg'])
        smf = eng.elbo_smf(p, smp, mode=MODE_FROZEN, reg=c['reg'])
        full = eng.elbo_grad(p, reg=c['reg'])                      # the full regime is untouched by the option
        outs.append((fr, fpi, smf, full, launches))
    (fr1, fpi1, smf1, full1, l1), (fr0, fpi0, smf0, full0, l0) = outs
    assert l1 < l0                                                 # no sweep launches in the tensor path
    scale = max(abs(fr0[0]), np.abs(fr0[1]).max())
    assert abs(fr1[0] - fr0[0]) <= 1e-9 * scale
    assert np.abs(fr1[2] - fr0[2]).max() <= 1e-8 * np.abs(fr0[2]).max()
    for a, b in zip(fpi1, fpi0):
        assert np.abs(a - b).max() <= 1e-6 * np.abs(b).max()
    assert abs(smf1[0] - smf0[0]) <= 1e-9 * scale and abs(smf1[2] - smf0[2]) <= 1e-7 * max(abs(smf0[2]), 1.0)
    assert full1[0] == full0[0] and np.array_equal(full1[2], full0[2])


def test_device_resident_buffers():
    c = make_case('toy_small')
    dev = lambda x: torch.tensor(x, dtype=torch.float64, device='cuda')
    eng = cgpcm_b200.Engine(c['nh'], c['nx'])
    eng.set_data(dev(c['t']), dev(c['y']), dev(c['th']), dev(c['tx']))
    want = _engine(c).elbo_grad(c['params'], reg=c['reg'])
    g = torch.zeros(c['params'].shape[0], dtype=torch.float64, device='cuda')
    got = eng.elbo_grad(dev(c['params']), reg=c['reg'], out_grad=g)
    assert got[0] == want[0]
    np.testing.assert_array_equal(g.cpu().numpy(), want[2])


def test_faithful_distance_formula():
    """Against the oracle with the reference's own distance formula the toy shape still meets 1e-9; the
    crude shape (inputs ~2010) only 1e-6 — the reference's rounding noise."""
    om.PW_DISTS_EXACT = False
    c = make_case('toy_test')
    want = om.elbo_and_grad(c['params'], c['t'], c['y'], c['th'], c['tx'], c['reg'])
    got = _engine(c).elbo_grad(c['params'], reg=c['reg'])
    assert abs(got[0] - want[0]) <= 1e-9 * abs(want[0])
    c = make_case('crude')
    want = om.elbo_and_grad(c['params'], c['t'], c['y'], c['th'], c['tx'], c['reg'])
    got = _engine(c).elbo_grad(c['params'], reg=c['reg'])
    assert abs(got[0] - want[0]) <= 1e-6 * abs(want[0])


def test_error_codes():
    c = make_case('toy_small')
    eng = _engine(c)
    p = c['params'].copy()
    p[7] = np.nan
    with pytest.raises(ValueError, match='non-finite'):
        eng.elbo_grad(p, reg=c['reg'])
    with pytest.raises(ValueError):
        eng.elbo_grad(c['params'], mode=MODE_FROZEN, reg=c['reg'])      # no precompute yet
    with pytest.raises(ValueError):
        eng.elbo_grad(c['params'][:-1], reg=c['reg'])
    # reg = 0 on a numerically singular Kh: the factorisation must fail with -3, naming the matrix
    c2 = make_case('toy_test')
    eng2 = _engine(c2)
    with pytest.raises(cgpcm_b200.CgpcmError, match='positive definite') as ei:
        eng2.elbo_grad(c2['params'], reg=0.0)
    assert ei.value.code == -3
    # the handle stays usable
    e, _, _ = eng2.elbo_grad(c2['params'], reg=c2['reg'])
    assert np.isfinite(e)


def test_repeatable():
    c = make_case('ou')
    eng = _engine(c)
    a = eng.elbo_grad(c['params'], reg=c['reg'])
    b = eng.elbo_grad(c['params'], reg=c['reg'])
    assert a[0] == b[0]
    np.testing.assert_array_equal(a[2], b[2])
    tm = eng.last_timing()
    assert tm['launches'] > 10 and tm['total_ms'] > 0
