"""BASELINE.json's full size (N = 1e5 observations, M = 200 inducing points) through size-independent
properties — the oracle cannot run there:
  * additivity: Psi sums of two half series add up to those of the whole series (the sharding identity),
  * dense and culled evaluation agree,
  * the analytic gradient matches central differences of the GPU ELBO along random directions,
  * symmetry of sum_Axx, and sum of the 7 terms == ELBO."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')
import cgpcm_b200
from tests.workload import sweep_workload

N, M = 100000, 200


@pytest.fixture(scope='module')
def wl():
    return sweep_workload(N, M, seed=0)


@pytest.fixture(scope='module')
def eng(wl):
    e = cgpcm_b200.Engine(M, M)
    e.set_data(wl['t'], wl['y'], wl['th'], wl['tx'])
    return e


def test_psi_additivity_over_shards(wl, eng):
    whole = eng.psi(*wl['hyp'])
    half = N // 2
    parts = []
    for sl in [slice(0, half), slice(half, N)]:
        e2 = cgpcm_b200.Engine(M, M)
        e2.set_data(wl['t'][sl], wl['y'][sl], wl['th'], wl['tx'])
        parts.append(e2.psi(*wl['hyp']))
        e2.close()
    for k in ['sum_Axx', 'sum_Ahx_y']:
        s = parts[0][k] + parts[1][k]
        assert np.abs(s - whole[k]).max() <= 1e-12 * np.abs(whole[k]).max()
    np.testing.assert_array_equal(whole['sum_Axx'], whole['sum_Axx'].T)


def test_dense_equals_culled_and_terms_sum(wl, eng):
    eng.set_option('cull', 80.0)
    culled = eng.elbo_grad(wl['params'], reg=wl['reg'])
    t_c = eng.last_timing()
    eng.set_option('cull', 0.0)
    dense = eng.elbo_grad(wl['params'], reg=wl['reg'])
    eng.set_option('cull', 80.0)
    # two evaluations that differ only in their summation order (other chunk plans, windows, K-splits) differ by up to
    # ~3e-10 relative here: cond(Kh) ~ 1/reg = 1e6 amplifies the order (observed spread over the plans of this round:
    # -1039747.43782 .. -1039747.43811)
    assert abs(dense[0] - culled[0]) <= 6e-10 * abs(dense[0])
    assert np.abs(dense[2] - culled[2]).max() <= 1e-9 * np.abs(dense[2]).max()
    assert dense[1].sum() == pytest.approx(dense[0], rel=1e-13)
    assert t_c['total_ms'] > 0
    # cull = 746: only windows whose elements are exactly 0.0 in IEEE double are skipped (exp(E) underflows for
    # E < -745.13) -- the same sums as the all-tiles evaluation in a different order
    t_d = eng.last_timing()
    eng.set_option('cull', 746.0)
    exact = eng.elbo_grad(wl['params'], reg=wl['reg'])
    t_e = eng.last_timing()
    eng.set_option('cull', 80.0)
    assert abs(dense[0] - exact[0]) <= 6e-10 * abs(dense[0])
    assert np.abs(dense[2] - exact[2]).max() <= 1e-9 * np.abs(dense[2]).max()
    assert t_e['gemm_flops'] < 0.5 * t_d['gemm_flops']


def test_gradient_directional_derivative(wl, eng):
    """Central differences of the GPU ELBO, Richardson-extrapolated (h = 2e-3, 1e-3: smaller steps drown in the
    ~1e-11 relative evaluation noise of an ELBO of magnitude 1e6), against the analytic gradient."""
    rng = np.random.default_rng(3)
    p = wl['params']
    e0, _, g = eng.elbo_grad(p, reg=wl['reg'])
    gnorm = np.linalg.norm(g)

    def fd(d, h):
        f1 = eng.elbo_grad(p + h * d, reg=wl['reg'], want_grad=False)[0]
        f2 = eng.elbo_grad(p - h * d, reg=wl['reg'], want_grad=False)[0]
        return (f1 - f2) / (2 * h)

    m = wl['nh']
    for sl in [slice(0, 5), slice(5, 5 + m), slice(5 + m, None), slice(0, None)]:
        d = np.zeros_like(p)
        d[sl] = rng.standard_normal(d[sl].shape[0])
        d /= np.linalg.norm(d)
        est = (4 * fd(d, 1e-3) - fd(d, 2e-3)) / 3
        an = float(g @ d)
        # evaluation noise ~1e-10 .. 5e-10 * |ELBO| = 1e-4 .. 5e-4 over 2h = 2e-3, amplified 5 / 3 by the extrapolation ->
        # up to ~0.8 absolute = 3e-7 * |g|.  (A coarse check of the full-size path; the sharp ones are the oracle / binary128
        # fixtures of the same shape at N = 1e4, tests/test_gpu_quad.py.)
        assert abs(est - an) <= 2e-5 * abs(an) + 6e-7 * gnorm, (sl, est, an)


def test_oracle_parity_at_m200_through_the_large_shape_kernels():
    """nh = nx = 200 with few observations: the evaluation goes through the persistent small-left kernel (dgemm_sl,
    N = 48 x 200 = 9600 columns per chunk) and the symmetric kernel (dgemm_sym, M = 200) -- the kernels of the bench
    shape -- at a size the oracle still handles.  Dense and culled, with and without the sweep stores."""
    from oracle import model as om
    from tests.cases import ulp_noise
    n = 96
    w = sweep_workload(4000, M, seed=3)
    sl = slice(1000, 1000 + n)                      # a window in the middle: inducing inputs on both sides
    t, y = np.ascontiguousarray(w['t'][sl]), np.ascontiguousarray(w['y'][sl])
    om.PW_DISTS_EXACT = True
    try:
        (e0, t0, g0), (en, tn, gn) = ulp_noise(
            lambda p, th: om.elbo_and_grad(p, t, y, th, w['tx'], w['reg']), w['params'], w['th'], trials=1)
    finally:
        om.PW_DISTS_EXACT = False
    scale = max(abs(e0), np.abs(t0).max())
    for opts in (dict(cull=0.0, chunk=48, store=1), dict(cull=80.0, chunk=64, store=0)):
        e = cgpcm_b200.Engine(M, M)
        for k, v in opts.items():
            e.set_option(k, v)
        e.set_data(t, y, w['th'], w['tx'])
        e1, t1, g1 = e.elbo_grad(w['params'], reg=w['reg'])
        tm = e.last_timing()
        assert tm['gemm_launches'] > 0
        assert abs(e1 - e0) <= 1e-9 * scale + 3 * en, opts
        assert np.abs(t1 - t0).max() <= 1e-9 * scale + 3 * tn, opts
        assert np.abs(g1 - g0).max() <= 1e-9 * np.abs(g0).max() + 3 * gn, opts
        e.close()


def test_chunk_planner_at_the_bench_shape(wl, eng):
    """Option chunk = 0 (default): with exact-zero windows the planner takes larger chunks than 512 (fewer, longer GEMM
    launches) for the same result up to the summation order; in the precomputed regime it plans the same chunks at
    every evaluation (the resident Ahx blocks are laid out by the plan)."""
    eng.set_option('cull', 746.0)
    eng.set_option('chunk', 512)
    fixed = eng.elbo_grad(wl['params'], reg=wl['reg'])
    t_fixed = eng.last_timing()
    eng.set_option('chunk', 0)
    auto = eng.elbo_grad(wl['params'], reg=wl['reg'])
    t_auto = eng.last_timing()
    assert t_auto['gemm_launches'] < 0.5 * t_fixed['gemm_launches']
    assert t_auto['gemm_flops'] < 1.2 * t_fixed['gemm_flops']
    assert abs(auto[0] - fixed[0]) <= 6e-10 * abs(fixed[0])
    assert np.abs(auto[2] - fixed[2]).max() <= 1e-9 * np.abs(fixed[2]).max()
    eng.precompute(*wl['hyp'], reg=wl['reg'])
    a = eng.elbo_grad(wl['params'], mode=cgpcm_b200.MODE_FROZEN, reg=wl['reg'])
    b = eng.elbo_grad(wl['params'], mode=cgpcm_b200.MODE_FROZEN, reg=wl['reg'])
    assert a[0] == b[0] and np.array_equal(a[2], b[2])
    assert abs(a[0] - auto[0]) <= 2e-9 * abs(auto[0])           # frozen == full at the freeze point
    eng.set_option('cull', 80.0)
