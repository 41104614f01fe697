"""Host-side mirror of the reference API: recipe values, variable packing, sharding, lazy handles."""
import numpy as np
import pytest

import cgpcm_b200
from cgpcm_b200 import cgpcm as cg
from cgpcm_b200 import util
from oracle import model as om


class _FakeEngine(object):
    """Stands in for the GPU engine so that the host logic can be exercised without a device."""
    calls = []

    def __init__(self, nh, nx, **kw):
        self.nh, self.nx = nh, nx

    def set_data(self, t, y, th, tx):
        self.t, self.y = t, y

    def precompute(self, *a):
        _FakeEngine.calls.append(('precompute',) + a)

    def elbo_grad(self, params, mode, grad_mask, reg, want_grad):
        _FakeEngine.calls.append(('elbo_grad', mode, grad_mask, reg, want_grad))
        g = np.arange(params.shape[0], dtype=np.float64) if want_grad else None
        return float(np.sum(params)), np.arange(7.), g


@pytest.fixture
def model(monkeypatch):
    monkeypatch.setattr(cg, 'Engine', _FakeEngine)
    _FakeEngine.calls = []
    np.random.seed(0)
    t = np.linspace(0, 1, 50)
    e = cgpcm_b200.Data(t, np.sin(9 * t))
    return cg.VCGPCM.from_recipe(cgpcm_b200.Session(device=0, rank=0, world=1), e, nx=20, nh=9, tau_w=.1,
                                 tau_f=.05, causal=True)


def test_recipe_matches_oracle_restatement(model):
    rec = om.recipe(model.e.x, nx=20, nh=9, tau_w=.1, tau_f=.05, causal=True)
    for k in ['alpha', 'gamma', 'omega', 's2', 's2_f']:
        assert getattr(model, k).eval() == pytest.approx(rec[k], rel=1e-15)
        assert float(model.vars[k].value) == pytest.approx(np.log(rec[k]), rel=1e-15)
    np.testing.assert_array_equal(model.th, rec['th'])
    np.testing.assert_array_equal(model.tx, rec['tx'])


def test_recipe_acausal_forces_odd_nh(monkeypatch):
    monkeypatch.setattr(cg, 'Engine', _FakeEngine)
    e = cgpcm_b200.Data(np.linspace(0, 1, 30), np.zeros(30))
    mod = cg.VCGPCM.from_recipe(cgpcm_b200.Session(device=0, rank=0, world=1), e, nx=10, nh=8, tau_w=.1, tau_f=.05,
                                causal=False)
    assert mod.nh == 9 and mod.th[4] == 0.0


def test_initial_var_u_is_cholesky_of_prior(model):
    rec = om.recipe(model.e.x, nx=20, nh=9, tau_w=.1, tau_f=.05, causal=True)
    _, var_u = om.init_q(rec['th'], rec['alpha'], rec['gamma'], cgpcm_b200.config.reg, np.random.default_rng(0))
    np.testing.assert_allclose(model.vars['var_u'].value, var_u, rtol=1e-6, atol=1e-8)
    assert model.vars['mu_u'].value.shape == (9, 1)


def test_pack_layout_and_grad_slicing(model):
    p = model._pack()
    assert p.shape[0] == cgpcm_b200.n_params(9)
    assert p[2] == float(model.vars['alpha'].value)
    np.testing.assert_array_equal(p[5:14], model.vars['mu_u'].value.ravel())
    g = np.arange(p.shape[0], dtype=np.float64)
    out = model._slice_grad(g, ['mu_u', 's2', 'var_u', 'omega'])
    np.testing.assert_array_equal(out, np.concatenate([g[5:14], g[0:1], g[14:], g[4:5]]))


def test_elbo_objective_modes_and_masks(model):
    sess = model.sess
    elbo, terms = model.elbo()
    assert [t['name'] for t in terms] == list(cgpcm_b200.TERM_NAMES) and terms[0]['modifier'] == '.2e'
    v = sess.run(elbo)
    assert v == pytest.approx(np.sum(model._pack()))
    assert sess.run(-elbo) == -v
    assert sess.run([t['tensor'] for t in terms]) == list(range(7))
    assert _FakeEngine.calls[-1][1] == cgpcm_b200.MODE_FULL
    model.precompute()
    assert _FakeEngine.calls[-1][0] == 'precompute'
    f, g = (-elbo).value_and_grad([model.vars['mu_u'], model.vars['s2']])
    call = _FakeEngine.calls[-1]
    assert call[1] == cgpcm_b200.MODE_FROZEN
    assert call[2] == cgpcm_b200.GRAD_MU_U | cgpcm_b200.GRAD_S2
    assert f == -v and g.shape == (10,) and g[-1] == -0.0
    model.undo_precompute()
    model.sess.run(model.vars['s2'].assign(np.log(.5)))
    assert model.s2.eval() == pytest.approx(.5)
    with pytest.raises(ValueError):
        model.vars['mu_u'].assign(np.zeros(3))
    ez, terms_z = model.elbo(z=False)                    # value and terms only (cgpcm_elbo_qz): no gradient
    assert len(terms_z) == 7 and terms_z[1]['name'] == 'p(u) complexity'
    with pytest.raises(NotImplementedError):
        ez.value_and_grad([model.vars['mu_u']])
    with pytest.raises(NotImplementedError):
        model.elbo(smf=True, z=False)


def test_minimise_lbfgs_drives_objective(model):
    class Quad(object):
        def __init__(self, mod):
            self.mod = mod

        def value_and_grad(self, vs):
            x = np.concatenate([v.value.ravel() for v in vs])
            return float(np.sum((x - 1.5) ** 2)), 2 * (x - 1.5)

    res = cgpcm_b200.learn.minimise_lbfgs(model.sess, Quad(model), [model.vars['mu_u'], model.vars['s2']], iters=30,
                                          quiet=True)
    np.testing.assert_allclose(model.vars['mu_u'].value, 1.5, atol=1e-5)
    assert model.vars['mu_u'].value.shape == (9, 1)
    assert cgpcm_b200.learn.minimise_lbfgs(model.sess, None, [], iters=0) is None
    assert res.nit <= 28


def test_required_parameters():
    with pytest.raises(RuntimeError, match='must specify'):
        cg.CGPCM(sess=None)


def test_shard_bounds_partition():
    for n in [0, 1, 7, 100, 100001]:
        for w in [1, 2, 3, 8]:
            b = [cg.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_shard_bounds_by_cost():
    """Cost-balanced shards: a partition, every rank non-empty, equal shares of the window-cost proxy; the proxy's
    radius is the one plan_chunks uses (csrc/cgpcm.cu: ahx_radius)."""
    rng = np.random.default_rng(0)
    t = np.sort(rng.uniform(0, 100, 5000))
    tx = np.linspace(0, 100, 200)
    R = cg.window_radius(39.27, 1237.0, 1.5708, 746.0)
    A = 39.27 + 1237.0 + 1.5708
    lam = 1.5708 * (39.27 + 1237.0) / A - (2 * 1237.0 * 1.5708 / A) ** 2 / (4 * ((39.27 + 1237.0) * A - 1237.0 ** 2) / A)
    assert R == pytest.approx(np.sqrt(746.0 / lam))
    assert cg.window_radius(1., 1., 1., 0.0) == float('inf')
    cost = cg.window_costs(t, tx, 200, R)
    assert cost.shape == t.shape and cost[0] < cost[2500]            # the ends of the series are cheaper
    for w in [2, 3, 8]:
        b = [cg.shard_bounds(len(t), r, w, cost) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == len(t)
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        shares = np.array([cost[lo:hi].sum() for lo, hi in b])
        assert shares.min() > 0 and shares.max() / shares.min() < 1.01
    assert b[0][1] - b[0][0] > b[3][1] - b[3][0]                      # edge shards hold more observations
    # degenerate costs fall back to equal counts; tiny n keeps every rank non-empty
    assert cg.shard_bounds(10, 1, 2, np.zeros(10)) == (5, 10)
    b = [cg.shard_bounds(3, r, 3, np.array([0., 0., 5.])) for r in range(3)]
    assert b == [(0, 1), (1, 2), (2, 3)]
    assert np.all(cg.window_costs(t, tx, 200, float('inf')) == 1.0)


def test_rebalance_costs_moves_cuts_towards_equal_measured_time():
    """rebalance_costs: a synthetic "true" cost the model misses (edge observations 30 % dearer than modelled); one
    round of measured shard times brings the true shares within a few per cent, a second within 1 %."""
    n, w = 8000, 8
    model = np.ones(n)
    true = np.ones(n)
    true[:1500] = 1.3
    true[-1000:] = 1.5
    cost = model.copy()
    spread = []
    for _ in range(3):
        b = [cg.shard_bounds(n, r, w, cost) for r in range(w)]
        times = np.array([true[lo:hi].sum() for lo, hi in b])
        spread.append(times.max() / times.mean())
        cost = cg.rebalance_costs(cost, b, times)
    assert spread[0] > 1.15 and spread[1] < 1.06 and spread[2] < 1.012
    with pytest.raises(ValueError):
        cg.rebalance_costs(cost, b, np.zeros(w))


def test_tril_packing_matches_numpy_order():
    L = np.tril(np.arange(1., 17.).reshape(4, 4))
    v = util.tril_to_vec(L)
    np.testing.assert_array_equal(v, L[np.tril_indices(4)])
    np.testing.assert_array_equal(util.vec_to_tril(v), L)


def test_elliptical_slice_sampler_on_a_gaussian_target():
    """cgpcm_b200.sample.ESS (interface of src/core/sample.py:8-130): prior N(0, I_2), likelihood N(x; m, s^2 I) ->
    posterior N(m / (1 + s^2), s^2 / (1 + s^2) I); the chain's moments match within Monte-Carlo error."""
    from cgpcm_b200.sample import ESS
    rng = np.random.RandomState(0)
    m, s2 = np.array([[1.5], [-0.5]]), 0.5
    ess = ESS(lambda x: float(-.5 * np.sum((x - m) ** 2) / s2), lambda: rng.randn(2, 1), rng=rng)
    ess.move(np.zeros((2, 1)))
    ess.sample(200)
    xs = np.concatenate(ess.sample(4000), 1)
    want_mean, want_var = m.ravel() / (1 + s2), s2 / (1 + s2)
    assert np.abs(xs.mean(1) - want_mean).max() < 0.06
    assert np.abs(xs.var(1) - want_var).max() < 0.05
    assert min(ess.attempts) >= 1 and np.mean(ess.attempts) < 6
    one = ess.sample(1)
    assert one.shape == (2, 1)


def test_fft_spectrum_follows_the_reference_conventions():
    """cgpcm_b200.util.fft_spectrum = Data.fft of the reference (src/core/data.py:184-209): 2000 zeros on both sides,
    fftshift, scaled by the spacing -> the spectrum of a Gaussian is the Gaussian with the continuous-time scaling."""
    from cgpcm_b200.util import fft_spectrum
    t = np.linspace(-2, 2, 401)
    y = np.exp(-np.pi * (t / .2) ** 2)                              # FT: 0.2 exp(-pi (0.2 f)^2)
    f, s = fft_spectrum(t, y)
    assert f.shape == (4401,) and f[2200] == 0 and np.all(np.diff(f) > 0)
    np.testing.assert_allclose(np.abs(s), .2 * np.exp(-np.pi * (.2 * f) ** 2), atol=1e-9)
    f2, s2 = fft_spectrum(t, np.stack([y, 2 * y], 1))
    np.testing.assert_allclose(np.abs(s2[:, 1]), 2 * np.abs(s), atol=1e-12)
    with pytest.raises(AssertionError):
        fft_spectrum(np.array([0., 1., 3.]), np.ones(3))


def test_signal_helpers_follow_the_reference():
    """minimum_phase / energy / autocorrelation of cgpcm_b200.util (src/core/data.py:136-171,293-303,354-359)."""
    from cgpcm_b200.util import minimum_phase, energy, autocorrelation
    x = np.linspace(0, 1, 101)
    y = np.exp(-30 * (x - .5) ** 2) * np.cos(40 * x)
    m = minimum_phase(y)
    np.testing.assert_allclose(np.abs(np.fft.fft(m)), np.abs(np.fft.fft(y)), rtol=1e-8, atol=1e-10)   # same magnitude
    assert np.sum(m[:20] ** 2) > np.sum(y[:20] ** 2)              # energy moved to the front
    assert energy(x, np.ones(101)) == pytest.approx(1.0)
    lags, ac = autocorrelation(x, y, normalise=True)
    assert lags.shape == (201,) and ac[100] == pytest.approx(1.0) and np.allclose(ac, ac[::-1])
    from cgpcm_b200.util import zero_phase                       # src/core/data.py:265-276
    xz, yz = zero_phase(x, y)
    assert xz[50] == 0.0 and xz[0] == pytest.approx(-.5) and xz.shape == (101,)
    np.testing.assert_allclose(np.abs(np.fft.fft(yz)), np.abs(np.fft.fft(y)), rtol=1e-8, atol=1e-10)  # same magnitude
    assert np.argmax(yz) == 50 and np.allclose(yz, yz[::-1], atol=1e-12)                      # even about the centre
    with pytest.raises(AssertionError):
        zero_phase(np.array([0., 1., 3.]), np.ones(3))
