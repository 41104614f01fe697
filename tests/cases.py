"""Synthetic inputs of the named configurations at sizes the oracle finishes in seconds.
Shared by the CPU oracle tests, the GPU parity tests and tools/make_golden.py."""
import numpy as np

from oracle import model as om


def _unit(y):
    return (y - y.mean()) / y.std()


def make_case(name, seed=0, n=None):
    """dict(t, y, th, tx, hyp (alpha, gamma, omega), s2, s2_f, reg, causal); `recipe` holds the arguments of
    `VCGPCM.from_recipe` (src/core/cgpcm.py:33-34) that reproduce th, tx and the initial hyper-parameters."""
    rng = np.random.default_rng(seed)
    causal = True
    rec_args = {}

    class _Rec(object):               # records the recipe arguments the case uses
        @staticmethod
        def recipe(t, **kw):
            rec_args.update(kw)
            return _om_recipe(t, **kw)
    _om_recipe = om.recipe
    if name == 'toy_test':          # src/tasks/toy.py:25-55, 'test' option
        n = n or 150
        t = np.linspace(0, 1, n)
        rec = _Rec.recipe(t, nx=60, nh=41, tau_w=.1, tau_f=.05, causal=True)
        y, reg = _unit(rng.standard_normal(n)), 1e-6
    elif name in ('toy_small', 'toy_small_cid'):   # a shrunken toy for finite differences ('_cid': causal_id=True,
                                                   #   the limits min(t, tx) of src/core/cgpcm.py:168-180,194-203)
        n = n or 40
        t = np.linspace(0, 1, n)
        rec = _Rec.recipe(t, nx=18, nh=11, tau_w=.1, tau_f=.05, causal=True)
        y, reg = _unit(rng.standard_normal(n)), 1e-6
    elif name == 'toy_acausal_model':
        n = n or 60
        t = np.linspace(0, 1, n)
        rec = _Rec.recipe(t, nx=24, nh=13, tau_w=.1, tau_f=.05, causal=False)
        y, reg, causal = _unit(rng.standard_normal(n)), 1e-6, False
    elif name == 'ou':              # src/tasks/ou.py:23-45 (shrunk)
        n = n or 120
        t = np.linspace(0, 1, n)
        rec = _Rec.recipe(t, nx=64, nh=25, tau_w=.15, tau_f=.025, causal=True)
        K = np.exp(-np.abs(t[:, None] - t[None, :]) / .05)
        y = _unit(np.linalg.cholesky(K + 1e-10 * np.eye(n)) @ rng.standard_normal(n))
        reg = 1e-5
    elif name == 'hrir':            # src/tasks/hrir.py:20-40 (shrunk)
        n = n or 100
        t = np.arange(n) / 44100.
        rec = _Rec.recipe(t, nx=56, nh=31, tau_w=1.5e-3, tau_f=5e-5, causal=True)
        filt = rng.standard_normal(40) * np.exp(-np.arange(40) / 8.)
        y = _unit(np.convolve(rng.standard_normal(n + 39), filt, mode='valid'))
        reg = 1e-8
    elif name in ('crude', 'crude_shifted'):   # src/tasks/crude.py:25-48: uneven time stamps (decimal years)
        n = n or 90
        stamps = 2010 + 4 * np.sort(rng.choice(1013, size=n, replace=False)) / 1013.
        # 'crude_shifted': the same series with the origin of time moved to 2012.  The model is invariant under the
        # shift; the reference's arithmetic is not (its integrals expand polynomials in absolute time and its
        # pw_dists2 forms |x|^2 - 2xy + |y|^2), so this is the case where the reference itself is accurate
        t = stamps if name == 'crude' else stamps - 2012.
        rec = _Rec.recipe(t, nx=50, nh=21, tau_w=1., tau_f=.1, causal=True)
        y = np.cumsum(rng.standard_normal(n))
        y = _unit(y - np.polyval(np.polyfit(t, y, 1), t))
        reg = 1e-4
    elif name == 'sweep':           # scaling sweep shape: rho ~ 0.9+ (high Genz branch), sparse windows
        n = n or 400
        t = np.linspace(0, n / 1000., n)
        rec = _Rec.recipe(t, nx=24, nh=16, tau_w=.1, tau_f=.025, causal=True)
        w = np.exp(-40 * np.linspace(-.3, .3, 61) ** 2)
        y = _unit(np.convolve(rng.standard_normal(n + 60), w, mode='valid'))
        reg = 1e-6
    elif name == 'sweep_hi':        # the bench's correlation: rho = 0.967 -> Genz's |rho| >= 0.925 branch (pair-hoisted
        n = n or 600                #   Chebyshev path of bvn.cuh), observations on both sides of every inducing input
        t = np.linspace(0, 6., n)
        rec = _Rec.recipe(t, nx=16, nh=12, tau_w=.1, tau_f=.025, causal=True)
        w = np.exp(-40 * np.linspace(-.3, .3, 61) ** 2)
        y = _unit(np.convolve(rng.standard_normal(n + 60), w, mode='valid'))
        reg = 1e-6
    elif name == 'sweep_wide':      # long series, few inducing inputs per unit time -> narrow windows
        n = n or 3000
        t = np.linspace(0, n / 100., n)
        rec = _Rec.recipe(t, nx=40, nh=12, tau_w=.1, tau_f=.025, causal=True)
        y, reg = _unit(rng.standard_normal(n)), 1e-6
    else:
        raise KeyError(name)
    hyp = (rec['alpha'], rec['gamma'], rec['omega'])
    mu_u, var_u = om.init_q(rec['th'], rec['alpha'], rec['gamma'], reg, rng)
    # move q(u) off the prior so that every gradient entry is exercised
    var_u = var_u * (1 + .1 * rng.standard_normal(var_u.shape[0]))
    s2 = 0.3 if name != 'hrir' else 0.05
    params = om.pack(s2, rec['s2_f'], hyp[0], hyp[1], hyp[2], mu_u, var_u)
    return dict(name=name, t=np.ascontiguousarray(t), y=np.ascontiguousarray(y), th=rec['th'], tx=rec['tx'],
                hyp=hyp, reg=reg, causal=causal, params=params, nh=len(rec['th']), nx=len(rec['tx']),
                recipe=dict(rec_args), causal_id=name.endswith('_cid'))


CASES = ['toy_test', 'toy_small', 'toy_acausal_model', 'ou', 'hrir', 'crude', 'sweep', 'sweep_hi']


def oracle_noise_floor(params, t, y, th, tx, reg, causal=True, trials=4, seed=0):
    """Conditioning noise of the path at ``params``: the spread of the oracle's own (elbo, grad) when every
    input (hyper-parameters, th) is moved by at most 2 ulp.  At trained points s2 is small and cond(Kh) ~ 1/reg,
    so two correct FP64 evaluations (the reference's TF graph on another machine, the oracle, the CUDA path)
    differ by this much; a parity bar below it is not attainable by *any* implementation.
    Returns (elbo_noise, grad_noise) as max-abs over the trials."""
    rng = np.random.default_rng(seed)
    e0, _, g0 = om.elbo_and_grad(params, t, y, th, tx, reg, causal)
    en, gn = 0.0, 0.0
    for _ in range(trials):
        p2 = np.array(params, dtype=np.float64)
        p2[:5] = p2[:5] * (1 + rng.integers(-2, 3, 5) * 1.1e-16)
        th2 = th * (1 + rng.integers(-2, 3, th.shape[0]) * 1.1e-16)
        e2, _, g2 = om.elbo_and_grad(p2, t, y, th2, tx, reg, causal)
        en = max(en, abs(e2 - e0))
        gn = max(gn, float(np.abs(g2 - g0).max()))
    return en, gn


def ulp_noise(fn, params, th, trials=3, seed=0):
    """Generic form of ``oracle_noise_floor``: ``fn(params, th)`` returns a tuple of scalars / arrays; the result is
    the per-entry max-abs deviation of that tuple when ``params[:5]`` and ``th`` move by at most 2 ulp."""
    rng = np.random.default_rng(seed)
    base = fn(params, th)
    noise = [0.0] * len(base)
    for _ in range(trials):
        p2 = np.array(params, dtype=np.float64)
        p2[:5] = p2[:5] * (1 + rng.integers(-2, 3, 5) * 1.1e-16)
        th2 = th * (1 + rng.integers(-2, 3, th.shape[0]) * 1.1e-16)
        out = fn(p2, th2)
        noise = [max(a, float(np.max(np.abs(np.asarray(x) - np.asarray(y))))) for a, x, y in zip(noise, out, base)]
    return base, noise
