"""The synthetic scaling-sweep workload of SURVEY.md §8d (BASELINE.json configs[4]); shared by bench.py
and the full-size GPU tests.  numpy only — no oracle, no reference."""
import numpy as np


def _length_scale(ls):
    return (.5 * np.pi) * (.5 / ls ** 2)


def _f32(v):
    return float(np.float32(v))


def sweep_workload(n, m, seed=0, tau_w=.1, tau_f=.025, reg=1e-6, s2=0.1):
    """t = linspace(0, n/1000, n) (1 kHz), nh = nx = m, recipe of src/core/cgpcm.py:59-98 (causal)."""
    rng = np.random.default_rng(seed)
    t = np.linspace(0, n / 1000., n)
    alpha = 2 * _length_scale(tau_w)
    gamma = _length_scale(tau_f) - .5 * alpha
    s2_f = _f32((2 * alpha / np.pi) ** .5)
    gamma += 3. * alpha / 8.
    alpha /= 4.
    alpha, gamma = _f32(alpha), _f32(gamma)
    dtx = (t.max() - t.min()) / m
    omega = _f32(.5 * _length_scale(dtx))
    tx = np.linspace(t.min(), t.max(), m)
    th = np.linspace(0, 2 * tau_w, m)
    th = th - (th[1] - th[0]) * 2
    w = np.exp(-40 * np.linspace(-.3, .3, 601) ** 2)
    y = np.convolve(rng.standard_normal(n + 600), w, mode='valid')
    y = (y - y.mean()) / y.std()
    # q(u): mean 0.1 N(0, I), covariance factor = chol(reg(iKh)) as in src/core/cgpcm.py:439-445
    Kh = np.exp(-alpha * (th[:, None] ** 2 + th[None, :] ** 2) - gamma * (th[:, None] - th[None, :]) ** 2)
    Lh = np.linalg.cholesky(Kh + reg * np.eye(m))
    iLh = np.linalg.solve(Lh, np.eye(m))
    Lp = np.linalg.cholesky(iLh.T @ iLh + reg * np.eye(m))
    mu_u = .1 * rng.standard_normal(m)
    var_u = Lp[np.tril_indices(m)]
    params = np.concatenate([np.log([s2, s2_f, alpha, gamma, omega]), mu_u, var_u])
    return dict(t=np.ascontiguousarray(t), y=np.ascontiguousarray(y), th=th, tx=tx, hyp=(alpha, gamma, omega),
                reg=reg, params=params, nh=m, nx=m, n=n)
