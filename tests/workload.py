"""The synthetic scaling-sweep workload of SURVEY.md §8d (BASELINE.json configs[4]); shared by bench.py
and the full-size GPU tests.  numpy only — no oracle, no reference."""
import numpy as np


def _length_scale(ls):
    return (.5 * np.pi) * (.5 / ls ** 2)


def _f32(v):
    return float(np.float32(v))


def _chol_fixed(a):
    """Lower Cholesky factor with a fixed operation order: column by column, every inner product an elementwise
    multiply + numpy's own (pairwise, single-threaded) reduction.  LAPACK's blocked factorisation sums in an order that
    depends on the BLAS thread count (OMP_NUM_THREADS is 1 under torchrun, the core count otherwise), which made the
    round-1 bench inputs differ in the last bits between the 1-GPU and the N-GPU runs."""
    a = np.array(a, dtype=np.float64)
    m = a.shape[0]
    L = np.zeros_like(a)
    for j in range(m):
        d = a[j, j] - np.sum(L[j, :j] * L[j, :j])
        L[j, j] = np.sqrt(d)
        if j + 1 < m:
            L[j + 1:, j] = (a[j + 1:, j] - np.sum(L[j + 1:, :j] * L[j, :j][None, :], axis=1)) / L[j, j]
    return L


def _tri_inv_fixed(L):
    """Inverse of a lower-triangular matrix by forward substitution, row by row, in a fixed order."""
    m = L.shape[0]
    X = np.zeros_like(L)
    for r in range(m):
        X[r, :] = -np.sum(L[r, :r][:, None] * X[:r, :], axis=0)
        X[r, r] += 1.0
        X[r, :] /= L[r, r]
    return X


def _gram_fixed(X):
    """X^T X with numpy's own sum-of-products loops (np.einsum without optimisation does not call BLAS)."""
    return np.einsum('ki,kj->ij', X, X)


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
    x = rng.standard_normal(n + 600)
    y = np.zeros(n)
    for j in range(601):                       # 'valid' convolution, tap by tap: a fixed summation order (no BLAS dot)
        y += w[600 - j] * x[j:j + n]
    y = (y - y.mean()) / y.std()
    # q(u): mean 0.1 N(0, I), covariance factor = chol(reg(iKh)) as in src/core/cgpcm.py:439-445
    Kh = np.exp(-alpha * (th[:, None] ** 2 + th[None, :] ** 2) - gamma * (th[:, None] - th[None, :]) ** 2)
    Lh = _chol_fixed(Kh + reg * np.eye(m))
    iLh = _tri_inv_fixed(Lh)
    Lp = _chol_fixed(_gram_fixed(iLh) + reg * np.eye(m))
    mu_u = .1 * rng.standard_normal(m)
    var_u = Lp[np.tril_indices(m)]
    params = np.concatenate([np.log([s2, s2_f, alpha, gamma, omega]), mu_u, var_u])
    return dict(t=np.ascontiguousarray(t), y=np.ascontiguousarray(y), th=th, tx=tx, hyp=(alpha, gamma, omega),
                reg=reg, params=params, nh=m, nx=m, n=n)


# The reference's own experiment shapes (BASELINE.json configs[0..3]; sizes from src/tasks/{toy,ou,hrir,crude}.py), synthetic
# data of that shape.  name: (n, nx, nh, tau_w, tau_f, t-grid, reg)
NAMED_SHAPES = {
    'toy': (400, 150, 41, .1, .05, lambda n: np.linspace(0, 1, n), 1e-6),
    'ou': (600, 300, 75, .15, .025, lambda n: np.linspace(0, 1, n), 1e-5),
    'hrir': (400, 300, 151, 1.5e-3, 5e-5, lambda n: np.arange(n) / 44100., 1e-8),
    'crude': (400, 300, 101, 1., .1,
              lambda n: 2010 + 4 * np.sort(np.random.default_rng(0).choice(1013, n, replace=False)) / 1013., 1e-4),
}


def named_workload(name, seed=0, s2=0.1):
    """Inputs of one evaluation at a named shape: recipe of src/core/cgpcm.py:59-98 (causal), white-noise observations
    of unit variance, q(u) at the prior (mean 0.1 N(0, I)).  Same keys as `sweep_workload`."""
    n, nx, nh, tau_w, tau_f, grid, reg = NAMED_SHAPES[name]
    rng = np.random.default_rng(seed)
    t = np.ascontiguousarray(grid(n))
    y = rng.standard_normal(n)
    y = (y - y.mean()) / y.std()
    alpha = 2 * _length_scale(tau_w)
    gamma = _length_scale(tau_f) - .5 * alpha
    s2_f = _f32((2 * alpha / np.pi) ** .5)
    gamma += 3. * alpha / 8.
    alpha /= 4.
    alpha, gamma = _f32(alpha), _f32(gamma)
    omega = _f32(.5 * _length_scale((t.max() - t.min()) / nx))
    tx = np.linspace(t.min(), t.max(), nx)
    th = np.linspace(0, 2 * tau_w, nh)
    th = th - (th[1] - th[0]) * 2
    Kh = np.exp(-alpha * (th[:, None] ** 2 + th[None, :] ** 2) - gamma * (th[:, None] - th[None, :]) ** 2)
    Lp = _chol_fixed(_gram_fixed(_tri_inv_fixed(_chol_fixed(Kh + reg * np.eye(nh)))) + reg * np.eye(nh))
    mu_u = .1 * rng.standard_normal(nh)
    params = np.concatenate([np.log([s2, s2_f, alpha, gamma, omega]), mu_u, Lp[np.tril_indices(nh)]])
    return dict(t=t, y=np.ascontiguousarray(y), th=th, tx=tx, hyp=(alpha, gamma, omega), reg=reg, params=params,
                nh=nh, nx=nx, n=n)
