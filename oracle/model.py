"""CPU oracle of the VCGPCM model path (torch-CPU float64).  TEST INFRASTRUCTURE ONLY.

Restates, operation by operation, the reference's

* recipe                       ``src/core/cgpcm.py:32-109``  (+ ``util.length_scale`` ``util.py:29-36``,
                               ``tf_util.to_float`` float32 round trip ``tf_util.py:108-115``),
* prior kernels                ``cgpcm.py:214-229`` (+ ``kernel.DEQ`` ``kernel.py:32-46``,
                               ``tf_util.pw_dists2`` ``tf_util.py:16-32``, ``reg`` ``:310-320``,
                               ``cholinv`` ``:271-278``),
* Psi statistics               ``cgpcm.py:111-203`` — (i) *generically*, by pushing the reference's own
                               integrands through :mod:`oracle.expq` on the reference's 6-D broadcast
                               layout, and (ii) by the closed forms of SURVEY.md App. A, chunked over n,
* model matrices               ``cgpcm.py:231-268``,
* q(u) and the optimal q(z)    ``cgpcm.py:435-477``,
* the ELBO and its 7 terms     ``cgpcm.py:518-575`` (+ ``Normal.m2/kl`` ``distribution.py:35-42,60-76``),

with gradients from torch autograd (the reference: ``tf.gradients``).  Parameter vector layout
(SURVEY.md §8b): ``[log s2, log s2_f, log alpha, log gamma, log omega, mu_u[nh], var_u[nh(nh+1)/2]]``.
"""
import math

import numpy as np
import torch

from . import expq
from .bvn_torch import bvn_cdf as _bvn_cdf_torch

DT = torch.float64
TERM_NAMES = ['s2 complexity', 'p(z) complexity', 'q*(z) complexity', 'q*(z) fit',
              'general conditioning penalty', 'q(u) conditioning penalty', '-KL[q(u)||p(u)]']


def T(x):
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x, dtype=np.float64))


# ----------------------------------------------------------------------------- recipe
def length_scale(ls):
    return (.5 * np.pi) * (.5 / ls ** 2)


def to_float(v):
    """``tf.cast(tf.to_float(v), float64)``: the value passes through float32."""
    return float(np.float32(v))


def recipe(t, nx, nh, tau_w, tau_f, causal, noise_init=1e-4, tx_range=None):
    """Hyper-parameters and inducing inputs of ``CGPCM.from_recipe`` (``cgpcm.py:51-98``)."""
    t = np.asarray(t, dtype=np.float64)
    alpha = 2 * length_scale(tau_w)
    gamma = length_scale(tau_f) - .5 * alpha
    s2_f = to_float((2 * alpha / np.pi) ** .5)
    if causal:
        gamma += 3. * alpha / 8.
        alpha /= 4.
    alpha, gamma = to_float(alpha), to_float(gamma)
    tx_range = (t.min(), t.max()) if tx_range is None else tx_range
    dtx = (tx_range[1] - tx_range[0]) / nx
    omega = to_float(.5 * length_scale(dtx))
    tx = np.linspace(tx_range[0], tx_range[1], nx)
    if not causal and nh % 2 == 0:
        nh += 1
    if causal:
        th = np.linspace(0, 2 * tau_w, nh)
        th = th - (th[1] - th[0]) * 2
    else:
        th = np.linspace(-tau_w, tau_w, nh)
    return dict(alpha=alpha, gamma=gamma, omega=omega, s2=to_float(noise_init), s2_f=s2_f,
                th=th, tx=tx, nh=nh, nx=nx, causal=bool(causal))


# ----------------------------------------------------------------------------- small helpers
def reg(x, r):
    return x + r * torch.eye(x.shape[-1], dtype=DT)


def cholinv(L):
    return torch.cholesky_solve(torch.eye(L.shape[-1], dtype=DT), L)


def log_det(L):
    return 2. * torch.sum(torch.log(torch.diagonal(L, dim1=-2, dim2=-1)), -1)


def trisolve(L, b):
    return torch.linalg.solve_triangular(L, b, upper=False)


def trmul(a, b):
    return torch.sum(a * b, dim=(-2, -1))


# The reference forms squared distances as |x|^2 - 2 x.y + |y|^2 (``tf_util.py:24-31``), which cancels
# catastrophically when the inputs carry a large offset (crude-oil time stamps are ~2010: ~7 digits
# of Kx are rounding noise).  ``PW_DISTS_EXACT = True`` switches the oracle to (x - y)^2, the exact value
# the reference approximates; tests state which variant they compare against and why.
PW_DISTS_EXACT = False


def deq(s2, alpha, gamma, x, y=None):
    """``DEQ._call`` through ``pw_dists2`` (``kernel.py:43-46``, ``tf_util.py:16-32``)."""
    y = x if y is None else y
    x, y = x[:, None], y[:, None]
    n2x = torch.sum(x ** 2, 1)[:, None]
    n2y = torch.sum(y ** 2, 1)[None, :]
    d2 = (x - y.T) ** 2 if PW_DISTS_EXACT else n2x - 2 * x @ y.T + n2y
    return s2 * torch.exp(-alpha * (n2x + n2y) - gamma * d2)


def tril_indices(m):
    return np.tril_indices(m)


def vec_to_tril(v):
    n = v.shape[0]
    m = int(((1 + 8 * n) ** .5 - 1) / 2)
    r, c = tril_indices(m)
    out = torch.zeros(m, m, dtype=DT)
    return out.index_put((torch.as_tensor(r), torch.as_tensor(c)), v)


def tril_to_vec(x):
    r, c = tril_indices(x.shape[-1])
    return x[torch.as_tensor(r), torch.as_tensor(c)]


def pack(s2, s2_f, alpha, gamma, omega, mu_u, var_u):
    head = np.log(np.array([s2, s2_f, alpha, gamma, omega], dtype=np.float64))
    return np.concatenate([head, np.asarray(mu_u, np.float64).ravel(), np.asarray(var_u, np.float64).ravel()])


def unpack(params, nh):
    p = T(params)
    s2, s2_f, alpha, gamma, omega = [torch.exp(p[i]) for i in range(5)]
    mu_u = p[5:5 + nh].reshape(nh, 1)
    var_u = p[5 + nh:]
    return s2, s2_f, alpha, gamma, omega, mu_u, var_u


# ----------------------------------------------------------------------------- Psi: generic route
def psi_generic(t, th, tx, alpha, gamma, omega, causal=True, causal_id=False):
    """The reference's own integrands (``cgpcm.py:111-121``) integrated by the restated
    ``integrate_box`` on the reference's 6-D axes (``cgpcm.py:129-154``): small sizes only.
    Returns ``a`` (scalar), ``Ahh`` [nh,nh], ``Axx`` [N,nx,nx], ``Ahx`` [N,nh,nx]."""
    t, th, tx = T(t), T(th), T(tx)
    v = expq.var
    tau1, tau2, t1, t2 = v('tau1'), v('tau2'), v('t1'), v('t2')
    th1, th2, tx1, tx2 = v('th1'), v('th2'), v('tx1'), v('tx2')
    kh = lambda x, y: expq.kh(alpha, gamma, x, y)
    kxs = lambda x, y: expq.kxs(omega, x, y)
    expq_a = kh(t1 - tau1, t2 - tau1)
    expq_Ahh = kh(t1 - tau1, th1) * kh(th2, t2 - tau1)
    expq_Axx = kh(t1 - tau1, t2 - tau2) * kxs(tau1, tx1) * kxs(tx2, tau2)
    expq_Ahx = kh(t1 - tau1, th1) * kxs(tau1, tx1)

    def ax(x, pre, post):
        return x.reshape((1,) * pre + (-1,) + (1,) * post)

    vm = {'th1': ax(th, 2, 3), 'th2': ax(th, 3, 2), 'tx1': ax(tx, 4, 1), 'tx2': ax(tx, 5, 0),
          't1': ax(t, 0, 5), 't2': ax(t, 1, 4)}
    vm['min_t1_tx1'] = torch.minimum(vm['t1'], vm['tx1'])
    vm['min_t1_tx2'] = torch.minimum(vm['t1'], vm['tx2'])
    ninf = -expq.inf
    if causal:
        up = t1
        up1, up2 = (v('min_t1_tx1'), v('min_t1_tx2')) if causal_id else (t1, t1)
        uphx = v('min_t1_tx1') if causal_id else t1
    else:
        up = up1 = up2 = uphx = expq.inf

    def sub(e, a, b):
        return e.substitute('t1', a).substitute('t2', b)

    n = t.shape[0]
    a = sub(expq_a, t1, t1).integrate_box(('tau1', ninf, up), **vm)
    a = (a * torch.ones(n, dtype=DT)).reshape(-1)[0]
    Ahh = sub(expq_Ahh, t1, t1).integrate_box(('tau1', ninf, up), **vm)
    Ahh = (Ahh * torch.ones(n, 1, 1, dtype=DT)).reshape(n, th.shape[0], th.shape[0])[0]
    Axx = sub(expq_Axx, t1, t1).integrate_box(('tau1', ninf, up1), ('tau2', ninf, up2), **vm)
    Axx = Axx.reshape(n, tx.shape[0], tx.shape[0])
    Ahx = sub(expq_Ahx, t1, t2).integrate_box(('tau1', ninf, uphx), **vm)
    Ahx = Ahx.reshape(n, th.shape[0], tx.shape[0])
    return a, Ahh, Axx, Ahx


# ----------------------------------------------------------------------------- Psi: closed forms
def psi_a(alpha, causal=True):
    a = torch.sqrt(math.pi / (2 * T(alpha)))
    return .5 * a if causal else a


def psi_Ahh(th, alpha, gamma, causal=True):
    th = T(th)
    B = alpha + gamma
    s = th[:, None] + th[None, :]
    b = -2 * gamma * s
    c = -B * (th[:, None] ** 2 + th[None, :] ** 2)
    pref = torch.sqrt(math.pi / (2 * B)) * torch.exp(c + b ** 2 / (8 * B))
    if causal:
        return .5 * pref * torch.special.erfc(b / (2 * torch.sqrt(2 * B)))
    return pref


def psi_Ahx(t, th, tx, alpha, gamma, omega, causal=True, causal_id=False):
    """[n, nh, nx]; SURVEY.md App. A.3.  ``causal_id`` (``cgpcm.py:194-203``): upper limit ``min(t, tx)``; the
    envelope ``exp(E)`` is the unconstrained maximum of the integrand and does not change, the argument of ``erfc``
    moves by the distance from the limit to ``t`` in conditional standard deviations: ``z += max(d, 0) sqrt(A)``."""
    t, th, tx = T(t), T(th), T(tx)
    A = alpha + gamma + omega
    d = t[:, None, None] - tx[None, None, :]
    thi = th[None, :, None]
    b = -2 * gamma * thi - 2 * omega * d
    E = -(alpha + gamma) * thi ** 2 - omega * d ** 2 + b ** 2 / (4 * A)
    pref = torch.sqrt(math.pi / A) * torch.exp(E)
    if causal:
        z = b / (2 * torch.sqrt(A))
        if causal_id:
            z = z + torch.clamp(d, min=0.) * torch.sqrt(A)
        return .5 * pref * torch.special.erfc(z)
    return pref


def psi_Axx(t, tx, alpha, gamma, omega, causal=True, causal_id=False):
    """[n, nx, nx]; SURVEY.md App. A.4.  ``causal_id`` (``cgpcm.py:168-180``): limits ``min(t, tx_k)``, ``min(t, tx_l)``;
    as for ``Ahx`` only the arguments of the CDF move: ``x_i -= max(d_i, 0) / sqrt(Sigma_11)``."""
    t, tx = T(t), T(tx)
    A = alpha + gamma + omega
    det = 4 * (A * A - gamma * gamma)
    S11, S12 = 2 * A / det, 2 * gamma / det
    d = t[:, None] - tx[None, :]
    dk, dl = d[:, :, None], d[:, None, :]
    g1 = omega * (1 - 2 * omega * S11)
    g2 = 4 * omega ** 2 * S12
    G = -g1 * (dk ** 2 + dl ** 2) + g2 * dk * dl
    pref = 2 * math.pi / torch.sqrt(det) * torch.exp(G)
    if not causal:
        return pref
    p = 2 * omega * torch.sqrt(S11)
    q = 2 * omega * S12 / torch.sqrt(S11)
    x1 = p * dk + q * dl
    x2 = q * dk + p * dl
    if causal_id:
        x1 = x1 - torch.clamp(dk, min=0.) / torch.sqrt(S11)
        x2 = x2 - torch.clamp(dl, min=0.) / torch.sqrt(S11)
    rho = (gamma / A) * torch.ones_like(x1)
    return pref * _bvn_cdf_torch(x1, x2, rho)


def psi_closed(t, th, tx, alpha, gamma, omega, causal=True, causal_id=False):
    alpha, gamma, omega = T(alpha), T(gamma), T(omega)
    return (psi_a(alpha, causal), psi_Ahh(th, alpha, gamma, causal),
            psi_Axx(t, tx, alpha, gamma, omega, causal, causal_id),
            psi_Ahx(t, th, tx, alpha, gamma, omega, causal, causal_id))


# ----------------------------------------------------------------------------- model matrices
def model_matrices(y, a, Ahh, Axx, Ahx, iKh, iKx):
    """``CGPCM._construct_model_matrices`` (``cgpcm.py:231-268``), the sums the ELBO reads."""
    y = T(y)
    n = Axx.shape[0]
    m = dict(a=a, Ahh=Ahh, Axx=Axx, Ahx=Ahx)
    m['sum_a'] = n * a
    m['sum_Axx'] = torch.sum(Axx, 0)
    m['sum_Ahh'] = n * Ahh
    m['sum_Ahx_y'] = torch.sum(y[:, None, None] * Ahx, 0)
    m['sum_b'] = (m['sum_a'] - trmul(iKh, m['sum_Ahh']) - trmul(iKx, m['sum_Axx'])
                  + torch.sum(trmul(iKh @ Ahx, Ahx @ iKx)))
    m['sum_Bxx'] = m['sum_Axx'] - torch.sum(Ahx.transpose(-1, -2) @ (iKh @ Ahx), 0)
    m['sum_Bhh'] = m['sum_Ahh'] - torch.sum(Ahx @ (iKx @ Ahx.transpose(-1, -2)), 0)
    return m


def prior_kernels(th, tx, alpha, gamma, omega, r):
    """``CGPCM._init_kernels`` (``cgpcm.py:214-229``)."""
    th, tx = T(th), T(tx)
    Kh = reg(deq(1., alpha, gamma, th), r)
    Lh = torch.linalg.cholesky(Kh)
    iKh = cholinv(Lh)
    Kx = reg(deq((.5 * math.pi / omega) ** .5, 0., .5 * omega, tx), r)
    Lx = torch.linalg.cholesky(Kx)
    iKx = cholinv(Lx)
    return dict(Kh=Kh, Lh=Lh, iKh=iKh, Kx=Kx, Lx=Lx, iKx=iKx)


def normal_kl(var_s, mean_s, var_o, mean_o):
    """``Normal.kl`` (``distribution.py:60-76``)."""
    cs = torch.linalg.cholesky(var_s)
    co = torch.linalg.cholesky(var_o)
    mu_diff = torch.sum(trisolve(co, mean_o - mean_s) ** 2)
    tr = torch.sum(trisolve(co, cs) ** 2)
    return .5 * (tr + mu_diff - var_s.shape[-1] + log_det(co) - log_det(cs))


def elbo_from_mats(m, k, n, sum_y2, s2, s2_f, mu_u, var_u, r):
    """``VCGPCM.elbo(z=True)`` with ``_optimal_q`` (``cgpcm.py:458-477,518-575``).  ``m``: model
    matrices (constants in the precomputed regime), ``k``: prior kernels."""
    Lq = vec_to_tril(var_u)
    h_var = reg(Lq @ Lq.T, r)
    h_mean = mu_u
    h_m2 = h_var + h_mean @ h_mean.T
    lam = s2_f ** .5 / s2 * (m['sum_Ahx_y'].T @ h_mean)
    S = m['sum_Bxx'] + torch.sum(m['Ahx'].transpose(-1, -2) @ (h_m2 @ m['Ahx']), 0)
    P = k['Kx'] + s2_f / s2 * S
    L = torch.linalg.cholesky(reg(P, r))
    trace_term = trmul(m['sum_Bhh'], h_m2)
    zero = torch.zeros(h_mean.shape, dtype=DT)
    terms = [-.5 * n * torch.log(2 * math.pi * s2) - .5 * sum_y2 / s2,
             .5 * log_det(k['Lx']),
             -.5 * log_det(L),
             .5 * torch.sum(trisolve(L, lam) ** 2),
             -.5 * s2_f / s2 * m['sum_b'],
             -.5 * s2_f / s2 * trace_term,
             -normal_kl(h_var, h_mean, reg(k['iKh'], r), zero)]
    return sum(terms), terms


def elbo_full(params, t, y, th, tx, r, causal=True, psi='closed', causal_id=False):
    """Full regime: Psi statistics rebuilt from the hyper-parameters in ``params``."""
    nh = len(th)
    p = T(params)
    s2, s2_f, alpha, gamma, omega, mu_u, var_u = unpack(p, nh)
    k = prior_kernels(th, tx, alpha, gamma, omega, r)
    fn = psi_closed if psi == 'closed' else psi_generic
    a, Ahh, Axx, Ahx = fn(t, th, tx, alpha, gamma, omega, causal, causal_id)
    m = model_matrices(y, a, Ahh, Axx, Ahx, k['iKh'], k['iKx'])
    y = T(y)
    return elbo_from_mats(m, k, y.shape[0], torch.sum(y ** 2), s2, s2_f, mu_u, var_u, r)


def elbo_and_grad(params, t, y, th, tx, r, causal=True, psi='closed', frozen=None, frozen_kernels='detached',
                  causal_id=False):
    """(elbo, terms[7], grad) as numpy.  ``frozen``: a ``(mats, kernels)`` pair of *detached*
    constants = the reference's precomputed regime (``cgpcm.py:270-284``).

    ``frozen_kernels='symbolic'`` is exactly what the reference does: ``precompute()`` replaces ``mats`` only; the
    prior kernels ``Kx, Lx, iKh`` (``cgpcm.py:214-229``) stay functions of the current ``alpha, gamma, omega``, so the
    value follows them and their gradient entries are not zero.  ``'detached'`` (the round-1 behaviour, kept for the
    tests of the q(u)-only training phases, which never move the hyper-parameters) also holds the kernels of the
    freeze point: then only ``log s2, log s2_f, mu_u, var_u`` receive gradient."""
    p = T(np.asarray(params, np.float64)).clone().requires_grad_(True)
    if frozen is None:
        e, terms = elbo_full(p, t, y, th, tx, r, causal, psi, causal_id)
    else:
        m, k = frozen
        nh = len(th)
        s2, s2_f, alpha, gamma, omega, mu_u, var_u = unpack(p, nh)
        if frozen_kernels == 'symbolic':
            k = prior_kernels(th, tx, alpha, gamma, omega, r)
        yy = T(y)
        e, terms = elbo_from_mats(m, k, yy.shape[0], torch.sum(yy ** 2), s2, s2_f, mu_u, var_u, r)
    g, = torch.autograd.grad(e, p, allow_unused=True)
    g = torch.zeros_like(p) if g is None else g
    return float(e.detach()), np.array([float(x.detach()) for x in terms]), g.numpy().copy()


def _chunk_sums(t_c, y_c, th, tx, alpha, gamma, omega, iKh, iKx, h_m2, causal, causal_id):
    """The sums over one chunk of observations that ``_construct_model_matrices`` and ``_optimal_q`` need
    (``cgpcm.py:240-267,473-475``), in the reference's operation order within the chunk."""
    a, Ahh, Axx, Ahx = psi_closed(t_c, th, tx, alpha, gamma, omega, causal, causal_id)
    AhxT = Ahx.transpose(-1, -2)
    return (torch.sum(Axx, 0), torch.sum(y_c[:, None, None] * Ahx, 0),
            torch.sum(trmul(iKh @ Ahx, Ahx @ iKx)),
            torch.sum(AhxT @ (iKh @ Ahx), 0), torch.sum(Ahx @ (iKx @ AhxT), 0), torch.sum(AhxT @ (h_m2 @ Ahx), 0))


def elbo_and_grad_chunked(params, t, y, th, tx, r, causal=True, causal_id=False, chunk=500):
    """``elbo_and_grad`` (full regime) for series too long to hold the ``N x nx x nx`` tensors and their autograd graph
    at once: the sums over observations are accumulated chunk by chunk under ``torch.utils.checkpoint`` (each chunk is
    recomputed in the backward pass).  Same operations as ``elbo_full``; only the order in which the per-chunk sums are
    added differs."""
    from torch.utils.checkpoint import checkpoint
    nh = len(th)
    p = T(np.asarray(params, np.float64)).clone().requires_grad_(True)
    s2, s2_f, alpha, gamma, omega, mu_u, var_u = unpack(p, nh)
    k = prior_kernels(th, tx, alpha, gamma, omega, r)
    Lq = vec_to_tril(var_u)
    h_var = reg(Lq @ Lq.T, r)
    h_m2 = h_var + mu_u @ mu_u.T
    tt, yy = T(t), T(y)
    n = tt.shape[0]
    acc = None
    for lo in range(0, n, chunk):
        out = checkpoint(_chunk_sums, tt[lo:lo + chunk], yy[lo:lo + chunk], T(th), T(tx), alpha, gamma, omega, k['iKh'],
                         k['iKx'], h_m2, causal, causal_id, use_reentrant=False)
        acc = out if acc is None else tuple(a_ + o_ for a_, o_ in zip(acc, out))
    sum_Axx, sum_Ahx_y, tr_cross, sum_AiKhA, sum_AiKxA, sum_Am2A = acc
    a = psi_a(alpha, causal)
    Ahh = psi_Ahh(th, alpha, gamma, causal)
    m = {'sum_Ahx_y': sum_Ahx_y,
         'sum_b': n * a - trmul(k['iKh'], n * Ahh) - trmul(k['iKx'], sum_Axx) + tr_cross,
         'sum_Bxx': sum_Axx - sum_AiKhA, 'sum_Bhh': n * Ahh - sum_AiKxA}
    lam = s2_f ** .5 / s2 * (m['sum_Ahx_y'].T @ mu_u)
    P = k['Kx'] + s2_f / s2 * (m['sum_Bxx'] + sum_Am2A)
    L = torch.linalg.cholesky(reg(P, r))
    zero = torch.zeros(mu_u.shape, dtype=DT)
    terms = [-.5 * n * torch.log(2 * math.pi * s2) - .5 * torch.sum(yy ** 2) / s2,
             .5 * log_det(k['Lx']),
             -.5 * log_det(L),
             .5 * torch.sum(trisolve(L, lam) ** 2),
             -.5 * s2_f / s2 * m['sum_b'],
             -.5 * s2_f / s2 * trmul(m['sum_Bhh'], h_m2),
             -normal_kl(h_var, mu_u, reg(k['iKh'], r), zero)]
    e = sum(terms)
    g, = torch.autograd.grad(e, p)
    return float(e.detach()), np.array([float(x.detach()) for x in terms]), g.numpy().copy()


def precompute(params, t, y, th, tx, r, causal=True, causal_id=False):
    """Detached model matrices + kernels at the hyper-parameters in ``params``."""
    with torch.no_grad():
        nh = len(th)
        s2, s2_f, alpha, gamma, omega, mu_u, var_u = unpack(T(np.asarray(params, np.float64)), nh)
        k = prior_kernels(th, tx, alpha, gamma, omega, r)
        a, Ahh, Axx, Ahx = psi_closed(t, th, tx, alpha, gamma, omega, causal, causal_id)
        m = model_matrices(y, a, Ahh, Axx, Ahx, k['iKh'], k['iKx'])
    return m, k


def init_q(th, alpha, gamma, r, rng, scale=0.1):
    """Deterministic stand-in for ``_init_inducing_points`` (``cgpcm.py:435-445``): the reference
    draws ``mu_u`` with the TF RNG (not reproducible), so tests *feed* ``mu_u``; ``var_u`` is
    ``tril_to_vec(chol(reg(iKh)))`` exactly as in the reference."""
    th = T(th)
    Kh = reg(deq(1., T(alpha), T(gamma), th), r)
    iKh = cholinv(torch.linalg.cholesky(Kh))
    var_u = tril_to_vec(torch.linalg.cholesky(reg(iKh, r))).numpy()
    mu_u = scale * rng.standard_normal(len(th))
    return mu_u, var_u


# ----------------------------------------------------------------------------- fixed-point iteration
def optimal_q(m, k, s2, s2_f, mean, m2, z):
    """``VCGPCM._optimal_q`` (``cgpcm.py:458-477``): natural parameters ``(lam, P)`` of the optimal q(z) given the
    moments of q(u) (``z=True``) or of the optimal q(u) given the moments of q(z) (``z=False``)."""
    if z:
        lam = s2_f ** .5 / s2 * (m['sum_Ahx_y'].T @ mean)
        S = m['sum_Bxx'] + torch.sum(m['Ahx'].transpose(-1, -2) @ (m2 @ m['Ahx']), 0)
        K = k['Kx']
    else:
        lam = s2_f ** .5 / s2 * (m['sum_Ahx_y'] @ mean)
        S = m['sum_Bhh'] + torch.sum(m['Ahx'] @ (m2 @ m['Ahx'].transpose(-1, -2)), 0)
        K = k['Kh']
    return lam, K + s2_f / s2 * S


def from_natural(P, lam, r):
    """``Normal.from_natural`` (``distribution.py:20-33``): ``(mean, var)`` of ``N(reg(P)^-1 lam, reg(P)^-1)``."""
    L = torch.linalg.cholesky(reg(P, r))
    return torch.cholesky_solve(lam, L), cholinv(L)


def fpi(params, t, y, th, tx, r, num, causal=True, high_reg=False, causal_id=False):
    """``VCGPCM.fpi(num, z=True, high_reg)`` (``cgpcm.py:479-516``): ``num`` rounds of
    q(u) -> optimal q(z) -> optimal q(u).  Returns ``(mu_u, var_u)`` as the reference assigns them
    (``var_u = tril_to_vec(cholesky(var))``) plus the last q(z) as ``(mu_z, var_z)`` in the same packing (what
    ``convert(z=True)``, ``cgpcm.py:577-592``, assigns when called with ``num = 0`` rounds before it)."""
    with torch.no_grad():
        nh = len(th)
        s2, s2_f, alpha, gamma, omega, mu_u, var_u = unpack(T(np.asarray(params, np.float64)), nh)
        k = prior_kernels(th, tx, alpha, gamma, omega, r)
        a, Ahh, Axx, Ahx = psi_closed(t, th, tx, alpha, gamma, omega, causal, causal_id)
        m = model_matrices(y, a, Ahh, Axx, Ahx, k['iKh'], k['iKx'])
        Lq = vec_to_tril(var_u)
        mean, var = mu_u, reg(Lq @ Lq.T, r)                       # self.h (cgpcm.py:444-445)
        for _ in range(num):
            lam, P = optimal_q(m, k, s2, s2_f, mean, var + mean @ mean.T, True)
            if high_reg:
                P = reg(P, 1e-4)
            mz, vz = from_natural(P, lam, r)
            lam, P = optimal_q(m, k, s2, s2_f, mz, vz + mz @ mz.T, False)
            if high_reg:
                P = reg(P, 1e-4)
            mean, var = from_natural(P, lam, r)
        lam, P = optimal_q(m, k, s2, s2_f, mean, var + mean @ mean.T, True)   # convert(z=True)
        mz, vz = from_natural(P, lam, r)
        out = [mean.numpy().ravel().copy(), tril_to_vec(torch.linalg.cholesky(var)).numpy().copy(),
               mz.numpy().ravel().copy(), tril_to_vec(torch.linalg.cholesky(vz)).numpy().copy()]
    return out


def elbo_qz(params, mu_z, var_z, t, y, th, tx, r, causal=True, causal_id=False):
    """``VCGPCM.elbo(z=False)`` (``cgpcm.py:518-575``, the ``z=False`` branches): the bound saturated for q(u) with the
    explicit ``q(z) = N(mu_z, reg(Lz Lz^T))``.  Returns ``(elbo, terms[7])``."""
    with torch.no_grad():
        nh = len(th)
        s2, s2_f, alpha, gamma, omega, _, _ = unpack(T(np.asarray(params, np.float64)), nh)
        k = prior_kernels(th, tx, alpha, gamma, omega, r)
        a, Ahh, Axx, Ahx = psi_closed(t, th, tx, alpha, gamma, omega, causal, causal_id)
        m = model_matrices(y, a, Ahh, Axx, Ahx, k['iKh'], k['iKx'])
        yy = T(y)
        n, sum_y2 = yy.shape[0], torch.sum(yy ** 2)
        mz = T(np.asarray(mu_z, np.float64)).reshape(-1, 1)
        Lz = vec_to_tril(T(np.asarray(var_z, np.float64)))
        x_var = reg(Lz @ Lz.T, r)
        x_m2 = x_var + mz @ mz.T
        lam, P = optimal_q(m, k, s2, s2_f, mz, x_m2, False)
        L = torch.linalg.cholesky(reg(P, r))
        zero = torch.zeros(mz.shape, dtype=DT)
        terms = [-.5 * n * torch.log(2 * math.pi * s2) - .5 * sum_y2 / s2,
                 .5 * log_det(k['Lh']),
                 -.5 * log_det(L),
                 .5 * torch.sum(trisolve(L, lam) ** 2),
                 -.5 * s2_f / s2 * m['sum_b'],
                 -.5 * s2_f / s2 * trmul(m['sum_Bxx'], x_m2),
                 -normal_kl(x_var, mz, reg(k['iKx'], r), zero)]
    return float(sum(terms)), np.array([float(x) for x in terms])


def fpi_qz(params, mu_z, var_z, t, y, th, tx, r, num, causal=True, high_reg=False, causal_id=False):
    """``VCGPCM.fpi(num, z=False, high_reg)`` followed by ``convert(z=False)`` (``cgpcm.py:479-516,577-592``):
    ``(mu_u, var_u, mu_z, var_z)`` in the reference's packing."""
    with torch.no_grad():
        nh = len(th)
        s2, s2_f, alpha, gamma, omega, _, _ = unpack(T(np.asarray(params, np.float64)), nh)
        k = prior_kernels(th, tx, alpha, gamma, omega, r)
        a, Ahh, Axx, Ahx = psi_closed(t, th, tx, alpha, gamma, omega, causal, causal_id)
        m = model_matrices(y, a, Ahh, Axx, Ahx, k['iKh'], k['iKx'])
        mean = T(np.asarray(mu_z, np.float64)).reshape(-1, 1)
        Lz = vec_to_tril(T(np.asarray(var_z, np.float64)))
        var = reg(Lz @ Lz.T, r)
        for _ in range(num):
            lam, P = optimal_q(m, k, s2, s2_f, mean, var + mean @ mean.T, False)
            if high_reg:
                P = reg(P, 1e-4)
            mu, vu = from_natural(P, lam, r)
            lam, P = optimal_q(m, k, s2, s2_f, mu, vu + mu @ mu.T, True)
            if high_reg:
                P = reg(P, 1e-4)
            mean, var = from_natural(P, lam, r)
        lam, P = optimal_q(m, k, s2, s2_f, mean, var + mean @ mean.T, False)    # convert(z=False)
        mu, vu = from_natural(P, lam, r)
        return [mu.numpy().ravel().copy(), tril_to_vec(torch.linalg.cholesky(vu)).numpy().copy(),
                mean.numpy().ravel().copy(), tril_to_vec(torch.linalg.cholesky(var)).numpy().copy()]


# ----------------------------------------------------------------------------- SMF bound and sampler target
def elbo_smf(params, t, y, th, tx, r, sample, causal=True, causal_id=False):
    """``VCGPCM.elbo(smf=True, sample=sample)`` (``cgpcm.py:518-575``) and the pseudo-log-likelihood of
    ``VCGPCM.sample`` (``cgpcm.py:848-866``) at ``h = sample``: ``(elbo, terms[7], log_lik)``."""
    with torch.no_grad():
        nh = len(th)
        s2, s2_f, alpha, gamma, omega, mu_u, var_u = unpack(T(np.asarray(params, np.float64)), nh)
        k = prior_kernels(th, tx, alpha, gamma, omega, r)
        a, Ahh, Axx, Ahx = psi_closed(t, th, tx, alpha, gamma, omega, causal, causal_id)
        m = model_matrices(y, a, Ahh, Axx, Ahx, k['iKh'], k['iKx'])
        yy = T(y)
        n, sum_y2 = yy.shape[0], torch.sum(yy ** 2)
        h = T(np.asarray(sample, np.float64)).reshape(-1, 1)
        lam, P = optimal_q(m, k, s2, s2_f, h, h @ h.T, True)
        L = torch.linalg.cholesky(reg(P, r))
        Lq = vec_to_tril(var_u)
        h_var = reg(Lq @ Lq.T, r)
        h_m2 = h_var + mu_u @ mu_u.T
        zero = torch.zeros(mu_u.shape, dtype=DT)
        terms = [-.5 * n * torch.log(2 * math.pi * s2) - .5 * sum_y2 / s2,
                 .5 * log_det(k['Lx']),
                 -.5 * log_det(L),
                 .5 * torch.sum(trisolve(L, lam) ** 2),
                 -.5 * s2_f / s2 * m['sum_b'],
                 -.5 * s2_f / s2 * trmul(m['sum_Bhh'], h_m2),
                 -normal_kl(h_var, mu_u, reg(k['iKh'], r), zero)]
        L0 = torch.linalg.cholesky(P)                                    # sample(): no jitter (cgpcm.py:856)
        ll = (-.5 * log_det(L0) + .5 * torch.sum(trisolve(L0, lam) ** 2)
              - .5 * s2_f / s2 * (h.T @ m['sum_Bhh'] @ h).squeeze())
    return float(sum(terms)), np.array([float(x) for x in terms]), float(ll)


# ----------------------------------------------------------------------------- function prediction
def predict_f(params, t, y, th, tx, r, t_star, samples_h, smf=False, causal=True, causal_id=False):
    """``VCGPCM.predict_f`` (``cgpcm.py:781-846``) for given filter samples ``samples_h`` ([B][nh]): posterior mean
    and variance of the function at ``t_star``, averaged over the samples.  ``smf=False``: q(z) is the optimal q(z) of
    q(u) (the reference's numeric ``samples_h``: draws from q(u)); ``smf=True``: q(z | h) per sample."""
    with torch.no_grad():
        nh = len(th)
        s2, s2_f, alpha, gamma, omega, mu_u, var_u = unpack(T(np.asarray(params, np.float64)), nh)
        k = prior_kernels(th, tx, alpha, gamma, omega, r)
        a, Ahh, Axx, Ahx = psi_closed(t, th, tx, alpha, gamma, omega, causal, causal_id)
        m = model_matrices(y, a, Ahh, Axx, Ahx, k['iKh'], k['iKx'])
        # test-side statistics (cgpcm.py:793: _construct_model_matrices(Data(t)))
        a_s, Ahh_s, Axx_s, Ahx_s = psi_closed(t_star, th, tx, alpha, gamma, omega, causal, causal_id)
        Lq = vec_to_tril(var_u)
        h_var = reg(Lq @ Lq.T, r)
        mus, vars_ = [], []
        for hs in samples_h:
            h = T(np.asarray(hs, np.float64)).reshape(-1, 1)
            if smf:
                lam, P = optimal_q(m, k, s2, s2_f, h, h @ h.T, True)
            else:
                lam, P = optimal_q(m, k, s2, s2_f, mu_u, h_var + mu_u @ mu_u.T, True)
            xm, xv = from_natural(P, lam, r)
            mu = s2_f ** .5 * (h.T @ Ahx_s @ xm).reshape(-1)                         # [n*]
            mh = h @ h.T - k['iKh']
            mx = xv + xm @ xm.T - k['iKx']
            m2 = s2_f * (a_s + trmul(Ahh_s, mh) + torch.sum(Axx_s * mx, (-1, -2))
                         + torch.sum((mh @ Ahx_s) * (Ahx_s @ mx), (-1, -2)))
            mus.append(mu)
            vars_.append(m2 - mu ** 2)
        mu = torch.mean(torch.stack(mus, 1), 1)
        var = torch.mean(torch.stack(vars_, 1), 1)
    return mu.numpy().copy(), var.numpy().copy()


# ----------------------------------------------------------------------------- kernel prediction (first part)
def psi_center_generic(t, th, alpha, gamma, causal=True):
    """``_a_center`` / ``_Ahh_center`` (``cgpcm.py:164-166,190-192``) through the restated ``integrate_box`` on the
    reference's integrands: ``t2 = 0``, upper limit ``min(t1, 0)`` (causal) or ``inf``.  ``a`` [n], ``Ahh`` [n,nh,nh]."""
    t, th = T(t), T(th)
    v = expq.var
    tau1, t1, th1, th2 = v('tau1'), v('t1'), v('th1'), v('th2')
    kh = lambda x, y: expq.kh(alpha, gamma, x, y)
    zero = expq.const(0)
    expq_a = kh(t1 - tau1, zero - tau1)
    expq_Ahh = kh(t1 - tau1, th1) * kh(th2, zero - tau1)
    n, nh = t.shape[0], th.shape[0]
    vm = {'t1': t.reshape(-1, 1, 1), 'th1': th.reshape(1, -1, 1), 'th2': th.reshape(1, 1, -1)}
    vm['min_t1_0'] = torch.minimum(vm['t1'], torch.zeros(1, dtype=DT))
    up = v('min_t1_0') if causal else expq.inf
    a = torch.as_tensor(expq_a.integrate_box(('tau1', -expq.inf, up), **vm))       # squeezed like tf.squeeze: [n]
    a = a.reshape(-1) * torch.ones(n, dtype=DT)
    Ahh = torch.as_tensor(expq_Ahh.integrate_box(('tau1', -expq.inf, up), **vm))
    return a, Ahh.reshape(n, nh, nh)


def psi_center_closed(t, th, alpha, gamma, causal=True):
    """Closed forms of the same: with ``B = alpha + gamma``, ``U = min(t, 0)``,
    ``a_c(t) = 1/2 sqrt(pi/2alpha) exp(-(alpha/2 + gamma) t^2) erfc(sqrt(2 alpha) |t| / 2)`` and
    ``Ahh_c(t)[i,j] = 1/2 sqrt(pi/2B) exp(c + b^2/8B) erfc(sqrt(2B) (b/4B - U))``, ``b = 2B t - 2 gamma (th_i + th_j)``,
    ``c = -alpha (t^2 + th_i^2) - gamma (t - th_i)^2 - B th_j^2``; acausal: twice the prefactor, no erfc."""
    t, th = T(t), T(th)
    alpha, gamma = T(alpha), T(gamma)
    B = alpha + gamma
    a = torch.sqrt(math.pi / (2 * alpha)) * torch.exp(-(.5 * alpha + gamma) * t ** 2)
    tt = t.reshape(-1, 1, 1)
    thi, thj = th.reshape(1, -1, 1), th.reshape(1, 1, -1)
    b = 2 * B * tt - 2 * gamma * (thi + thj)
    c = -alpha * (tt ** 2 + thi ** 2) - gamma * (tt - thi) ** 2 - B * thj ** 2
    Ahh = torch.sqrt(math.pi / (2 * B)) * torch.exp(c + b ** 2 / (8 * B))
    if causal:
        U = torch.minimum(tt, torch.zeros(1, dtype=DT))
        a = .5 * a * torch.special.erfc(torch.sqrt(2 * alpha) * torch.abs(t) / 2)
        Ahh = .5 * Ahh * torch.special.erfc(torch.sqrt(2 * B) * (b / (4 * B) - U))
    return a, Ahh


def kernel_samples(params, th, r, t, samples_h, causal=True):
    """The Monte-Carlo kernel samples of ``VCGPCM.predict_k`` (``cgpcm.py:610-634``), before normalisation:
    ``k[n, b] = s2_f (a_c(t_n) + tr((h_b h_b^T - iKh) Ahh_c(t_n)))``."""
    with torch.no_grad():
        nh = len(th)
        s2, s2_f, alpha, gamma, omega, _, _ = unpack(T(np.asarray(params, np.float64)), nh)
        thT = T(th)
        Kh = reg(deq(1., alpha, gamma, thT), r)
        iKh = cholinv(torch.linalg.cholesky(Kh))
        a, Ahh = psi_center_closed(t, th, alpha, gamma, causal)
        cols = []
        for hs in samples_h:
            h = T(np.asarray(hs, np.float64)).reshape(-1, 1)
            cols.append(s2_f * (a + torch.sum((h @ h.T - iKh) * Ahh, (-1, -2))))
    return torch.stack(cols, 1).numpy().copy()


# ----------------------------------------------------------------------------- filter prediction (second part)
def filter_samples(params, th, r, t, samples_h, noise):
    """The posterior draws of the filter that ``VCGPCM.predict_h`` / ``predict_psd`` (``cgpcm.py:663-779``) post-process:
    ``Kuh = k_h(th, t)``, ``A = Lh^-1 Kuh``, ``L = chol(reg(k_h(t, t) - A^T A))``,
    ``sample_b = Kuh^T h_b + L eps_b`` with ``noise[:, b] = eps_b`` (the reference draws ``randn`` in the graph).
    Returns [n, B]."""
    with torch.no_grad():
        nh = len(th)
        s2, s2_f, alpha, gamma, omega, _, _ = unpack(T(np.asarray(params, np.float64)), nh)
        thT, tt = T(th), T(t)
        Kh = reg(deq(1., alpha, gamma, thT), r)
        Lh = torch.linalg.cholesky(Kh)
        Kuh = deq(1., alpha, gamma, thT, tt)
        A = trisolve(Lh, Kuh)
        L = torch.linalg.cholesky(reg(deq(1., alpha, gamma, tt) - A.T @ A, r))
        H = T(np.asarray(samples_h, np.float64)).reshape(-1, nh).T          # [nh, B]
        out = Kuh.T @ H + L @ T(np.asarray(noise, np.float64))
    return out.numpy().copy()


# ----------------------------------------------------------------------------- AKM sampler (third part)
def psi_pairs_generic(t, th, alpha, gamma, causal=True):
    """``_a(t)`` / ``_Ahh(t)`` at all pairs of inputs (``cgpcm.py:156-158,182-184``) through the restated
    ``integrate_box`` on the reference's integrands: ``t1``, ``t2`` on axes 0, 1, upper limit ``min(t1, t2)`` (causal)
    or ``inf``.  ``a`` [n,n], ``Ahh`` [n,n,nh,nh].  Small sizes only."""
    t, th = T(t), T(th)
    v = expq.var
    tau1, t1, t2, th1, th2 = v('tau1'), v('t1'), v('t2'), v('th1'), v('th2')
    kh = lambda x, y: expq.kh(alpha, gamma, x, y)
    expq_a = kh(t1 - tau1, t2 - tau1)
    expq_Ahh = kh(t1 - tau1, th1) * kh(th2, t2 - tau1)
    n, nh = t.shape[0], th.shape[0]
    vm = {'t1': t.reshape(-1, 1, 1, 1), 't2': t.reshape(1, -1, 1, 1), 'th1': th.reshape(1, 1, -1, 1),
          'th2': th.reshape(1, 1, 1, -1)}
    vm['min_t1_t2'] = torch.minimum(vm['t1'], vm['t2'])
    up = v('min_t1_t2') if causal else expq.inf
    a = torch.as_tensor(expq_a.integrate_box(('tau1', -expq.inf, up), **vm))
    a = a.reshape(n, n) if a.numel() == n * n else a.reshape(-1)[0] * torch.ones(n, n, dtype=DT)
    Ahh = torch.as_tensor(expq_Ahh.integrate_box(('tau1', -expq.inf, up), **vm))
    return a, (Ahh * torch.ones(n, n, nh, nh, dtype=DT)).reshape(n, n, nh, nh)


def akm_f(params, th, r, t, h, e, causal=True):
    """``AKM.f()`` (``cgpcm.py:382-392``) on the pair statistics of ``psi_pairs_generic``:
    ``K = reg(a + tr((h h^T - iKh) Ahh))``, ``f = sqrt(s2_f) chol(K) e``.  Returns ``(f, K)``."""
    with torch.no_grad():
        nh = len(th)
        s2, s2_f, alpha, gamma, omega, _, _ = unpack(T(np.asarray(params, np.float64)), nh)
        Kh = reg(deq(1., alpha, gamma, T(th)), r)
        iKh = cholinv(torch.linalg.cholesky(Kh))
        a, Ahh = psi_pairs_generic(t, th, alpha, gamma, causal)
        hv = T(np.asarray(h, np.float64)).reshape(-1, 1)
        K = reg(a + torch.sum((hv @ hv.T - iKh) * Ahh, (-1, -2)), r)
        f = torch.sqrt(s2_f) * (torch.linalg.cholesky(K) @ T(np.asarray(e, np.float64)).reshape(-1, 1))
    return f.reshape(-1).numpy().copy(), K.numpy().copy()
