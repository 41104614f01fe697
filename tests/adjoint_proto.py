"""numpy prototype of the *product's* formulation of the ELBO + gradient (reduced contractions and
the hand-written adjoint chain that the CUDA library implements).  It exists only so
that ``tests/test_adjoint_chain.py`` can check that formulation against autograd of the literal
oracle on the CPU, independently of any GPU.  Not imported by the product.
"""
import math
import numpy as np
from scipy.special import erfc
from oracle import bvn as obvn


def tril_unpack(v, m):
    L = np.zeros((m, m))
    L[np.tril_indices(m)] = v
    return L


def psi_scalars(alpha, gamma, omega):
    """Per-launch scalars of App. A.4/E and their derivatives w.r.t. (alpha, gamma, omega)."""
    A = alpha + gamma + omega
    det = 4 * (A * A - gamma * gamma)
    ddet = np.array([8 * A, 8 * A - 8 * gamma, 8 * A])
    S11 = 2 * A / det
    dS11 = 2 * np.ones(3) / det - 2 * A * ddet / det ** 2
    S12 = 2 * gamma / det
    dS12 = 2 * np.array([0., 1., 0.]) / det - 2 * gamma * ddet / det ** 2
    dom = np.array([0., 0., 1.])
    g1 = omega * (1 - 2 * omega * S11)
    dg1 = dom * (1 - 2 * omega * S11) + omega * (-2 * dom * S11 - 2 * omega * dS11)
    g2 = 4 * omega ** 2 * S12
    dg2 = 8 * omega * dom * S12 + 4 * omega ** 2 * dS12
    sq = math.sqrt(S11)
    p = 2 * omega * sq
    dp = 2 * dom * sq + omega * dS11 / sq
    q = 2 * omega * S12 / sq
    dq = 2 * dom * S12 / sq + 2 * omega * dS12 / sq - omega * S12 * dS11 / (S11 * sq)
    rho = gamma / A
    drho = np.array([0., 1., 0.]) / A - gamma / A ** 2
    return dict(A=A, det=det, ddet=ddet, g1=g1, dg1=dg1, g2=g2, dg2=dg2, p=p, dp=dp, q=q, dq=dq,
                rho=rho, drho=drho)


def axx_with_tangents(t, tx, alpha, gamma, omega):
    """sum_n Axx[n] and its three tangents (forward mode, App. E)."""
    s = psi_scalars(alpha, gamma, omega)
    d = t[:, None] - tx[None, :]
    dk, dl = d[:, :, None], d[:, None, :]
    G = -s['g1'] * (dk ** 2 + dl ** 2) + s['g2'] * dk * dl
    pref = 2 * math.pi / math.sqrt(s['det']) * np.exp(G)
    x1 = s['p'] * dk + s['q'] * dl
    x2 = s['q'] * dk + s['p'] * dl
    rho = np.full(x1.shape, s['rho'])
    cdf = obvn.bvn_cdf(x1, x2, rho)
    d1, d2, dr = obvn.bvn_cdf_partials(x1, x2, rho)
    V = pref * cdf
    out = [V.sum(0)]
    for i in range(3):
        Gt = -s['dg1'][i] * (dk ** 2 + dl ** 2) + s['dg2'][i] * dk * dl
        dV = V * (Gt - s['ddet'][i] / (2 * s['det'])) + pref * (
            d1 * (s['dp'][i] * dk + s['dq'][i] * dl) + d2 * (s['dq'][i] * dk + s['dp'][i] * dl)
            + dr * s['drho'][i])
        out.append(dV.sum(0))
    return out


def half_line(D, dD, b, db, c, dc):
    """F = 1/2 sqrt(pi/D) exp(c + b^2/4D) erfc(b / 2 sqrt D) and tangents (App. E)."""
    E = c + b * b / (4 * D)
    z = b / (2 * math.sqrt(D))
    F = .5 * math.sqrt(math.pi / D) * np.exp(E) * erfc(z)
    dF = []
    for i in range(len(dD)):
        Et = dc[i] + b * db[i] / (2 * D) - b * b * dD[i] / (4 * D * D)
        zt = db[i] / (2 * math.sqrt(D)) - z * dD[i] / (2 * D)
        dF.append(F * (Et - dD[i] / (2 * D)) - np.exp(E - z * z) / math.sqrt(D) * zt)
    return F, dF


def ahx_with_tangents(t, th, tx, alpha, gamma, omega):
    A = alpha + gamma + omega
    d = t[:, None, None] - tx[None, None, :]
    thi = th[None, :, None] + 0 * d
    b = -2 * gamma * thi - 2 * omega * d
    c = -(alpha + gamma) * thi ** 2 - omega * d ** 2
    return half_line(A, [1., 1., 1.], b, [0 * d, -2 * thi, -2 * d], c, [-thi ** 2, -thi ** 2, -d ** 2])


def ahh_with_tangents(th, alpha, gamma):
    B2 = 2 * (alpha + gamma)
    s = th[:, None] + th[None, :]
    q2 = th[:, None] ** 2 + th[None, :] ** 2
    b = -2 * gamma * s
    c = -(alpha + gamma) * q2
    F, dF = half_line(B2, [2., 2.], b, [0 * s, -2 * s], c, [-q2, -q2])
    return F, dF + [0 * F]


def windowed_gram(A, W, chunk, sign=1.0):
    """``sum_n A_n W A_n^T`` the way the CUDA path contracts it with option ``tri`` (csrc/cgpcm.cu: window_factors,
    right_mul_tri): per chunk of observations, the window ``w`` = the columns where the chunk's ``A`` is not exactly 0,
    ``sign * W[w, w] = L L^T`` (Cholesky: the block must be positive definite), ``V' = A[:, :, w] L`` and
    ``sign * sum V' V'^T``.  Returns the sum and the widest window."""
    N, nh, nx = A.shape
    out = np.zeros((nh, nh))
    widest = 0
    for n0 in range(0, N, chunk):
        Ac = A[n0:n0 + chunk]
        live = np.nonzero(np.any(Ac != 0.0, axis=(0, 1)))[0]
        if live.size == 0:
            continue
        lo, hi = live[0], live[-1] + 1
        widest = max(widest, hi - lo)
        Lw = np.linalg.cholesky(sign * W[lo:hi, lo:hi])
        V = Ac[:, :, lo:hi] @ Lw
        out += np.einsum('nil,njl->ij', V, V)
    return sign * out, widest


def elbo_grad(params, t, y, th, tx, reg, mode=1, frozen=None, tri_chunk=None):
    """ELBO, terms[7], gradient.  mode 1 = full regime, mode 0 = Psi frozen (``frozen`` = dict of
    sum_Axx, Q, Y, Ahh, a, A, iKh, iKx, Kx, logdetKx computed at the freeze-time hypers).  ``tri_chunk``: contract
    ``Q`` and ``Hbar`` through the Cholesky factors of the window blocks, in chunks of that many observations."""
    nh, nx, N = len(th), len(tx), len(t)
    s2, s2_f, alpha, gamma, omega = np.exp(params[:5])
    mu = params[5:5 + nh]
    L = tril_unpack(params[5 + nh:], nh)
    r, c0 = s2_f / s2, math.sqrt(s2_f) / s2
    I_h, I_x = np.eye(nh), np.eye(nx)
    if mode == 1:
        Kh0 = np.exp(-alpha * (th[:, None] ** 2 + th[None, :] ** 2) - gamma * (th[:, None] - th[None, :]) ** 2)
        dx2 = (tx[:, None] - tx[None, :]) ** 2
        Kx0 = math.sqrt(math.pi / (2 * omega)) * np.exp(-.5 * omega * dx2)
        Kh, Kx = Kh0 + reg * I_h, Kx0 + reg * I_x
        iKh, iKx = np.linalg.inv(Kh), np.linalg.inv(Kx)
        logdetKx = np.linalg.slogdet(Kx)[1]
        a = .5 * math.sqrt(math.pi / (2 * alpha))
        Ahh, dAhh = ahh_with_tangents(th, alpha, gamma)
        sum_Axx, dAxx_a, dAxx_g, dAxx_o = axx_with_tangents(t, tx, alpha, gamma, omega)
        dAxx = [dAxx_a, dAxx_g, dAxx_o]
        A, dA = ahx_with_tangents(t, th, tx, alpha, gamma, omega)
        Q = windowed_gram(A, iKx, tri_chunk)[0] if tri_chunk else np.einsum('nik,kl,njl->ij', A, iKx, A)
        Y = np.einsum('n,nik->ik', y, A)
    else:
        f = frozen
        Kx, iKh, iKx, logdetKx, a, Ahh = f['Kx'], f['iKh'], f['iKx'], f['logdetKx'], f['a'], f['Ahh']
        sum_Axx, A, Q, Y = f['sum_Axx'], f['A'], f['Q'], f['Y']
    var = L @ L.T + reg * I_h
    m2 = var + np.outer(mu, mu)
    H = m2 - iKh
    T1 = np.einsum('ij,njk->nik', H, A)
    C1 = np.einsum('nik,nil->kl', A, T1)
    S = sum_Axx + C1
    Pr = Kx + r * S + reg * I_x
    Pinv = np.linalg.inv(Pr)
    logdetP = np.linalg.slogdet(Pr)[1]
    lam = c0 * Y.T @ mu
    lbar = Pinv @ lam
    sum_b = N * a - N * np.sum(iKh * Ahh) - np.sum(iKx * sum_Axx) + np.sum(iKh * Q)
    sum_Bhh = N * Ahh - Q
    trace_term = np.sum(sum_Bhh * m2)
    So = iKh + reg * I_h
    iSo, ivar = np.linalg.inv(So), np.linalg.inv(var)
    KL = .5 * (np.sum(iSo * var) + mu @ iSo @ mu - nh + np.linalg.slogdet(So)[1] - np.linalg.slogdet(var)[1])
    terms = np.array([-.5 * N * math.log(2 * math.pi * s2) - .5 * np.sum(y ** 2) / s2,
                      .5 * logdetKx, -.5 * logdetP, .5 * lam @ lbar,
                      -.5 * r * sum_b, -.5 * r * trace_term, -KL])
    # ---- adjoints
    Pbar = -.5 * Pinv - .5 * np.outer(lbar, lbar)
    rbar = np.sum(Pbar * S) - .5 * sum_b - .5 * trace_term
    C1bar = r * Pbar
    c0bar = lbar @ (Y.T @ mu)
    Ybar = c0 * np.outer(mu, lbar)
    if tri_chunk:
        Hbar = windowed_gram(A, C1bar, tri_chunk, sign=-1.0)[0]      # -C1bar = r (Pinv / 2 + lbar lbar^T / 2) > 0
    else:
        U1 = np.einsum('nik,kl->nil', A, C1bar)
        Hbar = np.einsum('nil,njl->ij', U1, A)
    m2bar = Hbar - .5 * r * sum_Bhh
    varbar = m2bar - .5 * (iSo - ivar)
    Lbar = np.tril(2 * varbar @ L)
    mubar = 2 * m2bar @ mu + c0 * Y @ lbar - iSo @ mu
    g = np.zeros_like(params)
    g[0] = -r * rbar - c0 * c0bar - .5 * N + .5 * np.sum(y ** 2) / s2
    g[1] = r * rbar + .5 * c0 * c0bar
    g[5:5 + nh] = mubar
    g[5 + nh:] = Lbar[np.tril_indices(nh)]
    if mode == 1:
        Wx = r * (2 * Pbar + iKx)
        Abar = np.einsum('nik,kl->nil', T1, Wx) + y[:, None, None] * Ybar[None]
        gA = np.array([np.sum(Abar * dA[i]) for i in range(3)])
        Sobar = .5 * (iSo @ var @ iSo + np.outer(iSo @ mu, iSo @ mu) - iSo)
        iKhbar = -Hbar + .5 * r * N * Ahh - .5 * r * Q + Sobar
        iKxbar = .5 * r * C1 + .5 * r * sum_Axx
        Khbar = -iKh @ iKhbar @ iKh
        Kxbar = -iKx @ iKxbar @ iKx + Pbar + .5 * iKx
        Axxbar = r * Pbar + .5 * r * iKx
        abar = -.5 * r * N
        Ahhbar = .5 * r * N * (iKh - m2)
        dKh = [-(th[:, None] ** 2 + th[None, :] ** 2) * Kh0, -(th[:, None] - th[None, :]) ** 2 * Kh0, 0 * Kh0]
        dKx = [0 * Kx0, 0 * Kx0, -(1 / (2 * omega) + .5 * dx2) * Kx0]
        da = [-a / (2 * alpha), 0., 0.]
        hyp = np.array([alpha, gamma, omega])
        for i in range(3):
            gi = (np.sum(Khbar * dKh[i]) + np.sum(Kxbar * dKx[i]) + abar * da[i]
                  + np.sum(Ahhbar * dAhh[i]) + np.sum(Axxbar * dAxx[i]) + gA[i])
            g[2 + i] = hyp[i] * gi
    return terms.sum(), terms, g
