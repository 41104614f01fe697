"""Generates tests/golden/quad/*.npz: the VCGPCM ELBO, its 7 terms and a set of directional derivatives evaluated in
IEEE binary128 (oracle/quad/elbo_quad.c) at the reference's own experiment shapes (src/tasks/{toy,ou,hrir,crude}.py
sizes, synthetic data of that shape), at two points each: the initial point of tools/named_shapes.py and a TRAINED
point (the variables after the reference's three-phase schedule, run on the GPU by tools/train_points.py), where FP64
evaluations carry conditioning noise.
These are the arbiter for "who is right" when two FP64 implementations differ by more than 1e-9
(tests/test_quad_truth.py on the CPU, tests/test_gpu_quad.py on the GPU).
Run:  python tools/make_quad_golden.py [shape ...]      (minutes per shape on 8 cores)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import model as om, quad  # noqa: E402

SHAPES = {  # name: (n, nx, nh, tau_w, tau_f, t-grid, reg)   -- tools/named_shapes.py
    'toy': (400, 150, 41, .1, .05, lambda n: np.linspace(0, 1, n), 1e-6),
    'ou': (600, 300, 75, .15, .025, lambda n: np.linspace(0, 1, n), 1e-5),
    'hrir': (400, 300, 151, 1.5e-3, 5e-5, lambda n: np.arange(n) / 44100., 1e-8),
    'crude': (400, 300, 101, 1., .1,
              lambda n: 2010 + 4 * np.sort(np.random.default_rng(0).choice(1013, n, replace=False)) / 1013., 1e-4),
}


def make_point(name, kind):
    if kind == 'trained':
        # a point the GPU trained with the reference's schedule (tools/train_points.py -> gpurun_out/points/<name>.npz)
        with np.load(os.path.join(ROOT, 'gpurun_out', 'points', name + '.npz')) as z:
            return dict(t=z['t'], y=z['y'], th=z['th'], tx=z['tx'], reg=float(z['reg']), params=z['params'],
                        nh=len(z['th']), nx=len(z['tx']))
    n, nx, nh, tau_w, tau_f, grid, reg = SHAPES[name]
    rng = np.random.default_rng(0)
    t = np.ascontiguousarray(grid(n))
    y = rng.standard_normal(n)
    y = (y - y.mean()) / y.std()
    rec = om.recipe(t, nx=nx, nh=nh, tau_w=tau_w, tau_f=tau_f, causal=True)
    hyp = (rec['alpha'], rec['gamma'], rec['omega'])
    mu_u, var_u = om.init_q(rec['th'], rec['alpha'], rec['gamma'], reg, rng)
    s2 = 0.1
    p = om.pack(s2, rec['s2_f'], hyp[0], hyp[1], hyp[2], mu_u, var_u)
    return dict(t=t, y=y, th=rec['th'], tx=rec['tx'], reg=reg, params=p, nh=len(rec['th']), nx=nx)


def directions(p, nh, rng):
    """Unit directions: the 5 hyper-parameter axes, 2 single entries of mu_u / var_u, 1-2 random full directions."""
    np_ = p.shape[0]
    out, names = [], []
    for i, nm in enumerate(['s2', 's2_f', 'alpha', 'gamma', 'omega']):
        v = np.zeros(np_); v[i] = 1; out.append(v); names.append(nm)
    for idx, nm in ((5 + nh // 2, 'mu_u[nh/2]'), (np_ - 1, 'var_u[-1]')):
        v = np.zeros(np_); v[idx] = 1; out.append(v); names.append(nm)
    for k in range(2 if nh < 100 else 1):
        v = rng.standard_normal(np_); v /= np.linalg.norm(v); out.append(v); names.append('random%d' % k)
    return np.stack(out), names


if __name__ == '__main__':
    out_dir = os.path.join(ROOT, 'tests', 'golden', 'quad')
    os.makedirs(out_dir, exist_ok=True)
    args = [a for a in sys.argv[1:] if not a.startswith('--')]
    kinds = ('trained',) if '--trained' in sys.argv else ('init',) if '--init' in sys.argv else ('init', 'trained')
    for name in (args or list(SHAPES)):
        for kind in kinds:
            if kind == 'trained' and not os.path.exists(os.path.join(ROOT, 'gpurun_out', 'points', name + '.npz')):
                continue
            c = make_point(name, kind)
            t0 = time.time()
            e_hi, e_lo, t_hi, t_lo = quad.elbo(c['params'], c['t'], c['y'], c['th'], c['tx'], c['reg'], True)
            dirs, names = directions(c['params'], c['nh'], np.random.default_rng(3))
            # Richardson pair on the first and the last direction only (it shows the truncation error: none at h = 1e-9)
            dd = np.array([quad.dderiv(c['params'], v, c['t'], c['y'], c['th'], c['tx'], c['reg'], True, h=1e-9,
                                       richardson=(k in (0, len(dirs) - 1))) for k, v in enumerate(dirs)])
            np.savez_compressed(os.path.join(out_dir, '%s_%s.npz' % (name, kind)), shape=name, kind=kind, t=c['t'],
                                y=c['y'], th=c['th'], tx=c['tx'], reg=c['reg'], params=c['params'], elbo=e_hi,
                                elbo_lo=e_lo, terms=t_hi, terms_lo=t_lo, dirs=dirs, dir_names=np.array(names),
                                dderiv=dd[:, 0], dderiv_h=dd[:, 1], dderiv_2h=dd[:, 2])
            print(name, kind, 'elbo %.15e' % e_hi, 'fd spread %.1e' % np.abs(dd[:, 1] - dd[:, 2]).max(),
                  '%.0f s' % (time.time() - t0), flush=True)
