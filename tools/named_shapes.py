"""Evaluation latency at the reference's own experiment shapes (BASELINE.json configs[0..3]; sizes from
src/tasks/{toy,ou,hrir,crude}.py, synthetic data of that shape): full and precomputed regime on the GPU, the CPU
oracle beside it.  One JSON line per shape."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, '.')
import cgpcm_b200
from oracle import model as om

SHAPES = {  # name: (n, nx, nh, tau_w, tau_f, t-grid, reg)
    'toy': (400, 150, 41, .1, .05, lambda n: np.linspace(0, 1, n), 1e-6),
    'ou': (600, 300, 75, .15, .025, lambda n: np.linspace(0, 1, n), 1e-5),
    'hrir': (400, 300, 151, 1.5e-3, 5e-5, lambda n: np.arange(n) / 44100., 1e-8),
    'crude': (400, 300, 101, 1., .1, lambda n: 2010 + 4 * np.sort(np.random.default_rng(0).choice(1013, n, replace=False)) / 1013., 1e-4),
}
cpu = '--no-cpu' not in sys.argv
for name, (n, nx, nh, tau_w, tau_f, grid, reg) in SHAPES.items():
    rng = np.random.default_rng(0)
    t = np.ascontiguousarray(grid(n))
    y = rng.standard_normal(n)
    y = (y - y.mean()) / y.std()
    rec = om.recipe(t, nx=nx, nh=nh, tau_w=tau_w, tau_f=tau_f, causal=True)
    hyp = (rec['alpha'], rec['gamma'], rec['omega'])
    mu_u, var_u = om.init_q(rec['th'], rec['alpha'], rec['gamma'], reg, rng)
    p = om.pack(0.1, rec['s2_f'], hyp[0], hyp[1], hyp[2], mu_u, var_u)
    eng = cgpcm_b200.Engine(len(rec['th']), nx)
    eng.set_data(t, y, rec['th'], rec['tx'])
    row = {'shape': name, 'n': n, 'nx': nx, 'nh': len(rec['th']), 'rho': hyp[1] / sum(hyp)}
    for mode, key in ((1, 'full'), (0, 'frozen')):
        if mode == 0:
            eng.precompute(*hyp, reg=reg)
        for _ in range(3):
            out = eng.elbo_grad(p, mode=mode, reg=reg)
        t0 = time.perf_counter()
        for _ in range(10):
            out = eng.elbo_grad(p, mode=mode, reg=reg)
        wall = (time.perf_counter() - t0) / 10
        tm = eng.last_timing()
        row[key] = {'wall_ms': round(wall * 1e3, 3), 'device_ms': round(tm['total_ms'], 3), 'launches': tm['launches'],
                    'elbo': out[0]}
    if cpu:
        om.PW_DISTS_EXACT = True
        t0 = time.perf_counter()
        e0, _, g0 = om.elbo_and_grad(p, t, y, rec['th'], rec['tx'], reg)
        row['cpu_oracle_full_ms'] = round((time.perf_counter() - t0) * 1e3, 1)
        full = eng.elbo_grad(p, mode=1, reg=reg)
        row['elbo_rel_diff_vs_oracle'] = abs(full[0] - e0) / abs(e0)
        row['grad_rel_diff_vs_oracle'] = float(np.abs(full[2] - g0).max() / np.abs(g0).max())
    print(json.dumps(row), flush=True)
    eng.close()
