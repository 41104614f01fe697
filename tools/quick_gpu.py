"""Scratch timing of one evaluation at the sweep shape (not the bench)."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import cgpcm_b200
from tests.workload import sweep_workload
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 200
wl = sweep_workload(n, m)
eng = cgpcm_b200.Engine(m, m)
eng.set_data(wl['t'], wl['y'], wl['th'], wl['tx'])
for cull, chunk in [(80.0, 256), (80.0, 1024), (0.0, 256), (0.0, 512)]:
    eng.set_option('cull', cull); eng.set_option('chunk', chunk)
    for it in range(2):
        t0 = time.time()
        e, terms, g = eng.elbo_grad(wl['params'], reg=wl['reg'])
        dt = time.time() - t0
    print('cull', cull, 'chunk', chunk, 'elbo', e, 'wall %.1f ms' % (dt * 1e3), eng.last_timing(), flush=True)
eng.precompute(*wl['hyp'], reg=wl['reg'])
for it in range(2):
    t0 = time.time(); e, terms, g = eng.elbo_grad(wl['params'], mode=0, reg=wl['reg']); dt = time.time() - t0
print('frozen elbo', e, 'wall %.1f ms' % (dt * 1e3), eng.last_timing())
