"""Step time against the chunk budget (option `chunk`; 0 = the planner) on the whole bench series and on one shard of an
8-way split."""
import os
import sys
sys.path.insert(0, '.')
import cgpcm_b200
from cgpcm_b200.cgpcm import shard_bounds, window_costs, window_radius
from tests.workload import sweep_workload
wl = sweep_workload(100000, 200)
cost = window_costs(wl['t'], wl['tx'], 200, window_radius(*wl['hyp'], 746.0))
for name, (lo, hi) in [('whole', (0, 100000)), ('shard 3/8', shard_bounds(100000, 3, 8, cost))]:
    eng = cgpcm_b200.Engine(200, 200)
    eng.set_option('cull', 746.0)
    eng.set_data(wl['t'][lo:hi], wl['y'][lo:hi], wl['th'], wl['tx'])
    for ch in [0, 768, 1024, 1280, 1536, 1792, 2048, 2560, 3072]:
        eng.set_option('chunk', ch)
        best = 1e9
        for _ in range(3):
            eng.elbo_grad(wl['params'], reg=wl['reg'])
            tm = eng.last_timing()
            best = min(best, tm['total_ms'])
        print(name, 'chunk', ch, 'step %.3f ms' % best, 'gemm launches', tm['gemm_launches'], 'flops %.4g' % tm['gemm_flops'], flush=True)
    eng.close()
