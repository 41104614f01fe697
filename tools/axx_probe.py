"""Axx kernel time against the number of observation slices (grid.y), on the whole bench series and on one shard of an
8-way split."""
import sys
sys.path.insert(0, '.')
import cgpcm_b200
from cgpcm_b200.cgpcm import shard_bounds, window_costs, window_radius
from tests.workload import sweep_workload
wl = sweep_workload(100000, 200)
cost = window_costs(wl['t'], wl['tx'], 200, window_radius(*wl['hyp'], 746.0))
for name, (lo, hi) in [('whole', (0, 100000)), ('shard 3/8', shard_bounds(100000, 3, 8, cost)), ('shard 0/8', shard_bounds(100000, 0, 8, cost))]:
    eng = cgpcm_b200.Engine(200, 200)
    eng.set_option('cull', 746.0)
    eng.set_data(wl['t'][lo:hi], wl['y'][lo:hi], wl['th'], wl['tx'])
    for s in [16, 24, 32, 48, 64]:
        eng.set_option('axx_slices', s)
        best = 1e9
        for _ in range(3):
            eng.elbo_grad(wl['params'], reg=wl['reg'])
            best = min(best, eng.last_timing()['axx_ms'])
        print(name, 'slices', s, 'axx %.3f ms' % best, flush=True)
    eng.close()
