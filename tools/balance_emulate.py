"""Emulates the shard balancing of a W-rank run on ONE GPU: every round evaluates the W shards one after the other
(no communicator: a shard's sweeps do not depend on the others), prints their own-sweep times and refines the
per-observation costs with cgpcm_b200.rebalance_costs, as bench.py does during its warm-up."""
import sys

import numpy as np

sys.path.insert(0, '.')
import cgpcm_b200
from cgpcm_b200.cgpcm import rebalance_costs, shard_bounds, window_costs, window_radius
from tests.workload import sweep_workload

W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 4
n, m, cull = 100000, 200, 746.0
wl = sweep_workload(n, m)
cost = window_costs(wl['t'], wl['tx'], m, window_radius(*wl['hyp'], cull))
eng = cgpcm_b200.Engine(m, m)
eng.set_option('cull', cull)
for rd in range(rounds):
    bounds = [shard_bounds(n, r, W, cost) for r in range(W)]
    times, totals = [], []
    for lo, hi in bounds:
        eng.set_data(wl['t'][lo:hi], wl['y'][lo:hi], wl['th'], wl['tx'])
        best = None
        for _ in range(3):
            eng.elbo_grad(wl['params'], reg=wl['reg'])
            tm = eng.last_timing()
            if best is None or tm['own_sweeps_ms'] < best[0]:
                best = (tm['own_sweeps_ms'], tm['total_ms'], tm['gemm_launches'])
        times.append(best[0])
        totals.append(best[1])
    times = np.array(times)
    print('round %d  sizes %s' % (rd, [hi - lo for lo, hi in bounds]))
    print('   own sweeps ms %s  max/mean %.4f' % (np.round(times, 3).tolist(), times.max() / times.mean()))
    print('   total ms      %s  max %.3f' % (np.round(totals, 3).tolist(), max(totals)), flush=True)
    cost = rebalance_costs(cost, bounds, times)
