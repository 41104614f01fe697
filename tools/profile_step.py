"""One full-regime ELBO + gradient evaluation at the bench shape for ncu (launch list / --set full).
Warm-up evaluations run before cudaProfilerStart so that `ncu --profile-from-start off` captures exactly one."""
import argparse
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import cgpcm_b200
from tests.workload import sweep_workload

ap = argparse.ArgumentParser()
ap.add_argument('--n', type=int, default=100000)
ap.add_argument('--m', type=int, default=200)
ap.add_argument('--cull', type=float, default=0.0)
ap.add_argument('--chunk', type=int, default=0)
ap.add_argument('--warmup', type=int, default=2)
ap.add_argument('--mode', type=int, default=1)
ap.add_argument('--shard', default='0/1', help='r/w: evaluate rank r\'s slice of the observations of a w-rank run (no communication)')
a, _unknown = ap.parse_known_args()
wl = sweep_workload(a.n, a.m)
eng = cgpcm_b200.Engine(a.m, a.m)
eng.set_option('cull', a.cull)
eng.set_option('chunk', a.chunk)
from cgpcm_b200.cgpcm import shard_bounds, window_costs, window_radius
_r, _w = [int(x) for x in a.shard.split('/')]
_cost = window_costs(wl['t'], wl['tx'], a.m, window_radius(*wl['hyp'], a.cull)) if _w > 1 else None
_lo, _hi = shard_bounds(a.n, _r, _w, _cost)
print('shard', _lo, _hi)
eng.set_data(wl['t'][_lo:_hi], wl['y'][_lo:_hi], wl['th'], wl['tx'])
ap_gram = [x for x in sys.argv if x.startswith('--gram=')]
if ap_gram:
    eng.set_option('gram', int(ap_gram[0].split('=')[1]))
if a.mode == 0:
    import time
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.precompute(*wl['hyp'], reg=wl['reg'])
    torch.cuda.synchronize()
    print('precompute wall %.1f ms' % ((time.perf_counter() - t0) * 1e3))
for _ in range(a.warmup):
    eng.elbo_grad(wl['params'], mode=a.mode, reg=wl['reg'])
torch.cuda.synchronize()
torch.cuda.profiler.start()
e, terms, g = eng.elbo_grad(wl['params'], mode=a.mode, reg=wl['reg'])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('elbo', e, eng.last_timing())
