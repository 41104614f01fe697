import sys, os
sys.path.insert(0, '.')
import numpy as np
import cgpcm_b200
from tests.workload import sweep_workload
wl = sweep_workload(100000, 200)
eng = cgpcm_b200.Engine(200, 200)
eng.set_option('cull', 746.0)
lo = 16740 + 11511 + 10834 + 11268 + 11205
for size in [10319, 10600, 10882, 11000, 11133, 11400]:
    eng.set_data(wl['t'][lo:lo + size], wl['y'][lo:lo + size], wl['th'], wl['tx'])
    os.environ['CGPCM_DEBUG_PLAN'] = '1'
    eng.elbo_grad(wl['params'], reg=wl['reg'])
    os.environ.pop('CGPCM_DEBUG_PLAN')
    best = 1e9
    for _ in range(3):
        eng.elbo_grad(wl['params'], reg=wl['reg'])
        tm = eng.last_timing()
        best = min(best, tm['own_sweeps_ms'])
    print(size, 'own sweeps %.3f ms' % best, 'gemm launches', tm['gemm_launches'], 'flops %.4g' % tm['gemm_flops'], flush=True)
