"""Scaling-sweep table (BASELINE.json configs[4]): one full-regime ELBO + gradient evaluation per (N, M) shape,
with exact-zero windows (cull = 746, the bench's setting), cull = 80 (library default) and every tile (cull = 0; skipped
where it is 10 x the bench shape), device time from cgpcm_last_timing.  Prints one JSON line per shape."""
import json
import sys

import numpy as np

sys.path.insert(0, '.')
import cgpcm_b200
from tests.workload import sweep_workload

shapes = [(10000, 50), (10000, 200), (10000, 400), (100000, 50), (100000, 200), (100000, 400), (1000000, 50),
          (1000000, 200)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split('x')) for a in sys.argv[1:]]
for n, m in shapes:
    wl = sweep_workload(n, m)
    eng = cgpcm_b200.Engine(m, m)
    eng.set_data(wl['t'], wl['y'], wl['th'], wl['tx'])
    row = {'n': n, 'm': m, 'rho': wl['hyp'][1] / sum(wl['hyp'])}
    ref = None
    for cull in (746.0, 80.0, 0.0):
        if cull == 0.0 and n * m * m > 4.1e9:
            continue                      # dense at N = 1e6, M = 200 is 10 x the bench shape: skipped in the table
        eng.set_option('cull', cull)
        for _ in range(2):
            e, terms, g = eng.elbo_grad(wl['params'], reg=wl['reg'])
        tm = eng.last_timing()
        key = {746.0: 'exact', 80.0: 'culled', 0.0: 'dense'}[cull]
        row[key + '_ms'] = round(tm['total_ms'], 3)
        row[key + '_evals_per_s'] = round(1e3 / tm['total_ms'], 3)
        row[key + '_gemm_gflop'] = round(tm['gemm_flops'] / 1e9, 1)
        if ref is None:
            ref = (e, g)
        else:
            row['elbo_rel_diff_%s_vs_exact' % key] = abs(e - ref[0]) / abs(ref[0])
            row['grad_rel_diff_%s_vs_exact' % key] = float(np.abs(g - ref[1]).max() / np.abs(ref[1]).max())
    row['elbo'] = ref[0]
    print(json.dumps(row), flush=True)
    eng.close()
