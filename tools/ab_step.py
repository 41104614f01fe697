"""A / B of library options at the bench shape: one full-regime evaluation per variant, results compared with the first.

    python tools/ab_step.py --cull 746 --variant sep=0 --variant sep=1

Prints, per variant, the step time (CUDA events inside the library, best of --reps), the ELBO and the relative
differences of ELBO / gradient against the first variant."""
import argparse
import json
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import cgpcm_b200
from tests.workload import sweep_workload

ap = argparse.ArgumentParser()
ap.add_argument('--n', type=int, default=100000)
ap.add_argument('--m', type=int, default=200)
ap.add_argument('--cull', type=float, default=746.0)
ap.add_argument('--reps', type=int, default=3)
ap.add_argument('--mode', type=int, default=1)
ap.add_argument('--variant', action='append', default=[], help='comma-separated key=value library options')
a = ap.parse_args()
wl = sweep_workload(a.n, a.m)
base = None
for var in a.variant or ['']:
    eng = cgpcm_b200.Engine(a.m, a.m)
    eng.set_option('cull', a.cull)
    for kv in [x for x in var.split(',') if x]:
        k, v = kv.split('=')
        eng.set_option(k, float(v))
    eng.set_data(wl['t'], wl['y'], wl['th'], wl['tx'])
    if a.mode == 0:
        eng.precompute(*wl['hyp'], reg=wl['reg'])
    best = None
    for _ in range(a.reps + 1):
        e, terms, g = eng.elbo_grad(wl['params'], mode=a.mode, reg=wl['reg'])
        tm = eng.last_timing()
        if best is None or tm['total_ms'] < best['total_ms']:
            best = tm
    if base is None:
        base = (e, terms, g)
    out = {'variant': var, 'elbo': e, 'total_ms': best['total_ms'], 'forward_ms': best['forward_ms'],
           'backward_ms': best['backward_ms'], 'axx_ms': best['axx_ms'], 'launches': best['launches'],
           'elbo_rel': abs(e - base[0]) / abs(base[0]),
           'terms_rel': float(np.abs(terms - base[1]).max() / np.abs(base[1]).max()),
           'grad_rel': float(np.abs(g - base[2]).max() / np.abs(base[2]).max())}
    print(json.dumps(out), flush=True)
    del eng
    torch.cuda.empty_cache()
