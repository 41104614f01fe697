"""Generates tests/golden/bench_m200.npz: the bench workload's own shape (scaling sweep, nh = nx = M = 200, causal,
full regime) at N = 1e4 observations -- inputs, the CPU oracle's ELBO / 7 terms / full gradient, and (with --quad) the
binary128 ELBO, terms and one random directional derivative (oracle/quad).  The N = 1e5 bench configuration itself is
out of the oracle's reach in minutes; cost and arithmetic per observation are identical.
Run:  python tools/make_bench_golden.py [--quad]      (~5 min oracle; ~1 h with --quad on 8 cores)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import model as om  # noqa: E402
from tests.workload import sweep_workload  # noqa: E402

N, M = 10000, 200
path = os.path.join(ROOT, 'tests', 'golden', 'bench_m200.npz')
w = sweep_workload(N, M)
keep = {}
if os.path.exists(path):
    with np.load(path) as z:
        keep = {k: z[k] for k in z.files}
if '--quad-only' not in sys.argv:
    om.PW_DISTS_EXACT = True
    t0 = time.time()
    e, terms, g = om.elbo_and_grad_chunked(w['params'], w['t'], w['y'], w['th'], w['tx'], w['reg'], True, chunk=400)
    print('oracle: elbo %.15e  %.0f s' % (e, time.time() - t0), flush=True)
    keep.update(t=w['t'], y=w['y'], th=w['th'], tx=w['tx'], reg=w['reg'], params=w['params'], elbo=e, terms=terms, grad=g)
    np.savez_compressed(path, **keep)
if '--quad' in sys.argv or '--quad-only' in sys.argv:
    from oracle import quad
    assert np.array_equal(keep['params'], w['params']), 'workload changed: regenerate without --quad-only'
    t0 = time.time()
    e_hi, e_lo, t_hi, t_lo = quad.elbo(w['params'], w['t'], w['y'], w['th'], w['tx'], w['reg'], True)
    print('quad: elbo %.15e  %.0f s' % (e_hi, time.time() - t0), flush=True)
    keep.update(quad_elbo=e_hi, quad_terms=t_hi)
    np.savez_compressed(path, **keep)
    v = np.random.default_rng(3).standard_normal(w['params'].shape[0])
    v /= np.linalg.norm(v)
    dd = quad.dderiv(w['params'], v, w['t'], w['y'], w['th'], w['tx'], w['reg'], True, h=1e-9)
    keep.update(quad_dir=v, quad_dderiv=dd[0], quad_dderiv_h=dd[1], quad_dderiv_2h=dd[2])
    np.savez_compressed(path, **keep)
    print('quad: dderiv %.15e (h vs 2h: %.1e)  %.0f s' % (dd[0], abs(dd[1] - dd[2]), time.time() - t0), flush=True)
