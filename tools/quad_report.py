"""|GPU - truth| and |oracle - truth| at the binary128 fixtures (tests/golden/quad/*.npz, tests/golden/bench_m200.npz):
one JSON line per fixture -> profiles/r02_quad_truth.json.  Run on the GPU box: python tools/quad_report.py"""
import glob
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cgpcm_b200  # noqa: E402
from oracle import model as om  # noqa: E402

om.PW_DISTS_EXACT = True
for f in sorted(glob.glob(os.path.join(ROOT, 'tests', 'golden', 'quad', '*.npz'))):
    with np.load(f) as z:
        d = {k: (z[k][()] if z[k].ndim == 0 else z[k]) for k in z.files}
    eng = cgpcm_b200.Engine(len(d['th']), len(d['tx']), causal=True)
    eng.set_option('cull', 746.0)
    eng.set_data(d['t'], d['y'], d['th'], d['tx'])
    e, terms, g = eng.elbo_grad(d['params'], reg=float(d['reg']))
    eo, to, go = om.elbo_and_grad(d['params'], d['t'], d['y'], d['th'], d['tx'], float(d['reg']), True)
    sc = float(np.abs(d['terms']).max())
    gm = float(np.abs(g).max())
    row = {'fixture': os.path.basename(f)[:-4], 'n': len(d['t']), 'nx': len(d['tx']), 'nh': len(d['th']),
           's2': float(np.exp(d['params'][0])), 'largest_term': sc, 'grad_max': gm, 'elbo_truth': float(d['elbo']),
           'gpu': {'elbo_err_rel_largest_term': abs(e - d['elbo']) / sc,
                   'terms_err_rel_largest_term': float(np.abs(terms - d['terms']).max() / sc),
                   'dderiv_err_rel_grad_max': float(np.abs(d['dirs'] @ g - d['dderiv']).max() / gm),
                   'dderiv_err_rel_largest_term': float(np.abs(d['dirs'] @ g - d['dderiv']).max() / sc)},
           'oracle': {'elbo_err_rel_largest_term': abs(eo - d['elbo']) / sc,
                      'terms_err_rel_largest_term': float(np.abs(to - d['terms']).max() / sc),
                      'dderiv_err_rel_grad_max': float(np.abs(d['dirs'] @ go - d['dderiv']).max() / gm),
                      'dderiv_err_rel_largest_term': float(np.abs(d['dirs'] @ go - d['dderiv']).max() / sc)},
           'gpu_vs_oracle': {'elbo_rel_largest_term': abs(e - eo) / sc, 'grad_rel_grad_max': float(np.abs(g - go).max() / gm)},
           'directions': [str(x) for x in d['dir_names']]}
    print(json.dumps(row), flush=True)
    eng.close()
