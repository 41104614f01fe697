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

# the bench shape's own fixture (M = 200, N = 1e4; tools/make_bench_golden.py --quad): oracle values are stored in it
f = os.path.join(ROOT, 'tests', 'golden', 'bench_m200.npz')
if os.path.exists(f):
    with np.load(f) as z:
        d = {k: (z[k][()] if z[k].ndim == 0 else z[k]) for k in z.files}
    if 'quad_elbo' in d:
        sc = float(np.abs(d['quad_terms']).max())
        gm = float(np.abs(d['grad']).max())
        row = {'fixture': 'bench_m200', 'n': len(d['t']), 'nx': len(d['tx']), 'nh': len(d['th']), 'largest_term': sc,
               'grad_max': gm, 'elbo_truth': float(d['quad_elbo']),
               'oracle': {'elbo_err_rel_largest_term': abs(float(d['elbo']) - d['quad_elbo']) / sc,
                          'terms_err': (d['terms'] - d['quad_terms']).tolist()}}
        if 'quad_dderiv' in d:
            row['oracle']['dderiv_err_rel_grad_max'] = abs(float(d['quad_dir'] @ d['grad']) - d['quad_dderiv']) / gm
        for name, opts in [('gpu', {'cull': 746.0}), ('gpu_all_tiles', {'cull': 0.0}), ('gpu_sep0', {'cull': 746.0, 'sep': 0})]:
            eng = cgpcm_b200.Engine(len(d['th']), len(d['tx']), causal=True)
            for k, v in opts.items():
                eng.set_option(k, v)
            eng.set_data(d['t'], d['y'], d['th'], d['tx'])
            e, terms, g = eng.elbo_grad(d['params'], reg=float(d['reg']))
            row[name] = {'elbo_err_rel_largest_term': abs(e - d['quad_elbo']) / sc,
                         'terms_err': (terms - d['quad_terms']).tolist(),
                         'grad_vs_oracle_rel_grad_max': float(np.abs(g - d['grad']).max() / gm)}
            if 'quad_dderiv' in d:
                row[name]['dderiv_err_rel_grad_max'] = abs(float(d['quad_dir'] @ g) - d['quad_dderiv']) / gm
            eng.close()
        print(json.dumps(row), flush=True)
