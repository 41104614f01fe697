"""Generates tests/golden/*.npz: seeded inputs and the CPU oracle's outputs (Psi sums, ELBO, 7 terms, gradient;
full and frozen regime) at the named shapes, with exact squared distances in the prior kernels (the product's
default arithmetic).  The vectors of the REFERENCE'S OWN CODE on the same inputs are tests/golden/ref/*.npz
(tools/make_ref_golden.py); tests/test_ref_parity.py holds the oracle to them.
Run:  python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import model as om  # noqa: E402
from tests.cases import make_case  # noqa: E402

om.PW_DISTS_EXACT = True     # exact squared distances (see oracle/model.py)
out_dir = os.path.join(ROOT, 'tests', 'golden')
os.makedirs(out_dir, exist_ok=True)
for name in ['toy_test', 'ou', 'hrir', 'crude', 'sweep', 'sweep_hi', 'toy_acausal_model']:
    c = make_case(name)
    e, terms, g = om.elbo_and_grad(c['params'], c['t'], c['y'], c['th'], c['tx'], c['reg'], c['causal'])
    fr = om.precompute(c['params'], c['t'], c['y'], c['th'], c['tx'], c['reg'], c['causal'])
    p2 = c['params'].copy()
    p2[0] += .25
    p2[5:] *= 1.03
    # the reference freezes `mats` only: the prior kernels stay functions of the hyper-parameters (oracle/model.py)
    ef, tf, gf = om.elbo_and_grad(p2, c['t'], c['y'], c['th'], c['tx'], c['reg'], c['causal'], frozen=fr,
                                  frozen_kernels='symbolic')
    m = fr[0]
    np.savez_compressed(os.path.join(out_dir, name + '.npz'), t=c['t'], y=c['y'], th=c['th'], tx=c['tx'],
                        hyp=np.array(c['hyp']), reg=c['reg'], causal=c['causal'], params=c['params'],
                        sum_Axx=m['sum_Axx'].numpy(), Ahh=m['Ahh'].numpy(), a=float(m['a']),
                        sum_Ahx_y=m['sum_Ahx_y'].numpy(), elbo=e, terms=terms, grad=g,
                        params_frozen=p2, elbo_frozen=ef, terms_frozen=tf, grad_frozen=gf)
    print(name, e, ef)

# AKM sampler (third part): covariance and draw of AKM.f() from the reference's pair integrands
akm_dir = os.path.join(out_dir, 'akm')
os.makedirs(akm_dir, exist_ok=True)
for name in ['toy_small', 'toy_acausal_model']:
    c = make_case(name)
    rng = np.random.default_rng(4)
    t = np.sort(rng.uniform(0., 1., 19))
    h = c['params'][5:5 + c['nh']] + .3 * rng.standard_normal(c['nh'])
    e = rng.standard_normal(19)
    f, K = om.akm_f(c['params'], c['th'], c['reg'], t, h, e, causal=c['causal'])
    np.savez_compressed(os.path.join(akm_dir, name + '.npz'), params=c['params'], th=c['th'], tx=c['tx'], reg=c['reg'],
                        causal=c['causal'], t=t, h=h, e=e, f=f, K=K)
    print('akm', name, np.abs(f).max())
