"""Generates tests/golden/ref/*.npz: outputs of the REFERENCE'S OWN CODE (oracle/_ref = py3-patched copies of
/root/reference/src on the TensorFlow stand-in, see oracle/build_ref.py) on the seeded inputs of tests/cases.py:
ELBO, 7 terms, gradient (full and precomputed regime), model matrices, the optimal q(z), fpi + convert, predict_f.
Needs /root/reference (this container); the fixtures travel to the GPU box.   Run:  python tools/make_ref_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from tests.cases import make_case  # noqa: E402

out_dir = os.path.join(ROOT, 'tests', 'golden', 'ref')
os.makedirs(out_dir, exist_ok=True)
names = sys.argv[1:] or ['toy_small', 'toy_small_cid', 'toy_test', 'toy_acausal_model', 'ou', 'hrir', 'crude', 'sweep', 'sweep_hi']
for name in names:
    c = make_case(name)
    rec = c['recipe']
    p2 = c['params'].copy()
    p2[0] += .25
    p2[5:] *= 1.03
    rng = np.random.default_rng(11)
    n = len(c['t'])
    t_star = np.sort(rng.uniform(c['t'].min(), c['t'].max(), 7))
    small = name in ('toy_small', 'toy_small_cid', 'toy_acausal_model')
    # an explicit q(z) for the z = False variants: N(mu_z, reg(Lz Lz^T)) with Lz a scaled, perturbed identity
    qz_mu = .3 * rng.standard_normal(c['nx'])
    Lz = np.tril(.05 * rng.standard_normal((c['nx'], c['nx']))) + .5 * np.eye(c['nx'])
    qz_var = Lz[np.tril_indices(c['nx'])]
    r = ref.call(qz_mu=qz_mu, qz_var=qz_var, qz_fpi_num=2, t=c['t'], y=c['y'], nx=rec['nx'], nh=rec['nh'], tau_w=rec['tau_w'], tau_f=rec['tau_f'],
                 causal=c['causal'], causal_id=c['causal_id'], reg=c['reg'], params=c['params'], params_frozen=p2,
                 want_per_n=small, fpi_num=3, t_star=t_star, num_samples=4, seed=5)
    assert np.array_equal(r['th'], c['th']) and np.array_equal(r['tx'], c['tx']), 'recipe mismatch'
    keep = {k: v for k, v in r.items() if k not in ('seconds_first', 'term_names', 'mat_m2_u')}
    np.savez_compressed(os.path.join(out_dir, name + '.npz'), t=c['t'], y=c['y'], reg=c['reg'], causal=c['causal'],
                        causal_id=c['causal_id'], params=c['params'], params_frozen=p2, t_star=t_star, qz_mu=qz_mu, qz_var=qz_var, **keep)
    print(name, 'elbo', r['elbo'], 'frozen', r['elbo_frozen'], 'fpi', r.get('fpi_elbo', r.get('fpi_error')))
