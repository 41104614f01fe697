"""The reference's toy experiment end to end on the GPU (BASELINE.json configs[0]): src/tasks/toy.py (causal sample,
causal model) through src/core/experiment.py `train` (:192-277) and `predict` (:302-340) — AKM draw, precompute,
three L-BFGS phases, fixed-point iterations, elliptical slice sampling, MF and SMF predictions — on the reference-facing
API of cgpcm_b200, with the wall time of every phase.  One JSON line.

    python tools/toy_experiment.py            # the task's 'test' sizes (n = 150, nx = 60)
    python tools/toy_experiment.py --full     # n = 400, nx = 150, the full iteration counts
"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, '.')
from cgpcm_b200 import VCGPCM, Session, config, learn
from cgpcm_b200.data import load_akm

full = '--full' in sys.argv
quick = '--quick' in sys.argv
cfg = dict(seed=1005, causal=True, causal_model=True, resample=0, nh=41, noise=0.0, noise_init=1e-4, data_scale=.5,
           iters_fpi_post=500)
if full:
    cfg.update(n=400, nx=150, iters_pre=400, iters=2000, iters_post=200, samps=500, tau_w=.1, tau_f=.1)
else:
    cfg.update(n=150, nx=60, iters_pre=200, iters=500, iters_post=50, samps=200, tau_w=.25, tau_f=.25)
if quick:
    cfg.update(iters_pre=20, iters=30, iters_post=10, samps=10, iters_fpi_post=20)
config.reg = 1e-6                                   # src/tasks/toy.py:7
np.random.seed(cfg['seed'])
sess = Session()
times, evals = {}, {}


class phase(object):
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        self.t0 = time.perf_counter()

    def __exit__(self, *a):
        times[self.name] = round(time.perf_counter() - self.t0, 4)


with phase('cuda_init'):                            # context creation, library load: not part of the experiment
    import cgpcm_b200
    cgpcm_b200.bvn_cdf(np.zeros(8), np.zeros(8), np.full(8, .5))
with phase('load_akm'):
    f, k, h = load_akm(sess=sess, n=cfg['n'], nh=cfg['nh'], tau_w=cfg['tau_w'] * cfg['data_scale'],
                       tau_f=cfg['tau_f'] * cfg['data_scale'], causal=cfg['causal'], resample=cfg['resample'])
    e = f.make_noisy(cfg['noise'])
with phase('construct'):
    mod = VCGPCM.from_recipe(sess=sess, e=e, nx=cfg['nx'], nh=cfg['nh'], tau_w=cfg['tau_w'],
                             tau_f=cfg['tau_f'] * cfg['data_scale'], causal=cfg['causal_model'],
                             noise_init=cfg['noise_init'])
with phase('precompute'):
    mod.precompute()
elbo, terms = mod.elbo()
e_start = sess.run(elbo)
V = mod.vars


def lbfgs(name, names, iters):
    with phase(name):
        res = learn.minimise_lbfgs(sess, -elbo, vars=[V[n] for n in names], iters=iters, name=name, quiet=True)
    evals[name] = int(res.nfev) if res is not None else 0


lbfgs('pretraining', ['mu_u', 'var_u'], cfg['iters_pre'])
lbfgs('training', ['mu_u', 'var_u', 's2_f', 's2'], cfg['iters'])
e_train = sess.run(elbo)
mod.undo_precompute()
elbo, terms = mod.elbo()
lbfgs('posttraining', ['mu_u', 'var_u', 's2_f', 's2', 'gamma', 'omega', 'alpha'], cfg['iters_post'])
with phase('precompute_2'):
    mod.precompute()
elbo = mod.elbo()[0]
e_post = sess.run(elbo)
with phase('fpi'):
    mod.fpi(cfg['iters_fpi_post'])
e_fpi = sess.run(elbo)
with phase('sample_smf'):
    samples = mod.sample(iters=cfg['samps'])
with phase('predict_mf'):
    f_pred = mod.predict_f(f.x)
    k_pred = mod.predict_k(k.x)
    psd_pred = mod.predict_psd(h.x)
    h_pred = mod.predict_h(h.x, phase_transform=None)
    h_mp = mod.predict_h(h.x, phase_transform='minimum_phase')
    h_zp = mod.predict_h(h.x, phase_transform='zero_phase')
with phase('predict_smf'):
    f_smf = mod.predict_f(f.x, samples_h=samples[-min(200, len(samples)):])
    k_smf = mod.predict_k(k.x, samples_h=samples)
    elbo_smf = mod.elbo_smf(samples)
smse = lambda pred, ref: float(np.mean((pred - ref) ** 2) / np.var(ref))
out = {'experiment': 'toy (causal sample, causal model)%s' % (' full sizes' if full else ' test sizes'),
       'n': cfg['n'], 'nx': cfg['nx'], 'nh': cfg['nh'],
       'iters': {k2: cfg[k2] for k2 in ('iters_pre', 'iters', 'iters_post', 'iters_fpi_post', 'samps')},
       'lbfgs_evaluations': evals, 'seconds': times,
       'seconds_total': round(sum(v for k2, v in times.items() if k2 != 'cuda_init'), 3),
       'elbo': {'start': e_start, 'after_training': e_train, 'after_posttraining': e_post, 'after_fpi': e_fpi,
                'smf': float(elbo_smf[0]), 'smf_stderr': float(elbo_smf[1])},
       'smse': {'f_mf': smse(f_pred.mean.y, f.y), 'f_smf': smse(f_smf.mean.y, f.y),
                'k_mf': smse(k_pred.mean.y, k.y), 'k_smf': smse(k_smf.mean.y, k.y)},
       'noise_learned': float(np.exp(V['s2'].value))}
assert e_start < e_train and e_post >= e_train - 1e-6 * abs(e_train) and e_fpi >= e_post - 1e-6 * abs(e_post)
print(json.dumps(out))
