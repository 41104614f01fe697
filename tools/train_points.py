"""Trains the reference's toy and HRIR experiment shapes on the GPU (the three-phase schedule of experiment.train,
src/core/experiment.py:192-262, on cgpcm_b200.experiment.train) and saves the TRAINED points -- inputs and variables --
to gpurun_out/points/<shape>.npz.  tools/make_quad_golden.py then evaluates the ELBO, its terms and directional
derivatives at those points in binary128 on the CPU (tests/golden/quad/*_trained.npz): trained points are where FP64
evaluations disagree beyond 1e-9 (s2 ~ 1e-3, cond(Kh) ~ 1 / reg), and the quad values say which one is right.
    gpurun -- python tools/train_points.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cgpcm_b200 import Data, Session, config, experiment  # noqa: E402
from cgpcm_b200.data import load_akm  # noqa: E402

out_dir = os.path.join(ROOT, 'gpurun_out', 'points')
os.makedirs(out_dir, exist_ok=True)
sess = Session()


def save(name, mod, rep, reg, t0):
    e, terms, g = mod.engine.elbo_grad(mod._pack(), reg=reg)
    np.savez_compressed(os.path.join(out_dir, name + '.npz'), t=mod.e.x, y=mod.e.y, th=mod.th, tx=mod.tx, reg=reg,
                        params=mod._pack(), gpu_elbo=e, gpu_terms=terms, gpu_grad=g)
    print(name, 'elbo %.6f' % e, 's2 %.3e' % np.exp(mod._pack()[0]), {k: round(v, 3) for k, v in rep['elbo'].items()},
          rep['evals'], '%.0f s' % (time.time() - t0), flush=True)


# toy (src/tasks/toy.py: causal sample, causal model): AKM draw, n = 400, nx = 150, nh = 41
np.random.seed(1005)
config.reg = 1e-6
t0 = time.time()
f, k, h = load_akm(sess=sess, n=400, nh=41, tau_w=.05, tau_f=.05, causal=True, resample=0)
mod, rep = experiment.train(sess, f, nx=150, nh=41, tau_w=.1, tau_f=.05, causal=True, reg=1e-6, iters_pre=400,
                            iters=2000, iters_post=200, iters_fpi_post=50)
save('toy', mod, rep, 1e-6, t0)

# hrir shape (src/tasks/hrir.py): white noise through a decaying random filter at 44.1 kHz, n = 400, nx = 300, nh = 151
rng = np.random.default_rng(0)
t0 = time.time()
n = 400
t = np.arange(n) / 44100.
filt = rng.standard_normal(177) * np.exp(-np.arange(177) / 30.)
y = np.convolve(rng.standard_normal(n + 176), filt, mode='valid')
y = (y - y.mean()) / y.std()
config.reg = 1e-8
mod, rep = experiment.train(sess, Data(t, y), nx=300, nh=151, tau_w=1.5e-3, tau_f=5e-5, causal=True, reg=1e-8,
                            iters_pre=200, iters=350, iters_post=300, iters_fpi_post=50)
save('hrir', mod, rep, 1e-8, t0)
