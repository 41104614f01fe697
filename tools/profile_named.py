"""One evaluation at a named experiment shape (toy / ou / hrir / crude; sizes of src/tasks/*.py, synthetic data) for ncu
launch lists:  ncu --metrics gpu__time_duration.sum --profile-from-start off python tools/profile_named.py hrir [mode]"""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import cgpcm_b200
from oracle import model as om

SHAPES = {  # name: (n, nx, nh, tau_w, tau_f, t-grid, reg)
    'toy': (400, 150, 41, .1, .05, lambda n: np.linspace(0, 1, n), 1e-6),
    'ou': (600, 300, 75, .15, .025, lambda n: np.linspace(0, 1, n), 1e-5),
    'hrir': (400, 300, 151, 1.5e-3, 5e-5, lambda n: np.arange(n) / 44100., 1e-8),
    'crude': (400, 300, 101, 1., .1, lambda n: 2010 + 4 * np.sort(np.random.default_rng(0).choice(1013, n, replace=False)) / 1013., 1e-4),
}
name = sys.argv[1] if len(sys.argv) > 1 else 'hrir'
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
n, nx, nh, tau_w, tau_f, grid, reg = SHAPES[name]
rng = np.random.default_rng(0)
t = np.ascontiguousarray(grid(n))
y = rng.standard_normal(n)
y = (y - y.mean()) / y.std()
rec = om.recipe(t, nx=nx, nh=nh, tau_w=tau_w, tau_f=tau_f, causal=True)
hyp = (rec['alpha'], rec['gamma'], rec['omega'])
mu_u, var_u = om.init_q(rec['th'], rec['alpha'], rec['gamma'], reg, rng)
p = om.pack(0.1, rec['s2_f'], hyp[0], hyp[1], hyp[2], mu_u, var_u)
eng = cgpcm_b200.Engine(len(rec['th']), nx)
eng.set_data(t, y, rec['th'], rec['tx'])
if mode == 0:
    eng.precompute(*hyp, reg=reg)
for _ in range(3):
    eng.elbo_grad(p, mode=mode, reg=reg)
torch.cuda.synchronize()
torch.cuda.profiler.start()
e, terms, g = eng.elbo_grad(p, mode=mode, reg=reg)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(name, 'elbo', e, eng.last_timing())
