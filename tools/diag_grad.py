import sys
import numpy as np
sys.path.insert(0, '.')
import cgpcm_b200
from tests.workload import sweep_workload
n = int(sys.argv[1]); m = int(sys.argv[2])
wl = sweep_workload(n, m)
eng = cgpcm_b200.Engine(m, m)
eng.set_data(wl['t'], wl['y'], wl['th'], wl['tx'])
p = wl['params']
e0, _, g = eng.elbo_grad(p, reg=wl['reg'])
print('elbo', e0, 'gmax', np.abs(g).max(), 'g head', g[:5])
rng = np.random.default_rng(3)
subsets = {'hyp': slice(2, 5), 's2': slice(0, 2), 'mu': slice(5, 5 + m), 'var': slice(5 + m, None), 'all': slice(0, None)}
for name, sl in subsets.items():
    d = np.zeros_like(p)
    d[sl] = rng.standard_normal(d[sl].shape[0])
    d /= np.linalg.norm(d)
    an = float(g @ d)
    for h in [1e-3, 1e-4, 1e-5, 1e-6]:
        f1 = eng.elbo_grad(p + h * d, reg=wl['reg'], want_grad=False)[0]
        f2 = eng.elbo_grad(p - h * d, reg=wl['reg'], want_grad=False)[0]
        fd = (f1 - f2) / (2 * h)
        print('%-4s h=%g fd=%.10e an=%.10e rel=%.2e' % (name, h, fd, an, abs(fd - an) / abs(an)), flush=True)
if len(sys.argv) > 3:
    from oracle import model as om
    om.PW_DISTS_EXACT = True
    e1, t1, g1 = om.elbo_and_grad(p, wl['t'], wl['y'], wl['th'], wl['tx'], wl['reg'])
    print('oracle elbo', e1, 'rel', abs(e1 - e0) / abs(e1), 'grad rel', np.abs(g - g1).max() / np.abs(g1).max(), 'head', g1[:5])
