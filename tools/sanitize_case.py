"""A small workload that touches every kernel family (for `compute-sanitizer --tool memcheck python tools/sanitize_case.py`):
the three DMMA GEMM kernels through the C-ABI exports, a full / frozen evaluation with and without the sweep stores at
nh = nx = 200 (so that dgemm_sl / dgemm_sym are on the path), fpi, the SMF bound, predict_f / predict_k / predict_h."""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import cgpcm_b200
from cgpcm_b200 import _lib
from tests.workload import sweep_workload

L = _lib.lib()
dev = lambda x: torch.tensor(np.ascontiguousarray(x), dtype=torch.float64, device='cuda')
rng = np.random.default_rng(0)
M, K, N = 200, 184, 64 * 148 + 40
S, B = dev(rng.standard_normal((M, K))), dev(rng.standard_normal((K, N)))
C = torch.empty(M, N, dtype=torch.float64, device='cuda')
assert L.cgpcm_dgemm(1, 0, 0, M, N, K, 1.0, S.data_ptr(), K, B.data_ptr(), N, 0.0, C.data_ptr(), N, 1, 0, 0, None) == 0
Bt = dev(rng.standard_normal((N, K)))
Ct = torch.empty(N, M, dtype=torch.float64, device='cuda')
assert L.cgpcm_dgemm(1, 1, 1, M, N, K, 1.0, S.data_ptr(), K, Bt.data_ptr(), K, 0.0, Ct.data_ptr(), M, 1, 0, 0, None) == 0
X = dev(rng.standard_normal((184, 1234)))
Cs = torch.empty(184, 184, dtype=torch.float64, device='cuda')
for kc, A_ in ((1, X), (0, dev(X.cpu().numpy().T))):
    assert L.cgpcm_dgemm_sym(kc, 184, 1234, A_.data_ptr(), 1234 if kc else 184, A_.data_ptr(), 1234 if kc else 184,
                             Cs.data_ptr(), 184, None, None) == 0
wl = sweep_workload(1500, 200)
for store, cull in ((1, 80.0), (0, 0.0)):
    eng = cgpcm_b200.Engine(200, 200)
    eng.set_option('store', store)
    eng.set_option('cull', cull)
    eng.set_option('chunk', 256)
    eng.set_data(wl['t'], wl['y'], wl['th'], wl['tx'])
    e, terms, g = eng.elbo_grad(wl['params'], reg=wl['reg'])
    eng.precompute(*wl['hyp'], reg=wl['reg'])
    e2, _, g2 = eng.elbo_grad(wl['params'], mode=0, reg=wl['reg'])
    mu_u, var_u, mu_z, var_z = eng.fpi(wl['params'], 1, reg=wl['reg'])
    smp = wl['params'][5:205] + .01 * rng.standard_normal(200)
    es, ts, ll = eng.elbo_smf(wl['params'], smp, mode=0, reg=wl['reg'])
    m, v = eng.predict_f(wl['params'], np.linspace(0, 1.5, 77), np.stack([smp, smp * 1.01]), reg=wl['reg'])
    k = eng.kernel_samples(wl['params'], np.linspace(-.3, .3, 33), np.stack([smp, smp * 1.01]), reg=wl['reg'])
    f = eng.filter_samples(wl['params'], np.linspace(-.05, .25, 41), np.stack([smp, smp * 1.01]),
                           rng.standard_normal((41, 2)), reg=wl['reg'])
    print('store', store, 'cull', cull, e, e2, ll, float(m[3]), float(k[5, 0]), float(f[7, 1]))
    eng.close()
torch.cuda.synchronize()
print('sanitize case done')
