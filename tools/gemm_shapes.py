"""Times the DMMA contraction kernels on the shapes the chunk planner produces at the bench shape (N = 1e5, nh = nx =
200, exact-zero windows: chunks of ~4256 observations with a window of 96 inducing inputs) through cgpcm_dgemm /
cgpcm_dgemm_sym, and checks each result against torch.matmul on a slice.

    T1    = H A          200 x 200 times 200 x (nc * kw)            dgemm_sl_kernel<13, 12, 0>
    V^T   = W  A2^T      kw x kw   times kw  x (200 * nc), C^T      dgemm_sl_kernel<12, 0, 1>   (Abar; V, U1 without `tri`)
    V'^T  = L^T A2^T     the same with an upper-triangular operand    dgemm_sl_kernel<12, 0, 1, 1, 1>  (V', U' with `tri`)
    Q    += A V^T        200 x 200 lower triangle over K = nc * kw   dgemm_sym_kernel<1, ...>    (also Hbar)
    C1   += A2^T T1_2    kw x kw   lower triangle over K = 200 * nc  dgemm_sym_kernel<0, ...>
"""
import sys

import torch

sys.path.insert(0, '.')
from cgpcm_b200 import _lib

L = _lib.lib()
nh = 200
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 4256
kw = int(sys.argv[2]) if len(sys.argv) > 2 else 96
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
PEAK = 37.16
cols = nc * kw
dev = 'cuda'
torch.manual_seed(0)
H = torch.randn(nh, nh, dtype=torch.float64, device=dev)
W = torch.randn(kw, kw, dtype=torch.float64, device=dev)
A = torch.randn(nh * cols, dtype=torch.float64, device=dev)          # [i][n][k]
T = torch.empty(nh * cols, dtype=torch.float64, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)


def timeit(run):
    run()
    best = 1e9
    for _ in range(reps):
        flush.zero_()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        run()
        ev1.record()
        torch.cuda.synchronize()
        best = min(best, ev0.elapsed_time(ev1))
    return best


def report(name, ms, flops, err):
    tf = flops / ms / 1e9
    print('%-34s %8.3f ms  %6.2f TFLOP/s  %.3f of %.2f   max err %.1e' % (name, ms, tf, tf / PEAK, PEAK, err), flush=True)


# T1 = H A  (A as 200 x cols, n-contiguous)
def run_t1():
    assert L.cgpcm_dgemm(1, 0, 0, nh, cols, nh, 1.0, H.data_ptr(), nh, A.data_ptr(), cols, 0.0, T.data_ptr(), cols, 1, 0, 0, None) == 0
ms = timeit(run_t1)
ref = H @ A.view(nh, cols)[:, :4096]
report('T1 = H A', ms, 2.0 * nh * nh * cols, float((T.view(nh, cols)[:, :4096] - ref).abs().max()))

# V = A2 W  (A2 = (200 nc) x kw rows), computed as V^T = W^T A2^T with the transposed store
V = torch.empty(nh * cols, dtype=torch.float64, device=dev)
def run_v():
    assert L.cgpcm_dgemm(1, 1, 1, kw, nh * nc, kw, 1.0, W.data_ptr(), kw, A.data_ptr(), kw, 0.0, V.data_ptr(), kw, 1, 0, 0, None) == 0
ms = timeit(run_v)
ref = A.view(nh * nc, kw)[:4096] @ W.t()
report('V = A2 W^T (right-multiply)', ms, 2.0 * kw * kw * nh * nc, float((V.view(nh * nc, kw)[:4096] - ref).abs().max()))

# V' = A2 L with a lower-triangular L (the resident operand is S = L^T, upper triangular; option `tri`)
Lt = torch.triu(torch.randn(kw, kw, dtype=torch.float64, device=dev))
Vt = torch.empty(nh * cols, dtype=torch.float64, device=dev)
def run_vt():
    assert L.cgpcm_dgemm_tri(kw, nh * nc, Lt.data_ptr(), kw, A.data_ptr(), kw, Vt.data_ptr(), kw, None) == 0
ms = timeit(run_vt)
ref = A.view(nh * nc, kw)[:4096] @ Lt.t()
report("V' = A2 L (triangular right-multiply)", ms, 1.0 * kw * (kw + 1) * nh * nc, float((Vt.view(nh * nc, kw)[:4096] - ref).abs().max()))

# Q = A V^T lower triangle (k contiguous), C1 = A2^T T1_2 (m contiguous)
work = torch.empty(148, nh, nh, dtype=torch.float64, device=dev)
C = torch.empty(nh, nh, dtype=torch.float64, device=dev)
def run_q():
    assert L.cgpcm_dgemm_sym(1, nh, cols, A.data_ptr(), cols, V.data_ptr(), cols, C.data_ptr(), nh, work.data_ptr(), None) == 0
ms = timeit(run_q)
ref = torch.tril(A.view(nh, cols) @ V.view(nh, cols).t())
report('Q = A V^T (sym, incl. slice sum)', ms, 1.0 * cols * nh * (nh + 1), float((torch.tril(C) - ref).abs().max() / ref.abs().max()))
C1 = torch.empty(kw, kw, dtype=torch.float64, device=dev)
def run_c1():
    assert L.cgpcm_dgemm_sym(0, kw, nh * nc, A.data_ptr(), kw, T.data_ptr(), kw, C1.data_ptr(), kw, work.data_ptr(), None) == 0
ms = timeit(run_c1)
ref = torch.tril(A.view(nh * nc, kw).t() @ T.view(nh * nc, kw))
report('C1 = A2^T T1_2 (sym, incl. sum)', ms, 1.0 * nh * nc * kw * (kw + 1), float((torch.tril(C1) - ref).abs().max() / ref.abs().max()))
