"""Times the DMMA GEMM on the contraction shapes of one dense chunk (nh = nx = 200, nc = 512) through
cgpcm_dgemm; under ncu (--set full) this is the capture of the dominant kernel."""
import sys
import torch
sys.path.insert(0, '.')
from cgpcm_b200 import _lib

L = _lib.lib()
nh = nx = 200
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cols = nc * nx
dev = 'cuda'
H = torch.randn(nh, nh, dtype=torch.float64, device=dev)
A = torch.randn(nh, cols, dtype=torch.float64, device=dev)
T = torch.empty(nh, cols, dtype=torch.float64, device=dev)
part = torch.empty(128, nh, nh, dtype=torch.float64, device=dev)
shapes = {
    'T1  [H (200x200)] x [A (200 x nc*200)]': (1, 0, 0, nh, cols, nh, H, nh, A, cols, T, cols, 1, 0, 0),
    'V^T [iKx (200x200)] x [A2^T (200 x 200*nc)], C transposed': (1, 1, 1, nx, nh * nc, nx, H, nh, A, nx, T, nx, 1, 0, 0),
    'Q   [A (200 x nc*200)] x [V^T], split-K 49, lower': (1, 1, 0, nh, nh, cols, A, cols, T, cols, part, nh, 49, nh * nh, 1),
    'C1  [A2^T (200 x 200*nc)] x [T1_2], split-K 49, lower': (0, 0, 0, nx, nx, nh * nc, A, nx, T, nx, part, nx, 49, nx * nx, 1),
}
for name, (akc, bkc, ctr, M, N, K, a, lda, b, ldb, c, ldc, splits, stride, lower) in shapes.items():
    def run():
        rc = L.cgpcm_dgemm(akc, bkc, ctr, M, N, K, 1.0, a.data_ptr(), lda, b.data_ptr(), ldb, 0.0, c.data_ptr(), ldc,
                           splits, stride, lower, None)
        assert rc == 0
    run()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        ev0.record()
        run()
        ev1.record()
        torch.cuda.synchronize()
        best = min(best, ev0.elapsed_time(ev1))
    flops = 2.0 * M * N * K * (0.75 if lower else 1.0)
    print('%-52s %8.3f ms  %6.2f TFLOP/s' % (name, best, flops / best / 1e9), flush=True)

# the symmetric-output kernel (dgemm_sym.cuh) on the same two contractions
C = torch.empty(nh, nh, dtype=torch.float64, device=dev)
work = torch.empty(148, nh, nh, dtype=torch.float64, device=dev)
for name, kc, K, lda in [('Q   sym kernel (k contiguous), 200 x 200, K = nc*200', 1, cols, cols),
                         ('C1  sym kernel (m contiguous), 200 x 200, K = 200*nc', 0, nh * nc, nx)]:
    def run():
        rc = L.cgpcm_dgemm_sym(kc, nh, K, A.data_ptr(), lda, T.data_ptr(), lda, C.data_ptr(), nh, work.data_ptr(), None)
        assert rc == 0
    run()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        ev0.record()
        run()
        ev1.record()
        torch.cuda.synchronize()
        best = min(best, ev0.elapsed_time(ev1))
    flops = 2.0 * K * nh * (nh + 1) / 2
    print('%-52s %8.3f ms  %6.2f TFLOP/s (lower triangle; incl. the reduction of the 148 K-slices)' % (name, best, flops / best / 1e9), flush=True)
