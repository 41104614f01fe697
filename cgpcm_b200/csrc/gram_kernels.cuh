// Fourth-order Psi tensor ("Gram" of the Ahx blocks) for the precomputed regime, sm_100a FP64.
//
// In the precomputed regime (src/core/cgpcm.py:270-292, where 55-92 % of the reference's L-BFGS iterations, every
// fixed-point round and every sampler proposal run) the Ahx blocks are constants and each evaluation only contracts
// them with small matrices:  C1 = sum_n A_n^T H A_n  (nx x nx)  and  D = sum_n A_n W A_n^T  (nh x nh).  Both are
// linear images of
//                 G[(k,i),(l,j)] = sum_n Ahx[n,i,k] Ahx[n,j,l]            ((nx nh) x (nx nh), symmetric)
//   C1[k,l] = sum_ij H[i,j] G[(k,i),(l,j)]           D[i,j] = sum_kl W[k,l] G[(k,i),(l,j)] ,
// so once G is resident (8 (nh nx)^2 bytes: 12.8 GB at nh = nx = 200 -- HBM capacity is what makes this possible) an
// evaluation costs two streaming passes over G (HBM-bound, ~2 ms each at the bench shape) instead of 6 N M^3 flops.
// G is accumulated chunk by chunk with the DMMA GEMM from blocks generated in the layout Ac[n][(k,i)] (row index
// r = k * nhp + i, so that a window of inducing inputs is a contiguous row range) and mirrored once.
//
// OPT-IN (option "gram", default 0).  H = m2 ~ iKh and W ~ iKx have entries ~1/reg that cancel against the smooth
// columns of A_n.  The sweeps cancel per observation (H A_n first) and then add N well-scaled terms; here the N terms are
// added first and G's rounding error (eps |G|) meets the 1/reg entries afterwards: ~sqrt(N) more rounding noise.
// Measured at the bench shape (N = 1e5, reg = 1e-6): ELBO 2.4e-8 relative from the sweeps' value (sweeps among
// themselves: 5e-11); at a trained toy point 6e-7.  Use it for throughput-bound frozen-regime work (sampler proposals,
// fixed-point rounds) where that noise is below the Monte-Carlo / iteration error.
#pragma once
#include "psi_kernels.cuh"

namespace cg {

// Ac[n][k * nhp + i] = Ahx[n0 + n, i, k_lo + k]   for n < nc (zero for n >= n_valid), k < kwp, i < nhp (zero padding)
__global__ void ahx_gen_nki_kernel(const double* __restrict__ t, int n_valid, int nc, const double* __restrict__ th,
                                   int nh, int nhp, const double* __restrict__ tx, int nx, int k_lo, int kwp,
                                   double* __restrict__ Ac, const PsiConst c) {
  const long row = (long)kwp * nhp;
  const long total = (long)nc * row;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int n = (int)(idx / row);
    const int e = (int)(idx - (long)n * row);
    const int k = e / nhp, i = e - k * nhp;
    const int kg = k_lo + k;
    double v = 0.0;
    if (n < n_valid && i < nh && kg < nx) v = ahx_value(th[i], __ldg(t + n) - tx[kg], c);
    Ac[idx] = v;
  }
}

// G[r][s] = G[s][r] for s > r (the accumulation only fills tiles on / below the diagonal).  32 x 32 tiles.
__global__ void __launch_bounds__(256) gram_mirror_kernel(double* __restrict__ G, long ldg, int n) {
  __shared__ double tile[32][33];
  const int bx = blockIdx.x, by = blockIdx.y;      // tile (by, bx) with bx > by is written from tile (bx, by)
  if (bx < by) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const long row = (long)bx * 32 + r, col = (long)by * 32 + tx;      // source: lower tile (bx, by)
    tile[r][tx] = (row < n && col < n) ? G[row * ldg + col] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const long row = (long)by * 32 + r, col = (long)bx * 32 + tx;      // destination: upper tile (by, bx)
    if (row < n && col < n && col > row) G[row * ldg + col] = tile[tx][r];
  }
}

// C1[k][l] = C1[l][k] = sum_{i,j} H[i][j] G[(k nhp + i) ldg + l nhp + j]   for l <= k.  One CTA per (k, l).
__global__ void __launch_bounds__(256) gram_c1_kernel(const double* __restrict__ G, long ldg, int nx, int nh, int nhp,
                                                     const double* __restrict__ H, long ldh, double* __restrict__ C1,
                                                     long ldc) {
  __shared__ double red[8];
  int tidx = blockIdx.x;
  int k = (int)((sqrt(8.0 * tidx + 1.0) - 1.0) * 0.5);
  while ((k + 1) * (k + 2) / 2 <= tidx) ++k;
  while (k * (k + 1) / 2 > tidx) --k;
  const int l = tidx - k * (k + 1) / 2;
  const double* base = G + (long)k * nhp * ldg + (long)l * nhp;
  double s = 0.0;
  for (int e = threadIdx.x; e < nh * nhp; e += blockDim.x) {
    const int i = e / nhp, j = e - i * nhp;
    if (j < nh) s += H[(long)i * ldh + j] * base[(long)i * ldg + j];
  }
  for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += red[w];
    C1[(long)k * ldc + l] = a;
    C1[(long)l * ldc + k] = a;
  }
}

// part[s][i][j] = sum_{k in slice s} sum_l W[k][l] G[(k nhp + i) ldg + l nhp + j].  grid (nh, slices), threads over j.
__global__ void __launch_bounds__(256) gram_q_kernel(const double* __restrict__ G, long ldg, int nx, int nh, int nhp,
                                                    const double* __restrict__ W, long ldw, double* __restrict__ part,
                                                    long ldp) {
  extern __shared__ double wrow[];                 // W[k][0..nx)
  const int i = blockIdx.x;
  const int slices = gridDim.y;
  const int kper = (nx + slices - 1) / slices;
  const int k0 = blockIdx.y * kper, k1 = min(nx, k0 + kper);
  const int j = threadIdx.x;
  double s0 = 0.0, s1 = 0.0;                       // j and j + blockDim.x (nhp <= 512)
  for (int k = k0; k < k1; ++k) {
    __syncthreads();
    for (int l = threadIdx.x; l < nx; l += blockDim.x) wrow[l] = W[(long)k * ldw + l];
    __syncthreads();
    const double* row = G + ((long)k * nhp + i) * ldg;
#pragma unroll 4
    for (int l = 0; l < nx; ++l) {
      const double w = wrow[l];
      if (j < nh) s0 += w * row[(long)l * nhp + j];
      if (j + (int)blockDim.x < nh) s1 += w * row[(long)l * nhp + j + blockDim.x];
    }
  }
  double* out = part + ((long)blockIdx.y * nhp + i) * ldp;
  if (j < nhp) out[j] = j < nh ? s0 : 0.0;
  if (j + (int)blockDim.x < nhp) out[j + blockDim.x] = j + (int)blockDim.x < nh ? s1 : 0.0;
}

}  // namespace cg
