// Test-side kernels of VCGPCM.predict_f (src/core/cgpcm.py:781-846): per test input t*_n the reference builds
// Ahx*_n (nh x nx), Axx*_n (nx x nx) and, per filter sample h,
//     mu_n = sqrt(s2_f) h^T Ahx*_n x_mean
//     m2_n = s2_f (a + <Ahh, mh> + <Axx*_n, mx> + tr((mh Ahx*_n) o (Ahx*_n mx))),   mh = h h^T - iKh,  mx = x_m2 - iKx
// Here the sample-independent part of the last term is folded into the per-point matrix
//     Bxx*_n = Axx*_n - Ahx*_n^T iKh Ahx*_n        (the reference's mats['Bxx'], cgpcm.py:249-251)
// so that per sample and point only  v = Ahx*_n^T h,  v^T mx v  and  <Bxx*_n, mx>  remain:
//     m2_n = s2_f (a + h^T Ahh h - tr(Ahh iKh) + <Bxx*_n, mx> + v^T mx v).
#pragma once
#include <cuda_runtime.h>

namespace cg {

// X[n][k][l] -= sum_i A[i][n][k] * T[i][n][l]   (A, T in the chunk layout [i][n][kw], X dense [n][nx][nx]).
// grid (tiles_l, tiles_k, n), block 16 x 16, each thread a 2 x 2 patch of a 32 x 32 tile.
__global__ void __launch_bounds__(256) bxx_star_kernel(const double* __restrict__ A, const double* __restrict__ T, int nh,
                                                      int nc, int kw, int nx, double* __restrict__ X) {
  __shared__ double sa[8][32], st[8][32];
  const int n = blockIdx.z;
  const int k0 = blockIdx.y * 32, l0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[2][2] = {{0, 0}, {0, 0}};
  for (int i0 = 0; i0 < nh; i0 += 8) {
    {
      const int r = threadIdx.x >> 5, c = threadIdx.x & 31;     // 8 rows (i) x 32 columns
      const int i = i0 + r;
      const long base = ((long)i * nc + n) * kw;
      sa[r][c] = (i < nh && k0 + c < nx) ? A[base + k0 + c] : 0.0;
      st[r][c] = (i < nh && l0 + c < nx) ? T[base + l0 + c] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const double a0 = sa[r][ty * 2], a1 = sa[r][ty * 2 + 1];
      const double t0 = st[r][tx * 2], t1 = st[r][tx * 2 + 1];
      acc[0][0] += a0 * t0; acc[0][1] += a0 * t1;
      acc[1][0] += a1 * t0; acc[1][1] += a1 * t1;
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const int k = k0 + ty * 2 + u, l = l0 + tx * 2 + v;
      if (k < nx && l < nx) X[((long)n * nx + k) * nx + l] -= acc[u][v];
    }
}

// One CTA per test point: accumulates this sample's mean and variance contribution (weight w = 1 / #samples).
//   A: chunk layout [i][n][kw];  hvec[nh];  xm[nx];  mx: nx x nx with leading dimension ld;  B: Bxx* dense [n][nx][nx]
//   q1 = a + h^T Ahh h - tr(Ahh iKh)
__global__ void __launch_bounds__(256) predict_point_kernel(const double* __restrict__ A, int nh, int nc, int kw, int nx,
                                                           const double* __restrict__ hvec, const double* __restrict__ xm,
                                                           const double* __restrict__ mx, long ld,
                                                           const double* __restrict__ B, const double* __restrict__ q1,
                                                           double sqrt_s2f, double s2f, double w,
                                                           double* __restrict__ acc_mu, double* __restrict__ acc_var) {
  extern __shared__ double pv[];          // v[nx]
  __shared__ double red[3][8];
  const int n = blockIdx.x;
  for (int k = threadIdx.x; k < nx; k += blockDim.x) {
    double s = 0.0;
    for (int i = 0; i < nh; ++i) s += hvec[i] * A[((long)i * nc + n) * kw + k];
    pv[k] = s;
  }
  __syncthreads();
  double mu = 0.0, qv = 0.0, qb = 0.0;
  for (int k = threadIdx.x; k < nx; k += blockDim.x) mu += pv[k] * xm[k];
  const double* Bn = B + (long)n * nx * nx;
  for (long e = threadIdx.x; e < (long)nx * nx; e += blockDim.x) {
    const int k = (int)(e / nx), l = (int)(e - (long)k * nx);
    const double m = mx[(long)k * ld + l];
    qv += pv[k] * m * pv[l];
    qb += Bn[e] * m;
  }
  for (int off = 16; off > 0; off >>= 1) {
    mu += __shfl_down_sync(0xffffffffu, mu, off);
    qv += __shfl_down_sync(0xffffffffu, qv, off);
    qb += __shfl_down_sync(0xffffffffu, qb, off);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = mu; red[1][warp] = qv; red[2][warp] = qb; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a0 = 0, a1 = 0, a2 = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a0 += red[0][i]; a1 += red[1][i]; a2 += red[2][i]; }
    const double m1 = sqrt_s2f * a0;
    const double m2 = s2f * (q1[0] + a2 + a1);
    acc_mu[n] += w * m1;
    acc_var[n] += w * (m2 - m1 * m1);
  }
}

}  // namespace cg
