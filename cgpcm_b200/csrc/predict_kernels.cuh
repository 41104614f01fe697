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

namespace cg {

// Off-diagonal ("center") Psi statistics of predict_k (src/core/cgpcm.py:164-166,190-192): t2 = 0, upper limit
// min(t, 0) for the causal model.  Closed forms (oracle.model.psi_center_closed):
//   a_c(t)        = pref_a exp(-(alpha/2 + gamma) t^2) erfc(sqrt(2 alpha) |t| / 2)
//   Ahh_c(t)[i,j] = pref_h exp(c + b^2 / 8B) erfc(sqrt(2B) (b / 4B - min(t, 0))),   B = alpha + gamma,
//                   b = 2B t - 2 gamma (th_i + th_j),  c = -alpha (t^2 + th_i^2) - gamma (t - th_i)^2 - B th_j^2
// (acausal: no erfc, twice the prefactor).  AC: [n][nh * nhl] (row i at stride nhl, zero padded), ac: [n].
__global__ void ahh_center_kernel(const double* __restrict__ t, int n, const double* __restrict__ th, int nh, int nhl,
                                  double alpha, double gamma, int causal, double* __restrict__ AC,
                                  double* __restrict__ ac) {
  const double PI = 3.14159265358979323846;
  const double B = alpha + gamma;
  const double pref_a = (causal ? 0.5 : 1.0) * sqrt(PI / (2.0 * alpha));
  const double pref_h = (causal ? 0.5 : 1.0) * sqrt(PI / (2.0 * B));
  const long per = (long)nh * nhl;
  const long total = (long)n * per;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int p = (int)(idx / per);
    const int e = (int)(idx - (long)p * per);
    const int i = e / nhl, j = e - i * nhl;
    const double tt = t[p];
    if (e == 0) {
      double v = pref_a * exp(-(0.5 * alpha + gamma) * tt * tt);
      if (causal) v *= erfc(sqrt(2.0 * alpha) * fabs(tt) * 0.5);
      ac[p] = v;
    }
    double v = 0.0;
    if (j < nh) {
      const double ti = th[i], tj = th[j];
      const double b = 2.0 * B * tt - 2.0 * gamma * (ti + tj);
      const double c = -alpha * (tt * tt + ti * ti) - gamma * (tt - ti) * (tt - ti) - B * tj * tj;
      v = pref_h * exp(c + b * b / (8.0 * B));
      if (causal) v *= erfc(sqrt(2.0 * B) * (b / (4.0 * B) - fmin(tt, 0.0)));
    }
    AC[idx] = v;
  }
}

// HH[b][i * nhl + j] = h_b[i] h_b[j] - iKh[i][j]   (rows b >= nb and the padding are zero)
__global__ void hh_build_kernel(const double* __restrict__ hs, long ldh, int nb, int nbp, const double* __restrict__ iKh,
                                long ld, int nh, int nhl, double* __restrict__ HH) {
  const long per = (long)nh * nhl;
  const long total = (long)nbp * per;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int b = (int)(idx / per);
    const int e = (int)(idx - (long)b * per);
    const int i = e / nhl, j = e - i * nhl;
    double v = 0.0;
    if (b < nb && j < nh) v = hs[(long)b * ldh + i] * hs[(long)b * ldh + j] - iKh[(long)i * ld + j];
    HH[idx] = v;
  }
}

// out[p][b] = s2f (ac[p] + G[p][b])
__global__ void kernel_finish_kernel(const double* __restrict__ G, long ldg, const double* __restrict__ ac, int n, int nb,
                                     double s2f, double* __restrict__ out) {
  const long total = (long)n * nb;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int p = (int)(idx / nb), b = (int)(idx - (long)p * nb);
    out[idx] = s2f * (ac[p] + G[(long)p * ldg + b]);
  }
}

}  // namespace cg

namespace cg {

// Filter-side kernel matrices of predict_h / predict_psd (src/core/cgpcm.py:685-688,738-741; DEQ kernel
// src/core/kernel.py:43-46):  Kuh[i][p] = k_h(th_i, t_p)  (nhp x ldn, zero padded),
// Ktt[p][q] = k_h(t_p, t_q) + reg [p == q]  (ldn x ldn; identity on the padding block so that its Cholesky exists).
__global__ void filter_kernels_kernel(const double* __restrict__ th, int nh, int nhp, const double* __restrict__ t, int n,
                                      int ldn, double alpha, double gamma, double reg, double* __restrict__ Kuh,
                                      double* __restrict__ Ktt) {
  const long n1 = (long)nhp * ldn, n2 = (long)ldn * ldn;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < n1 + n2; idx += (long)gridDim.x * blockDim.x) {
    if (idx < n1) {
      const int i = (int)(idx / ldn), p = (int)(idx - (long)i * ldn);
      double v = 0.0;
      if (i < nh && p < n) {
        const double a = th[i], b = t[p];
        v = exp(-alpha * (a * a + b * b) - gamma * (a - b) * (a - b));
      }
      Kuh[idx] = v;
    } else {
      const long e = idx - n1;
      const int p = (int)(e / ldn), q = (int)(e - (long)p * ldn);
      double v = (p == q) ? 1.0 : 0.0;
      if (p < n && q < n) {
        const double a = t[p], b = t[q];
        v = exp(-alpha * (a * a + b * b) - gamma * (a - b) * (a - b)) + (p == q ? reg : 0.0);
      } else if (p < n || q < n) {
        v = 0.0;
      }
      Ktt[e] = v;
    }
  }
}

}  // namespace cg
