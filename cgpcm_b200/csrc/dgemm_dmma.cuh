// FP64 tensor-core GEMM for sm_100a (DMMA.8x8x4 via mma.sync.m8n8k4.f64), the contraction engine
// of the VCGPCM hot path: the reference's batched `tf.matmul` contractions over observations
// (src/core/cgpcm.py:255-267, :473-475) become four plain GEMMs per chunk of observations
// (DESIGN.md §3).  tcgen05 has no f64 kind, so DMMA is the FP64 tensor path on B200; its measured
// peak (tools/fp64_peaks.cu) is 37.2 TFLOP/s = 64 FMA/clk/SM, fed comfortably by cp.async.
//
// C[M x N] = alpha * op(A)[M x K] * op(B)[K x N] + beta * C, all FP64.
//   A_KC : A(m,k) at A[m*lda + k]  (k contiguous)    else A[k*lda + m]  (m contiguous)
//   B_KC : B(k,n) at B[n*ldb + k]  (k contiguous)    else B[k*ldb + n]  (n contiguous)
//   C_TR : C(m,n) at C[n*ldc + m]                    else C[m*ldc + n]
// Requirements: M, N multiples of 8; K, lda, ldb, ldc even; 16-byte aligned pointers.
//
// Tiling: CTA tile = bm x 128 x 16 with bm <= 104 a multiple of 8 chosen by the caller so that M
// splits evenly (M = 200 -> two tiles of 104/96 rows: 96 % useful instead of 78 % with 128-row
// tiles).  8 warps side by side along N; each warp owns all (<= 13) 8-row blocks x two 8-column
// blocks = 26 DMMA per k4 step with 52 FP64 accumulators, so every SM sub-partition carries the
// same tensor work.  Shared-memory strides are == 4 (mod 8) doubles: conflict-free fragment loads.
// 3-stage cp.async pipeline (zero-filled at every edge).  Split-K over blockIdx.z writes partial
// results C + z * c_split_stride (reduced by reduce_partials) so that sums stay deterministic.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cg {

constexpr int G_BM = 104;      // max rows per CTA tile
constexpr int G_MB = 13;       // max 8-row blocks per CTA tile
constexpr int G_BN = 128;
constexpr int G_BK = 16;
constexpr int G_STAGES = 3;
constexpr int G_THREADS = 256;
constexpr int G_AS = (G_BM * (G_BK + 4) > G_BK * (G_BM + 4)) ? G_BM * (G_BK + 4) : G_BK * (G_BM + 4);
constexpr int G_BS = (G_BN * (G_BK + 4) > G_BK * (G_BN + 4)) ? G_BN * (G_BK + 4) : G_BK * (G_BN + 4);
constexpr int G_SMEM_BYTES = G_STAGES * (G_AS + G_BS) * 8;

struct GemmArgs {
  const double* A;
  const double* B;
  double* C;
  int M, N, K;
  long lda, ldb, ldc;
  double alpha, beta;
  int bm;               // rows per CTA tile (multiple of 8, <= 104)
  int k_per_split;      // multiple of G_BK
  long c_split_stride;  // elements between split-K partial results
  int lower_only;       // skip CTA tiles that lie strictly above the diagonal (symmetric results)
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

template <bool A_KC, bool B_KC, bool C_TR>
__global__ void __launch_bounds__(G_THREADS, 1) dgemm_dmma_kernel(const GemmArgs g) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = lane >> 2, tig = lane & 3;
  const int m0 = blockIdx.y * g.bm;
  const int n0 = blockIdx.x * G_BN;
  if (g.lower_only && n0 > m0 + g.bm - 1) return;
  const int rows = min(g.bm, g.M - m0);       // valid rows of this tile (multiple of 8)
  const int cols = min(G_BN, g.N - n0);       // valid columns (multiple of 8)
  const int mb_count = rows >> 3;
  const int nw = warp * 16;                   // this warp's first column inside the tile
  const bool nb_ok0 = nw < cols, nb_ok1 = nw + 8 < cols;
  const int kbeg = blockIdx.z * g.k_per_split;
  const int kend = min(g.K, kbeg + g.k_per_split);
  const int ktiles = (kend - kbeg + G_BK - 1) / G_BK;

  double acc[G_MB][2][2];
#pragma unroll
  for (int i = 0; i < G_MB; ++i) { acc[i][0][0] = acc[i][0][1] = acc[i][1][0] = acc[i][1][1] = 0.0; }

  auto load_tile = [&](int stage, int kt) {
    double* As = smem + stage * (G_AS + G_BS);
    double* Bs = As + G_AS;
    const int k0 = kbeg + kt * G_BK;
    if (A_KC) {
      const int nch = rows * (G_BK / 2);
      for (int c = tid; c < nch; c += G_THREADS) {
        int m = c >> 3, kc = (c & 7) * 2;
        bool ok = (k0 + kc) < kend;
        const double* src = ok ? g.A + (long)(m0 + m) * g.lda + k0 + kc : g.A;
        cp_async16(As + m * (G_BK + 4) + kc, src, ok);
      }
    } else {
      const int half = rows >> 1;
      const int nch = G_BK * half;
      for (int c = tid; c < nch; c += G_THREADS) {
        int kk = c / half, mc = (c - kk * half) * 2;
        bool ok = (k0 + kk) < kend;
        const double* src = ok ? g.A + (long)(k0 + kk) * g.lda + m0 + mc : g.A;
        cp_async16(As + kk * (G_BM + 4) + mc, src, ok);
      }
    }
    if (B_KC) {
      const int nch = cols * (G_BK / 2);
      for (int c = tid; c < nch; c += G_THREADS) {
        int n = c >> 3, kc = (c & 7) * 2;
        bool ok = (k0 + kc) < kend;
        const double* src = ok ? g.B + (long)(n0 + n) * g.ldb + k0 + kc : g.B;
        cp_async16(Bs + n * (G_BK + 4) + kc, src, ok);
      }
    } else {
      const int half = cols >> 1;
      const int nch = G_BK * half;
      for (int c = tid; c < nch; c += G_THREADS) {
        int kk = c / half, nc = (c - kk * half) * 2;
        bool ok = (k0 + kk) < kend;
        const double* src = ok ? g.B + (long)(k0 + kk) * g.ldb + n0 + nc : g.B;
        cp_async16(Bs + kk * (G_BN + 4) + nc, src, ok);
      }
    }
  };

#pragma unroll
  for (int s = 0; s < G_STAGES - 1; ++s) {
    if (s < ktiles) load_tile(s, s);
    cp_async_commit();
  }

  for (int kt = 0; kt < ktiles; ++kt) {
    cp_async_wait<G_STAGES - 2>();
    __syncthreads();
    {
      int nt = kt + G_STAGES - 1;
      if (nt < ktiles) load_tile(nt % G_STAGES, nt);
      cp_async_commit();
    }
    const double* As = smem + (kt % G_STAGES) * (G_AS + G_BS);
    const double* Bs = As + G_AS;
    if (nb_ok0) {
#pragma unroll
      for (int k4 = 0; k4 < G_BK / 4; ++k4) {
        double b0, b1;
        if (B_KC) {
          b0 = Bs[(nw + grp) * (G_BK + 4) + k4 * 4 + tig];
          b1 = Bs[(nw + 8 + grp) * (G_BK + 4) + k4 * 4 + tig];
        } else {
          b0 = Bs[(k4 * 4 + tig) * (G_BN + 4) + nw + grp];
          b1 = Bs[(k4 * 4 + tig) * (G_BN + 4) + nw + 8 + grp];
        }
        double a[G_MB];
#pragma unroll
        for (int mb = 0; mb < G_MB; ++mb) {
          if (mb < mb_count) {
            a[mb] = A_KC ? As[(mb * 8 + grp) * (G_BK + 4) + k4 * 4 + tig]
                         : As[(k4 * 4 + tig) * (G_BM + 4) + mb * 8 + grp];
          } else {
            a[mb] = 0.0;
          }
        }
#pragma unroll
        for (int mb = 0; mb < G_MB; ++mb) {
          if (mb < mb_count) {
            dmma_8x8x4(acc[mb][0][0], acc[mb][0][1], a[mb], b0);
            dmma_8x8x4(acc[mb][1][0], acc[mb][1][1], a[mb], b1);
          }
        }
      }
    }
  }
  cp_async_wait<0>();

  double* C = g.C + (long)blockIdx.z * g.c_split_stride;
#pragma unroll
  for (int mb = 0; mb < G_MB; ++mb) {
    if (mb >= mb_count) continue;
    const int row = m0 + mb * 8 + grp;
#pragma unroll
    for (int nb = 0; nb < 2; ++nb) {
      if (nb == 0 ? !nb_ok0 : !nb_ok1) continue;
      const int col = n0 + nw + nb * 8 + tig * 2;
      double v0 = g.alpha * acc[mb][nb][0], v1 = g.alpha * acc[mb][nb][1];
      if (!C_TR) {
        double2* p = reinterpret_cast<double2*>(C + (long)row * g.ldc + col);
        if (g.beta != 0.0) {
          double2 o = *p;
          v0 += g.beta * o.x;
          v1 += g.beta * o.y;
        }
        *p = make_double2(v0, v1);
      } else {
        double* p0 = C + (long)col * g.ldc + row;
        double* p1 = p0 + g.ldc;
        if (g.beta != 0.0) {
          v0 += g.beta * *p0;
          v1 += g.beta * *p1;
        }
        *p0 = v0;
        *p1 = v1;
      }
    }
  }
}

// out[r][c] (+)= sum_s part[s][r][c]  (deterministic split-K reduction); optional mirroring of the
// lower triangle into the upper one for results computed with lower_only.
__global__ void reduce_partials_kernel(const double* __restrict__ part, long split_stride, int splits,
                                       double* __restrict__ out, int rows, int cols, long ld_part,
                                       long ld_out, double beta, int mirror_lower) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long total = (long)rows * cols;
  if (idx >= total) return;
  int r = (int)(idx / cols), c = (int)(idx % cols);
  int rr = r, cc = c;
  if (mirror_lower && c > r) { rr = c; cc = r; }
  double s = 0.0;
  for (int k = 0; k < splits; ++k) s += part[(long)k * split_stride + (long)rr * ld_part + cc];
  double* o = out + (long)r * ld_out + c;
  *o = (beta != 0.0 ? beta * *o : 0.0) + s;
}

inline int pick_bm(int M) {
  int tiles = (M + G_BM - 1) / G_BM;
  int bm = (M + tiles - 1) / tiles;
  bm = (bm + 7) & ~7;
  return bm > G_BM ? G_BM : bm;
}

// Launch.  Returns cudaError_t of the launch.
inline cudaError_t dgemm(cudaStream_t st, bool a_kc, bool b_kc, bool c_tr, int M, int N, int K, double alpha,
                         const double* A, long lda, const double* B, long ldb, double beta, double* C,
                         long ldc, int splits = 1, long c_split_stride = 0, int lower_only = 0) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  GemmArgs g;
  g.A = A; g.B = B; g.C = C; g.M = M; g.N = N; g.K = K;
  g.lda = lda; g.ldb = ldb; g.ldc = ldc; g.alpha = alpha; g.beta = beta;
  g.bm = pick_bm(M);
  if (splits < 1) splits = 1;
  int kt = (K + G_BK - 1) / G_BK;
  int kt_per = (kt + splits - 1) / splits;
  if (kt_per < 1) kt_per = 1;
  g.k_per_split = kt_per * G_BK;
  g.c_split_stride = c_split_stride;
  g.lower_only = lower_only;
  dim3 grid((N + G_BN - 1) / G_BN, (M + g.bm - 1) / g.bm, splits);
  static bool attr_done = false;
  auto set_attr = [](const void* f) {
    cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM_BYTES);
  };
  if (!attr_done) {
    set_attr((const void*)dgemm_dmma_kernel<true, true, false>);
    set_attr((const void*)dgemm_dmma_kernel<true, false, false>);
    set_attr((const void*)dgemm_dmma_kernel<false, true, false>);
    set_attr((const void*)dgemm_dmma_kernel<false, false, false>);
    set_attr((const void*)dgemm_dmma_kernel<true, true, true>);
    set_attr((const void*)dgemm_dmma_kernel<true, false, true>);
    set_attr((const void*)dgemm_dmma_kernel<false, true, true>);
    set_attr((const void*)dgemm_dmma_kernel<false, false, true>);
    attr_done = true;
  }
#define CG_LAUNCH(AK, BK, CT) dgemm_dmma_kernel<AK, BK, CT><<<grid, G_THREADS, G_SMEM_BYTES, st>>>(g)
  if (!c_tr) {
    if (a_kc && b_kc) CG_LAUNCH(true, true, false);
    else if (a_kc && !b_kc) CG_LAUNCH(true, false, false);
    else if (!a_kc && b_kc) CG_LAUNCH(false, true, false);
    else CG_LAUNCH(false, false, false);
  } else {
    if (a_kc && b_kc) CG_LAUNCH(true, true, true);
    else if (a_kc && !b_kc) CG_LAUNCH(true, false, true);
    else if (!a_kc && b_kc) CG_LAUNCH(false, true, true);
    else CG_LAUNCH(false, false, true);
  }
#undef CG_LAUNCH
  return cudaGetLastError();
}

}  // namespace cg
