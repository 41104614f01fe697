// Symmetric-output split-K contraction on the FP64 tensor path (DMMA m8n8k4) for sm_100a:
//
//     C[z] (+)= op(A)[M x Kz] * op(B)[Kz x M]      lower triangle only, one K-slice z per CTA
//
// This is the shape of three of the per-chunk contractions of the VCGPCM path, whose results are symmetric
// M x M matrices (M = nh or nx = 200 at the target) summed over an enormous K (= nh * n_chunk or
// n_chunk * nx ~ 1e5):  C1 = sum_n A_n^T (H A_n),  Q = sum_n A_n (iKx A_n^T),  Hbar = sum_n (A_n C1bar) A_n^T
// (src/core/cgpcm.py:255-267,473-475 and the adjoint of SURVEY.md App. D).
//
// Tiling: the output is cut into 40 x 40 warp blocks (5 x 5 DMMA tiles: 10 fragment loads feed 25 DMMAs);
// only the 15 blocks of the lower triangle exist (the 5 diagonal ones compute 15 of their 25 tiles), one warp
// each => one CTA owns the whole output for its
// K-slice, and the grid is the K-split (one CTA per SM, one wave).  Both operand panels (M x 16 k each) are
// staged in shared memory by a 3-stage cp.async pipeline (3 x 64 KB) fed by a 16th, producer warp and shared
// by the 15 consumer warps.  Strides are == 4 (mod 8) doubles: every fragment load is bank-conflict free.
// Each K-slice accumulates into its own private M x M buffer across all launches of a sweep (C[z] += ...),
// and one fixed-order reduction at the end of the sweep adds the slices: deterministic, and the per-launch
// split-K reduction disappears.
//
// Orders below 200 (the windows of inducing inputs that survive when all-zero tiles are skipped) run the same kernel
// with G = 1..4 warp-block rows.  With few blocks (G <= 3) every block is shared by KSUB = 2 or 4 warps that take
// alternate k4 steps of each k-tile, so that the SM still carries >= 12 DMMA-issuing warps; their accumulators are
// added in a fixed order through shared memory before the epilogue.
#pragma once
#include "dgemm_dmma.cuh"

namespace cg {

constexpr int SY_BLK = 40;                 // rows / columns per warp block
constexpr int SY_G = 5;                    // warp-block rows
constexpr int SY_MP = SY_G * SY_BLK;       // 200: largest supported order (15 consumer warps + 1 producer warp)
constexpr int SY_BK = 16;
constexpr int SY_MAX_SPLITS = 148;

// Per-order staging: G warp-block rows of TB x TB DMMA tiles stage GM = 8 TB G panel rows; smaller panels get a deeper pipeline so that the
// bytes in flight per SM stay ~100 KB (the low-order contractions are close to HBM-bound: 6 flop / byte at M = 96).
template <int G, int TB = 5>
struct SymCfg {
  static constexpr int GM = G * TB * 8;
  static constexpr int LDN = GM + 4;                     // row stride of the m-contiguous layout, == 4 (mod 8)
  static constexpr int PANEL = GM * (SY_BK + 4);         // doubles (>= SY_BK * LDN)
  static constexpr int STAGE = 2 * PANEL;
  static constexpr int STAGES = G >= 5 ? 3 : G == 4 ? 4 : G == 3 ? 5 : G == 2 ? 7 : 8;
  static constexpr int SMEM_BYTES = STAGES * STAGE * 8;
};

struct SymArgs {
  const double* A;
  const double* B;
  double* C;             // slice z writes C + z * c_split_stride (leading dimension ldc)
  int M, K;
  long lda, ldb, ldc;
  long c_split_stride;
  int k_per_split;       // multiple of SY_BK
  int accumulate;        // 1: C[z] += result, 0: C[z] = result
};

// DMMA tiles (i, j) of a warp block that a warp computes.  DIAG: only i >= j.  PART splits one diagonal 5 x 5 block
// between two warps (balanced kernel below): part 1 = rows 0..2 and (3,0), (3,1) = 8 tiles, part 2 = (3,2), (3,3) and
// row 4 = 7 tiles.
template <bool DIAG, int PART>
__device__ __forceinline__ constexpr bool sym_tile(int i, int j) {
  return (!DIAG || i >= j) &&
         (PART == 0 || (PART == 1 ? (i <= 2 || (i == 3 && j <= 1)) : (i == 4 || (i == 3 && j >= 2))));
}

// Copies of one operand panel (PANEL_B = false: A, true: B) of one k-tile into a stage, by one warp.
template <bool KC, int G, int TB>
struct SymLoader {
  using Cfg = SymCfg<G, TB>;
  const double* src;
  const double* safe;
  long ld;
  int dst, kofs, kleft, lane, M;
  __device__ __forceinline__ void init(const SymArgs& g, bool panel_b, int kbeg, int kend, int lane_) {
    lane = lane_;
    M = g.M;
    ld = panel_b ? g.ldb : g.lda;
    safe = panel_b ? g.B : g.A;
    const int pofs = panel_b ? Cfg::PANEL : 0;
    if (KC) {
      const int r0 = lane >> 3, kc = (lane & 7) * 2;
      src = safe + (long)r0 * ld + kbeg + kc;
      dst = pofs + r0 * (SY_BK + 4) + kc;
      kofs = kc;
    } else {
      src = safe + (long)kbeg * ld + lane * 2;
      dst = pofs + lane * 2;
      kofs = 0;
    }
    kleft = kend - kbeg;
  }
  __device__ __forceinline__ void issue(double* smem, int stage) {
    double* base = smem + stage * Cfg::STAGE + dst;
    if (KC) {
      const bool kok = kofs < kleft;
      const int r0 = lane >> 3;
#pragma unroll 10
      for (int r = 0; r < Cfg::GM / 4; ++r) {
        const bool v = kok && r * 4 + r0 < M;
        cp_async16(base + r * 4 * (SY_BK + 4), v ? src + (long)r * 4 * ld : safe, v);
      }
      src += SY_BK;
    } else {
#pragma unroll 4
      for (int kk = 0; kk < SY_BK; ++kk) {
        const bool kok = kk < kleft;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int m = (q * 32 + lane) * 2;
          if (m < Cfg::GM) {
            const bool v = kok && m < M;
            cp_async16(base + kk * Cfg::LDN + q * 64, v ? src + (long)kk * ld + q * 64 : safe, v);
          }
        }
      }
      src += (long)SY_BK * ld;
    }
    kleft -= SY_BK;
  }
};

// One consumer warp: block (gi, gj) of the lower triangle.  DIAG: gi == gj, only the DMMA tiles i >= j are computed.
// PART / PROD (balanced kernel): the warp computes one part of a split diagonal block and also feeds one operand panel
// (PROD = 1: A, 2: B) through the cp.async pipeline.
template <bool KC, bool DIAG, int G, int KSUB, int NCONS, int TB, int PART = 0, int PROD = 0>
__device__ __forceinline__ void sym_consume(const SymArgs& g, double* smem, int ktiles, int gi, int gj, int lane,
                                            int sub, int blk, int kbeg = 0, int kend = 0) {
  constexpr int STEPS = (SY_BK / 4) / KSUB;      // k4 steps of a k-tile taken by this warp
  using Cfg = SymCfg<G, TB>;
  constexpr int LDN = Cfg::LDN, PANEL = Cfg::PANEL, STAGE = Cfg::STAGE, STAGES = Cfg::STAGES;
  constexpr int BLK = TB * 8;                    // rows / columns of a warp block
  constexpr int RED = TB * TB * 64;              // doubles of one warp's accumulators
  const int grp = lane >> 2, tig = lane & 3;
  const int a_off = KC ? (gi * BLK + grp) * (SY_BK + 4) + tig : tig * LDN + gi * BLK + grp;
  const int b_off = PANEL + (KC ? (gj * BLK + grp) * (SY_BK + 4) + tig : tig * LDN + gj * BLK + grp);

  double acc[TB][TB][2];
#pragma unroll
  for (int i = 0; i < TB; ++i)
#pragma unroll
    for (int j = 0; j < TB; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // Fragments are single-buffered (25 accumulator pairs leave no room for a second set).  The 5 x 5 DMMA block
  // is walked row-major on even k4 steps and column-major on odd ones, and every fragment of the next step is
  // loaded as soon as its register dies: each load then has at least four DMMAs of this warp (and all the
  // DMMAs of the SMSP's other warps) between issue and first use.
  auto ldf = [&](const double* P, int k4, int x) {
    return KC ? P[x * 8 * (SY_BK + 4) + k4 * 4] : P[k4 * 4 * LDN + x * 8];
  };

  SymLoader<KC, G, TB> ldr;
  if (PROD) {
    ldr.init(g, PROD == 2, kbeg, kend, lane);
#pragma unroll
    for (int s0 = 0; s0 < STAGES - 1; ++s0) {
      if (s0 < ktiles) ldr.issue(smem, s0);
      cp_async_commit();
    }
  }
  int stage = 0;
  for (int kt = 0; kt < ktiles; ++kt) {
    if (PROD) cp_async_wait<STAGES - 2>();
    __syncthreads();
    if (PROD) {
      int st2 = stage + STAGES - 1;
      if (st2 >= STAGES) st2 -= STAGES;
      if (kt + STAGES - 1 < ktiles) ldr.issue(smem, st2);
      cp_async_commit();
    }
    const double* Ap = smem + stage * STAGE + a_off;
    const double* Bp = smem + stage * STAGE + b_off;
    double fa[TB], fb[TB];
    const int k40 = KSUB > 1 ? sub * STEPS : 0;
#pragma unroll
    for (int x = 0; x < TB; ++x) { fa[x] = ldf(Ap, k40, x); fb[x] = ldf(Bp, k40, x); }
#pragma unroll
    for (int s4 = 0; s4 < STEPS; ++s4) {
      const int k4 = k40 + s4;
      const bool more = s4 + 1 < STEPS;
      if ((s4 & 1) == 0) {
#pragma unroll
        for (int i = 0; i < TB; ++i) {
#pragma unroll
          for (int j = 0; j < TB; ++j) {
            if (sym_tile<DIAG, PART>(i, j)) dmma_8x8x4(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
            if (more && i == TB - 1) fb[j] = ldf(Bp, k4 + 1, j);
          }
          if (more) fa[i] = ldf(Ap, k4 + 1, i);
        }
      } else {
#pragma unroll
        for (int j = 0; j < TB; ++j) {
#pragma unroll
          for (int i = 0; i < TB; ++i) {
            if (sym_tile<DIAG, PART>(i, j)) dmma_8x8x4(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
            if (more && j == TB - 1) fa[i] = ldf(Ap, k4 + 1, i);
          }
          if (more) fb[j] = ldf(Bp, k4 + 1, j);
        }
      }
    }
    if (++stage == STAGES) stage = 0;
  }
  if (PROD) cp_async_wait<0>();

  if (KSUB > 1) {
    // add the KSUB partial accumulators of this block in a fixed order (sub 1, 2, .. into sub 0) through the drained
    // pipeline buffers; named barrier 1 = the consumer warps only (the producer warp has left)
    asm volatile("bar.sync 1, %0;\n" ::"n"(NCONS * 32) : "memory");
    double* red = smem + (long)blk * (KSUB - 1) * RED;
    if (sub > 0) {
      double* r = red + (sub - 1) * RED + lane;
#pragma unroll
      for (int i = 0; i < TB; ++i)
#pragma unroll
        for (int j = 0; j < TB; ++j) {
          r[(i * TB + j) * 64] = acc[i][j][0];
          r[(i * TB + j) * 64 + 32] = acc[i][j][1];
        }
    }
    asm volatile("bar.sync 1, %0;\n" ::"n"(NCONS * 32) : "memory");
    if (sub > 0) return;
#pragma unroll
    for (int q = 0; q < KSUB - 1; ++q) {
      const double* r = red + q * RED + lane;
#pragma unroll
      for (int i = 0; i < TB; ++i)
#pragma unroll
        for (int j = 0; j < TB; ++j) {
          acc[i][j][0] += r[(i * TB + j) * 64];
          acc[i][j][1] += r[(i * TB + j) * 64 + 32];
        }
    }
  }

  double* C = g.C + (long)blockIdx.x * g.c_split_stride;
#pragma unroll
  for (int i = 0; i < TB; ++i) {
    const int row = gi * BLK + i * 8 + grp;
    if (row >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TB; ++j) {
      const int col = gj * BLK + j * 8 + tig * 2;
      if (col >= g.M || !sym_tile<DIAG, PART>(i, j)) continue;
      double2* p = reinterpret_cast<double2*>(C + (long)row * g.ldc + col);
      double v0 = acc[i][j][0], v1 = acc[i][j][1];
      if (g.accumulate) {
        const double2 o = *p;
        v0 += o.x;
        v1 += o.y;
      }
      *p = make_double2(v0, v1);
    }
  }
}

// KC = true : A(m,k) at A[m*lda + k], B(n,k) at B[n*ldb + k]   (k contiguous)
// KC = false: A(m,k) at A[k*lda + m], B(n,k) at B[k*ldb + n]   (m / n contiguous)
//
// Warp roles: warps 0..14 are consumers (one 40 x 40 output block each: 25 accumulator pairs + 10 fragments,
// no loader state, so they fit the 128-register budget of a 512-thread CTA); warp 15 is the producer and
// issues every cp.async of both panels.  One __syncthreads per k-tile hands a filled stage to the consumers
// and a drained one back to the producer.
template <bool KC, int G = SY_G, int KSUB = 1, int TB = 5>
__global__ void __launch_bounds__((G * (G + 1) / 2 * KSUB + 1) * 32, 1) dgemm_sym_kernel(const SymArgs g) {
  extern __shared__ __align__(16) double smem[];
  constexpr int NB = G * (G + 1) / 2;          // warp blocks of the lower triangle
  constexpr int NCONS = NB * KSUB;             // consumer warps; warp NCONS is the producer
  using Cfg = SymCfg<G, TB>;
  constexpr int GM = Cfg::GM;                  // panel rows this instantiation stages
  constexpr int LDN = Cfg::LDN, PANEL = Cfg::PANEL, STAGE = Cfg::STAGE, STAGES = Cfg::STAGES;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int kbeg = blockIdx.x * g.k_per_split;
  const int kend = min(g.K, kbeg + g.k_per_split);
  if (kend <= kbeg) return;
  const int ktiles = (kend - kbeg + SY_BK - 1) / SY_BK;

  if (warp == NCONS) {
    // ------------------------------------------------------------------ producer warp
    const double* srcA;
    const double* srcB;
    int dst, kofs;
    if (KC) {            // one instruction = 4 panel rows x 128 B
      const int r0 = lane >> 3, kc = (lane & 7) * 2;
      srcA = g.A + (long)r0 * g.lda + kbeg + kc;
      srcB = g.B + (long)r0 * g.ldb + kbeg + kc;
      dst = r0 * (SY_BK + 4) + kc;
      kofs = kc;
    } else {             // one instruction = 32 chunks (64 columns) of one k row
      srcA = g.A + (long)kbeg * g.lda + lane * 2;
      srcB = g.B + (long)kbeg * g.ldb + lane * 2;
      dst = lane * 2;
      kofs = 0;
    }
    int kleft = kend - kbeg;
    auto issue = [&](int stage) {
      double* base = smem + stage * STAGE + dst;
      if (KC) {
        const bool kok = kofs < kleft;
        const int r0 = lane >> 3;
#pragma unroll 10
        for (int r = 0; r < GM / 4; ++r) {
          const bool v = kok && r * 4 + r0 < g.M;
          cp_async16(base + r * 4 * (SY_BK + 4), v ? srcA + (long)r * 4 * g.lda : g.A, v);
          cp_async16(base + PANEL + r * 4 * (SY_BK + 4), v ? srcB + (long)r * 4 * g.ldb : g.B, v);
        }
        srcA += SY_BK;
        srcB += SY_BK;
      } else {
#pragma unroll 4
        for (int kk = 0; kk < SY_BK; ++kk) {
          const bool kok = kk < kleft;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int m = (q * 32 + lane) * 2;
            if (m < GM) {
              const bool v = kok && m < g.M;
              cp_async16(base + kk * LDN + q * 64, v ? srcA + (long)kk * g.lda + q * 64 : g.A, v);
              cp_async16(base + PANEL + kk * LDN + q * 64, v ? srcB + (long)kk * g.ldb + q * 64 : g.B, v);
            }
          }
        }
        srcA += (long)SY_BK * g.lda;
        srcB += (long)SY_BK * g.ldb;
      }
      kleft -= SY_BK;
    };
#pragma unroll
    for (int s0 = 0; s0 < STAGES - 1; ++s0) {
      if (s0 < ktiles) issue(s0);
      cp_async_commit();
    }
    int stage = 0;
    for (int kt = 0; kt < ktiles; ++kt) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      int st2 = stage + STAGES - 1;
      if (st2 >= STAGES) st2 -= STAGES;
      if (kt + STAGES - 1 < ktiles) issue(st2);
      cp_async_commit();
      if (++stage == STAGES) stage = 0;
    }
    cp_async_wait<0>();
    return;
  }

  // -------------------------------------------------------------------- consumer warps
  // Block of this warp.  The five diagonal blocks only need their lower 15 of 25 DMMA tiles; blocks are dealt to
  // warps so that every SM sub-partition (warp % 4) carries about the same number of DMMAs per k-step:
  // {F,F,D,D} = 80, {F,F,D,D} = 80, {F,F,F,D} = 90, {F,F,F + producer} = 75   (F = 25, D = 15).
  const int blk = (G == SY_G && KSUB == 1 && TB == 5)
                      ? (int)((0xE92DA50C863B741ull >> (4 * warp)) & 15)   // warp -> block {1,4,7,11,3,6,8,12,0,5,10,13,2,9,14}
                      : warp / KSUB;
  const int sub = KSUB > 1 ? warp % KSUB : 0;
  const int gi = blk >= 10 ? 4 : blk >= 6 ? 3 : blk >= 3 ? 2 : blk >= 1 ? 1 : 0;
  const int gj = blk - gi * (gi + 1) / 2;
  if (gi == gj) sym_consume<KC, true, G, KSUB, NCONS, TB>(g, smem, ktiles, gi, gj, lane, sub, blk);
  else sym_consume<KC, false, G, KSUB, NCONS, TB>(g, smem, ktiles, gi, gj, lane, sub, blk);
}

// Balanced variant for the full order (G = 5, TB = 5: 160 < M <= 200).  The kernel above deals 10 full (25 DMMA tiles)
// and 5 diagonal (15) warp blocks to the four SM sub-partitions as {80, 80, 90, 75} tiles per k4 step: the sub-partition
// with 90 sets the pace (ncu: 23 % of the warp time at the k-tile barrier).  Here there is no dedicated producer warp:
// 16 warps compute, one diagonal block is split 8 + 7 between the two warps that also feed the pipeline (one operand
// panel each), and the sub-partitions carry {83, 82, 80, 80}:
//   warp % 4 = 0:  F(1,0) F(2,0) F(2,1)  D(4,4) part 1 + panel A
//   warp % 4 = 1:  F(3,0) F(3,1) F(3,2)  D(4,4) part 2 + panel B
//   warp % 4 = 2:  F(4,0) F(4,1) D(0,0) D(1,1)
//   warp % 4 = 3:  F(4,2) F(4,3) D(2,2) D(3,3)
template <bool KC>
__global__ void __launch_bounds__(512, 1) dgemm_sym16_kernel(const SymArgs g) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int kbeg = blockIdx.x * g.k_per_split;
  const int kend = min(g.K, kbeg + g.k_per_split);
  if (kend <= kbeg) return;
  const int ktiles = (kend - kbeg + SY_BK - 1) / SY_BK;
  // (gi << 4 | gj) per warp
  const unsigned long long tab = warp < 8 ? 0x4341312042403010ull : 0x3311444422003221ull;   // one byte per warp
  const int ent = (int)((tab >> (8 * (warp & 7))) & 255);
  const int gi = ent >> 4, gj = ent & 15;
  if (warp == 12) sym_consume<KC, true, 5, 1, 16, 5, 1, 1>(g, smem, ktiles, gi, gj, lane, 0, 0, kbeg, kend);
  else if (warp == 13) sym_consume<KC, true, 5, 1, 16, 5, 2, 2>(g, smem, ktiles, gi, gj, lane, 0, 0, kbeg, kend);
  else if (gi == gj) sym_consume<KC, true, 5, 1, 16, 5>(g, smem, ktiles, gi, gj, lane, 0, 0);
  else sym_consume<KC, false, 5, 1, 16, 5>(g, smem, ktiles, gi, gj, lane, 0, 0);
}

inline bool dgemm_sym_supported(int M) { return M >= 8 && M <= SY_MP && (M % 8) == 0; }

// Warp blocks are TB x TB DMMA tiles (TB = 5: 40 rows, TB = 4: 32 rows), G <= 5 block rows; the launch for order M
// uses the pair that executes the fewest DMMA cells (ties: the larger block, fewer fragment loads per DMMA).
inline double dgemm_sym_cells_of(int G, int TB) {
  return G * (G - 1) / 2 * (double)(64 * TB * TB) + G * (TB * (TB + 1) / 2) * 64.0;   // full + diagonal blocks
}
inline void dgemm_sym_pick(int M, int* G, int* TB) {
  const int g5 = (M + 39) / 40, g4 = (M + 31) / 32;
  if (g4 <= SY_G && dgemm_sym_cells_of(g4, 4) < dgemm_sym_cells_of(g5, 5)) { *G = g4; *TB = 4; }
  else { *G = g5; *TB = 5; }
}
inline double dgemm_sym_cells(int M) {
  int G, TB;
  dgemm_sym_pick(M, &G, &TB);
  return dgemm_sym_cells_of(G, TB);
}

// Number of K-slices a launch with this K uses (<= SY_MAX_SPLITS).
inline int dgemm_sym_splits(int K, int* k_per_split = nullptr) {
  const int kt = (K + SY_BK - 1) / SY_BK;
  int splits = kt < SY_MAX_SPLITS ? kt : SY_MAX_SPLITS;
  if (splits < 1) splits = 1;
  const int kt_per = (kt + splits - 1) / splits;
  if (k_per_split) *k_per_split = kt_per * SY_BK;
  return (kt + kt_per - 1) / kt_per;
}

inline cudaError_t dgemm_sym(cudaStream_t st, bool kc, int M, int K, const double* A, long lda, const double* B,
                             long ldb, double* C, long ldc, long c_split_stride, int accumulate) {
  if (M <= 0 || K <= 0) return cudaSuccess;
  SymArgs g;
  g.A = A; g.B = B; g.C = C; g.M = M; g.K = K; g.lda = lda; g.ldb = ldb; g.ldc = ldc;
  g.c_split_stride = c_split_stride; g.accumulate = accumulate;
  const int splits = dgemm_sym_splits(K, &g.k_per_split);
  int G, TB;
  dgemm_sym_pick(M, &G, &TB);
#define CG_SYM_LAUNCH(GG, KS, TT)                                                                                      \
  do {                                                                                                             \
    static DeviceOnce attr_once;                                                                                   \
    unsigned long long attr_bit;                                                                                   \
    if (attr_once.need(&attr_bit)) {                                                                               \
      cudaFuncSetAttribute((const void*)dgemm_sym_kernel<true, GG, KS, TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                           SymCfg<GG, TT>::SMEM_BYTES);                                                                \
      cudaFuncSetAttribute((const void*)dgemm_sym_kernel<false, GG, KS, TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                           SymCfg<GG, TT>::SMEM_BYTES);                                                                \
      attr_once.done(attr_bit);                                                                                    \
    }                                                                                                              \
    constexpr int NT = (GG * (GG + 1) / 2 * KS + 1) * 32;                                                          \
    if (kc) dgemm_sym_kernel<true, GG, KS, TT><<<splits, NT, SymCfg<GG, TT>::SMEM_BYTES, st>>>(g);                         \
    else dgemm_sym_kernel<false, GG, KS, TT><<<splits, NT, SymCfg<GG, TT>::SMEM_BYTES, st>>>(g);                           \
  } while (0)
  switch (G * 10 + TB) {
    case 15: CG_SYM_LAUNCH(1, 4, 5); break;
    case 25: CG_SYM_LAUNCH(2, 4, 5); break;
    case 35: CG_SYM_LAUNCH(3, 2, 5); break;
    case 45: CG_SYM_LAUNCH(4, 1, 5); break;
    case 14: CG_SYM_LAUNCH(1, 4, 4); break;
    case 24: CG_SYM_LAUNCH(2, 4, 4); break;
    case 34: CG_SYM_LAUNCH(3, 2, 4); break;
    case 44: CG_SYM_LAUNCH(4, 1, 4); break;
    case 54: CG_SYM_LAUNCH(5, 1, 4); break;
    default: {
      static DeviceOnce attr_once;
      unsigned long long attr_bit;
      if (attr_once.need(&attr_bit)) {
        cudaFuncSetAttribute((const void*)dgemm_sym16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             SymCfg<5, 5>::SMEM_BYTES);
        cudaFuncSetAttribute((const void*)dgemm_sym16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             SymCfg<5, 5>::SMEM_BYTES);
        attr_once.done(attr_bit);
      }
      if (kc) dgemm_sym16_kernel<true><<<splits, 512, SymCfg<5, 5>::SMEM_BYTES, st>>>(g);
      else dgemm_sym16_kernel<false><<<splits, 512, SymCfg<5, 5>::SMEM_BYTES, st>>>(g);
      break;
    }
  }
#undef CG_SYM_LAUNCH
  return cudaGetLastError();
}

}  // namespace cg
