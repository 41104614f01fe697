// "Small-left" FP64 tensor-core GEMM for sm_100a:  C[M x N] = alpha * S[M x K] * B[K x N]  with a small left
// operand (M, K <= 200: H = m2 - iKh, iKx, C1bar, Wx of the VCGPCM path) and an enormous N (= observations of
// the chunk x inducing inputs ~ 1e5).  Four of the seven per-chunk contractions have this shape
// (src/core/cgpcm.py:255-267,473-475: T1 = H A, V = A iKx; adjoint: U1 = A C1bar, Abar = T1 Wx).
//
// Persistent, warp-specialised design (one CTA per SM, 10 warps):
//   * the CTA's half of S (<= 104 rows x K) is loaded into shared memory once and stays resident; only B streams;
//   * two consumer groups of 4 warps each own a 104 x 64 output tile at a time (13 x 2 DMMA m8n8k4 blocks per
//     warp) and work on alternate tiles of the CTA's list, so one group's epilogue overlaps the other's main loop;
//   * two producer warps (one per group) stream the B tiles through a 3-stage ring -- 1-D bulk copies
//     (cp.async.bulk, SASS UBLKCP) for n-contiguous B, 16-byte cp.async for k-contiguous B -- that complete on
//     mbarriers; consumers hand stages back through mbarriers.
//     No __syncthreads in the main loop.
// Layouts (S always k-contiguous: S[m*lds + k]):
//   B_KC = false: B(k,n) at B[k*ldb + n], C(m,n) at C[m*ldc + n]               (left multiply,  T1 = H A)
//   B_KC = true : B(k,n) at B[n*ldb + k], C(m,n) at C[n*ldc + m]  (transposed)  (right multiply, out^T = W X^T)
// Requirements: M, N, K multiples of 8; 16 <= K <= 200; 16 <= M <= 208; lds, ldb, ldc even; 16-byte aligned pointers.
// M <= 104 (the windows of inducing inputs that survive when all-zero tiles are skipped) runs as one "half" owned by
// every CTA; otherwise rows [0, 104) and [104, M) are two halves dealt to CTAs in proportion to their DMMA work.
// The row blocks per half are compile-time (5, 9, 12 or 13 blocks of 8 rows; the smallest that covers the half).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <type_traits>

#include "dgemm_dmma.cuh"

namespace cg {

constexpr int SL_BN = 64;          // output columns per consumer group tile (4 warps x 16)
constexpr int SL_BK = 16;
constexpr int SL_STAGES = 3;
constexpr int SL_MAX_STAGES = 8;
constexpr int SL_BAR_BYTES = 384;   // [2 groups][full | empty][SL_MAX_STAGES] + stagger
constexpr int SL_MB = 13;          // 8-row blocks of the resident half (104 rows)
constexpr int SL_KMAX = 200;
constexpr int SL_SK = SL_KMAX + 4;   // row stride of the resident operand: == 4 (mod 8), conflict-free fragment loads
constexpr int SL_NT = 320;         // 8 consumer warps + 2 producer warps
constexpr int SL_TILE_N = SL_BK * (SL_BN + 4);   // doubles per stage, B n-contiguous
constexpr int SL_TILE_K = SL_BN * (SL_BK + 4);   // doubles per stage, B k-contiguous

struct SlArgs {
  const double* S;
  const double* B;
  double* C;
  int M, N, K;
  long lds, ldb, ldc;
  double alpha;
  int ctas0;      // CTAs [0, ctas0) own rows [0, 8 MB0), the rest rows [8 MB0, M)
  int stages;     // ring depth per consumer group (SL_STAGES .. SL_MAX_STAGES: whatever the resident operand leaves)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
// 1-D bulk copy global -> shared, completion (bytes) signalled on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// One consumer warp's work on one output tile: MB x 2 blocks, K in k-tiles streamed through the ring.
// TRI (one-half variants, K <= 112): the resident operand is UPPER TRIANGULAR, S[m][k] = 0 for k < m -- the transposed
// Cholesky factor of a window block (cgpcm.cu: window_factors).  Row block mb is zero in every k4 step that ends before
// column 8 mb, so its DMMAs and fragment loads are not issued: 54 % of the DMMAs of the full product at K = 96.  The
// k-tile index is a template argument (the set of active blocks must be known at compile time: a run-time condition
// around mma.sync costs a WARPSYNC and a branch per step).
template <int MB, bool B_KC, int SKT, bool TRI = false>
__device__ __forceinline__ void sl_consume_tile(const SlArgs& g, const double* Ssm, const double* ring,
                                                uint64_t* full, uint64_t* empty, int& stage, uint32_t& phase,
                                                int nkt, int m_base, int rows, int n0, int lane, int wq,
                                                uint64_t* stagger) {
  const int grp = lane >> 2, tig = lane & 3;
  const int nw = wq * 16;
  constexpr int TILE = B_KC ? SL_TILE_K : SL_TILE_N;
  double acc[MB][2][2];
#pragma unroll
  for (int i = 0; i < MB; ++i) acc[i][0][0] = acc[i][0][1] = acc[i][1][0] = acc[i][1][1] = 0.0;
  constexpr int sk = SKT;
  const double* Sp = Ssm + grp * sk + tig;
  const int b_off = B_KC ? (nw + grp) * (SL_BK + 4) + tig : tig * (SL_BN + 4) + nw + grp;

  // B fragments are double-buffered; the 13 S fragments are single-buffered and each is reloaded for the next
  // k4 step right after its two DMMAs have issued (24 DMMAs of slack before its next use).
  auto load_b = [&](const double* Bp, int k4, double (&b)[2]) {
    b[0] = B_KC ? Bp[k4 * 4] : Bp[k4 * 4 * (SL_BN + 4)];
    b[1] = B_KC ? Bp[8 * (SL_BK + 4) + k4 * 4] : Bp[k4 * 4 * (SL_BN + 4) + 8];
  };

  // One k-tile.  STEPS = 4 for full tiles (no run-time condition anywhere near the DMMA stream: a conditional
  // around mma.sync costs a WARPSYNC + branch per k4 step), 2 for the short last tile of K % 16 == 8.
  auto ktile = [&](int kt, auto steps_tag) {
    constexpr int STEPS = decltype(steps_tag)::value;
    const double* Sk = Sp + kt * SL_BK;
    double fa[MB], fb[2][2];
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) fa[mb] = Sk[mb * 8 * sk];   // resident operand: no need to wait for the stage
    mbar_wait(full + stage, phase);
    const double* Bp = ring + stage * TILE + b_off;
    load_b(Bp, 0, fb[0]);
#pragma unroll
    for (int k4 = 0; k4 < STEPS; ++k4) {
      const bool more = k4 + 1 < STEPS;
      if (more) load_b(Bp, k4 + 1, fb[(k4 + 1) & 1]);
      if (k4 == 1 && stagger && kt == min(6, nkt - 1)) {      // group 0, first tile: release group 1 (see kernel)
        __syncwarp();
        if (lane == 0) mbar_arrive(stagger);
      }
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) {
        dmma_8x8x4(acc[mb][0][0], acc[mb][0][1], fa[mb], fb[k4 & 1][0]);
        dmma_8x8x4(acc[mb][1][0], acc[mb][1][1], fa[mb], fb[k4 & 1][1]);
        // reload one block behind: the DMMAs that still read fa[mb] have a head start on the overwriting load
        if (more && mb >= 1) fa[mb - 1] = Sk[(mb - 1) * 8 * sk + (k4 + 1) * 4];
      }
      if (more) fa[MB - 1] = Sk[(MB - 1) * 8 * sk + (k4 + 1) * 4];
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + stage);
    if (++stage == g.stages) { stage = 0; phase ^= 1u; }
  };
  auto ktile_tri = [&](auto kt_tag, auto steps_tag) {
    constexpr int KT = decltype(kt_tag)::value, STEPS = decltype(steps_tag)::value;
    const double* Sk = Sp + KT * SL_BK;
    double fa[MB], fb[2][2];
#pragma unroll
    for (int mb = 0; mb < MB; ++mb)
      if (mb <= ((KT * SL_BK + 3) >> 3)) fa[mb] = Sk[mb * 8 * sk];
    mbar_wait(full + stage, phase);
    const double* Bp = ring + stage * TILE + b_off;
    load_b(Bp, 0, fb[0]);
#pragma unroll
    for (int k4 = 0; k4 < STEPS; ++k4) {
      const bool more = k4 + 1 < STEPS;
      const int mx = (KT * SL_BK + k4 * 4 + 3) >> 3;        // last row block with a non-zero in this k4 step
      const int mx1 = (KT * SL_BK + k4 * 4 + 7) >> 3;       // ... in the next one
      if (more) load_b(Bp, k4 + 1, fb[(k4 + 1) & 1]);
      if (k4 == 1 && stagger && KT == min(6, nkt - 1)) {
        __syncwarp();
        if (lane == 0) mbar_arrive(stagger);
      }
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) {
        if (mb <= mx) {
          dmma_8x8x4(acc[mb][0][0], acc[mb][0][1], fa[mb], fb[k4 & 1][0]);
          dmma_8x8x4(acc[mb][1][0], acc[mb][1][1], fa[mb], fb[k4 & 1][1]);
        }
        if (more && mb >= 1 && mb - 1 <= mx1) fa[mb - 1] = Sk[(mb - 1) * 8 * sk + (k4 + 1) * 4];
      }
      if (more && MB - 1 <= mx1) fa[MB - 1] = Sk[(MB - 1) * 8 * sk + (k4 + 1) * 4];
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + stage);
    if (++stage == g.stages) { stage = 0; phase ^= 1u; }
  };
  const int nfull = g.K / SL_BK;
  if (TRI) {
#define CG_SL_KT(KT_)                                                                                    \
  if (KT_ < nfull) ktile_tri(std::integral_constant<int, KT_>(), std::integral_constant<int, 4>());     \
  else if (KT_ < nkt) ktile_tri(std::integral_constant<int, KT_>(), std::integral_constant<int, 2>());
    CG_SL_KT(0) CG_SL_KT(1) CG_SL_KT(2) CG_SL_KT(3) CG_SL_KT(4) CG_SL_KT(5) CG_SL_KT(6)
#undef CG_SL_KT
  } else {
    for (int kt = 0; kt < nfull; ++kt) ktile(kt, std::integral_constant<int, 4>());
    if (nfull < nkt) ktile(nfull, std::integral_constant<int, 2>());
  }

  // epilogue
  const int cols = g.N - n0;
  if (g.alpha == 1.0 && rows == MB * 8 && cols >= SL_BN) {
    // Full tile, no scaling (every contraction of the sweeps): straight-line stores.  The general path below costs a
    // bounds test, a branch and a reconvergence point per DMMA block plus a DMUL per accumulator that queues behind the
    // other group's DMMAs in the shared FP64 pipe -- ncu: 11 % of the consumer warps' samples at K = 96.
    if (!B_KC) {
      double* p = g.C + (long)(m_base + grp) * g.ldc + n0 + nw + tig * 2;
#pragma unroll
      for (int mb = 0; mb < MB; ++mb, p += 8 * g.ldc) {
        *reinterpret_cast<double2*>(p) = make_double2(acc[mb][0][0], acc[mb][0][1]);
        *reinterpret_cast<double2*>(p + 8) = make_double2(acc[mb][1][0], acc[mb][1][1]);
      }
    } else {
      double* p0 = g.C + (long)(n0 + nw + tig * 2) * g.ldc + m_base + grp;
      double* p1 = p0 + g.ldc;
      double* p2 = p0 + 8 * g.ldc;
      double* p3 = p2 + g.ldc;
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) {
        p0[mb * 8] = acc[mb][0][0];
        p1[mb * 8] = acc[mb][0][1];
        p2[mb * 8] = acc[mb][1][0];
        p3[mb * 8] = acc[mb][1][1];
      }
    }
    return;
  }
#pragma unroll
  for (int mb = 0; mb < MB; ++mb) {
    if (mb * 8 >= rows) continue;
    const int row = m_base + mb * 8 + grp;
#pragma unroll
    for (int nb = 0; nb < 2; ++nb) {
      if (nw + nb * 8 >= cols) continue;
      const int col = n0 + nw + nb * 8 + tig * 2;
      const double v0 = g.alpha * acc[mb][nb][0], v1 = g.alpha * acc[mb][nb][1];
      if (!B_KC) {
        *reinterpret_cast<double2*>(g.C + (long)row * g.ldc + col) = make_double2(v0, v1);
      } else {
        double* p0 = g.C + (long)col * g.ldc + row;
        p0[0] = v0;
        p0[g.ldc] = v1;
      }
    }
  }
}

// NARROW (one-half variants whose K fits the half's own rows, e.g. the windowed right-multiplies M = K = window): the
// resident operand is stored at row stride 8 MB0 + 4 instead of 204 (both == 4 or 12 mod 16: conflict-free fragment
// loads), and the shared memory it leaves goes to a deeper ring (dgemm_sl_stages).
template <int MB0, bool NARROW>
struct SlSk { static constexpr int value = NARROW ? MB0 * 8 + 4 : SL_SK; };

template <int MB0, int MB1, bool B_KC, bool NARROW, bool TRI = false>
__global__ void __launch_bounds__(SL_NT, 1) dgemm_sl_kernel(const SlArgs g) {
  extern __shared__ __align__(128) unsigned char sl_smem_raw[];
  constexpr int TILE = B_KC ? SL_TILE_K : SL_TILE_N;
  constexpr int sk = SlSk<MB0, NARROW>::value;
  const int NS = g.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sl_smem_raw);          // [2 groups][full NS | empty NS], stagger
  double* rings = reinterpret_cast<double*>(sl_smem_raw + SL_BAR_BYTES);   // [2][NS][TILE]
  double* Ssm = rings + 2 * NS * TILE;                                // [8 MB][sk]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int half = (MB1 > 0 && (int)blockIdx.x >= g.ctas0) ? 1 : 0;
  const int m_base = half ? MB0 * 8 : 0;
  const int rows = min((half ? MB1 : MB0) * 8, g.M - m_base);   // valid rows of this half
  const int q = half ? blockIdx.x - g.ctas0 : blockIdx.x; // index among the CTAs of this half
  const int cq = half ? gridDim.x - g.ctas0 : g.ctas0;
  const int ntiles = (g.N + SL_BN - 1) / SL_BN;
  const int nkt = (g.K + SL_BK - 1) / SL_BK;

  // ---- resident half of S (rows beyond the matrix zero-filled), barriers
  for (int e = tid; e < (half ? MB1 : MB0) * 8 * (g.K / 2); e += SL_NT) {
    const int r = e / (g.K / 2), c2 = (e - r * (g.K / 2)) * 2;
    double2 v = make_double2(0.0, 0.0);
    if (r < rows) v = *reinterpret_cast<const double2*>(g.S + (long)(m_base + r) * g.lds + c2);
    *reinterpret_cast<double2*>(Ssm + r * sk + c2) = v;
  }
  if (tid == 0) {
    for (int gq = 0; gq < 2; ++gq)
      for (int s = 0; s < NS; ++s) {
        mbar_init(bars + gq * 2 * NS + s, B_KC ? 32 : 1);             // full: expect_tx arrival | 32 cp.async lanes
        mbar_init(bars + gq * 2 * NS + NS + s, 4);                    // empty: one arrival per consumer warp
      }
    mbar_init(bars + 4 * NS, 4);                                      // stagger: group 0's four warps
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  const int grpid = warp < 8 ? (warp >> 2) : warp - 8;    // consumer group / producer's group
  uint64_t* full = bars + grpid * 2 * NS;
  uint64_t* empty = full + NS;
  double* ring = rings + grpid * NS * TILE;
  int stage = 0;
  uint32_t phase = 0;

  // Group 1 starts half a tile (and half a k-tile) after group 0: its producer holds the first copy back until
  // group 0 is in the middle of k-tile 6 of its first tile.  From then on one group's epilogue and k-tile
  // hand-overs fall into the other's DMMA stream instead of coinciding with them.
  uint64_t* stagger = bars + 4 * NS;

  if (warp >= 8) {
    // ---------------------------------------------------------------- producer warp of group grpid
    if (grpid == 1) {
      if (lane == 0) mbar_wait(stagger, 0u);
      __syncwarp();
    }
    for (int it = grpid; q + (long)it * cq < ntiles; it += 2) {
      const int n0 = (q + it * cq) * SL_BN;
      const int cols = min(SL_BN, g.N - n0);
      for (int kt = 0; kt < nkt; ++kt) {
        const int kk = min(SL_BK, g.K - kt * SL_BK);      // valid k of this k-tile
        if (lane == 0) {
          mbar_wait(empty + stage, phase ^ 1u);
          if (!B_KC) mbar_expect_tx(full + stage, (uint32_t)(kk * cols * 8));
        }
        __syncwarp();
        double* dst = ring + stage * TILE;
        if (!B_KC) {
          // kk rows of `cols` contiguous doubles: one bulk copy per row
          if (lane < kk)
            bulk_g2s(dst + lane * (SL_BN + 4), g.B + (long)(kt * SL_BK + lane) * g.ldb + n0, (uint32_t)(cols * 8),
                     full + stage);
        } else {
          // `cols` rows of kk contiguous doubles (128 B): too small for bulk copies -- 16-byte cp.async, 16 per lane,
          // whose completion arrives on the stage's mbarrier (count 32, one arrival per lane)
          const int r0 = lane >> 3, kc = (lane & 7) * 2;
          const double* src = g.B + (long)(n0 + r0) * g.ldb + kt * SL_BK + kc;
          double* d = dst + r0 * (SL_BK + 4) + kc;
          const bool kok = kc < kk;
#pragma unroll
          for (int j = 0; j < SL_BN / 4; ++j) {
            const bool v = kok && r0 + 4 * j < cols;
            cp_async16(d + j * 4 * (SL_BK + 4), v ? src + (long)j * 4 * g.ldb : g.B, v);
          }
          asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(full + stage)) : "memory");
        }
        if (++stage == NS) { stage = 0; phase ^= 1u; }
      }
    }
    return;
  }

  // ------------------------------------------------------------------ consumer warps
  const int wq = warp & 3;
  for (int it = grpid; q + (long)it * cq < ntiles; it += 2) {
    const int n0 = (q + it * cq) * SL_BN;
    uint64_t* sg = it == 0 ? stagger : nullptr;
    if constexpr (MB1 == 0 || MB1 == MB0) {
      sl_consume_tile<MB0, B_KC, sk, TRI>(g, Ssm, ring, full, empty, stage, phase, nkt, m_base, rows, n0, lane, wq, sg);
    } else {
      if (half == 0)
        sl_consume_tile<MB0, B_KC, sk>(g, Ssm, ring, full, empty, stage, phase, nkt, m_base, rows, n0, lane, wq, sg);
      else
        sl_consume_tile<MB1, B_KC, sk>(g, Ssm, ring, full, empty, stage, phase, nkt, m_base, rows, n0, lane, wq, sg);
    }
  }
}

// Shared memory: barriers, the two rings, the resident operand at its own row stride K + 4.  A narrow resident operand
// (the windowed right-multiplies: 96 x 96 instead of 104 x 200) leaves room for a deeper ring, which is what hides the
// DRAM latency of the streamed operand (3 stages of 16 k = ~2 us of DMMA work; ncu showed 5 % of the consumer warps'
// time waiting on `full` barriers at K = 96).
inline int dgemm_sl_stages(int mb_rows8, int sk, bool b_kc) {
  const size_t tile = (size_t)(b_kc ? SL_TILE_K : SL_TILE_N) * 8;
  const size_t fixed = SL_BAR_BYTES + (size_t)mb_rows8 * 8 * sk * 8;
  int st = (int)((232448 - fixed) / (2 * tile));
  return st < SL_STAGES ? SL_STAGES : st > SL_MAX_STAGES ? SL_MAX_STAGES : st;
}
inline size_t dgemm_sl_smem(int mb_rows8, int sk, bool b_kc, int stages) {
  const size_t tile = (size_t)(b_kc ? SL_TILE_K : SL_TILE_N) * 8;
  return SL_BAR_BYTES + (size_t)2 * stages * tile + (size_t)mb_rows8 * 8 * sk * 8;
}

inline bool dgemm_sl_supported(int M, int N, int K) {
  return M >= 16 && M <= 2 * SL_MB * 8 && K >= 16 && K <= SL_KMAX && (M % 8) == 0 && (N % 8) == 0 && (K % 8) == 0 &&
         N >= 64 * 148;
}

// smallest instantiated number of 8-row blocks that covers mb blocks
inline int dgemm_sl_blocks(int mb) { return mb <= 5 ? 5 : mb <= 9 ? 9 : mb <= 12 ? 12 : 13; }

template <int MB0, int MB1, bool NARROW>
inline void dgemm_sl_launch(cudaStream_t st, bool b_kc, SlArgs g, int sms) {
  static DeviceOnce attr_once;
  unsigned long long attr_bit;
  if (attr_once.need(&attr_bit)) {
    const int mx = 232448;
    cudaFuncSetAttribute((const void*)dgemm_sl_kernel<MB0, MB1, false, NARROW>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    cudaFuncSetAttribute((const void*)dgemm_sl_kernel<MB0, MB1, true, NARROW>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    attr_once.done(attr_bit);
  }
  constexpr int sk = SlSk<MB0, NARROW>::value;
  g.stages = dgemm_sl_stages(MB0, sk, b_kc);
  const size_t smem = dgemm_sl_smem(MB0, sk, b_kc, g.stages);
  if (b_kc) dgemm_sl_kernel<MB0, MB1, true, NARROW><<<sms, SL_NT, smem, st>>>(g);
  else dgemm_sl_kernel<MB0, MB1, false, NARROW><<<sms, SL_NT, smem, st>>>(g);
}

template <int MB0>
inline void dgemm_sl_launch_tri(cudaStream_t st, SlArgs g, int sms) {
  static DeviceOnce attr_once;
  unsigned long long attr_bit;
  if (attr_once.need(&attr_bit)) {
    cudaFuncSetAttribute((const void*)dgemm_sl_kernel<MB0, 0, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    attr_once.done(attr_bit);
  }
  constexpr int sk = SlSk<MB0, true>::value;
  g.stages = dgemm_sl_stages(MB0, sk, true);
  const size_t smem = dgemm_sl_smem(MB0, sk, true, g.stages);
  dgemm_sl_kernel<MB0, 0, true, true, true><<<sms, SL_NT, smem, st>>>(g);
}

inline bool dgemm_sl_tri_supported(int M, int N) { return dgemm_sl_supported(M, N, M) && M <= SL_MB * 8; }

// C[n*ldc + m] = sum_{k >= m} S[m*lds + k] B[n*ldb + k]  for an UPPER-TRIANGULAR M x M operand S (everything below the
// diagonal is taken as 0 without being read as such: those DMMA blocks are skipped).  DMMAs executed relative to the
// full product: dgemm_sl_tri_fraction(M).
inline cudaError_t dgemm_sl_tri(cudaStream_t st, int M, int N, const double* S, long lds, const double* B, long ldb,
                                double* C, long ldc, int sms = 148) {
  SlArgs g;
  g.S = S; g.B = B; g.C = C; g.M = M; g.N = N; g.K = M; g.lds = lds; g.ldb = ldb; g.ldc = ldc; g.alpha = 1.0;
  g.ctas0 = sms;
  switch (dgemm_sl_blocks(M / 8)) {
    case 5: dgemm_sl_launch_tri<5>(st, g, sms); break;
    case 9: dgemm_sl_launch_tri<9>(st, g, sms); break;
    case 12: dgemm_sl_launch_tri<12>(st, g, sms); break;
    default: dgemm_sl_launch_tri<13>(st, g, sms); break;
  }
  return cudaGetLastError();
}
inline double dgemm_sl_tri_fraction(int M) {
  const int mb = M / 8;
  long act = 0;
  for (int k4 = 0; k4 * 4 < M; ++k4) act += std::min(mb, ((k4 * 4 + 3) >> 3) + 1);
  return (double)act / ((double)mb * (M / 4));
}

// b_kc = false: C[m*ldc + n] = alpha sum_k S[m*lds + k] B[k*ldb + n];  b_kc = true: C[n*ldc + m] = alpha sum_k S[m*lds + k] B[n*ldb + k]
inline cudaError_t dgemm_sl(cudaStream_t st, bool b_kc, int M, int N, int K, double alpha, const double* S, long lds,
                            const double* B, long ldb, double* C, long ldc, int sms = 148) {
  SlArgs g;
  g.S = S; g.B = B; g.C = C; g.M = M; g.N = N; g.K = K; g.lds = lds; g.ldb = ldb; g.ldc = ldc; g.alpha = alpha;
  const int mb = M / 8;
  if (mb <= SL_MB) {
    g.ctas0 = sms;
    const int blocks = dgemm_sl_blocks(mb);
    const bool narrow = K <= blocks * 8;
#define CG_SL(MB)                                                           \
  if (narrow) dgemm_sl_launch<MB, 0, true>(st, b_kc, g, sms);               \
  else dgemm_sl_launch<MB, 0, false>(st, b_kc, g, sms)
    switch (blocks) {
      case 5: CG_SL(5); break;
      case 9: CG_SL(9); break;
      case 12: CG_SL(12); break;
      default: CG_SL(13); break;
    }
#undef CG_SL
  } else {
    // CTAs are shared between the two halves in proportion to their DMMA blocks (+1: per-tile overhead)
    const int mb1 = dgemm_sl_blocks(mb - SL_MB);
    int ctas0 = (int)((double)sms * (SL_MB + 1) / (SL_MB + mb1 + 2) + 0.5);
    if (ctas0 >= sms) ctas0 = sms - 1;
    if (ctas0 < 1) ctas0 = 1;
    g.ctas0 = ctas0;
    switch (mb1) {
      case 5: dgemm_sl_launch<13, 5, false>(st, b_kc, g, sms); break;
      case 9: dgemm_sl_launch<13, 9, false>(st, b_kc, g, sms); break;
      case 12: dgemm_sl_launch<13, 12, false>(st, b_kc, g, sms); break;
      default: dgemm_sl_launch<13, 13, false>(st, b_kc, g, sms); break;
    }
  }
  return cudaGetLastError();
}

}  // namespace cg
