// What limits DMMA throughput under GEMM-like conditions on B200?  Variants of a register-tile inner loop.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/dmma_probe tools/dmma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(2);} } while (0)

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double* d, const double* a, double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* d, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* d, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

extern __shared__ double sm[];

// V0: m8n8k4, MB x 2 register tile, operands loaded from shared memory each k4 step (LDS = 1) or kept in registers.
template <int MB, int LDS>
__global__ void k884(double* out, int iters) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, grp = lane >> 2, tig = lane & 3;
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 1e-3 * i;
  __syncthreads();
  double acc[MB][2][2];
#pragma unroll
  for (int i = 0; i < MB; ++i) acc[i][0][0] = acc[i][0][1] = acc[i][1][0] = acc[i][1][1] = 0.0;
  const double* Ap = sm + grp * 20 + tig;
  const double* Bp = sm + 2200 + (warp * 16 + grp) * 20 + tig;
  double a[MB], b0, b1;
#pragma unroll
  for (int mb = 0; mb < MB; ++mb) a[mb] = Ap[mb * 160];
  b0 = Bp[0]; b1 = Bp[160];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4) {
      if (LDS) {
        b0 = Bp[k4 * 4]; b1 = Bp[160 + k4 * 4];
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) a[mb] = Ap[mb * 160 + k4 * 4];
      }
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) {
        dmma884(acc[mb][0][0], acc[mb][0][1], a[mb], b0);
        dmma884(acc[mb][1][0], acc[mb][1][1], a[mb], b1);
      }
    }
    if (LDS == 2) __syncthreads();
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < MB; ++i) s += acc[i][0][0] + acc[i][0][1] + acc[i][1][0] + acc[i][1][1];
  if (s == 123.456) out[0] = s;
}

// V1: m16n8k{4,8,16}, MB16 x 2 register tile, A fragments loaded per m-block from shared memory.
template <int MB16, int KK, int LDS>
__global__ void k16(double* out, int iters) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, grp = lane >> 2, tig = lane & 3;
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 1e-3 * i;
  __syncthreads();
  double acc[MB16][2][4];
#pragma unroll
  for (int i = 0; i < MB16; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][0][j] = acc[i][1][j] = 0.0;
  const double* Ap = sm + grp * 20 + tig;
  const double* Bp = sm + 2400 + (warp % 4 * 16 + grp) * 20 + tig;
  constexpr int NA = KK / 2, NB = KK / 4;
  double a[NA], b[2][NB];
#pragma unroll
  for (int j = 0; j < NA; ++j) a[j] = Ap[(j & 1) * 160 + (j >> 1) * 4];
#pragma unroll
  for (int j = 0; j < NB; ++j) { b[0][j] = Bp[j * 4]; b[1][j] = Bp[160 + j * 4]; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int ks = 0; ks < 16 / KK; ++ks) {
      if (LDS) {
#pragma unroll
        for (int j = 0; j < NB; ++j) { b[0][j] = Bp[ks * KK + j * 4]; b[1][j] = Bp[160 + ks * KK + j * 4]; }
      }
#pragma unroll
      for (int mb = 0; mb < MB16; ++mb) {
        if (LDS) {
#pragma unroll
          for (int j = 0; j < NA; ++j) a[j] = Ap[mb * 320 + (j & 1) * 160 + ks * KK + (j >> 1) * 4];
        }
        if (KK == 4) { dmma1684(acc[mb][0], a, b[0][0]); dmma1684(acc[mb][1], a, b[1][0]); }
        if (KK == 8) { dmma1688(acc[mb][0], a, b[0]); dmma1688(acc[mb][1], a, b[1]); }
        if (KK == 16) { dmma16816(acc[mb][0], a, b[0]); dmma16816(acc[mb][1], a, b[1]); }
      }
    }
    if (LDS == 2) __syncthreads();
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < MB16; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += acc[i][0][j] + acc[i][1][j];
  if (s == 123.456) out[0] = s;
}


// V2: m8n8k4, MB x NB register tile per warp, fragments double-buffered in registers, barrier every 4 k4 steps.
template <int MB, int NB, int BAR>
__global__ void k884g(double* out, int iters) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, grp = lane >> 2, tig = lane & 3;
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = 1e-3 * i;
  __syncthreads();
  double acc[MB][NB][2];
#pragma unroll
  for (int i = 0; i < MB; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const double* Ap = sm + ((warp & 3) * 8 + grp) * 20 + tig;
  const double* Bp = sm + 4400 + ((warp >> 2) * 8 + grp) * 20 + tig;
  double a[2][MB], b[2][NB];
  auto ld = [&](int k4, double (&aa)[MB], double (&bb)[NB]) {
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) aa[mb] = Ap[mb * 160 + k4 * 4];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) bb[nb] = Bp[nb * 160 + k4 * 4];
  };
  for (int it = 0; it < iters; ++it) {
    ld(0, a[0], b[0]);
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4) {
      if (k4 + 1 < 4) ld(k4 + 1, a[(k4 + 1) & 1], b[(k4 + 1) & 1]);
#pragma unroll
      for (int mb = 0; mb < MB; ++mb)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[k4 & 1][mb], b[k4 & 1][nb]);
    }
    if (BAR) __syncthreads();
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < MB; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) s += acc[i][j][0] + acc[i][j][1];
  if (s == 123.456) out[0] = s;
}

template <class K>
void run(const char* name, K kern, int threads, int smem_bytes, double flop_per_thread_iter, int iters) {
  double* out; CK(cudaMalloc(&out, 8));
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  int nb = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, threads, smem_bytes));
  int grid = 148 * nb;
  kern<<<grid, threads, smem_bytes>>>(out, 10);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    kern<<<grid, threads, smem_bytes>>>(out, iters);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  double fl = flop_per_thread_iter * iters * (double)grid * threads;
  printf("%-48s blocks/SM=%d warps/SM=%2d  %7.2f TFLOP/s\n", name, nb, nb * threads / 32, fl / best / 1e9);
  cudaFree(out);
}

int main() {
  const int it = 4000;
  // flops per thread per iter: m8n8k4 = 512 flop per warp-instr = 16 per thread
  run("884 MB13 regs          1x256thr", k884<13, 0>, 256, 120000, 13 * 2 * 4 * 16.0, it);
  run("884 MB13 LDS           1x256thr", k884<13, 1>, 256, 120000, 13 * 2 * 4 * 16.0, it);
  run("884 MB13 LDS+barrier   1x256thr", k884<13, 2>, 256, 120000, 13 * 2 * 4 * 16.0, it);
  run("884 MB6  LDS           2x256thr", k884<6, 1>, 256, 100000, 6 * 2 * 4 * 16.0, it);
  run("884 MB6  LDS+barrier   2x256thr", k884<6, 2>, 256, 100000, 6 * 2 * 4 * 16.0, it);
  run("884 MB6  LDS+barrier   1x512thr", k884<6, 2>, 512, 120000, 6 * 2 * 4 * 16.0, it);
  run("884 MB4  LDS+barrier   3x256thr", k884<4, 2>, 256, 70000, 4 * 2 * 4 * 16.0, it);
  run("1684 MB7 LDS+barrier   1x256thr", k16<7, 4, 2>, 256, 120000, 7 * 2 * 4 * 32.0, it);
  run("1688 MB7 LDS+barrier   1x256thr", k16<7, 8, 2>, 256, 120000, 7 * 2 * 2 * 64.0, it);
  run("16816 MB7 regs         1x256thr", k16<7, 16, 0>, 256, 120000, 7 * 2 * 1 * 128.0, it);
  run("16816 MB7 LDS          1x256thr", k16<7, 16, 1>, 256, 120000, 7 * 2 * 1 * 128.0, it);
  run("16816 MB7 LDS+barrier  1x256thr", k16<7, 16, 2>, 256, 120000, 7 * 2 * 1 * 128.0, it);
  run("16816 MB3 LDS+barrier  2x256thr", k16<3, 16, 2>, 256, 100000, 3 * 2 * 1 * 128.0, it);
  run("16816 MB4 LDS+barrier  1x512thr", k16<4, 16, 2>, 512, 120000, 4 * 2 * 1 * 128.0, it);
  run("1688 MB4 LDS+barrier   1x512thr", k16<4, 8, 2>, 512, 120000, 4 * 2 * 2 * 64.0, it);
  // generic MB x NB warp tiles, double-buffered fragments
  run("884g 13x2 bar  2x128thr (8 warps/SM)", k884g<13, 2, 1>, 128, 100000, 13 * 2 * 4 * 16.0, it);
  run("884g 13x2 bar  1x128thr (4 warps/SM)", k884g<13, 2, 1>, 128, 120000, 13 * 2 * 4 * 16.0, it);
  run("884g 13x2 bar  1x256thr (8 warps/SM)", k884g<13, 2, 1>, 256, 120000, 13 * 2 * 4 * 16.0, it);
  run("884g 13x2 nobar 2x128thr", k884g<13, 2, 0>, 128, 100000, 13 * 2 * 4 * 16.0, it);
  run("884g 5x5 bar   1x480thr (15 warps/SM)", k884g<5, 5, 1>, 480, 120000, 5 * 5 * 4 * 16.0, it);
  run("884g 5x5 bar   1x256thr (8 warps/SM)", k884g<5, 5, 1>, 256, 120000, 5 * 5 * 4 * 16.0, it);
  run("884g 5x5 bar   2x256thr (16 warps/SM)", k884g<5, 5, 1>, 256, 100000, 5 * 5 * 4 * 16.0, it);
  run("884g 5x5 nobar 1x480thr", k884g<5, 5, 0>, 480, 120000, 5 * 5 * 4 * 16.0, it);
  run("884g 7x4 bar   1x256thr (8 warps/SM)", k884g<7, 4, 1>, 256, 120000, 7 * 4 * 4 * 16.0, it);
  run("884g 7x4 bar   2x128thr (8 warps/SM)", k884g<7, 4, 1>, 128, 100000, 7 * 4 * 4 * 16.0, it);
  run("884g 6x4 bar   1x256thr (8 warps/SM)", k884g<6, 4, 1>, 256, 120000, 6 * 4 * 4 * 16.0, it);
  run("884g 6x4 bar   1x384thr (12 warps/SM)", k884g<6, 4, 1>, 384, 120000, 6 * 4 * 4 * 16.0, it);
  run("884g 4x4 bar   1x512thr (16 warps/SM)", k884g<4, 4, 1>, 512, 120000, 4 * 4 * 4 * 16.0, it);
  run("884g 13x1 bar  1x256thr (8 warps/SM)", k884g<13, 1, 1>, 256, 120000, 13 * 1 * 4 * 16.0, it);
  run("884g 13x1 bar  1x512thr (16 warps/SM)", k884g<13, 1, 1>, 512, 120000, 13 * 1 * 4 * 16.0, it);
  return 0;
}
