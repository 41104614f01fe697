// FP64 peak micro-benchmark for B200 (sm_100a): the roofline denominators that
// MEASURED_PEAKS.json does not carry ("the builder must measure sustained
// DFMA and DMMA peaks on the box").  Prints one JSON object on stdout.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peaks tools/fp64_peaks.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2);} } while (0)

constexpr int ILP = 8;

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* d, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* d, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

__global__ void __launch_bounds__(256) k_dmma884(double* out, int iters, double a, double b) {
  double d[ILP][2];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { d[i][0] = threadIdx.x * 1e-9; d[i][1] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma884(d[i][0], d[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += d[i][0] + d[i][1];
  if (s == 123.456) out[0] = s;
}
__global__ void __launch_bounds__(256) k_dmma1688(double* out, int iters, double a, double b) {
  double d[ILP][4]; double af[4] = {a, a + 1, a + 2, a + 3}; double bf[2] = {b, b + 1};
#pragma unroll
  for (int i = 0; i < ILP; ++i) { d[i][0] = threadIdx.x * 1e-9; d[i][1] = i; d[i][2] = 1; d[i][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma1688(d[i], af, bf);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  if (s == 123.456) out[0] = s;
}
__global__ void __launch_bounds__(256) k_dmma16816(double* out, int iters, double a, double b) {
  double d[ILP][4]; double af[8]; double bf[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) af[i] = a + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) bf[i] = b + i;
#pragma unroll
  for (int i = 0; i < ILP; ++i) { d[i][0] = threadIdx.x * 1e-9; d[i][1] = i; d[i][2] = 1; d[i][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma16816(d[i], af, bf);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  if (s == 123.456) out[0] = s;
}
// half the warps issue DFMA, the other half DMMA: do the two pipes add?
__global__ void __launch_bounds__(256) k_mixed(double* out, int iters, double a, double b) {
  int warp = threadIdx.x >> 5;
  double s = 0;
  if ((warp >> 2) & 1) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
  } else {
    double d[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { d[i][0] = threadIdx.x * 1e-9; d[i][1] = i; }
    for (int it = 0; it < iters / 8; ++it) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) dmma884(d[i][0], d[i][1], a, b);
    }
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += d[i][0] + d[i][1];
  }
  if (s == 123.456) out[0] = s;
}
template <int OP>
__global__ void __launch_bounds__(256) k_special(double* out, int iters, double x0) {
  double x = x0 + threadIdx.x * 1e-3, s = 0;
  for (int it = 0; it < iters; ++it) {
    double v;
    if (OP == 0) v = exp(-x); else if (OP == 1) v = erfc(x); else if (OP == 2) v = sqrt(x + 1.0); else v = 1.0 / (x + 1.0);
    s += v; x += 1e-6;
  }
  if (s == 123.456) out[0] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, 64));
  const int blocks = sms * 8, threads = 256;
  const double warps = (double)blocks * threads / 32.0;
  printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
  {
    int iters = 1 << 14;
    double ms = time_ms([&] { k_dfma<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 10);
    double fl = (double)blocks * threads * iters * ILP * 2.0;
    printf(", \"dfma_tflops\": %.3f", fl / ms / 1e9);
    // sustained: back to back for ~3 s
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    int n = (int)(3000.0 / ms) + 1; CK(cudaEventRecord(e0));
    for (int i = 0; i < n; ++i) k_dfma<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float t; CK(cudaEventElapsedTime(&t, e0, e1));
    printf(", \"dfma_tflops_sustained\": %.3f", fl * n / t / 1e9);
  }
  {
    int iters = 1 << 12;
    double ms = time_ms([&] { k_dmma884<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 10);
    double fl = warps * iters * ILP * 8.0 * 8 * 4 * 2;
    printf(", \"dmma_m8n8k4_tflops\": %.3f", fl / ms / 1e9);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    int n = (int)(3000.0 / ms) + 1; CK(cudaEventRecord(e0));
    for (int i = 0; i < n; ++i) k_dmma884<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float t; CK(cudaEventElapsedTime(&t, e0, e1));
    printf(", \"dmma_m8n8k4_tflops_sustained\": %.3f", fl * n / t / 1e9);
    ms = time_ms([&] { k_dmma1688<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 10);
    printf(", \"dmma_m16n8k8_tflops\": %.3f", warps * iters * ILP * 16.0 * 8 * 8 * 2 / ms / 1e9);
    ms = time_ms([&] { k_dmma16816<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 10);
    printf(", \"dmma_m16n8k16_tflops\": %.3f", warps * iters * ILP * 16.0 * 8 * 16 * 2 / ms / 1e9);
  }
  {
    int iters = 1 << 14;
    double ms = time_ms([&] { k_mixed<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 10);
    double fl_fma = (double)blocks * threads / 2 * iters * ILP * 2.0;
    double fl_mma = warps / 2 * (iters / 8) * ILP * 512.0;
    printf(", \"mixed_dfma_tflops\": %.3f, \"mixed_dmma_tflops\": %.3f", fl_fma / ms / 1e9, fl_mma / ms / 1e9);
  }
  {
    int iters = 1 << 12; double n = (double)blocks * threads * iters;
    double ms = time_ms([&] { k_special<0><<<blocks, threads>>>(out, iters, 0.5); }, 5);
    printf(", \"exp_gops\": %.2f", n / ms / 1e6);
    ms = time_ms([&] { k_special<1><<<blocks, threads>>>(out, iters, 0.5); }, 5);
    printf(", \"erfc_gops\": %.2f", n / ms / 1e6);
    ms = time_ms([&] { k_special<2><<<blocks, threads>>>(out, iters, 0.5); }, 5);
    printf(", \"sqrt_gops\": %.2f", n / ms / 1e6);
    ms = time_ms([&] { k_special<3><<<blocks, threads>>>(out, iters, 0.5); }, 5);
    printf(", \"rcp_gops\": %.2f", n / ms / 1e6);
  }
  printf("}\n");
  return 0;
}
