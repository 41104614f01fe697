// Per-device one-time setup guard.  cudaFuncSetAttribute (the > 48 KB dynamic shared memory opt-in) applies to the
// CURRENT device only, and one process may hold handles on several GPUs (one thread per GPU in the multi-restart
// driver): a process-wide `static bool` would leave every device but the first without the attribute.
// Two threads racing through the same guard both set the (idempotent) attribute; no lock is needed.
#pragma once
#include <cuda_runtime.h>

#include <atomic>

namespace cg {

struct DeviceOnce {
  std::atomic<unsigned long long> mask{0};
  // true when the current device has not been set up through this guard yet; call done() after the setup
  bool need(unsigned long long* bit) {
    int d = 0;
    cudaGetDevice(&d);
    *bit = 1ull << (d & 63);
    return (mask.load(std::memory_order_acquire) & *bit) == 0;
  }
  void done(unsigned long long bit) { mask.fetch_or(bit, std::memory_order_release); }
};

}  // namespace cg
