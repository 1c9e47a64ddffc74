// Per-device scratch arena of the stateless entry points (lva_trim_batch, lva_posteriors_batch):
// device buffers that grow on demand and are kept between calls, and one private non-blocking
// stream per device, so that a call costs no cudaMalloc / cudaFree and never touches the default
// stream.  A call holds the device's mutex from first use of the arena to its last.
#pragma once
#include <cstddef>
#include <mutex>

#include <cuda_runtime.h>

namespace lva {

struct DeviceScratch {
  static constexpr int kBufs = 8;
  std::mutex mu;
  cudaStream_t stream = nullptr;
  void *buf[kBufs] = {};
  size_t cap[kBufs] = {};

  // buffer `k` with room for n bytes (nullptr when the device is out of memory)
  void *get(int k, size_t n) {
    if (n <= cap[k] && buf[k]) return buf[k];
    if (buf[k]) cudaFree(buf[k]);
    buf[k] = nullptr;
    cap[k] = 0;
    const size_t want = n < 4096 ? 4096 : n + n / 4;
    if (cudaMalloc(&buf[k], want) != cudaSuccess) { cudaGetLastError(); buf[k] = nullptr; return nullptr; }
    cap[k] = want;
    return buf[k];
  }
  bool ready() {
    if (stream) return true;
    return cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) == cudaSuccess;
  }
};

inline DeviceScratch &scratch_for(int device) {
  static DeviceScratch arenas[16];
  return arenas[device < 0 || device >= 16 ? 0 : device];
}

}  // namespace lva
