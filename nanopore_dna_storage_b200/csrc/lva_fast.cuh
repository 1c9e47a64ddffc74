// Register fast path of the trellis step (specialised on list slots and history words).
// Placeholder until the specialised kernels land: everything runs through lva_step_generic.
#pragma once
#include "lva_kernels.cuh"

namespace lva {
inline bool fast_supported(uint32_t, uint32_t) { return false; }
inline cudaError_t fast_prepare(uint32_t, uint32_t) { return cudaSuccess; }
inline cudaError_t fast_launch_step(uint32_t, cudaStream_t, const StepArgs &, const CodeSpec &, uint32_t,
                                    uint32_t, uint32_t, uint64_t *) {
  return cudaErrorNotSupported;
}
}  // namespace lva
