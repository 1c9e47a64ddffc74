// Flip-flop transition posteriors on the GPU: from the CRF transition
// weights the flappie network emits (T x 40 per read) to the T x 40 log-posteriors the list decoder
// reads -- forward-backward over the 8 flip-flop states, then a per-block log-normalisation.
//
// Restates, for a batch of reads,
//   transpost_crf_flipflop(trans, return_log = true)   /root/reference/flappie/src/decode.c:377-497
//   logsumexpf                                         /root/reference/flappie/src/util.h:188-190
//   log_row_normalise_inplace                          /root/reference/flappie/src/flappie_matrix.c:450-467
// with every sum folded in the reference's order (so the only difference to the CPU result is the
// last-bit behaviour of expf / log1pf).  Input and output use the .post layout the decoder expects
// (block-major; [to*8 + from] for flip targets, [32 + from] for the flop target of `from`), so the
// result can be handed to lva_decode_batch_device without leaving the GPU.
//
// Mapping: 8 lanes per read (lane = flip-flop state), 4 reads per warp; the recursions are serial in
// T by nature, the batch supplies the parallelism.  The forward vectors are kept in a scratch array
// (8 floats per block), read back by the backward pass.
#include <cstdint>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/lva_b200.h"
#include "device_scratch.hpp"

namespace {

__device__ __forceinline__ float lse(float x, float y) { return fmaxf(x, y) + log1pf(expf(-fabsf(x - y))); }

__global__ void __launch_bounds__(128) flipflop_posteriors(const float *trans, const uint64_t *offs, uint32_t n_reads,
                                                           float *out, float *fwd_all) {
  const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;  // read
  if (g >= n_reads) return;
  const int s = threadIdx.x & 7;                                    // my flip-flop state
  const unsigned gm = 0xffu << (threadIdx.x & 24);                  // my 8 lanes (threadIdx.x & 31 & ~7)
  const int l0 = threadIdx.x & 24;                                  // first lane of the group
  const uint64_t b0 = offs[g];
  const uint32_t T = (uint32_t)(offs[g + 1] - b0);
  const float *t0 = trans + b0 * 40;
  float *o0 = out + b0 * 40;
  float *fwd = fwd_all + (b0 + g) * 8;                              // (T + 1) x 8

  // ---- forward (decode.c:398-424) ----
  float v = 0.f;                                                    // fwd column 0 is zero-filled
  fwd[s] = v;
  for (uint32_t blk = 0; blk < T; blk++) {
    const float *t = t0 + (uint64_t)blk * 40;
    float prev[8];
#pragma unroll
    for (int k = 0; k < 8; k++) prev[k] = __shfl_sync(gm, v, l0 + k);
    float nv;
    if (s >= 4) {  // flop: stay, then the move from the flip state of the same base
      nv = prev[s] + t[32 + s];
      nv = lse(nv, prev[s - 4] + t[32 + s - 4]);
    } else {       // flip: every state, in state order
      nv = t[s * 8] + prev[0];
#pragma unroll
      for (int from = 1; from < 8; from++) nv = lse(nv, t[s * 8 + from] + prev[from]);
    }
    v = nv;
    fwd[(uint64_t)(blk + 1) * 8 + s] = v;
  }

  // ---- backward, posteriors, normalisation (decode.c:436-483, flappie_matrix.c:456-466) ----
  float bw = 0.f;                                                   // backward vector after the last block
  for (uint32_t blk = T; blk > 0; blk--) {
    const float *t = t0 + (uint64_t)(blk - 1) * 40;
    const float f = fwd[(uint64_t)(blk - 1) * 8 + s];
    float prev[8];
#pragma unroll
    for (int k = 0; k < 8; k++) prev[k] = __shfl_sync(gm, bw, l0 + k);
    // my five entries of the block: rows b1*8 + s (ending in flip b1 from state s) and 32 + s
    float o[5];
#pragma unroll
    for (int b1 = 0; b1 < 4; b1++) o[b1] = f + prev[b1] + t[b1 * 8 + s];
    o[4] = f + prev[s >= 4 ? s : s + 4] + t[32 + s];
    // new backward vector
    float nb = prev[s >= 4 ? s : s + 4] + t[32 + s];
#pragma unroll
    for (int b1 = 0; b1 < 4; b1++) nb = lse(nb, t[b1 * 8 + s] + prev[b1]);
    bw = nb;
    // log-sum of the 40 entries, folded in row order
    float sum = 0.f;
#pragma unroll
    for (int r = 0; r < 40; r++) {
      const float x = __shfl_sync(gm, o[r >> 3], l0 + (r & 7));
      sum = r == 0 ? x : lse(sum, x);
    }
    float *ob = o0 + (uint64_t)(blk - 1) * 40;
#pragma unroll
    for (int b1 = 0; b1 < 4; b1++) ob[b1 * 8 + s] = o[b1] - sum;
    ob[32 + s] = o[4] - sum;
  }
}

int run(int32_t device, const float *trans, bool on_device, const uint64_t *offs, uint32_t n_reads, float *out) {
  if (n_reads == 0) return LVA_OK;
  if (!trans || !offs || !out) return LVA_ERR_BAD_ARG;
  for (uint32_t i = 0; i < n_reads; i++)
    if (offs[i + 1] < offs[i] || offs[i + 1] - offs[i] > 0xffffffffull) return LVA_ERR_BAD_ARG;
  if (device < 0 || device >= 16 || cudaSetDevice(device) != cudaSuccess) return LVA_ERR_NO_DEVICE;
  const uint64_t first = offs[0], nblk = offs[n_reads] - first;
  std::vector<uint64_t> rel(n_reads + 1);
  for (uint32_t i = 0; i <= n_reads; i++) rel[i] = offs[i] - first;
  lva::DeviceScratch &S = lva::scratch_for(device);
  std::lock_guard<std::mutex> hold(S.mu);
  if (!S.ready()) return LVA_ERR_CUDA;
  uint64_t *d_offs = (uint64_t *)S.get(0, (n_reads + 1) * 8);
  float *d_fwd = (float *)S.get(1, (nblk + n_reads) * 32 + 32);
  float *d_trans = on_device ? nullptr : (float *)S.get(2, nblk * 160 + 16);
  float *d_out = on_device ? nullptr : (float *)S.get(3, nblk * 160 + 16);
  if (!d_offs || !d_fwd || (!on_device && (!d_trans || !d_out))) return LVA_ERR_OOM;
  cudaMemcpyAsync(d_offs, rel.data(), (n_reads + 1) * 8, cudaMemcpyHostToDevice, S.stream);
  const float *tin = on_device ? trans + first * 40 : d_trans;
  float *tout = on_device ? out + first * 40 : d_out;
  if (!on_device) cudaMemcpyAsync(d_trans, trans + first * 40, nblk * 160, cudaMemcpyHostToDevice, S.stream);
  const uint32_t groups_per_block = 128 / 8;
  flipflop_posteriors<<<(n_reads + groups_per_block - 1) / groups_per_block, 128, 0, S.stream>>>(tin, d_offs, n_reads, tout, d_fwd);
  if (cudaGetLastError() != cudaSuccess) return LVA_ERR_CUDA;
  if (!on_device) cudaMemcpyAsync(out + first * 40, d_out, nblk * 160, cudaMemcpyDeviceToHost, S.stream);
  return cudaStreamSynchronize(S.stream) == cudaSuccess ? LVA_OK : LVA_ERR_CUDA;
}

}  // namespace

extern "C" int lva_posteriors_batch(int32_t device, const float *trans, const uint64_t *block_offsets, uint32_t n_reads,
                                    float *out_post) {
  return run(device, trans, false, block_offsets, n_reads, out_post);
}

extern "C" int lva_posteriors_batch_device(int32_t device, const float *trans_dev, const uint64_t *block_offsets,
                                           uint32_t n_reads, float *out_post_dev) {
  return run(device, trans_dev, true, block_offsets, n_reads, out_post_dev);
}
