// Posterior trimming on the GPU: where the payload starts and ends in a
// read's posterior matrix, and in which orientation, from the barcodes found in its basecall.
//
// Restates, for a batch of reads,
//   helper.find_barcode_pos_in_post      /root/reference/helper.py:157-209
//   the orientation / size rule          /root/reference/generate_decoded_lists.py:68-83
// (helper.truncate_post_file, helper.py:211-224, becomes a pair of block offsets: lva_decode_batch
// takes [first, last+1) per read, so nothing is copied).
//
// One thread per (read, barcode, window): Levenshtein distance (unit costs; PyPI `distance`,
// distance.levenshtein) between the barcode and the equally long window of the basecall, two DP
// rows in local memory; the first window with the smallest distance wins through a 64-bit
// atomicMin on (distance << 32 | window).  A second small kernel turns the four winners of a
// read into block positions via the read's .trans table and applies the reference's rules.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/lva_b200.h"
#include "device_scratch.hpp"

namespace {

constexpr int kMaxBarcode = LVA_TRIM_MAX_BARCODE;

struct TrimArgs {
  const uint8_t *bases;             // concatenated basecalls
  const uint64_t *base_off;         // [n_reads + 1]
  const uint32_t *trans;            // concatenated .trans tables (block index per base)
  const uint64_t *trans_off;        // [n_reads + 1]
  unsigned long long *best;         // [n_reads][4]  (distance << 32 | window), start/end x fwd/rc
  uint32_t n_reads, min_blocks;
  uint32_t blen[4];                 // start fwd, end fwd, start rc, end rc
  uint8_t bc[4][kMaxBarcode];
  int32_t *out_start, *out_end, *out_rc, *out_status;
  uint32_t *out_dist;               // [n_reads][4]
};

__device__ __forceinline__ uint32_t window_count(uint32_t L, uint32_t blen, bool is_end) {
  // helper.py:183 range(basecall_len//2+1-start_barcode_len); helper.py:187 range(basecall_len//2, basecall_len-end_barcode_len)
  const long long n = is_end ? (long long)L - blen - (long long)(L / 2) : (long long)(L / 2) + 1 - blen;
  return n > 0 ? (uint32_t)n : 0u;
}

__global__ void trim_search(const TrimArgs A) {
  const uint32_t r = blockIdx.y, which = blockIdx.z;
  const bool is_end = which & 1u;
  const uint32_t blen = A.blen[which];
  const uint64_t b0 = A.base_off[r];
  const uint32_t L = (uint32_t)(A.base_off[r + 1] - b0);
  // helper.py:177-179: both barcodes of the orientation must fit
  if (A.blen[which & 2u] + A.blen[(which & 2u) | 1u] > L) return;
  const uint32_t nwin = window_count(L, blen, is_end);
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nwin) return;
  const uint8_t *w = A.bases + b0 + (is_end ? L / 2 : 0u) + k;
  const uint8_t *bc = A.bc[which];
  // Levenshtein, rows = barcode, columns = window; prev[j] = D[i-1][j]
  uint8_t prev[kMaxBarcode + 1];
  for (uint32_t j = 0; j <= blen; j++) prev[j] = (uint8_t)j;
  for (uint32_t i = 1; i <= blen; i++) {
    uint32_t diag = prev[0];
    prev[0] = (uint8_t)i;
    const uint8_t a = bc[i - 1];
    for (uint32_t j = 1; j <= blen; j++) {
      const uint32_t up = prev[j], left = prev[j - 1];
      const uint32_t sub = diag + (a != w[j - 1] ? 1u : 0u);
      uint32_t v = up + 1u < left + 1u ? up + 1u : left + 1u;
      v = sub < v ? sub : v;
      diag = up;
      prev[j] = (uint8_t)v;
    }
  }
  atomicMin(A.best + (unsigned long long)r * 4 + which, ((unsigned long long)prev[blen] << 32) | k);
}

__global__ void trim_finish(const TrimArgs A) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= A.n_reads) return;
  const uint32_t L = (uint32_t)(A.base_off[r + 1] - A.base_off[r]);
  const uint32_t *trans = A.trans + A.trans_off[r];
  const uint32_t ntrans = (uint32_t)(A.trans_off[r + 1] - A.trans_off[r]);
  const uint32_t INF = 0xffffffffu;
  int32_t start[2], end[2], st[2];
  uint32_t ds[2], de[2];
  for (int o = 0; o < 2; o++) {
    start[o] = end[o] = -1;
    ds[o] = de[o] = INF;
    st[o] = LVA_TRIM_OK;
    const uint32_t sl = A.blen[2 * o], el = A.blen[2 * o + 1];
    if (sl + el > L) { st[o] = LVA_TRIM_TOO_SHORT; continue; }           // helper.py:177-179
    if (window_count(L, sl, false) == 0 || window_count(L, el, true) == 0 || ntrans < L) {
      st[o] = LVA_TRIM_NO_WINDOW;                                          // (the reference raises here)
      continue;
    }
    const unsigned long long bs = A.best[(unsigned long long)r * 4 + 2 * o], be = A.best[(unsigned long long)r * 4 + 2 * o + 1];
    const uint32_t s_first = (uint32_t)bs, e_first = L / 2 + (uint32_t)be;  // helper.py:191-192
    const uint32_t s_last = s_first + sl - 1;                               // helper.py:193
    const int32_t sp = (int32_t)trans[s_last + 1] - 1;                      // helper.py:194
    const int32_t ep = (int32_t)trans[e_first - 1] - 1;                     // helper.py:195
    if (ep < sp) { st[o] = LVA_TRIM_BARCODE_ORDER; continue; }              // helper.py:207-209
    start[o] = sp;
    end[o] = ep;
    ds[o] = (uint32_t)(bs >> 32);
    de[o] = (uint32_t)(be >> 32);
  }
  // generate_decoded_lists.py:70-74: rc iff dist_start + dist_end > dist_start_RC + dist_end_RC (inf on failure)
  const bool f_ok = ds[0] != INF, r_ok = ds[1] != INF;
  const unsigned long long sf = (unsigned long long)ds[0] + de[0], sr = (unsigned long long)ds[1] + de[1];
  const bool rc = r_ok && (!f_ok || sf > sr);
  const int o = rc ? 1 : 0;
  int32_t status = st[o];
  // generate_decoded_lists.py:76: too few blocks left for the trellis
  if (status == LVA_TRIM_OK && (uint32_t)(end[o] - start[o] + 1) < A.min_blocks) status = LVA_TRIM_TOO_FEW_BLOCKS;
  A.out_start[r] = start[o];
  A.out_end[r] = end[o];
  A.out_rc[r] = rc ? 1 : 0;
  A.out_status[r] = status;
  if (A.out_dist) {
    A.out_dist[4 * r + 0] = ds[0];
    A.out_dist[4 * r + 1] = de[0];
    A.out_dist[4 * r + 2] = ds[1];
    A.out_dist[4 * r + 3] = de[1];
  }
}

bool reverse_complement(const char *s, uint32_t n, uint8_t *out) {  // helper.py:226-229
  for (uint32_t i = 0; i < n; i++) {
    const char c = s[n - 1 - i];
    char d;
    switch (c) {
      case 'A': d = 'T'; break;
      case 'C': d = 'G'; break;
      case 'G': d = 'C'; break;
      case 'T': d = 'A'; break;
      case 'N': d = 'N'; break;
      default: return false;
    }
    out[i] = (uint8_t)d;
  }
  return true;
}

}  // namespace

extern "C" int lva_trim_batch(int32_t device, const char *basecalls, const uint64_t *base_offsets,
                              const uint32_t *trans, const uint64_t *trans_offsets, uint32_t n_reads,
                              const char *start_barcode, const char *end_barcode, uint32_t min_blocks,
                              int32_t *out_start, int32_t *out_end, int32_t *out_rc, int32_t *out_status,
                              uint32_t *out_dist) {
  if (n_reads == 0) return LVA_OK;
  if (!basecalls || !base_offsets || !trans || !trans_offsets || !start_barcode || !end_barcode || !out_start ||
      !out_end || !out_rc || !out_status)
    return LVA_ERR_BAD_ARG;
  const size_t sl = strlen(start_barcode), el = strlen(end_barcode);
  if (sl == 0 || el == 0 || sl > (size_t)kMaxBarcode || el > (size_t)kMaxBarcode) return LVA_ERR_BAD_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return LVA_ERR_NO_DEVICE;

  TrimArgs a{};
  a.n_reads = n_reads;
  a.min_blocks = min_blocks;
  a.blen[0] = (uint32_t)sl;
  a.blen[1] = (uint32_t)el;
  a.blen[2] = (uint32_t)el;  // START_BARCODE_RC = rc(END_BARCODE)   generate_decoded_lists.py:33
  a.blen[3] = (uint32_t)sl;  // END_BARCODE_RC   = rc(START_BARCODE) generate_decoded_lists.py:34
  memcpy(a.bc[0], start_barcode, sl);
  memcpy(a.bc[1], end_barcode, el);
  if (!reverse_complement(end_barcode, (uint32_t)el, a.bc[2]) || !reverse_complement(start_barcode, (uint32_t)sl, a.bc[3]))
    return LVA_ERR_BAD_ARG;

  const uint64_t nb = base_offsets[n_reads] - base_offsets[0], nt = trans_offsets[n_reads] - trans_offsets[0];
  uint32_t max_len = 0;
  std::vector<uint64_t> boff(n_reads + 1), toff(n_reads + 1);
  for (uint32_t i = 0; i <= n_reads; i++) {
    boff[i] = base_offsets[i] - base_offsets[0];
    toff[i] = trans_offsets[i] - trans_offsets[0];
    if (i && boff[i] < boff[i - 1]) return LVA_ERR_BAD_ARG;
    if (i && toff[i] < toff[i - 1]) return LVA_ERR_BAD_ARG;
    if (i && boff[i] - boff[i - 1] > 0x7fffffffull) return LVA_ERR_BAD_ARG;
    if (i) max_len = std::max<uint32_t>(max_len, (uint32_t)(boff[i] - boff[i - 1]));
  }
  if (device < 0 || device >= 16) return LVA_ERR_NO_DEVICE;
  lva::DeviceScratch &S = lva::scratch_for(device);
  std::lock_guard<std::mutex> hold(S.mu);
  if (!S.ready()) return LVA_ERR_CUDA;
  struct Ref { void *p; } d_bases{S.get(0, nb + 16)}, d_boff{S.get(1, (n_reads + 1) * 8)}, d_trans{S.get(2, nt * 4 + 16)},
      d_toff{S.get(3, (n_reads + 1) * 8)}, d_best{S.get(4, (size_t)n_reads * 4 * 8)};
  const size_t out_words = (size_t)n_reads * 8;
  Ref d_out{S.get(5, out_words * 4)};
  if (!d_bases.p || !d_boff.p || !d_trans.p || !d_toff.p || !d_best.p || !d_out.p) return LVA_ERR_OOM;
  cudaMemcpyAsync(d_bases.p, basecalls + base_offsets[0], nb, cudaMemcpyHostToDevice, S.stream);
  cudaMemcpyAsync(d_boff.p, boff.data(), (n_reads + 1) * 8, cudaMemcpyHostToDevice, S.stream);
  cudaMemcpyAsync(d_trans.p, trans + trans_offsets[0], nt * 4, cudaMemcpyHostToDevice, S.stream);
  cudaMemcpyAsync(d_toff.p, toff.data(), (n_reads + 1) * 8, cudaMemcpyHostToDevice, S.stream);
  cudaMemsetAsync(d_best.p, 0xff, (size_t)n_reads * 4 * 8, S.stream);
  a.bases = (const uint8_t *)d_bases.p;
  a.base_off = (const uint64_t *)d_boff.p;
  a.trans = (const uint32_t *)d_trans.p;
  a.trans_off = (const uint64_t *)d_toff.p;
  a.best = (unsigned long long *)d_best.p;
  int32_t *o = (int32_t *)d_out.p;
  a.out_start = o;
  a.out_end = o + n_reads;
  a.out_rc = o + 2 * (size_t)n_reads;
  a.out_status = o + 3 * (size_t)n_reads;
  a.out_dist = (uint32_t *)(o + 4 * (size_t)n_reads);

  const uint32_t max_win = max_len / 2 + 2;
  for (uint32_t r0 = 0; r0 < n_reads; r0 += 65535u) {  // gridDim.y limit
    TrimArgs b = a;
    const uint32_t nr = std::min<uint32_t>(65535u, n_reads - r0);
    b.base_off = a.base_off + r0;
    b.best = a.best + (size_t)r0 * 4;
    trim_search<<<dim3((max_win + 127) / 128, nr, 4), 128, 0, S.stream>>>(b);
  }
  trim_finish<<<(n_reads + 127) / 128, 128, 0, S.stream>>>(a);
  if (cudaGetLastError() != cudaSuccess) return LVA_ERR_CUDA;
  std::vector<int32_t> host(out_words);
  cudaMemcpyAsync(host.data(), d_out.p, out_words * 4, cudaMemcpyDeviceToHost, S.stream);
  if (cudaStreamSynchronize(S.stream) != cudaSuccess) return LVA_ERR_CUDA;
  memcpy(out_start, host.data(), (size_t)n_reads * 4);
  memcpy(out_end, host.data() + n_reads, (size_t)n_reads * 4);
  memcpy(out_rc, host.data() + 2 * (size_t)n_reads, (size_t)n_reads * 4);
  memcpy(out_status, host.data() + 3 * (size_t)n_reads, (size_t)n_reads * 4);
  if (out_dist) memcpy(out_dist, host.data() + 4 * (size_t)n_reads, (size_t)n_reads * 16);
  return LVA_OK;
}
