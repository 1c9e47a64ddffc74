// Trellis step for list sizes 17..64 (32 or 64 list slots): the reference takes any -l (VCC:743-800) and
// the paper runs 64 for m = 8.
//
// Same geometry as lva_step_fast (lva_fast.cuh): one tile of 32 predecessor conv states per CTA, one target
// state per thread, warps 0-3 the flip targets, warps 4-7 the flop targets.  What differs is where the lists
// are read: at these sizes a tile is 96 / 192 KB and the own lists as much again -- staging them would leave
// one CTA per SM -- while a k-way merge that stops after L distinct messages touches about a quarter of the
// entries.  So nothing is staged but the 40 transition scores: every candidate is read where it lies in HBM
// when the merge gets to it (ld.global.nc; the lists are read-only during the launch and the lanes of a warp
// sit in neighbouring 32-byte sectors) and the SM runs as many warps as the registers allow, which is what
// hides the latency of those loads.
//
// The merge is the pop-by-pop k-way merge of merge_lists() (arg-max tree over the list heads, lazy tie
// detection, exact_merge() as the fallback) with one change: the accept log keeps one-byte signatures only, and a
// signature match is verified against the entry it points at in the NEW list, which the thread itself has just
// written (program order; plain loads, mostly L1 hits).  Measured on the paper's m = 8, L = 64 configuration:
// 7.6 reads/s against 1.6 through lva_step_generic (64 registers, 4 CTAs per SM); the verification of true duplicates (about every third pop
// at this list size, two lanes of a warp at a time) is what the launch waits for (profiles/README.md).  A 32-bit
// hash screen in shared memory in front of the verification was measured and dropped (4.8 reads/s).
#pragma once
#include "lva_fast.cuh"

namespace lva {

constexpr int kWideThreads = 256;

template <int LP, int NL>
struct WideView {
  static constexpr uint32_t GP = 64u * LP;  // words between the planes of a piece, of an own list and of a nid piece alike
  const uint32_t *gsrc;     // my predecessor column: metric of (crf 0, slot 0); crf stride LP*8 words, slot stride 8
  const uint32_t *gsrc_id;  // ... its chain-id plane: gsrc + GP, or the source row's nid plane on post-boundary rows
  const uint32_t *gown;     // my own list in the source parity buffer: metric of slot 0; slot stride 8
  uint32_t wrow;            // shared address of my score row
  uint32_t keep, app;       // identity of a move candidate: (id, (partial & keep) | app)
  int s;

  __device__ __forceinline__ int crf_of(int i) const { return NL == 8 ? i : (i == 0 ? s : s - 4); }
  __device__ __forceinline__ bool is_stay(int i) const { return NL == 8 ? i == s : i == 0; }
  __device__ __forceinline__ float score(int i, int j) const {
    const uint32_t crf = (uint32_t)crf_of(i);
    const float w = lds_f32(wrow + 4u * crf);
    const uint32_t m = is_stay(i) ? __ldg(gown + (uint32_t)j * 8u) : __ldg(gsrc + (crf * LP + (uint32_t)j) * 8u);
    return __uint_as_float(m) + w;
  }
  __device__ __forceinline__ void identity(int i, int j, uint32_t &id, uint32_t &pt) const {
    if (is_stay(i)) {
      id = __ldg(gown + GP + (uint32_t)j * 8u);
      pt = __ldg(gown + 2 * GP + (uint32_t)j * 8u);
    } else {
      const uint32_t o = ((uint32_t)crf_of(i) * LP + (uint32_t)j) * 8u;
      id = __ldg(gsrc_id + o);
      pt = (__ldg(gsrc + 2 * GP + o) & keep) | app;  // keep == 0 on post-boundary rows: the load is masked out
    }
  }
};

// One-byte signatures of the accepted entries, four per register, byte k of register r = the (4r+k)-th most
// recent accept (top bit always set: empty bytes match nothing).
template <int LP>
struct SigLog {
  static constexpr int NR = LP / 4;
  uint32_t sig[NR];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int r = 0; r < NR; r++) sig[r] = 0;
  }
  static __device__ __forceinline__ uint32_t signature(uint32_t x) { return ((x * 0x01010101u) >> 25) | 0x80u; }
  __device__ __forceinline__ void push(uint32_t sig8) {
#pragma unroll
    for (int r = NR - 1; r > 0; r--) sig[r] = __funnelshift_l(sig[r - 1], sig[r], 8);
    sig[0] = (sig[0] << 8) | sig8;
  }
};

// Rows without any history bit: the new list is the best head alone (see merge_first_only()).
template <int LP, int NL>
__device__ __forceinline__ int wide_first_only(uint32_t *dst, const WideView<LP, NL> &V) {
  constexpr uint32_t GP = 64u * LP;
  float h[NL];
#pragma unroll
  for (int i = 0; i < NL; i++) h[i] = V.score(i, 0);
  float best;
  int i;
  select_head<NL>(h, best, i);
  if (best == neg_inf()) return 0;
  bool tie = false;
#pragma unroll
  for (int i2 = 0; i2 < NL; i2++) tie = tie || (i2 != i && h[i2] == best);
  uint32_t cid, cpt;
  V.identity(i, 0, cid, cpt);
  dst[0] = __float_as_uint(best);
  dst[GP] = cid;
  dst[2 * GP] = cpt;
  return tie ? -1 : 1;
}

// Pop-by-pop k-way merge with de-duplication (VCC:743-800); returns the number of accepted entries or -1 on a
// tie the heap order of the reference could decide (the same rule as merge_lists()).
template <int LP, int NL>
__device__ __forceinline__ int merge_wide(uint32_t *dst, const WideView<LP, NL> &V, int L) {
  constexpr uint32_t GP = 64u * LP;
  const float NEG = neg_inf();
  float h[NL];
#pragma unroll
  for (int i = 0; i < NL; i++) h[i] = V.score(i, 0);
  float best;
  int i;
  select_head<NL>(h, best, i);
  if (best == NEG) return 0;
  SigLog<LP> log;
  log.clear();
  unsigned long long ptrs = 0;  // one byte of read position per list
  int l = 0, last_i = -1;
  float last_acc = __int_as_float(0x7f800000);  // +inf: never equals a metric
  bool tie = false;
#pragma unroll 1
  while (true) {
    const int j = (int)((ptrs >> (8 * i)) & 255ull);
    uint32_t cid, cpt;
    V.identity(i, j, cid, cpt);
    const uint32_t csig = SigLog<LP>::signature(cid ^ cpt);
    float nh = NEG;
    if (j + 1 < L) {
      nh = V.score(i, j + 1);
      ptrs += 1ull << (8 * i);
    }
#pragma unroll
    for (int i2 = 0; i2 < NL; i2++) h[i2] = (i2 == i) ? nh : h[i2];
    float nbest;
    int ni;
    select_head<NL>(h, nbest, ni);
    tie = tie || (best == last_acc && i != last_i);
    bool dup = false;
    const uint32_t bc = csig * 0x01010101u;
#pragma unroll
    for (int r = 0; r < SigLog<LP>::NR; r++) {
      if (4 * r < l && !dup) {
        const uint32_t x = log.sig[r] ^ bc;
        uint32_t m = (x - 0x01010101u) & ~x & 0x80808080u;  // may also flag the byte above a true match: verified below
        while (m != 0u && !dup) {
          const int k = 4 * r + ((__ffs((int)m) - 1) >> 3);  // k-th most recent accept = slot l-1-k of the new list
          m &= m - 1u;
          if (k < l) {
            const uint32_t *acc = dst + (uint32_t)(l - 1 - k) * 8u;
            dup = acc[GP] == cid && acc[2 * GP] == cpt;  // the thread's own stores
          }
        }
      }
    }
    if (!dup) {
      last_acc = best;
      last_i = i;
      uint32_t *d = dst + (uint32_t)l * 8u;
      d[0] = __float_as_uint(best);
      d[GP] = cid;
      d[2 * GP] = cpt;
      log.push(csig);
      l++;
      if (l == L) break;
    }
    if (nbest == NEG) break;
    best = nbest;
    i = ni;
  }
#pragma unroll
  for (int i2 = 0; i2 < NL; i2++) tie = tie || (h[i2] == last_acc);
  return tie ? -1 : l;
}

template <int LP>
__global__ void __launch_bounds__(kWideThreads, 4) lva_step_wide(const StepArgs A) {
  constexpr uint32_t GP = 64u * LP;
  extern __shared__ __align__(16) uint32_t wsm[];
  float *post = reinterpret_cast<float *>(wsm);               // 40 transition scores (64 words reserved)
  uint32_t *scratch = wsm + 64;                               // one list per warp for the exact fallback
  const CodeDev &K = c_code[A.code_slot];
  const uint32_t z = blockIdx.z;
  const uint32_t a = A.band_a[z];
  uint32_t bend = a + A.dev2;
  if (bend > K.nrows) bend = K.nrows;
  const uint32_t p = a + blockIdx.y;
  if (p >= bend) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t tile = blockIdx.x;
  if (p == 0) {
    if (tile == (K.init >> 5) && tid < 8) row0_update(A, K, z, tid);
    return;
  }
  const RowDev R = K.row[p];
  uint32_t wp = A.band_slot[z] + blockIdx.y;
  if (wp >= K.window) wp -= K.window;
  const uint32_t wp1 = wp == 0 ? K.window - 1 : wp - 1;
  const bool one_bit = R.shift == 1;
  // ---- which target state is mine: exactly the mapping of lva_step_fast ----
  uint32_t c, run, idx, lo, xy;
  int role, s, src_g;
  if (!one_bit) {
    role = warp >> 2;
    run = warp & 3; idx = lane >> 2; lo = tile * 8u + idx;
    c = (run << (K.mem - 2)) | lo;
    const uint32_t b = lane & 3u;
    const uint32_t t = b ^ emitted_base(K, R.block, lo << 2, run & 1u, run >> 1);
    xy = (R.inv_lin >> (2 * t)) & 3u;
    src_g = (int)((idx << 2) | xy);
    s = (int)b + 4 * role;
  } else {
    role = (warp >> 1) & 1;
    const uint32_t tg = (tid & 63) >> 1;
    run = tg >> 4; idx = tg & 15u; lo = tile * 16u + idx;
    c = (run << (K.mem - 1)) | lo;
    xy = tid & 1u;
    const uint32_t b = emitted_base(K, 0, lo << 1, run, 0) ^ (xy ? R.lin1 : 0u);
    s = (int)b + 4 * role;
    src_g = (int)((idx << 1) | xy);
  }
  const bool valid = (c & R.vmask) == R.vval && !(one_bit && warp >= 4);
  if (tid < 40) post[tid] = A.post[(unsigned long long)A.post_blk[z] * 40 + tid];
  if (!__syncthreads_or(valid)) return;
  if (!valid) return;

  const uint32_t prv_row0 = (z * 2u + (A.t & 1u)) * K.window, cur_row0 = (z * 2u + ((A.t + 1u) & 1u)) * K.window;
  const uint32_t own_off = (c >> 3) * A.piece_words + (uint32_t)s * LP * 8 + (c & 7u);
  uint32_t *dst = A.state + (unsigned long long)(cur_row0 + wp) * A.row_words + own_off;
  const uint32_t *gown = A.state + (unsigned long long)(prv_row0 + wp) * A.row_words + own_off;
  // my predecessor column: conv state 32*tile + src_g of row p-1
  const uint32_t src_piece = tile * 4u + (uint32_t)(src_g >> 3);
  const uint32_t *gsrc = A.state + (unsigned long long)(prv_row0 + wp1) * A.row_words + (unsigned long long)src_piece * A.piece_words +
                         (uint32_t)(src_g & 7);
  const uint32_t *gsrc_id = R.post_boundary ? A.nid + nid_base(A, K, z, A.t & 1u, R.src_nid_slot) + (unsigned long long)src_piece * GP +
                                                  (uint32_t)(src_g & 7)
                                            : gsrc + GP;
  const uint32_t appv = appended_bits(xy, R.shift, R.app) << R.app_pos;
  const uint32_t keep = R.post_boundary ? 0u : 0xffffffffu;
  const int L = (int)K.L;
  const int row = s >= 4 ? 4 : s;
  const uint32_t wrow = smem_u32(post + row * 8);

  int n_acc;
  if (role == 0) {
    WideView<LP, 8> V{gsrc, gsrc_id, gown, wrow, keep, appv, s};
    n_acc = R.nohist ? wide_first_only<LP, 8>(dst, V) : merge_wide<LP, 8>(dst, V, L);
  } else {
    WideView<LP, 2> V{gsrc, gsrc_id, gown, wrow, keep, appv, s};
    n_acc = R.nohist ? wide_first_only<LP, 2>(dst, V) : merge_wide<LP, 2>(dst, V, L);
  }

  uint32_t *nidw = nullptr;
  unsigned long long *tab = nullptr;
  if (R.boundary) {
    nidw = A.nid + nid_base(A, K, z, (A.t + 1u) & 1u, R.nid_slot) + (unsigned long long)(c >> 3) * GP + (unsigned long long)s * LP * 8 + (c & 7u);
    tab = A.tab + (unsigned long long)z * ((unsigned long long)A.tab_mask + 1ull);
  }
  // metrics tie where libstdc++'s heap order could show: redo the state exactly, lane by lane through the warp's scratch list
  {
    uint32_t *scr = scratch + warp * 3 * LP;
    const unsigned grp = __activemask();
    unsigned need = __ballot_sync(grp, n_acc < 0);
    while (need != 0u) {
      if (lane == __ffs((int)need) - 1) {
        for (int l = 0; l < L; l++) {
          scr[l] = __ldg(gown + l * 8);
          scr[LP + l] = __ldg(gown + GP + l * 8);
          scr[2 * LP + l] = __ldg(gown + 2 * GP + l * 8);
        }
        n_acc = exact_merge<LP>(gsrc, scr, post, L, 8, 1, keep, appv, s, 0, gsrc_id);
        atomicAdd(A.counters, 1ull);
        for (int l = 0; l < L; l++) {
          dst[l * 8] = scr[l];
          dst[GP + l * 8] = scr[LP + l];
          dst[2 * GP + l * 8] = scr[2 * LP + l];
        }
      }
      need &= need - 1u;
      __syncwarp(grp);
    }
  }
  for (int l = n_acc; l < L; l++) dst[l * 8] = 0xff800000u;  // -inf; identity of a dead slot is never read
  if (R.boundary)  // publish the interned id of every new entry for the row above (insert-or-get: canonical either way)
    for (int l = 0; l < n_acc; l++) nidw[l * 8] = intern_chunk(tab, A.tab_mask, dst[GP + l * 8], dst[2 * GP + l * 8], A.counters, z);
}

inline bool wide_supported(uint32_t LP) { return LP == 32 || LP == 64; }
inline bool wide_possible(uint32_t LP) { return LP == 8 || LP == 16 || wide_supported(LP); }  // 8, 16: developer A/B against lva_step_fast
inline size_t wide_smem_bytes(uint32_t LP) { return (64 + (size_t)(kWideThreads / 32) * 3 * LP) * 4; }
inline cudaError_t wide_prepare(uint32_t LP) {
  if (LP == 8) return cudaFuncSetAttribute(lva_step_wide<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide_smem_bytes(8));
  if (LP == 16) return cudaFuncSetAttribute(lva_step_wide<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide_smem_bytes(16));
  if (LP == 32) return cudaFuncSetAttribute(lva_step_wide<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide_smem_bytes(32));
  if (LP == 64) return cudaFuncSetAttribute(lva_step_wide<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide_smem_bytes(64));
  return cudaErrorInvalidValue;
}

inline cudaError_t wide_launch_step(uint32_t LP, cudaStream_t st, const StepArgs &a, const CodeSpec &k, uint32_t span, uint32_t active,
                                    bool all_one_bit) {
  dim3 grid(k.nconv / 32, span, active);
  const uint32_t threads = all_one_bit ? 128 : kWideThreads;
  if (LP == 8) lva_step_wide<8><<<grid, threads, wide_smem_bytes(8), st>>>(a);
  else if (LP == 16) lva_step_wide<16><<<grid, threads, wide_smem_bytes(16), st>>>(a);
  else if (LP == 32) lva_step_wide<32><<<grid, threads, wide_smem_bytes(32), st>>>(a);
  else if (LP == 64) lva_step_wide<64><<<grid, threads, wide_smem_bytes(64), st>>>(a);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

}  // namespace lva
