// Fast path of the trellis step, specialised on list slots (LP) and on the number of
// message-history words of the target row (NWD).
//
// Geometry: one HBM tile of 32 predecessor conv states per CTA.  The tile feeds 128 flip targets
// (8-list merges) and 128 flop targets (2-list merges) of row p:
//   LVA_MERGED_ROLES = 1 (default): one 256-thread CTA does both, warps 0-3 the flip targets and
//     warps 4-7 the flop targets, from ONE staged copy of the tile (measured +3 % over the split:
//     20 % fewer tile bytes through L2 and less shared-memory residency per unit of work);
//   LVA_MERGED_ROLES = 0: a 128-thread CTA per role (blockIdx.x & 1), the flop CTA staging only the
//     four flip lists (first half of the tile).
//   * the predecessor tile is ONE contiguous chunk -> a single TMA bulk copy (cp.async.bulk, SASS
//     UBLKCP) into shared memory, completion on an mbarrier; the step's 40 transition scores ride
//     the same barrier;
//   * each thread owns one target state (p, c, s); the 32 lanes of a warp read 32 DIFFERENT
//     predecessor conv states, which in the tile-transposed layout are 32 different banks ->
//     no bank conflicts whatever the (crf, slot, plane) pattern;
//   * the new list is built by a flat k-way merge (one candidate popped per trip by every lane)
//     with de-duplication on the message history: a one-byte signature and a (list, slot)
//     descriptor per accepted entry are kept in byte shift registers, signature matches (about one
//     per state) are verified word by word against the staged lists; an accepted entry is stored
//     to the new list right away;
//   * states that no conv transition can produce (two of the four bases of every conv state on
//     rows entered by a one-bit step) are never touched: they keep their initial -inf;
//   * whenever two list heads tie exactly for the maximum the thread re-runs that state through
//     exact_merge(), which reproduces libstdc++'s heap order (VCC:762-795).  Without a tie the
//     pop sequence of any correct k-way merge equals the reference's, so the fast path is
//     bit-identical by construction.
#pragma once
#include "lva_kernels.cuh"

namespace lva {

#ifndef LVA_MERGED_ROLES
#define LVA_MERGED_ROLES 1
#endif
constexpr int kFastThreads = LVA_MERGED_ROLES ? 256 : 128;


// Everything a thread needs to address the lists it merges.  Staged lists -- the predecessor
// tile and the warp's own (stay) lists -- share one shape: [plane][slot] rows of 32 words.
//   NL = 8 (flip target s < 4): list i is predecessor crf i, except list s which is the state's
//                               own (stay) list; all weights come from posterior row s.
//   NL = 2 (flop target s >= 4): list 0 = own list (column s of row 4), list 1 = the flip state
//                               s-4 of the predecessor conv state (column s-4 of row 4).
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) { return __uint_as_float(lds_u32(addr)); }

// All list accesses go through 32-bit shared-window addresses computed once per thread (plain
// generic pointers made the compiler rebuild the shared window base inside the merge loop).
template <int LP, int NWD, int NL>
struct ListView {
  static constexpr uint32_t SS = 32 * 4, PS = LP * 32 * 4;  // slot / plane stride in bytes
  uint32_t srcg, mine, wrow;  // shared addresses: my predecessor column, my own list, my score row
  uint32_t list_s;            // bytes between the crf lists of the predecessor tile
  int nws, sh, s;
  uint32_t drop;

  // keep the addresses in registers: left alone, the compiler rebuilds them from the thread's
  // geometry inside the merge loop
  __device__ __forceinline__ void pin() { asm volatile("" : "+r"(srcg), "+r"(mine), "+r"(wrow), "+r"(list_s)); }
  __device__ __forceinline__ int crf_of(int i) const { return NL == 8 ? i : (i == 0 ? s : s - 4); }
  __device__ __forceinline__ bool is_stay(int i) const { return NL == 8 ? i == s : i == 0; }
  __device__ __forceinline__ uint32_t base(int i) const {
    return is_stay(i) ? mine : srcg + (uint32_t)crf_of(i) * list_s;
  }
  __device__ __forceinline__ float score(int i, int j) const {
    return lds_f32(base(i) + (uint32_t)j * SS) + lds_f32(wrow + 4u * (uint32_t)crf_of(i));
  }
  // message history entry j of list i would carry at the target (NWD words)
  __device__ __forceinline__ void history(int i, int j, uint32_t (&cand)[NWD > 0 ? NWD : 1]) const {
    if (NWD == 0) return;
    const bool st = is_stay(i);
    const uint32_t b = base(i) + (uint32_t)j * SS;
    const int sh_e = st ? 0 : sh;
    const uint32_t drop_e = st ? 0u : drop;
    const bool top_word = st || nws == NWD;  // does this list carry NWD words already?
    uint32_t old[NWD > 0 ? NWD : 1];
#pragma unroll
    for (int w = 0; w < NWD; w++) {
      old[w] = 0;
      if (w < NWD - 1 || top_word) old[w] = lds_u32(b + (1 + w) * PS);
    }
    cand[0] = (old[0] << sh_e) | drop_e;
#pragma unroll
    for (int w = 1; w < NWD; w++) cand[w] = __funnelshift_l(old[w - 1], old[w], sh_e);
  }
};

// Log of the accepted entries, kept as byte shift registers: byte k of register r describes the
// (4r+k)-th most recent accept.  `pk` holds its (list << 4 | slot) descriptor -- a signature
// match is verified by re-deriving that entry's history from the staged lists -- and `sig` a
// one-byte signature of its history whose top bit is always set, so that the zero bytes of the
// not yet filled positions never match.  Appending is two instructions per register; nothing is
// indexed dynamically.
template <int LP>
struct AcceptLog {
  static_assert(LP <= 16, "four registers cover at most 16 slots");
  static constexpr int NR = (LP + 3) / 4;
  uint32_t sig[NR], pk[NR];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int r = 0; r < NR; r++) sig[r] = pk[r] = 0;
  }
  static __device__ __forceinline__ uint32_t signature(uint32_t x) {  // byte sum of x, top bit forced
    return ((x * 0x01010101u) >> 25) | 0x80u;
  }
  __device__ __forceinline__ void push(uint32_t sig8, uint32_t d) {
#pragma unroll
    for (int r = NR - 1; r > 0; r--) {
      sig[r] = __funnelshift_l(sig[r - 1], sig[r], 8);
      pk[r] = __funnelshift_l(pk[r - 1], pk[r], 8);
    }
    sig[0] = (sig[0] << 8) | sig8;
    pk[0] = (pk[0] << 8) | d;
  }
  // bit (8k + 7 - r) is set when byte k of register r MAY equal sig8 (the zero-byte test can flag
  // the byte above a true match; callers verify)
  __device__ __forceinline__ uint32_t match(uint32_t sig8) const {
    const uint32_t bc = sig8 * 0x01010101u;
    uint32_t m = 0;
#pragma unroll
    for (int r = 0; r < NR; r++) {
      const uint32_t x = sig[r] ^ bc;
      m |= ((x - 0x01010101u) & ~x & 0x80808080u) >> r;
    }
    return m;
  }
  __device__ __forceinline__ uint32_t desc_at(int pos) const {  // pos: a bit index of match()
    uint32_t v = pk[0];
#pragma unroll
    for (int r = 1; r < NR; r++) v = (pos & 7) == 7 - r ? pk[r] : v;
    return (v >> (pos & 24)) & 0xffu;
  }
};

// arg-max over the list heads as a balanced tree (3 levels for 8 lists)
template <int NL>
__device__ __forceinline__ void select_head(const float (&h)[NL], float &best, int &idx) {
  if (NL == 2) {
    best = fmaxf(h[0], h[1]);
    idx = h[0] >= h[1] ? 0 : 1;
  } else {
    float m[4];
    int k[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      m[q] = fmaxf(h[2 * q], h[2 * q + 1]);
      k[q] = h[2 * q] >= h[2 * q + 1] ? 2 * q : 2 * q + 1;
    }
    const float m01 = fmaxf(m[0], m[1]), m23 = fmaxf(m[2], m[3]);
    const int k01 = m[0] >= m[1] ? k[0] : k[1], k23 = m[2] >= m[3] ? k[2] : k[3];
    best = fmaxf(m01, m23);
    idx = m01 >= m23 ? k01 : k23;
  }
}

// Flat k-way merge with de-duplication (VCC:743-800).  Each trip handles one popped candidate;
// the selection of the NEXT candidate (advance the popped list, arg-max over the heads) does not
// depend on the accept/duplicate decision of the current one, so both chains sit in one loop
// body and overlap.  An accepted candidate is stored to the new list in HBM right away: lanes
// accept in lock step unless one of them met a duplicate (about once per state), so the stores
// of a warp still land as complete sectors almost always.  Exact ties are detected lazily and only
// where the heap order of the reference could show in the output: an entry accepted with the same
// metric as the previously accepted one but from a different list, or a head equal to the last
// accepted metric when the loop ends.  (A tie with a candidate that is dropped as a duplicate
// cannot change the list: whichever copy of a message pops first, the same (message, metric)
// pair is kept.)
// Returns the number of accepted entries, or -1 on a tie (caller falls back to exact_merge and
// overwrites whatever was stored).
template <int LP, int NWD, int NL>
__device__ __forceinline__ int merge_lists(uint32_t *dst, const ListView<LP, NWD, NL> &V, int L, bool has_src) {
  const float NEG = neg_inf();
  float h[NL];
#pragma unroll
  for (int i = 0; i < NL; i++) h[i] = (V.is_stay(i) || has_src) ? V.score(i, 0) : NEG;
  float best;
  int i;
  select_head<NL>(h, best, i);
  if (best == NEG) return 0;
  AcceptLog<LP> log;
  log.clear();
  uint32_t ptrs = 0;  // 4-bit read position per list
  int l = 0, last_i = -1;
  float last_acc = __int_as_float(0x7f800000);  // +inf: never equals a metric
  bool tie = false;
  while (true) {
    const int j = (ptrs >> (4 * i)) & 15;
    // ---- chain A: history and signature of the current candidate ----
    uint32_t cand[NWD > 0 ? NWD : 1];
    V.history(i, j, cand);
    uint32_t csig = 0;
#pragma unroll
    for (int w = 0; w < NWD; w++) csig ^= cand[w];
    csig = AcceptLog<LP>::signature(csig);
    // ---- chain B: advance the popped list and select the next candidate ----
    float nh = NEG;
    if (j + 1 < L) {
      nh = V.score(i, j + 1);
      ptrs += 1u << (4 * i);
    }
#pragma unroll
    for (int i2 = 0; i2 < NL; i2++) h[i2] = (i2 == i) ? nh : h[i2];
    float nbest;
    int ni;
    select_head<NL>(h, nbest, ni);
    // ---- chain A continued: de-duplicate, accept ----
    bool dup;
    if (NWD == 0) {
      dup = l > 0;
    } else {
      uint32_t m = log.match(csig);
      dup = false;
      while (m != 0 && !dup) {
        const uint32_t d = log.desc_at(__ffs(m) - 1);
        m &= m - 1u;
        uint32_t acc[NWD > 0 ? NWD : 1];
        V.history((int)(d >> 4), (int)(d & 15u), acc);
        uint32_t diff = 0;
#pragma unroll
        for (int w = 0; w < NWD; w++) diff |= acc[w] ^ cand[w];
        dup = diff == 0;
      }
    }
    if (!dup) {
      tie = tie || (best == last_acc && i != last_i);
      last_acc = best;
      last_i = i;
      __stcs(dst, __float_as_uint(best));  // dst walks the slots of the new list; streaming stores: the new
                                           // list is not read again before 2-3 GB of other traffic went by
#pragma unroll
      for (int w = 0; w < NWD; w++) __stcs(dst + (1 + w) * LP * 32, cand[w]);
      dst += 32;
      log.push(csig, (uint32_t)((i << 4) | j));
      l++;
      if (l == L || NWD == 0) break;  // NWD == 0: every further candidate carries the same (empty) history
    }
    if (nbest == NEG) break;
    best = nbest;
    i = ni;
  }
#pragma unroll
  for (int i2 = 0; i2 < NL; i2++) tie = tie || (h[i2] == last_acc);
  return tie ? -1 : l;
}

template <int LP, int NWD>
__global__ void __launch_bounds__(kFastThreads) lva_step_fast(const StepArgs A, const __grid_constant__ CUtensorMap own_map) {
  constexpr int NT = kFastThreads, TG = 32;
  extern __shared__ __align__(128) uint32_t smem[];
  const CodeDev &K = c_code[A.code_slot];
  const uint32_t a = A.band_a[blockIdx.z];
  uint32_t bend = a + A.dev2;
  if (bend > K.nrows) bend = K.nrows;
  if (bend > A.row_hi) bend = A.row_hi;
  const uint32_t p = (a > A.row_lo ? a : A.row_lo) + blockIdx.y;
  if (p >= bend) return;

  const RowDev R = K.row[p];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#if LVA_MERGED_ROLES
  const uint32_t tile = blockIdx.x;
  const int cta_role = tid >> 7;  // (p == 0 only) threads 0..127 flip, 128..255 flop
  const int warp4 = warp & 3, tid7 = tid & 127;
#else
  const int cta_role = blockIdx.x & 1;
  const uint32_t tile = blockIdx.x >> 1;
  const int warp4 = warp, tid7 = tid;
#endif
  // Rows entered by a one-bit step (puncturing block 0) can produce only two of the four bases
  // per conv state: the other two flip states and their flop states have no predecessor, are
  // never reached and keep the -inf they were initialised with.  For such rows ONE CTA per tile
  // handles the 64 live flip targets (warps 0-1) and the 64 live flop targets (warps 2-3).
  const bool one_bit = p > 0 && R.shift == 1;
#if !LVA_MERGED_ROLES
  if (one_bit && cta_role == 1) return;
#endif
  // ---- which target state is mine (cheap part first: the tile copy is issued before the rest) ----
  uint32_t c, drop = 0, run, idx, lo;
  int s, src_g = -1, role, pos;  // pos: my word inside the warp's 32-word list rows
  if (p == 0) {
    role = cta_role;
    run = tid7 >> 6; idx = (tid7 >> 2) & 15u; lo = tile * 16u + idx;
    c = (run << (K.mem - 1)) | lo;
    pos = lane;
  } else if (!one_bit) {
    role = LVA_MERGED_ROLES ? (warp >> 2) : cta_role;  // 0: the 128 flip targets of the tile, 1: the 128 flop targets
    run = warp4; idx = lane >> 2; lo = tile * 8u + idx;  // a warp = 8 consecutive conv states x 4 bases
    c = (run << (K.mem - 2)) | lo;
    pos = (int)((lane & 3u) * 8u + idx);
  } else {
    role = (warp >> 1) & 1;
    const uint32_t tg = (tid & 63) >> 1;
    run = tg >> 4; idx = tg & 15u; lo = tile * 16u + idx;
    c = (run << (K.mem - 1)) | lo;
    pos = lane;
  }
  const bool valid = (c & R.vmask) == R.vval && !(LVA_MERGED_ROLES && one_bit && warp >= 4);

  uint64_t *bar = reinterpret_cast<uint64_t *>(smem);          // bytes 0..7
  float *post = reinterpret_cast<float *>(smem + 4);           // bytes 16..175
  uint32_t *src = smem + 64;                                   // byte 256
  const int nws = p > 0 ? K.row[p - 1].idwords : 0;
  // flop-only CTAs read just the flip lists 0..3 = the first half of the tile
  const int n_src_lists = (!LVA_MERGED_ROLES && p > 0 && !one_bit && role == 1) ? 4 : 8;
  uint32_t *own = src + n_src_lists * (1 + nws) * LP * TG;     // stay lists, per warp [plane][slot][32]
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (!__syncthreads_or(valid)) return;

  const int L = (int)K.L;
  uint32_t *prv = A.state + ((unsigned long long)blockIdx.z * 2 + (A.t & 1u)) * A.slot_words;
  uint32_t *cur = A.state + ((unsigned long long)blockIdx.z * 2 + ((A.t + 1u) & 1u)) * A.slot_words;

  // ---- stage ----
  constexpr int ROWS = (1 + NWD) * LP;                         // [plane][slot] rows of one list
  // Own lists of a two-bit row: warp w needs, of target tile tt(w), the 8 conv columns 8*(tile&3).. of
  // the four crf lists of its role -- a 3-D box {8 columns, 4 crf, ROWS rows} of the state buffer seen
  // as (32 columns, crf stride ROWS*128 B, 128-byte rows), which lands in shared memory as
  // [row][crf][8] = exactly the [row][pos] layout the merge reads.  One TMA copy per warp instead of
  // ROWS*8/32 16-byte cp.async per lane (which cost the LSU two L2 requests per 32-byte sector).
  const bool own_by_tma = A.own_tma != 0 && p > 0 && !one_bit;
  if (tid == 0) {
    const uint32_t tile_bytes = p > 0 ? (uint32_t)n_src_lists * (1 + nws) * LP * TG * 4u : 0u;
    const uint32_t nwarps = blockDim.x >> 5;
    mbar_expect_tx(bar, tile_bytes + 160u + (own_by_tma ? nwarps * (uint32_t)ROWS * 32u * 4u : 0u));
    bulk_g2s(post, A.post + (unsigned long long)A.post_blk[blockIdx.z] * 40, 160u, bar);
    if (p > 0) {
      const RowDev S = K.row[p - 1];
      bulk_g2s(src, prv + S.off + (unsigned long long)tile * S.tile_words, tile_bytes, bar);
    }
    if (own_by_tma) {
      const unsigned long long buf = ((unsigned long long)blockIdx.z * 2 + (A.t & 1u)) * A.slot_words + R.off;
      for (uint32_t w = 0; w < nwarps; w++) {
        const uint32_t w4 = LVA_MERGED_ROLES ? (w & 3u) : w, wrole = LVA_MERGED_ROLES ? (w >> 2) : (uint32_t)cta_role;
        const uint32_t cw = (w4 << (K.mem - 2)) | (tile * 8u);  // first conv state of the warp
        const unsigned long long row128 = (buf + (unsigned long long)(cw >> 5) * R.tile_words) >> 5;
        tensor3d_g2s(own + w * ROWS * 32, &own_map, (int)(cw & 31u), (int)(4u * wrole), (int)row128, bar);
      }
    }
  }
  // ---- the rest of the geometry, while the copies fly: my crf state and predecessor column ----
  if (p == 0) {
    s = (tid & 3) + 4 * role;
  } else if (!one_bit) {
    const uint32_t b = lane & 3u;
    const uint32_t t = b ^ emitted_base(K, R.block, lo << 2, run & 1u, run >> 1);
    const uint32_t xy = (R.inv_lin >> (2 * t)) & 3u;
    src_g = (int)((idx << 2) | xy);
    drop = (((xy & 1u) << 1) | (xy >> 1)) & R.dropmask;  // bits falling off the register, oldest first
    s = (int)b + 4 * role;
  } else {
    const uint32_t x = tid & 1u;
    s = (int)(emitted_base(K, 0, lo << 1, run, 0) ^ (x ? R.lin1 : 0u)) + 4 * role;
    src_g = (int)((idx << 1) | x);
    drop = x & R.dropmask;
  }
  uint32_t *wown = own + warp * ROWS * 32;
  const unsigned long long own_base = R.off + (unsigned long long)(c >> 5) * R.tile_words + (c & 31u) +
                                      (unsigned long long)s * ROWS * 32;
  if (own_by_tma) {
    // (issued above by thread 0)
  } else if (p > 0 && !one_bit) {
    // the warp's 8 conv states are consecutive: every (crf, plane, slot) row is 32 contiguous
    // bytes in HBM -> 16-byte copies, 8 per lane, instead of 32 4-byte ones
    const uint32_t c0 = ((uint32_t)warp4 << (K.mem - 2)) | (tile * 8u);
    const unsigned long long wbase = R.off + (unsigned long long)(c0 >> 5) * R.tile_words + (c0 & 31u);
#pragma unroll
    for (int k = 0; k < (ROWS * 8 + 31) / 32; k++) {
      const int q = lane + 32 * k;
      const int rowi = q >> 3, piece = (q >> 1) & 3, half = q & 1;
      if (ROWS % 4 == 0 || q < ROWS * 8)
        cp_async16(wown + rowi * 32 + piece * 8 + half * 4,
                   prv + wbase + ((unsigned long long)(piece + 4 * role) * ROWS + rowi) * 32 + half * 4);
    }
  } else if (valid) {
#pragma unroll
    for (int r = 0; r < ROWS; r++) cp_async4(wown + r * 32 + pos, prv + own_base + (unsigned long long)r * 32);
  }
  cp_async_wait_all();
  __syncwarp();
  mbar_wait(bar, 0);
  if (!valid) return;

  const uint32_t *mine = wown + pos;
  uint32_t *dst = cur + own_base;
  const int row = s >= 4 ? 4 : s;
  const float *wrow = post + row * 8;
  if (p == 0) {  // VCC:706-713: stay edge only, slot 0 only
    dst[0] = __float_as_uint(__uint_as_float(mine[0]) + wrow[s]);
#pragma unroll
    for (int l = 1; l < LP; l++)
      if (l < L) dst[l * 32] = 0xff800000u;
    return;
  }
  const uint32_t *srcg = src + src_g;
  if (LP == 1) {
    // VCC:715-742: first strict maximum, own list first then predecessor crf ascending
    float best = neg_inf();
    int bk = -1;
    const float s0 = __uint_as_float(mine[0]) + wrow[s];
    if (s0 > best) { best = s0; bk = s; }
    const int list_s = (1 + nws) * TG;
    const int k0 = s < 4 ? 0 : s - 4, k1 = s < 4 ? 7 : s - 4;
    for (int k = k0; k <= k1; k++) {
      if (k == s) continue;
      const float sc = __uint_as_float(srcg[k * list_s]) + wrow[k];
      if (sc > best) { best = sc; bk = k; }
    }
    dst[0] = __float_as_uint(best);
    if (A.dec != nullptr) {
      // survivor mode: which predecessor won, 4 bits per state (crf of the source; == s means
      // "stayed"), the 8 states of a conv state packed into one word: flips low half, flops high
      const uint32_t code = (uint32_t)(bk < 0 ? s : bk) & 7u;
      uint32_t v = code << (4 * (s & 3));
      uint32_t *word = A.dec + (unsigned long long)blockIdx.z * A.dec_slot_words +
                       ((unsigned long long)A.t * A.dec_rows + (p - a)) * K.nconv + c;
      unsigned short *half = reinterpret_cast<unsigned short *>(word) + role;
      if (!one_bit) {  // the 4 lanes lane&~3 .. +3 hold the 4 bases of one conv state
        const uint32_t gm = 0xFu << (lane & ~3);
        v |= __shfl_xor_sync(gm, v, 1);
        v |= __shfl_xor_sync(gm, v, 2);
        if ((lane & 3) == 0) *half = (unsigned short)v;
      } else {         // lanes lane&~1, +1 hold the two producible bases of one conv state
        const uint32_t gm = 0x3u << (lane & ~1);
        v |= __shfl_xor_sync(gm, v, 1);
        if ((lane & 1) == 0) *half = (unsigned short)v;
      }
    }
    if (NWD > 0) {
      if (bk < 0 || bk == s) {  // (metric -inf: history is don't-care)
#pragma unroll
        for (int w = 0; w < NWD; w++) dst[(1 + w) * 32] = mine[(1 + w) * 32];
      } else {
        uint32_t prev = 0;
#pragma unroll
        for (int w = 0; w < NWD; w++) {
          uint32_t old = 0;
          if (w < nws) old = srcg[bk * list_s + (1 + w) * TG];
          dst[(1 + w) * 32] = (old << R.shift) | (w == 0 ? drop : (prev >> (32 - R.shift)));
          prev = old;
        }
      }
    }
    return;
  }
  int n_acc;
  if (role == 0) {
    ListView<LP, NWD, 8> V{smem_u32(srcg), smem_u32(mine), smem_u32(wrow), (uint32_t)(1 + nws) * LP * TG * 4u,
                           nws, (int)R.shift, s, drop};
    V.pin();
    n_acc = merge_lists<LP, NWD, 8>(dst, V, L, true);
  } else {
    ListView<LP, NWD, 2> V{smem_u32(srcg), smem_u32(mine), smem_u32(wrow), (uint32_t)(1 + nws) * LP * TG * 4u,
                           nws, (int)R.shift, s, drop};
    V.pin();
    n_acc = merge_lists<LP, NWD, 2>(dst, V, L, true);
  }
  if (n_acc < 0) {  // exact score tie between two list heads: libstdc++ heap order decides
    uint32_t *mine_rw = wown + pos;
    exact_merge<LP>(src, mine_rw, post, L, nws, NWD, TG, 32, R.shift, drop, s, src_g);
    atomicAdd(A.counters, 1ull);
    for (int l = 0; l < L; l++)
      for (int w = 0; w < 1 + NWD; w++) dst[(w * LP + l) * 32] = mine_rw[(w * LP + l) * 32];
  } else {
    for (int l = n_acc; l < L; l++) __stcs(dst + l * 32, 0xff800000u);  // -inf; history of a dead slot is never read
  }
}

// ---------------------------------------------------------------------------------------------
// lva_step_l1: the whole step for list size 1 in survivor mode (metrics only, VCC:715-742), the
// reference tool's default.  A list-size-1 state is one float: staging tiles through shared memory
// and one thread per product state (lva_step_fast<1,0>) spend ten times more instructions on
// geometry than on the update.  Here one thread owns a conv state (p, c): it reads the 8 metrics of
// the state itself and of each predecessor conv state straight from L2 (the band state of every
// read in flight fits there), updates the up to 8 crf states in the reference's order (stay first,
// then predecessor crf ascending, first strict maximum) and writes the 8 four-bit decisions as ONE
// 32-bit word, coalesced over c.  grid = (band rows x conv states / 256, 1, reads); the per-read
// control data (band start, posterior block) comes from the tables the traceback uses.
// ---------------------------------------------------------------------------------------------
struct L1Args {  // small on purpose: 800+ launches per batch, the launch itself is what costs
  uint32_t *state;
  unsigned long long slot_words;
  const float *post;
  uint32_t *dec;
  unsigned long long dec_slot_words;
  const ReadDev *reads;   // per slot: first posterior block, T, offset of its band table
  const uint16_t *band;   // band starts a_t of every read in flight (VCC:677-678)
  uint32_t dec_rows, t, dev2, code_slot;
};

__global__ void __launch_bounds__(256) lva_step_l1(const L1Args A) {
  __shared__ float post[40];
  const CodeDev &K = c_code[A.code_slot];
  const uint32_t z = blockIdx.z;
  const ReadDev rd = A.reads[z];
  const uint32_t a = A.band[rd.band_off + A.t];
  uint32_t bend = a + A.dev2;
  if (bend > K.nrows) bend = K.nrows;
  if (threadIdx.x < 40) post[threadIdx.x] = A.post[(rd.post_block + A.t) * 40 + threadIdx.x];
  __syncthreads();
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;  // (band row, conv state), conv state fastest
  const uint32_t p = a + idx / K.nconv, c = idx & (K.nconv - 1u);
  if (p >= bend) return;
  const RowDev R = K.row[p];
  if ((c & R.vmask) != R.vval) return;
  const uint32_t *prv = A.state + ((unsigned long long)z * 2 + (A.t & 1u)) * A.slot_words;
  uint32_t *cur = A.state + ((unsigned long long)z * 2 + ((A.t + 1u) & 1u)) * A.slot_words;
  const unsigned long long own = R.off + (unsigned long long)(c >> 5) * R.tile_words + (c & 31u);  // + s * 32
  float mine[8];
#pragma unroll
  for (int s = 0; s < 8; s++) mine[s] = __uint_as_float(prv[own + s * 32]);
  if (p == 0) {  // VCC:706-713: stay edge only
#pragma unroll
    for (int s = 0; s < 8; s++) cur[own + s * 32] = __float_as_uint(mine[s] + post[(s >= 4 ? 4 : s) * 8 + s]);
    return;
  }
  const RowDev S = K.row[p - 1];
  const uint32_t m = K.mem, sh = R.shift;
  const uint32_t run = c >> (m - sh), lo = c & ((1u << (m - sh)) - 1u);
  const uint32_t e0 = sh == 2 ? emitted_base(K, R.block, lo << 2, run & 1u, run >> 1) : emitted_base(K, 0, lo << 1, run, 0);
  uint32_t word = 0;
  const int nb = sh == 2 ? 4 : 2;  // producible bases of this conv state
  for (int q = 0; q < nb; q++) {
    uint32_t b, cp;  // the base, and the predecessor conv state that emits it on the way to c
    if (sh == 2) {
      b = (uint32_t)q;
      cp = (lo << 2) | ((R.inv_lin >> (2 * (b ^ e0))) & 3u);
    } else {
      b = e0 ^ (q ? R.lin1 : 0u);
      cp = (lo << 1) | (uint32_t)q;
    }
    const unsigned long long srcb = S.off + (unsigned long long)(cp >> 5) * S.tile_words + (cp & 31u);
    float src[8];
#pragma unroll
    for (int k = 0; k < 8; k++) src[k] = __uint_as_float(prv[srcb + k * 32]);
    {  // flip state b: stay, then every other crf state of the predecessor
      float best = neg_inf();
      int bk = -1;
      const float s0 = mine[b] + post[b * 8 + b];
      if (s0 > best) { best = s0; bk = (int)b; }
#pragma unroll
      for (int k = 0; k < 8; k++) {
        if (k == (int)b) continue;
        const float sc = src[k] + post[b * 8 + k];
        if (sc > best) { best = sc; bk = k; }
      }
      cur[own + b * 32] = __float_as_uint(best);
      word |= ((uint32_t)(bk < 0 ? (int)b : bk) & 7u) << (4 * b);
    }
    {  // flop state b+4: stay, then the flip state b of the predecessor
      const uint32_t s = b + 4;
      float best = neg_inf();
      int bk = -1;
      const float s0 = mine[s] + post[32 + s];
      if (s0 > best) { best = s0; bk = (int)s; }
      const float sc = src[b] + post[32 + b];
      if (sc > best) { best = sc; bk = (int)b; }
      cur[own + s * 32] = __float_as_uint(best);
      word |= ((uint32_t)(bk < 0 ? (int)s : bk) & 7u) << (4 * s);
    }
  }
  A.dec[(unsigned long long)z * A.dec_slot_words + ((unsigned long long)A.t * A.dec_rows + (p - a)) * K.nconv + c] = word;
}

inline cudaError_t l1_launch_step(cudaStream_t st, const L1Args &a, const CodeSpec &k, uint32_t active) {
  const uint32_t rows = a.dev2 < k.nrows ? a.dev2 : k.nrows;  // band rows of one read at one step
  lva_step_l1<<<dim3((rows * k.nconv + 255) / 256, 1, active), 256, 0, st>>>(a);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// host-side dispatch
// ---------------------------------------------------------------------------------------------
inline size_t fast_smem_bytes(uint32_t LP, uint32_t nwd, uint32_t threads = kFastThreads) {
  // predecessors carrying as many words as the target row: tile + own (stay) lists of every thread
  return 256 + ((size_t)8 * (1 + nwd) * LP * 32 + (size_t)(1 + nwd) * LP * threads) * 4;
}

#define LVA_FAST_LP_CASES(X) X(1) X(2) X(4) X(8) X(16)
#define LVA_FAST_NW_CASES(X, lp) X(lp, 0) X(lp, 1) X(lp, 2) X(lp, 3) X(lp, 4) X(lp, 5) X(lp, 6) X(lp, 7) X(lp, 8)

inline bool fast_supported(uint32_t LP, uint32_t nw_last) {
  if (LP > 16) return false;
  return fast_smem_bytes(LP, nw_last) <= 227 * 1024;
}

template <int LP, int NWD>
inline cudaError_t fast_prepare_one() {
  return cudaFuncSetAttribute(lva_step_fast<LP, NWD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)fast_smem_bytes(LP, NWD));
}

inline cudaError_t fast_prepare(uint32_t LP, uint32_t nw_last) {
  cudaError_t e = cudaSuccess;
#define Y(lp, nw) \
  if (LP == lp && nw <= (int)nw_last && e == cudaSuccess) e = fast_prepare_one<lp, nw>();
#define X(lp) LVA_FAST_NW_CASES(Y, lp)
  LVA_FAST_LP_CASES(X)
#undef X
#undef Y
  return e;
}

template <int LP, int NWD>
inline cudaError_t fast_launch_one(cudaStream_t st, dim3 grid, uint32_t threads, const StepArgs &a, const CUtensorMap &own_map) {
  lva_step_fast<LP, NWD><<<grid, threads, fast_smem_bytes(LP, NWD, threads), st>>>(a, own_map);
  return cudaGetLastError();
}

// One time step: one launch per distinct history width among rows [lo, hi).
inline cudaError_t fast_launch_step(uint32_t LP, cudaStream_t st, StepArgs a, const CodeSpec &k, uint32_t lo,
                                    uint32_t hi, uint32_t active, uint64_t *n_launches, const CUtensorMap *own_maps) {
  uint32_t p = lo;
  while (p < hi) {
    const uint32_t nw = k.idwords[p];
    uint32_t q = p;
    while (q < hi && k.idwords[q] == nw) q++;
    a.row_lo = p;
    a.row_hi = q;
    dim3 grid((LVA_MERGED_ROLES ? 1 : 2) * (k.nconv / 32), q - p, active);  // x: (tile, role)
    // rows entered by a one-bit step have 64 + 64 live targets per tile (warps 0-3): when a launch
    // holds nothing else (every row of a rate-1/2 code), half the merged CTA would sit idle
    uint32_t threads = kFastThreads;
    if (LVA_MERGED_ROLES && p > 0) {
      bool all_one_bit = true;
      for (uint32_t r = p; r < q; r++) all_one_bit = all_one_bit && k.shift[r] == 1;
      if (all_one_bit) threads = 128;
    }
    cudaError_t e = cudaErrorInvalidValue;
#define Y(lp, nwc) \
  if (LP == lp && nw == nwc) e = fast_launch_one<lp, nwc>(st, grid, threads, a, own_maps[nwc]);
#define X(lp) LVA_FAST_NW_CASES(Y, lp)
    LVA_FAST_LP_CASES(X)
#undef X
#undef Y
    if (e != cudaSuccess) return e;
    (*n_launches)++;
    p = q;
  }
  return cudaSuccess;
}

}  // namespace lva
