// Fast path of the trellis step, specialised on list slots (LP) and on the number of
// message-history words of the target row (NWD).
//
// Geometry: one HBM tile of 32 predecessor conv states per CTA, 128 threads, and a ROLE per CTA
// (blockIdx.x & 1): role 0 computes the 128 flip targets fed by the tile, role 1 the 128 flop
// targets.  Flip targets merge 8 lists, flop targets 2, so giving them separate CTAs lets the
// short ones retire and free their SM slot instead of idling beside the long ones.
//   * only the PLANES region of the predecessor tile (metric + history hash, 16 KB whatever the
//     history width; flop CTAs need just its first half, the four flip lists) is staged: ONE
//     contiguous TMA bulk copy (cp.async.bulk, SASS UBLKCP) into shared memory, completion on an
//     mbarrier; the step's 40 transition scores ride the same barrier.  ~25 KB of shared memory
//     per CTA -> 9 CTAs / 36 warps per SM for every row of the trellis;
//   * each thread owns one target state (p, c, s); the 32 lanes of a warp read 32 DIFFERENT
//     predecessor conv states, which in the tile-transposed layout are 32 different banks ->
//     no bank conflicts whatever the (crf, slot, plane) pattern;
//   * the new list is built by a flat k-way merge (one candidate popped per trip by every lane).
//     De-duplication works on the stored 32-bit history hash: a one-byte signature per accepted
//     entry is kept in registers, signature matches are checked against the full hash; equal
//     hashes are PRESUMED duplicates and verified word by word against the histories in HBM, off
//     the critical path (loads issued at detection, compared one event later);
//   * history words never enter shared memory: when an entry is accepted its predecessor's words
//     are loaded from HBM (one sector), extended by the bits leaving the register, and stored to
//     the new list one accept later, so the merge never waits for them;
//   * whenever two list heads tie exactly for the maximum, or a presumed duplicate fails its
//     verification, the thread re-runs that state through exact_merge(), which reproduces
//     libstdc++'s heap order (VCC:762-795).  Without a tie the pop sequence of any correct k-way
//     merge equals the reference's, so the fast path is bit-identical by construction;
//   * states that no conv transition can produce (two of the four bases of every conv state on
//     rows entered by a one-bit step) are never touched: they keep their initial -inf.
#pragma once
#include "lva_kernels.cuh"

namespace lva {

constexpr int kFastThreads = 128;

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) { return __uint_as_float(lds_u32(addr)); }
__device__ __forceinline__ uint32_t ldg_nc(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// Everything a thread needs to address the lists it merges.  Shared-memory accesses go through
// 32-bit shared-window addresses computed once per thread.
//   NL = 8 (flip target s < 4): list i is predecessor crf i, except list s which is the state's
//                               own (stay) list; all weights come from posterior row s.
//   NL = 2 (flop target s >= 4): list 0 = own list (column s of row 4), list 1 = the flip state
//                               s-4 of the predecessor conv state (column s-4 of row 4).
template <int LP, int NWD, int NL>
struct ListView {
  static constexpr uint32_t SS = 32 * 4, PS = LP * 32 * 4;  // slot / plane stride in bytes
  uint32_t srcg, mine, wrow;  // shared addresses: my predecessor column, my own list, my score row
  uint32_t list_s;            // bytes between the crf lists of the staged predecessor planes
  const uint32_t *gsrc;       // HBM: history of entry (crf k, slot j) of my predecessor at gsrc[((k*LP+j)*32)*nws + w]
  const uint32_t *gown;       // HBM: history of entry j of my own previous list at gown[j*32*NWD + w]
  int nws, sh, s;
  uint32_t drop;

  __device__ __forceinline__ int crf_of(int i) const { return NL == 8 ? i : (i == 0 ? s : s - 4); }
  __device__ __forceinline__ bool is_stay(int i) const { return NL == 8 ? i == s : i == 0; }
  __device__ __forceinline__ uint32_t base(int i) const {
    return is_stay(i) ? mine : srcg + (uint32_t)crf_of(i) * list_s;
  }
  __device__ __forceinline__ float score(int i, int j) const {
    return lds_f32(base(i) + (uint32_t)j * SS) + lds_f32(wrow + 4u * (uint32_t)crf_of(i));
  }
  // hash of the history entry j of list i would carry at the target
  __device__ __forceinline__ uint32_t hash(int i, int j) const {
    const bool st = is_stay(i);
    uint32_t h = 0;
    if (st || nws > 0) h = lds_u32(base(i) + PS + (uint32_t)j * SS);
    return st ? h : mix_hash(h, drop);
  }
};

// History words of one list entry on their way from HBM: loaded when the entry is accepted (or a
// presumed duplicate needs verifying), finished (extended by the bits that left the register) when
// they are consumed one event later.
template <int NWD>
struct Fetched {
  uint32_t old[NWD > 0 ? NWD : 1];
  int sh_e;
  uint32_t drop_e;
  template <int LP, int NL>
  __device__ __forceinline__ void load(const ListView<LP, NWD, NL> &V, int i, int j) {
    const bool st = V.is_stay(i);
    const uint32_t *p = st ? V.gown + j * 32 * NWD : V.gsrc + ((V.crf_of(i) * LP + j) * 32) * V.nws;
    const int n = st ? NWD : V.nws;
#pragma unroll
    for (int w = 0; w < NWD; w++) {
      old[w] = 0;
      if (w < NWD - 1 || n == NWD) old[w] = ldg_nc(p + w);
    }
    sh_e = st ? 0 : V.sh;
    drop_e = st ? 0u : V.drop;
  }
  __device__ __forceinline__ void finish(uint32_t (&out)[NWD > 0 ? NWD : 1]) const {
    out[0] = (old[0] << sh_e) | drop_e;
#pragma unroll
    for (int w = 1; w < NWD; w++) out[w] = __funnelshift_l(old[w - 1], old[w], sh_e);
  }
};

// Accepted entries are remembered as (list, slot) descriptors, 8 bits each.
template <int LP>
struct Picks {
  static_assert(LP <= 16, "descriptor words cover at most 16 slots");
  uint32_t w0, w1, w2, w3;  // scalars on purpose: an indexed array would be demoted to local memory
  __device__ __forceinline__ void clear() { w0 = w1 = w2 = w3 = 0; }
  __device__ __forceinline__ void set(int l, uint32_t d) {
    const uint32_t v = d << (8 * (l & 3));
    const int q = l >> 2;
    w0 |= q == 0 ? v : 0u;
    if (LP > 4) w1 |= q == 1 ? v : 0u;
    if (LP > 8) w2 |= q == 2 ? v : 0u;
    if (LP > 12) w3 |= q == 3 ? v : 0u;
  }
  __device__ __forceinline__ uint32_t get(int l) const {
    const int q = l >> 2;
    uint32_t v = w0;
    if (LP > 4) v = q == 1 ? w1 : v;
    if (LP > 8) v = q == 2 ? w2 : v;
    if (LP > 12) v = q == 3 ? w3 : v;
    return (v >> (8 * (l & 3))) & 0xffu;
  }
};

// One-byte signatures of the accepted entries, four per register.  match() returns a mask with
// bit (8*(a&3) + (a>>2)) set for every accepted slot a whose signature MAY equal `sig8` (the
// zero-byte test can flag a byte above a true match; callers check the full hash), live slots only.
template <int LP>
struct SigBytes {
  static constexpr int NR = (LP + 3) / 4;
  uint32_t r0, r1, r2, r3, live;
  __device__ __forceinline__ void clear() { r0 = r1 = r2 = r3 = live = 0; }
  static __device__ __forceinline__ uint32_t zero_bytes(uint32_t x) {
    return ((x - 0x01010101u) & ~x & 0x80808080u) >> 7;
  }
  __device__ __forceinline__ uint32_t match(uint32_t sig8) const {
    const uint32_t bc = sig8 * 0x01010101u;
    uint32_t m = zero_bytes(r0 ^ bc);
    if (NR > 1) m |= zero_bytes(r1 ^ bc) << 1;
    if (NR > 2) m |= zero_bytes(r2 ^ bc) << 2;
    if (NR > 3) m |= zero_bytes(r3 ^ bc) << 3;
    return m & live;
  }
  static __device__ __forceinline__ int slot_of_bit(int pos) { return ((pos & 7) << 2) | (pos >> 3); }
  __device__ __forceinline__ void insert(int l, uint32_t sig8) {
    const uint32_t v = sig8 << (8 * (l & 3));
    const int q = l >> 2;
    r0 |= q == 0 ? v : 0u;
    if (NR > 1) r1 |= q == 1 ? v : 0u;
    if (NR > 2) r2 |= q == 2 ? v : 0u;
    if (NR > 3) r3 |= q == 3 ? v : 0u;
    live |= 1u << (8 * (l & 3) + q);
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
// body and overlap.  An accepted candidate's metric and hash are stored to the new list right
// away (lanes accept in lock step unless one of them met a duplicate, so a warp's stores still
// land as complete sectors almost always); its history follows one accept later.
// dplanes / dhist: my new list in the current parity buffer (planes region / history region).
// Returns the number of accepted entries, or -1 if the state must be redone exactly (two heads
// tied for the maximum, or a presumed duplicate turned out to be a hash collision).
template <int LP, int NWD, int NL>
__device__ __forceinline__ int merge_lists(uint32_t *dplanes, uint32_t *dhist, const ListView<LP, NWD, NL> &V,
                                           int L) {
  const float NEG = neg_inf();
  float h[NL];
#pragma unroll
  for (int i = 0; i < NL; i++) h[i] = V.score(i, 0);
  float best;
  int i;
  select_head<NL>(h, best, i);
  if (best == NEG) return 0;
  Picks<LP> picks;
  SigBytes<LP> sigs;
  picks.clear();
  sigs.clear();
  uint32_t ptrs = 0;  // 4-bit read position per list
  int l = 0, prev_i = -1;
  float prev_best = __int_as_float(0x7f800000);  // +inf: never equals a metric
  bool redo = false;
  Fetched<NWD> em, va, vb;       // history in flight: of the last accepted entry / of a presumed duplicate pair
  bool em_on = false, v_on = false;
  uint32_t *em_dst = dhist;
  while (true) {
    const int j = (ptrs >> (4 * i)) & 15;
    redo = redo || (best == prev_best && i != prev_i);
    prev_best = best;
    prev_i = i;
    // ---- chain A: hash of the current candidate ----
    uint32_t hcur = 0, sig8 = 0;
    if (NWD > 0) {
      hcur = V.hash(i, j);
      sig8 = hcur ^ (hcur >> 16);
      sig8 = (sig8 ^ (sig8 >> 8)) & 0xffu;
    }
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
    bool dup = false;
    uint32_t dmatch = 0;
    if (NWD == 0) {
      dup = l > 0;
    } else {
      uint32_t m = sigs.match(sig8);
      while (m != 0 && !dup) {
        const int a = SigBytes<LP>::slot_of_bit(__ffs(m) - 1);
        m &= m - 1u;
        dmatch = picks.get(a);
        dup = V.hash((int)(dmatch >> 4), (int)(dmatch & 15u)) == hcur;
      }
    }
    if (dup) {
      if (NWD > 0) {  // presumed duplicate: verify the histories, one event later
        if (v_on) {
          uint32_t x[NWD > 0 ? NWD : 1], y[NWD > 0 ? NWD : 1];
          va.finish(x);
          vb.finish(y);
          uint32_t diff = 0;
#pragma unroll
          for (int w = 0; w < NWD; w++) diff |= x[w] ^ y[w];
          redo = redo || diff != 0;
        }
        va.load(V, i, j);
        vb.load(V, (int)(dmatch >> 4), (int)(dmatch & 15u));
        v_on = true;
      }
    } else {
      dplanes[0] = __float_as_uint(best);  // dplanes walks the slots of the new list
      if (NWD > 0) {
        dplanes[LP * 32] = hcur;
        if (em_on) {
          uint32_t x[NWD > 0 ? NWD : 1];
          em.finish(x);
#pragma unroll
          for (int w = 0; w < NWD; w++) em_dst[w] = x[w];
        }
        em.load(V, i, j);
        em_dst = dhist + l * 32 * NWD;
        em_on = true;
        picks.set(l, (uint32_t)((i << 4) | j));
        sigs.insert(l, sig8);
      }
      dplanes += 32;
      l++;
      if (l == L || NWD == 0) break;  // NWD == 0: every further candidate carries the same (empty) history
    }
    if (nbest == NEG) break;
    best = nbest;
    i = ni;
  }
  if (NWD > 0) {
    if (em_on) {
      uint32_t x[NWD > 0 ? NWD : 1];
      em.finish(x);
#pragma unroll
      for (int w = 0; w < NWD; w++) em_dst[w] = x[w];
    }
    if (v_on) {
      uint32_t x[NWD > 0 ? NWD : 1], y[NWD > 0 ? NWD : 1];
      va.finish(x);
      vb.finish(y);
      uint32_t diff = 0;
#pragma unroll
      for (int w = 0; w < NWD; w++) diff |= x[w] ^ y[w];
      redo = redo || diff != 0;
    }
  }
#pragma unroll
  for (int i2 = 0; i2 < NL; i2++) redo = redo || (h[i2] == prev_best);
  return redo ? -1 : l;
}

template <int LP, int NWD>
__global__ void __launch_bounds__(kFastThreads) lva_step_fast(const StepArgs A) {
  constexpr int TG = 32;
  constexpr int PD = NWD > 0 ? 2 : 1;  // planes of the target row: metric (+ hash)
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
  const int cta_role = blockIdx.x & 1;
  const uint32_t tile = blockIdx.x >> 1;
  // Rows entered by a one-bit step (puncturing block 0) can produce only two of the four bases
  // per conv state: the other two flip states and their flop states have no predecessor, are
  // never reached and keep the -inf they were initialised with.  For such rows ONE CTA per tile
  // handles the 64 live flip targets (warps 0-1) and the 64 live flop targets (warps 2-3).
  const bool one_bit = p > 0 && R.shift == 1;
  if (one_bit && cta_role == 1) return;
  // ---- which target state is mine (cheap part first: the tile copy is issued before the rest) ----
  uint32_t c, drop = 0, run, idx, lo;
  int s, src_g = 0, role, pos;  // pos: my word inside the warp's 32-word list rows
  if (p == 0) {
    role = cta_role;
    run = tid >> 6; idx = (tid >> 2) & 15u; lo = tile * 16u + idx;
    c = (run << (K.mem - 1)) | lo;
    pos = lane;
  } else if (!one_bit) {
    role = cta_role;  // 0: the 128 flip targets of the tile, 1: the 128 flop targets
    run = warp; idx = lane >> 2; lo = tile * 8u + idx;  // a warp = 8 consecutive conv states x 4 bases
    c = (run << (K.mem - 2)) | lo;
    pos = (int)((lane & 3u) * 8u + idx);
  } else {
    role = warp >> 1;
    const uint32_t tg = (tid & 63) >> 1;
    run = tg >> 4; idx = tg & 15u; lo = tile * 16u + idx;
    c = (run << (K.mem - 1)) | lo;
    pos = lane;
  }
  const bool valid = (c & R.vmask) == R.vval;

  uint64_t *bar = reinterpret_cast<uint64_t *>(smem);          // bytes 0..7
  float *post = reinterpret_cast<float *>(smem + 4);           // bytes 16..175
  uint32_t *src = smem + 64;                                   // byte 256
  const RowDev S = K.row[p > 0 ? p - 1 : 0];
  const int nws = p > 0 ? S.idwords : 0;
  const int PSRC = nws > 0 ? 2 : 1;                            // planes of the predecessor row
  // flop-only CTAs read just the flip lists 0..3 = the first half of the planes region
  const int n_src_lists = (p > 0 && !one_bit && role == 1) ? 4 : 8;
  uint32_t *own = src + n_src_lists * PSRC * LP * TG;          // own lists, per warp [plane][slot][32]
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (!__syncthreads_or(valid)) return;

  const int L = (int)K.L;
  uint32_t *prv = A.state + ((unsigned long long)blockIdx.z * 2 + (A.t & 1u)) * A.slot_words;
  uint32_t *cur = A.state + ((unsigned long long)blockIdx.z * 2 + ((A.t + 1u) & 1u)) * A.slot_words;

  // ---- stage ----
  if (tid == 0) {
    const uint32_t tile_bytes = p > 0 ? (uint32_t)n_src_lists * PSRC * LP * TG * 4u : 0u;
    mbar_expect_tx(bar, tile_bytes + 160u);
    bulk_g2s(post, A.post + (unsigned long long)A.post_blk[blockIdx.z] * 40, 160u, bar);
    if (p > 0) bulk_g2s(src, prv + S.off + (unsigned long long)tile * S.tile_words, tile_bytes, bar);
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
  constexpr int ROWS = PD * LP;                                // [plane][slot] rows of one list's planes
  uint32_t *wown = own + warp * ROWS * 32;
  const unsigned long long tile_d = R.off + (unsigned long long)(c >> 5) * R.tile_words;
  const unsigned long long own_planes = tile_d + (unsigned long long)s * ROWS * 32 + (c & 31u);
  const unsigned long long own_hist = tile_d + R.hist_off + ((unsigned long long)s * LP * 32 + (c & 31u)) * NWD;
  if (p > 0 && !one_bit) {
    // the warp's 8 conv states are consecutive: every (crf, plane, slot) row is 32 contiguous
    // bytes in HBM -> 16-byte copies
    const uint32_t c0 = ((uint32_t)warp << (K.mem - 2)) | (tile * 8u);
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
    for (int r = 0; r < ROWS; r++) cp_async4(wown + r * 32 + pos, prv + own_planes + (unsigned long long)r * 32);
  }
  cp_async_wait_all();
  __syncwarp();
  mbar_wait(bar, 0);
  if (!valid) return;

  const uint32_t *mine = wown + pos;
  uint32_t *dplanes = cur + own_planes;
  const int row = s >= 4 ? 4 : s;
  const float *wrow = post + row * 8;
  if (p == 0) {  // VCC:706-713: stay edge only, slot 0 only
    dplanes[0] = __float_as_uint(__uint_as_float(mine[0]) + wrow[s]);
#pragma unroll
    for (int l = 1; l < LP; l++)
      if (l < L) dplanes[l * 32] = 0xff800000u;
    return;
  }
  const uint32_t *srcg = src + src_g;
  const int list_w = PSRC * LP * TG;  // words between the crf lists of the staged planes
  const uint32_t *gsrc = prv + S.off + (unsigned long long)tile * S.tile_words + S.hist_off + (uint32_t)src_g * nws;
  if (LP == 1) {
    // VCC:715-742: first strict maximum, own list first then predecessor crf ascending
    float best = neg_inf();
    int bk = -1;
    const float s0 = __uint_as_float(mine[0]) + wrow[s];
    if (s0 > best) { best = s0; bk = s; }
    const int k0 = s < 4 ? 0 : s - 4, k1 = s < 4 ? 7 : s - 4;
    for (int k = k0; k <= k1; k++) {
      if (k == s) continue;
      const float sc = __uint_as_float(srcg[k * list_w]) + wrow[k];
      if (sc > best) { best = sc; bk = k; }
    }
    dplanes[0] = __float_as_uint(best);
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
    if (NWD > 0 && bk >= 0) {  // (metric -inf: hash and history are don't-care)
      uint32_t *dh = cur + own_hist;
      if (bk == s) {
        dplanes[32] = mine[32];
#pragma unroll
        for (int w = 0; w < NWD; w++) dh[w] = ldg_nc(prv + own_hist + w);
      } else {
        dplanes[32] = mix_hash(nws > 0 ? srcg[bk * list_w + TG] : 0u, drop);
        uint32_t prev = 0;
#pragma unroll
        for (int w = 0; w < NWD; w++) {
          uint32_t old = 0;
          if (w < nws) old = ldg_nc(gsrc + (bk * 32) * nws + w);
          dh[w] = (old << R.shift) | (w == 0 ? drop : (prev >> (32 - R.shift)));
          prev = old;
        }
      }
    }
    return;
  }
  int n_acc;
  if (role == 0) {
    ListView<LP, NWD, 8> V{smem_u32(srcg), smem_u32(mine), smem_u32(wrow), (uint32_t)list_w * 4u,
                           gsrc, prv + own_hist, nws, (int)R.shift, s, drop};
    n_acc = merge_lists<LP, NWD, 8>(dplanes, cur + own_hist, V, L);
  } else {
    ListView<LP, NWD, 2> V{smem_u32(srcg), smem_u32(mine), smem_u32(wrow), (uint32_t)list_w * 4u,
                           gsrc, prv + own_hist, nws, (int)R.shift, s, drop};
    n_acc = merge_lists<LP, NWD, 2>(dplanes, cur + own_hist, V, L);
  }
  if (n_acc < 0) {  // exact score tie between two list heads (libstdc++ heap order decides), or hash collision
    exact_merge<LP>(prv, cur, post, &K, p, c, s, (int)(tile * TG + (uint32_t)src_g), drop);
    atomicAdd(A.counters, 1ull);
  } else {
    for (int l = n_acc; l < L; l++) dplanes[l * 32] = 0xff800000u;  // -inf; the rest of a dead slot is never read
  }
}

// ---------------------------------------------------------------------------------------------
// host-side dispatch
// ---------------------------------------------------------------------------------------------
inline size_t fast_smem_bytes(uint32_t LP, uint32_t nwd) {
  // flip CTA: staged predecessor planes (8 lists) + own lists' planes; independent of the history width
  const size_t planes = nwd > 0 ? 2 : 1;
  return 256 + ((size_t)8 * planes * LP * 32 + planes * LP * kFastThreads) * 4;
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
inline cudaError_t fast_launch_one(cudaStream_t st, dim3 grid, const StepArgs &a) {
  lva_step_fast<LP, NWD><<<grid, kFastThreads, fast_smem_bytes(LP, NWD), st>>>(a);
  return cudaGetLastError();
}

// One time step: one launch per distinct history width among rows [lo, hi).
inline cudaError_t fast_launch_step(uint32_t LP, cudaStream_t st, StepArgs a, const CodeSpec &k, uint32_t lo,
                                    uint32_t hi, uint32_t active, uint64_t *n_launches) {
  uint32_t p = lo;
  while (p < hi) {
    const uint32_t nw = k.idwords[p];
    uint32_t q = p;
    while (q < hi && k.idwords[q] == nw) q++;
    a.row_lo = p;
    a.row_hi = q;
    dim3 grid(2 * (k.nconv / 32), q - p, active);  // x: (tile, role)
    cudaError_t e = cudaErrorInvalidValue;
#define Y(lp, nwc) \
  if (LP == lp && nw == nwc) e = fast_launch_one<lp, nwc>(st, grid, a);
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
