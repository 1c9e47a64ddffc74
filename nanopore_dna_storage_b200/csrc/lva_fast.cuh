// Fast path of the trellis step, specialised on list slots (LP).
//
// Geometry: one tile of 32 predecessor conv states of row p-1 per CTA.  The tile feeds 128 flip
// targets (8-list merges, warps 0-3) and 128 flop targets (2-list merges, warps 4-7) of row p;
// on rows entered by a one-bit step only two of the four bases of a conv state are producible,
// so 64 + 64 targets are live (warps 0-1 flip, 2-3 flop, warps 4-7 retire at once).
//   * every list entry is 3 words (metric, chain id, partial chunk -- lva_device.cuh), so ONE
//     kernel serves every row of every code: 6 KB of shared memory per warp, 4 CTAs per SM at
//     list size 8;
//   * the predecessor tile (4 octet pieces of row p-1) arrives as four contiguous TMA bulk copies,
//     piece by piece as it lies in HBM (SASS: UBLKCP); the CTA's own (stay) lists are whole pieces
//     of row p, one 5-D tensor copy per warp re-shaped to [plane][slot][4 bases x 8 conv states]
//     (UTMALDG).  Three mbarriers in the order the data is needed: transition scores + tile, own
//     lists of the flop warps, own lists of the flip warps (which merge their seven predecessor
//     lists first);
//   * each thread owns one target state (p, c, s).  Own lists: bank = lane, no conflicts.  Tile:
//     a lane's column is (piece, conv state & 7), the pieces padded 8 words apart: lanes reading
//     the same slot never conflict, lanes at different slots may (two-way on average);
//   * list size == slot count <= 16: merge_two_stage() -- the best move-in candidates by a plain
//     7-way merge, twins of the own entries found by exact compares in registers, two-way merge
//     of the survivors; every stage runs a fixed number of trips in lock step.  Other list sizes
//     and the first rows of the trellis: merge_lists(), a pop-by-pop k-way merge with
//     de-duplication on (chain id, partial chunk) through one-byte signatures;
//   * whenever the reference's heap order could show in the output (exactly equal metrics from
//     different lists) the thread re-runs that state through exact_merge(), which reproduces
//     libstdc++'s heap order (VCC:762-795).  Without a tie the pop sequence of any correct k-way
//     merge equals the reference's, so the fast path is bit-identical by construction;
//   * boundary rows publish the interned id of every new entry (lva_device.cuh) after the merge.
#pragma once
#include "lva_kernels.cuh"

namespace lva {

constexpr int kFastThreads = 256;

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) { return __uint_as_float(lds_u32(addr)); }

// Conditional moves on the FMA pipe.  The merges are made of compares and selects, all of them ALU-pipe
// instructions (ncu: ALU pipe 70 % busy, FMA pipe 15 %); `x = p ? y : x` written as a predicated integer
// multiply-add `@p x = one * y + 0` is the same move issued on the other pipe.  `one` is a 1 the compiler
// cannot see through: a word of shared memory (kOneWord, written from StepArgs::one) read with a volatile
// load inside the loop that uses it -- a multiplier known to be loop-invariant lets ptxas hoist the products
// and turn the moves back into selects.  An integer multiply by 1 is exact for any 32-bit pattern, metrics
// included.  (profiles/r02c_pmov_ab.log: the four rewrites of this kind together, 51.0 -> 54.3 reads/s on C2.)
constexpr uint32_t kOneWord = 56;  // word of the CTA's shared memory that holds 1 (bytes 224..227: between the barriers and the tile)
__device__ __forceinline__ uint32_t lds_volatile_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
// (a0 == b0 && a1 == b1) ? (x0 = y0, x1 = y1) : nothing
__device__ __forceinline__ void cond_move2_if_pair_eq(uint32_t &x0, uint32_t &x1, uint32_t a0, uint32_t b0, uint32_t a1, uint32_t b1,
                                                      uint32_t y0, uint32_t y1, uint32_t one) {
  asm("{\n\t.reg .pred q;\n\tsetp.eq.u32 q, %2, %3;\n\tsetp.eq.and.u32 q, %4, %5, q;\n\t"
      "@q mad.lo.u32 %0, %8, %6, 0;\n\t@q mad.lo.u32 %1, %8, %7, 0;\n\t}"
      : "+r"(x0), "+r"(x1)
      : "r"(a0), "r"(b0), "r"(a1), "r"(b1), "r"(y0), "r"(y1), "r"(one));
}

// (a == b) ? x = y + w : nothing -- a predicated FADD (FMA pipe) in place of an FADD and a select
__device__ __forceinline__ void cond_add_if_eq(float &x, int a, int b, float y, float w) {
  asm("{\n\t.reg .pred q;\n\tsetp.eq.s32 q, %1, %2;\n\t@q add.f32 %0, %3, %4;\n\t}" : "+f"(x) : "r"(a), "r"(b), "f"(y), "f"(w));
}

// Everything a thread needs to address the lists it merges.  Staged lists share one shape:
// rows of 32 words per (plane, crf, slot); the predecessor tile is [plane][crf][slot], the
// warp's own lists [plane][slot].
//   NL = 8 (flip target s < 4): list i is predecessor crf i, except list s which is the state's
//                               own (stay) list; all weights come from posterior row s.
//   NL = 2 (flop target s >= 4): list 0 = own list (column s of row 4), list 1 = the flip state
//                               s-4 of the predecessor conv state (column s-4 of row 4).
// All list accesses go through 32-bit shared-window addresses computed once per thread (plain
// generic pointers made the compiler rebuild the shared window base inside the merge loop).
// The predecessor tile is staged piece by piece as it lies in HBM ([plane][crf][slot][8 conv states], one
// bulk copy per piece, 8 words of padding between pieces so that the four pieces sit in different banks):
// a lane's column is (piece, conv state & 7).  The own lists keep rows of 32 words per (plane, slot).
template <int LP>
struct TileGeom {
  static constexpr uint32_t kPieceWords = 3 * 8 * LP * 8;       // = StepArgs::piece_words
  static constexpr uint32_t kPieceStride = kPieceWords + 8;     // words, staged
  static constexpr uint32_t kTileWords = 4 * kPieceStride;
};

template <int LP, int NL>
struct ListView {
  static constexpr uint32_t SS = 32 * 4;              // own list: slot stride in bytes
  static constexpr uint32_t LS = LP * SS;             // own list: between the planes
  static constexpr uint32_t TSS = 8 * 4;              // tile: slot stride in bytes
  static constexpr uint32_t TLS = LP * TSS;           // tile: between the crf lists
  static constexpr uint32_t TPS = 8 * TLS;            // tile: between the planes
  uint32_t srcg, mine, wrow;  // shared addresses: my predecessor column, my own list, my score row
  uint32_t keep, app;         // identity of a move candidate: (id, (partial & keep) | app)
  int s;
  uint32_t one_addr;          // shared address of kOneWord (cond_move*)

  // keep the addresses in registers: left alone, the compiler rebuilds them from the thread's
  // geometry inside the merge loop
  __device__ __forceinline__ void pin() { asm volatile("" : "+r"(srcg), "+r"(mine), "+r"(wrow), "+r"(keep), "+r"(app)); }
  __device__ __forceinline__ int crf_of(int i) const { return NL == 8 ? i : (i == 0 ? s : s - 4); }
  __device__ __forceinline__ bool is_stay(int i) const { return NL == 8 ? i == s : i == 0; }
  __device__ __forceinline__ uint32_t entry(int i, int j) const {  // address of the metric of entry j of list i
    return is_stay(i) ? mine + (uint32_t)j * SS : srcg + (uint32_t)crf_of(i) * TLS + (uint32_t)j * TSS;
  }
  __device__ __forceinline__ float score(int i, int j) const {
    return lds_f32(entry(i, j)) + lds_f32(wrow + 4u * (uint32_t)crf_of(i));
  }
  // (chain id, partial chunk) that entry j of list i would carry at the target
  __device__ __forceinline__ void identity(int i, int j, uint32_t &id, uint32_t &pt) const {
    const bool st = is_stay(i);
    const uint32_t b = entry(i, j), ps = st ? LS : TPS;
    id = lds_u32(b + ps);
    const uint32_t raw = lds_u32(b + 2 * ps);
    pt = st ? raw : ((raw & keep) | app);
  }
};

// Log of the accepted entries, kept as byte shift registers: byte k of register r describes the
// (4r+k)-th most recent accept.  `pk` holds its (list << 4 | slot) descriptor -- a signature
// match is verified by re-deriving that entry's identity from the staged lists -- and `sig` a
// one-byte signature of its identity whose top bit is always set, so that the zero bytes of the
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
  __device__ __forceinline__ uint32_t desc_recent(int e) const {  // e-th most recent accept, e a compile-time index
    return (pk[e >> 2] >> (8 * (e & 3))) & 0xffu;
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
// where the heap order of the reference could show in the output: a candidate popped with the
// metric of the last accepted entry but from a different list (accepted or not: a dropped copy
// may be the one the reference accepts first), or a head equal to the last accepted metric when
// the loop ends.
// Returns the number of accepted entries, or -1 on a tie (caller falls back to exact_merge and
// overwrites whatever was stored).
// Rows without any history bit yet (the first rows of the trellis): every candidate of a state carries
// the same (empty) message, so the new list is the best head alone (VCC:743-800 drops all the others as
// duplicates).  Returns 0 / 1, or -1 when another head ties with the best (heap order decides).
template <int LP, int NL>
__device__ __forceinline__ int merge_first_only(uint32_t *dst, const ListView<LP, NL> &V, AcceptLog<LP> &log) {
  constexpr uint32_t GP = 64u * LP;
  float h[NL];
#pragma unroll
  for (int i = 0; i < NL; i++) h[i] = V.score(i, 0);
  float best;
  int i;
  select_head<NL>(h, best, i);
  log.clear();
  if (best == neg_inf()) return 0;
  bool tie = false;
#pragma unroll
  for (int i2 = 0; i2 < NL; i2++) tie = tie || (i2 != i && h[i2] == best);
  uint32_t cid, cpt;
  V.identity(i, 0, cid, cpt);
  __stcs(dst, __float_as_uint(best));
  __stcs(dst + GP, cid);
  __stcs(dst + 2 * GP, cpt);
  log.push(AcceptLog<LP>::signature(cid ^ cpt), (uint32_t)(i << 4));
  return tie ? -1 : 1;
}

// FULL: the list size equals the slot count (L == LP), the usual case: one constant less in the loop.
template <int LP, int NL, bool FULL>
__device__ __forceinline__ int merge_lists(uint32_t *dst, const ListView<LP, NL> &V, int L_rt, AcceptLog<LP> &log) {
  const int L = FULL ? LP : L_rt;
  constexpr uint32_t GP = 64u * LP;  // words between the planes of a list in HBM
  const float NEG = neg_inf();
  float h[NL];
#pragma unroll
  for (int i = 0; i < NL; i++) h[i] = V.score(i, 0);
  float best;
  int i;
  select_head<NL>(h, best, i);
  log.clear();
  if (best == NEG) return 0;
  uint32_t ptrs = 0;  // 4-bit read position per list
  int l = 0, last_i = -1;
  float last_acc = __int_as_float(0x7f800000);  // +inf: never equals a metric
  bool tie = false;
  while (true) {
    const int j = (ptrs >> (4 * i)) & 15;
    // ---- chain A: identity and signature of the current candidate ----
    uint32_t cid, cpt;
    V.identity(i, j, cid, cpt);
    const uint32_t csig = AcceptLog<LP>::signature(cid ^ cpt);
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
    tie = tie || (best == last_acc && i != last_i);
    uint32_t m = log.match(csig);
    bool dup = false;
    while (m != 0 && !dup) {
      const uint32_t d = log.desc_at(__ffs(m) - 1);
      m &= m - 1u;
      uint32_t aid, apt;
      V.identity((int)(d >> 4), (int)(d & 15u), aid, apt);
      dup = aid == cid && apt == cpt;
    }
    if (!dup) {
      last_acc = best;
      last_i = i;
      __stcs(dst, __float_as_uint(best));  // dst walks the slots of the new list; streaming stores: the new
                                           // list is not read again before gigabytes of other traffic went by
      __stcs(dst + GP, cid);
      __stcs(dst + 2 * GP, cpt);
      dst += 8;
      log.push(csig, (uint32_t)((i << 4) | j));
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

// LP one-byte descriptors (tile crf << 4 | slot), entry e at byte e: one 64-bit word up to 8 slots, two for 16.
template <int LP>
struct DescVec {
  static constexpr int NW = LP <= 8 ? 1 : 2;
  unsigned long long w[NW];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < NW; i++) w[i] = 0;
  }
  // shift every entry one place down and put d in the last place: after LP pushes the first push sits at entry 0
  __device__ __forceinline__ void push_back(uint32_t d) {
    if (NW == 2) w[0] = (w[0] >> 8) | (w[NW - 1] << 56);
    constexpr int top = LP <= 8 ? 8 * LP - 8 : 56;
    w[NW - 1] = (w[NW - 1] >> 8) | ((unsigned long long)d << top);
  }
  __device__ __forceinline__ uint32_t get(int e) const {
    const unsigned long long v = (NW == 2 && e >= 8) ? w[NW - 1] : w[0];
    return (uint32_t)(v >> (8 * (e & 7))) & 0xffu;
  }
};

// Two-stage merge (list size == slot count <= 16; every row whose source row cannot hold one message in
// two crf lists -- RowDev::xdup, code_tables.cpp first_run_rows()).
//
// A message reaches a state at most twice per step: through the state's own (stay) list A and through
// ONE predecessor list: the crf state a message sits in at row p-1 follows from its base sequence
// (last base; flip or flop by the parity of the closing homopolymer run), the only freedom being the
// parity of the FIRST run, because all eight crf states start at row 0 (VCC:657-663).  Away from those
// first rows the move-in candidates B are pairwise distinct, so
//   stage B  the LP best move-in candidates are the first LP pops of a plain 7-way merge -- no
//            de-duplication, every lane pops in lock step, exactly LP trips (flop targets have one
//            predecessor list: nothing to merge).  Later move-in candidates can never be accepted:
//            LP distinct messages with better metrics come before them;
//   stage T  twins: each of the LP own entries is compared with the LP move-in candidates, held in
//            registers (exact two-word compare, no signatures, no divergence); of a twin pair the entry
//            with the lower metric is marked dropped;
//   stage F  the survivors of A and B are disjoint and sorted: a two-way merge writes the new list, at
//            most LP trips, in lock step.
// Without ties this is the reference's list: sort all candidates by metric, keep the first occurrence of
// every message, stop at LP.  Ties (equal metrics of successive pops from different lists in stage B,
// of the candidate behind the last one kept, of a twin pair, of the two heads in stage F) send the state
// to exact_merge() as before: a superset of the cases where libstdc++'s heap order could show.
// Returns the number of accepted entries, or -1 on a tie.
template <int LP, int NL>
__device__ __forceinline__ int merge_two_stage(uint32_t *dst, const ListView<LP, NL> &V, AcceptLog<LP> &log,
                                               bool check, unsigned long long *counters, uint64_t *own_bar, bool want_log) {
  static_assert(LP <= 16, "stage T keeps 3*LP words of move-in candidates in registers (16 slots: 2 CTAs per SM by shared memory anyway)");
  using View = ListView<LP, NL>;
  constexpr uint32_t SS = View::SS, LS = View::LS, TSS = View::TSS, TLS = View::TLS, TPS = View::TPS, GP = 64u * LP;
  constexpr uint32_t FULL = (1u << LP) - 1u;
  const float NEG = neg_inf();
  bool tie = false;
  DescVec<LP> desc;  // entry e = (tile crf << 4 | slot) of the e-th best move-in candidate
  desc.clear();
  int nB = LP;                  // live move-in candidates among them

  // ---- stage B ----
  if constexpr (NL == 8) {
    float h[8];
    float w[8];  // transition scores of the seven predecessor lists, in registers while stage B runs
#pragma unroll
    for (int i = 0; i < 8; i++) {
      w[i] = lds_f32(V.wrow + 4u * (uint32_t)i);
      h[i] = i == V.s ? NEG : lds_f32(V.srcg + (uint32_t)i * TLS) + w[i];
    }
    float best, last_m = __int_as_float(0x7f800000);  // +inf: never equals a metric
    int i, last_i = -1;
    select_head<8>(h, best, i);
    uint32_t ptrs = 0;
    nB = 0;
    int e = 0;
#pragma unroll 1
    for (; e < LP; e++) {  // always LP trips (no vote for an early exit: lists are full almost everywhere); a lane out of
                           // candidates pushes descriptors of slots that exist and counts nothing
      const bool live = best != NEG;
      const int j = (ptrs >> (4 * i)) & 15;
      desc.push_back((uint32_t)((i << 4) | j));
      nB += live ? 1 : 0;
      tie = tie || (live && best == last_m && i != last_i);
      last_m = best;
      last_i = i;
      // the popped list's next head: its raw metric (-inf when the list is exhausted: -inf + score = -inf) is added
      // to the list's own score inside the conditional move, one predicated FADD per list
      float raw = NEG;
      if (live && j + 1 < LP) {
        raw = lds_f32(V.srcg + (uint32_t)i * TLS + (uint32_t)(j + 1) * TSS);
        ptrs += 1u << (4 * i);
      }
#pragma unroll
      for (int i2 = 0; i2 < 8; i2++) cond_add_if_eq(h[i2], i, i2, raw, w[i2]);
      select_head<8>(h, best, i);
    }
    tie = tie || (best != NEG && best == last_m);  // the candidate behind the LP-th: which of the two is in B?
  } else {
    const unsigned long long crfs = 0x1010101010101010ull * (unsigned long long)(uint32_t)(V.s - 4);
    desc.w[0] = (0x0706050403020100ull | crfs) & (~0ull >> (LP >= 8 ? 0 : 64 - 8 * LP));
    if (DescVec<LP>::NW == 2) desc.w[DescVec<LP>::NW - 1] = 0x0f0e0d0c0b0a0908ull | crfs;
  }

  // ---- the move-in candidates, into registers ----
  float mB[LP];
  uint32_t idB[LP], ptB[LP];
#pragma unroll
  for (int e = 0; e < LP; e++) {
    const uint32_t d = desc.get(e), crf = d >> 4;
    const uint32_t addr = V.srcg + crf * TLS + (d & 15u) * TSS;
    const float m = lds_f32(addr) + lds_f32(V.wrow + 4u * crf);
    const bool live = e < nB && m != NEG;
    mB[e] = live ? m : NEG;
    // a dead slot may still hold the identity of an entry that lives elsewhere: it must match nothing
    idB[e] = live ? lds_u32(addr + TPS) : 0xffffffffu;  // no chain id has this value (table slots < 2^27, kRootId)
    ptB[e] = (lds_u32(addr + 2 * TPS) & V.keep) | V.app;
  }
  if (check) {  // developer switch (LVA_DBG bit 3): the structural argument above, tested
    bool bad = false;
#pragma unroll
    for (int e = 1; e < LP; e++)
#pragma unroll
      for (int e2 = 0; e2 < e; e2++) bad = bad || (mB[e] != NEG && idB[e] == idB[e2] && ptB[e] == ptB[e2]);
    if (bad) atomicAdd(counters + 2, 1ull);
  }

  // ---- stage T ----
  mbar_wait(own_bar, 0);  // the own lists were the last copies issued: stage B ran under them
  const float wA = lds_f32(V.wrow + 4u * (uint32_t)V.s);
  uint32_t dropA = 0, dropB = 0;
#pragma unroll 1
  for (int a = 0; a < LP; a++) {
    const uint32_t addr = V.mine + (uint32_t)a * SS;
    const float mA = lds_f32(addr) + wA;
    const uint32_t idA = lds_u32(addr + LS), ptA = lds_u32(addr + 2 * LS);
    uint32_t tb = 0, tmb = __float_as_uint(NEG);  // the twin among the move-in candidates (at most one): its bit and metric
    const uint32_t one = lds_volatile_u32(V.one_addr);
#pragma unroll
    for (int e = 0; e < LP; e++) cond_move2_if_pair_eq(tmb, tb, idA, idB[e], ptA, ptB[e], __float_as_uint(mB[e]), 1u << e, one);
    const float tm = __uint_as_float(tmb);
    const bool has = tb != 0u && tm != NEG && mA != NEG;
    tie = tie || (has && mA == tm);
    if (has) {
      if (mA > tm) dropB |= tb;
      else dropA |= 1u << a;
    }
  }

  // ---- stage F ----
  log.clear();
  uint32_t keepA = ~dropA & FULL, keepB = ~dropB & FULL;
  uint32_t addrA = V.mine, addrB = V.srcg, jA = 0, dB = 0;
  float hA = NEG, hB = NEG;
  // heads without a branch: an empty side reads a slot that exists (its last one; descriptor 0 behind the live
  // move-in candidates) and the value is replaced by -inf
  auto head_a = [&]() {
    jA = ((uint32_t)__ffs((int)keepA) - 1u) & (uint32_t)(LP - 1);
    addrA = V.mine + jA * SS;
    const float m = lds_f32(addrA) + wA;
    hA = keepA != 0u ? m : NEG;
  };
  auto head_b = [&]() {
    const int eB = __ffs((int)keepB) - 1;
    dB = desc.get(eB & (LP - 1));
    const uint32_t crf = dB >> 4;
    addrB = V.srcg + crf * TLS + (dB & 15u) * TSS;
    const float m = lds_f32(addrB) + lds_f32(V.wrow + 4u * crf);
    hB = (keepB != 0u && eB < nB) ? m : NEG;
  };
  head_a();
  head_b();
  int l = 0;
  float last = __int_as_float(0x7f800000);
  uint32_t *d = dst;
  const uint32_t own_list = NL == 8 ? (uint32_t)V.s : 0u;
#pragma unroll 1
  while (true) {
    const float best = fmaxf(hA, hB);
    if (best == NEG) break;
    tie = tie || hA == hB;
    const bool pop_a = hA >= hB;
    const uint32_t addr = pop_a ? addrA : addrB, ps = pop_a ? LS : TPS;
    const uint32_t id = lds_u32(addr + ps), raw = lds_u32(addr + 2 * ps);
    const uint32_t pt = pop_a ? raw : ((raw & V.keep) | V.app);
    __stcs(d, __float_as_uint(best));
    __stcs(d + GP, id);
    __stcs(d + 2 * GP, pt);
    d += 8;
    if (want_log) log.push(0u, pop_a ? ((own_list << 4) | jA) : (NL == 8 ? dB : (0x10u | (dB & 15u))));  // boundary rows: publish_ids()
    l++;
    last = best;
    // no divergent branch: only the popped side loses its head, and re-deriving the head of the other side gives the
    // value it already has, so both are derived on every trip (a warp whose lanes pop from different sides ran both
    // branches one after the other anyway)
    const uint32_t popped = pop_a ? keepA : keepB, rest = popped & (popped - 1u);
    keepA = pop_a ? rest : keepA;
    keepB = pop_a ? keepB : rest;
    head_a();
    head_b();
    if (l == LP) break;  // after the heads: the candidate behind the last one kept is tested for a tie below
  }
  tie = tie || (fmaxf(hA, hB) == last);  // the candidate behind the last one kept (l == 0: last = +inf)
  return tie ? -1 : l;
}

// Which merge serves this state: rows without history, the two-stage merge, or the pop-by-pop loop.
template <int LP, int NL>
__device__ __forceinline__ int merge_state(uint32_t *dst, const ListView<LP, NL> &V, int L, AcceptLog<LP> &log, bool nohist,
                                           bool two_stage, bool check,
                                           unsigned long long *counters, uint64_t *own_bar, bool want_log) {
  if constexpr (LP <= 16) {
    if (two_stage && !nohist) return merge_two_stage<LP, NL>(dst, V, log, check, counters, own_bar, want_log);
  }
  mbar_wait(own_bar, 0);  // every other merge reads the own list from its first instruction on
  if (nohist) return merge_first_only<LP, NL>(dst, V, log);
  if (L == LP) return merge_lists<LP, NL, true>(dst, V, L, log);
  return merge_lists<LP, NL, false>(dst, V, L, log);
}

// Boundary rows: publish the interned id of (chain id, partial chunk) of every entry just written.
// An entry taken over from the state's own list keeps the id that list published one step ago;
// only entries that moved in are looked up in (or added to) the read's chain table.
template <int LP, int NL>
__device__ __forceinline__ void publish_ids(const ListView<LP, NL> &V, const AcceptLog<LP> &log, int n_acc, uint32_t *nidw,
                                            const uint32_t *nid_own, unsigned long long *tab, const StepArgs &A, uint32_t z) {
#pragma unroll
  for (int e = 0; e < LP; e++)
    if (e < n_acc) {
      const uint32_t d = log.desc_recent(e);
      const int i = (int)(d >> 4), j = (int)(d & 15u);
      uint32_t v;
      if (V.is_stay(i)) {
        v = nid_own[j * 8];
      } else {
        uint32_t id, pt;
        V.identity(i, j, id, pt);
        v = intern_chunk(tab, A.tab_mask, id, pt, A.counters, z);
      }
      nidw[(n_acc - 1 - e) * 8] = v;
    }
}

template <int LP>
__global__ void __launch_bounds__(kFastThreads, LP <= 8 ? 4 : 2) lva_step_fast(const StepArgs A, const __grid_constant__ CUtensorMap own_map) {
  constexpr int TG = 32;
  constexpr uint32_t GP = 64u * LP;
  extern __shared__ __align__(128) uint32_t smem[];
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
  // window slots of rows p and p-1 (p % window without a division: the band never spans more than the window)
  uint32_t wp = A.band_slot[z] + blockIdx.y;
  if (wp >= K.window) wp -= K.window;
  const uint32_t wp1 = wp == 0 ? K.window - 1 : wp - 1;
  // Rows entered by a one-bit step (puncturing block 0) can produce only two of the four bases
  // per conv state: the other two flip states and their flop states have no predecessor, are
  // never reached and keep their -inf.  On such rows warps 0-1 handle the 64 live flip targets
  // and warps 2-3 the 64 live flop targets of the tile.
  const bool one_bit = R.shift == 1;
  // ---- which target state is mine (cheap part first: the copies are issued before the rest) ----
  uint32_t c, run, idx, lo;
  int role;
  if (!one_bit) {
    role = warp >> 2;  // 0: the 128 flip targets of the tile, 1: the 128 flop targets
    run = warp & 3; idx = lane >> 2; lo = tile * 8u + idx;  // a warp = 8 consecutive conv states x 4 bases
    c = (run << (K.mem - 2)) | lo;
  } else {
    role = (warp >> 1) & 1;
    const uint32_t tg = (tid & 63) >> 1;
    run = tg >> 4; idx = tg & 15u; lo = tile * 16u + idx;
    c = (run << (K.mem - 1)) | lo;
  }
  const bool valid = (c & R.vmask) == R.vval && !(one_bit && warp >= 4);

  // three load barriers, in the order the data is needed: [0] posterior block + predecessor tile (every warp, at
  // once), [1] own lists of the flop warps (right after), [2] own lists of the flip warps (only after their stage B)
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 48);     // bytes 192..215
  float *post = reinterpret_cast<float *>(smem + 4);           // bytes 16..175
  uint32_t *src = smem + 64;                                   // byte 256: 4 x [plane][crf][slot][8] + padding (TileGeom)
  uint32_t *own = src + TileGeom<LP>::kTileWords;              // per warp [plane][slot][32] (one-bit rows: 2 pieces per warp)
  // The copies are issued by lane 0 of the LAST warp: a flop warp (it has some 600 instructions of slack against
  // the flip warps, the critical path of the CTA), or on one-bit rows of a 256-thread launch a warp with nothing else to do
  const bool issuer = tid == (int)blockDim.x - 32;
  if (issuer) {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
    mbar_init(bar + 2, 1);
    fence_barrier_init();
    smem[kOneWord] = A.one;
  }
  if (!__syncthreads_or(valid)) return;

  const uint32_t ppr = K.nconv >> 3;  // pieces per row
  const uint32_t prv_row0 = (z * 2u + (A.t & 1u)) * K.window, cur_row0 = (z * 2u + ((A.t + 1u) & 1u)) * K.window;
  // ---- stage ----
  if (issuer) {
    constexpr uint32_t plane_bytes = 8u * LP * 8u * 4u;  // one plane of one piece
    const uint32_t warp_bytes = 3u * LP * TG * 4u;
    mbar_expect_tx(bar, 160u + 4u * (R.post_boundary ? 2u : 3u) * plane_bytes);
    mbar_expect_tx(bar + 1, 4u * warp_bytes);
    mbar_expect_tx(bar + 2, 4u * warp_bytes);
    bulk_g2s(post, A.post + (unsigned long long)A.post_blk[z] * 40, 160u, bar);
    // the predecessor tile: four consecutive pieces of row p-1, each one contiguous copy (SASS: UBLKCP)
    const uint32_t *tile_g = A.state + (unsigned long long)(prv_row0 + wp1) * A.row_words + (unsigned long long)tile * 4u * A.piece_words;
    if (!R.post_boundary) {
      for (uint32_t q = 0; q < 4; q++)
        bulk_g2s(src + q * TileGeom<LP>::kPieceStride, tile_g + q * TileGeom<LP>::kPieceWords, 3u * plane_bytes, bar);
    } else {  // metrics only; the chain ids of a completed chunk come from the source row's nid plane, the partial chunk starts empty
      const uint32_t *nid_g = A.nid + nid_base(A, K, z, A.t & 1u, R.src_nid_slot) + (unsigned long long)tile * 4u * (64u * LP);
      for (uint32_t q = 0; q < 4; q++) {
        bulk_g2s(src + q * TileGeom<LP>::kPieceStride, tile_g + q * TileGeom<LP>::kPieceWords, plane_bytes, bar);
        bulk_g2s(src + q * TileGeom<LP>::kPieceStride + 64u * LP, nid_g + q * (64u * LP), plane_bytes, bar);
      }
    }
    const int own_piece0 = (int)((prv_row0 + wp) * ppr);
    if (!one_bit) {  // flop warps (4..7) first
      for (uint32_t i = 0; i < 8; i++) {
        const uint32_t w = i ^ 4u;
        tensor5d_g2s(own + w * 3 * LP * TG, &own_map, 0, (int)(4u * (w >> 2)), 0, 0,
                     own_piece0 + (int)(((w & 3u) << (K.mem - 5)) + tile), bar + 1 + (i >> 2));
      }
    } else {         // flop warps are 2..3
      for (uint32_t i = 0; i < 4; i++) {
        const uint32_t w = i ^ 2u;
        for (uint32_t h = 0; h < 2; h++)
          tensor5d_g2s(own + (w * 2 + h) * 3 * LP * TG, &own_map, 0, (int)(4u * (w >> 1)), 0, 0,
                       own_piece0 + (int)(((w & 1u) << (K.mem - 4)) + tile * 2u + h), bar + 1 + (i >> 1));
      }
    }
  }
  // ---- the rest of the geometry, while the copies fly: my crf state and predecessor column ----
  int s, src_g, pos;
  uint32_t xy;
  if (!one_bit) {
    const uint32_t b = lane & 3u;
    const uint32_t t = b ^ emitted_base(K, R.block, lo << 2, run & 1u, run >> 1);
    xy = (R.inv_lin >> (2 * t)) & 3u;
    src_g = (int)((idx << 2) | xy);
    s = (int)b + 4 * role;
    pos = warp * 3 * LP * TG + (int)(b * 8u + idx);
  } else {
    xy = tid & 1u;
    const uint32_t b = emitted_base(K, 0, lo << 1, run, 0) ^ (xy ? R.lin1 : 0u);
    s = (int)b + 4 * role;
    src_g = (int)((idx << 1) | xy);
    pos = (warp * 2 + (int)(idx >> 3)) * 3 * LP * TG + (int)(b * 8u + (idx & 7u));
  }
  const uint32_t own_off = (c >> 3) * A.piece_words + (uint32_t)s * LP * 8 + (c & 7u);  // inside a row (< 2^32 words)
  uint32_t *dst = A.state + (unsigned long long)(cur_row0 + wp) * A.row_words + own_off;
  const uint32_t appv = appended_bits(xy, R.shift, R.app) << R.app_pos;
  const uint32_t keep = R.post_boundary ? 0u : 0xffffffffu;
  mbar_wait(bar, 0);
  if (A.dbg & 32u) {  // developer switch: every copy awaited before any work (A/B of the barrier split)
    mbar_wait(bar + 1, 0);
    mbar_wait(bar + 2, 0);
  }
  if (!valid) return;
  uint64_t *own_bar = bar + (role == 0 ? 2 : 1);

  uint32_t *mine = own + pos;
  if (A.dbg & 2u) {  // developer switch: data movement only -- the own list is copied to the new list (WRONG results)
    mbar_wait(own_bar, 0);
    const uint32_t *col = src + (uint32_t)(src_g >> 3) * TileGeom<LP>::kPieceStride + (uint32_t)(src_g & 7);
    if (A.dbg & 16u) { if (mine[0] == 0x12345678u) dst[0] = col[0]; return; }  // ... and no stores
#pragma unroll
    for (int l = 0; l < LP; l++) {
      __stcs(dst + l * 8, mine[l * 32]);
      __stcs(dst + GP + l * 8, mine[(LP + l) * 32] ^ col[l * 8]);
      __stcs(dst + 2 * GP + l * 8, mine[(2 * LP + l) * 32]);
    }
    return;
  }
  const int L = (int)K.L;
  const bool two_stage = LP <= 16 && L == LP && !R.xdup && !(A.dbg & 4u);  // LVA_DBG bit 2: the pop-by-pop merge of merge_lists() everywhere
  const int row = s >= 4 ? 4 : s;
  const float *wrow = post + row * 8;
  const uint32_t *srcg = src + (uint32_t)(src_g >> 3) * TileGeom<LP>::kPieceStride + (uint32_t)(src_g & 7);
  AcceptLog<LP> log;
  int n_acc;
  uint32_t *nidw = nullptr;
  const uint32_t *nid_own = nullptr;
  unsigned long long *tab = nullptr;
  if (R.boundary) {  // this row's entries also publish their interned id (lva_device.cuh)
    const unsigned long long o = (unsigned long long)(c >> 3) * GP + (unsigned long long)s * LP * 8 + (c & 7u);
    nidw = A.nid + nid_base(A, K, z, (A.t + 1u) & 1u, R.nid_slot) + o;
    nid_own = A.nid + nid_base(A, K, z, A.t & 1u, R.nid_slot) + o;
    tab = A.tab + (unsigned long long)z * ((unsigned long long)A.tab_mask + 1ull);
  }
  if (role == 0) {
    ListView<LP, 8> V{smem_u32(srcg), smem_u32(mine), smem_u32(wrow), keep, appv, s, smem_u32(smem + kOneWord)};
    V.pin();
    n_acc = merge_state<LP, 8>(dst, V, L, log, R.nohist, two_stage, (A.dbg & 8u) != 0u, A.counters, own_bar, R.boundary != 0);
    if (R.boundary && n_acc > 0) publish_ids<LP, 8>(V, log, n_acc, nidw, nid_own, tab, A, z);
  } else {
    ListView<LP, 2> V{smem_u32(srcg), smem_u32(mine), smem_u32(wrow), keep, appv, s, smem_u32(smem + kOneWord)};
    V.pin();
    n_acc = merge_state<LP, 2>(dst, V, L, log, R.nohist, two_stage, (A.dbg & 8u) != 0u, A.counters, own_bar, R.boundary != 0);
    if (R.boundary && n_acc > 0) publish_ids<LP, 2>(V, log, n_acc, nidw, nid_own, tab, A, z);
  }
  if (n_acc < 0) {  // metrics tie where libstdc++'s heap order could show: redo this state exactly
    const int n = exact_merge<LP>(src + (uint32_t)(src_g >> 3) * TileGeom<LP>::kPieceStride, mine, post, L, 8, 32, keep, appv, s, src_g & 7);
    atomicAdd(A.counters, 1ull);
    for (int l = 0; l < L; l++) {
      dst[l * 8] = mine[l * 32];
      dst[GP + l * 8] = mine[(LP + l) * 32];
      dst[2 * GP + l * 8] = mine[(2 * LP + l) * 32];
    }
    if (R.boundary)
      for (int l = 0; l < n; l++)
        nidw[l * 8] = intern_chunk(tab, A.tab_mask, mine[(LP + l) * 32], mine[(2 * LP + l) * 32], A.counters, z);
  } else {
    for (int l = n_acc; l < L; l++) __stcs(dst + l * 8, 0xff800000u);  // -inf; identity of a dead slot is never read
  }
}

// ---------------------------------------------------------------------------------------------
// lva_step_l1: the whole step for list size 1 in survivor mode (metrics only, VCC:715-742), the
// reference tool's default.  A list-size-1 state is one float: one thread owns a conv state
// (p, c): it reads the 8 metrics of the state itself and of each predecessor conv state straight
// from L2 (the state of every read in flight fits there), updates the up to 8 crf states in the
// reference's order (stay first, then predecessor crf ascending, first strict maximum) and writes
// the 8 four-bit decisions as ONE 32-bit word, coalesced over c.  grid = (band rows x conv
// states / 256, 1, reads); the per-read control data (band start, posterior block) comes from the
// tables the traceback uses.  Layout: l1_word() in lva_kernels.cuh, every row resident.
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
  const unsigned long long own = l1_word(K, p, c, 0);  // + s * 32
  float mine[8];
#pragma unroll
  for (int s = 0; s < 8; s++) mine[s] = __uint_as_float(prv[own + s * 32]);
  if (p == 0) {  // VCC:706-713: stay edge only
#pragma unroll
    for (int s = 0; s < 8; s++) cur[own + s * 32] = __float_as_uint(mine[s] + post[(s >= 4 ? 4 : s) * 8 + s]);
    return;
  }
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
    const unsigned long long srcb = l1_word(K, p - 1, cp, 0);
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
inline size_t fast_smem_bytes(uint32_t LP) {
  // predecessor tile (4 pieces + padding) + own (stay) lists of every thread: 3 planes each
  return 256 + (size_t)4 * (3 * 8 * LP * 8 + 8) * 4 + (size_t)3 * 8 * LP * 32 * 4;
}

#define LVA_FAST_LP_CASES(X) X(2) X(4) X(8) X(16)

inline bool fast_supported(uint32_t LP) { return LP >= 2 && LP <= 16; }

inline cudaError_t fast_prepare(uint32_t LP) {
  cudaError_t e = cudaErrorInvalidValue;
#define X(lp) \
  if (LP == lp) e = cudaFuncSetAttribute(lva_step_fast<lp>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast_smem_bytes(lp));
  LVA_FAST_LP_CASES(X)
#undef X
  return e;
}

struct FastMaps {
  CUtensorMap own;
};

// One time step of every read in flight: one launch.
inline cudaError_t fast_launch_step(uint32_t LP, cudaStream_t st, const StepArgs &a, const CodeSpec &k, uint32_t span,
                                    uint32_t active, const FastMaps &maps, bool all_one_bit) {
  dim3 grid(k.nconv / 32, span, active);
  // rows entered by a one-bit step have 64 + 64 live targets per tile (warps 0-3): when a launch
  // holds nothing else (every row of a rate-1/2 code), half the CTA would sit idle
  const uint32_t threads = all_one_bit ? 128 : kFastThreads;
  cudaError_t e = cudaErrorInvalidValue;
#define X(lp) \
  if (LP == lp) { lva_step_fast<lp><<<grid, threads, fast_smem_bytes(lp), st>>>(a, maps.own); e = cudaGetLastError(); }
  LVA_FAST_LP_CASES(X)
#undef X
  return e;
}

}  // namespace lva
