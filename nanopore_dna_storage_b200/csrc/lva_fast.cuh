// Register fast path of the trellis step, specialised on list slots (LP) and on the number of
// message-history words of the target row (NWD).
//
// Same CTA geometry as lva_step_generic with TG = 32 (one HBM tile of predecessors per CTA):
//   * the predecessor tile is ONE contiguous chunk -> a single TMA bulk copy (cp.async.bulk,
//     SASS UBLKCP) into shared memory, completion on an mbarrier; the step's 40 transition
//     scores ride on the same barrier;
//   * each thread owns one target state (p, c, s): warps are role-homogeneous (4 flip warps,
//     4 flop warps) and the 32 lanes of a warp read 32 DIFFERENT predecessor conv states, which
//     in the tile-transposed layout are 32 different banks -> no bank conflicts whatever the
//     (crf, slot, plane) pattern;
//   * the new list is built in registers by a k-way merge over the list heads with
//     de-duplication on the message history; whenever two heads tie exactly for the maximum
//     the thread re-runs that state through exact_merge(), which reproduces libstdc++'s heap
//     order (VCC:762-795).  Without a tie the pop sequence of any correct k-way merge equals
//     the reference's, so the fast path is bit-identical by construction.
#pragma once
#include "lva_kernels.cuh"

namespace lva {

template <int LP, int NWD>
struct OutList {
  float sc[LP];
  uint32_t id[LP][NWD > 0 ? NWD : 1];
};

// k-way merge of NL lists into `out`.
//   NL = 8 (flip target s < 4): list i is predecessor crf i, except list s which is the state's
//                               own (stay) list; all weights come from posterior row s.
//   NL = 2 (flop target s >= 4): list 0 = own list (column s of row 4), list 1 = the flip state
//                               s-4 of the predecessor conv state (column s-4 of row 4).
// Returns false if two heads tie exactly for the maximum (caller falls back to exact_merge).
template <int LP, int NWD, int NL>
__device__ __forceinline__ bool merge_lists(OutList<LP, NWD> &out, const uint32_t *srcg, const uint32_t *mine,
                                            const float *wrow, int L, int nws, int sh, uint32_t drop, int s,
                                            bool has_src) {
  constexpr int NT = 256, TGc = 32;
  const float NEG = neg_inf();
  const int plane_s = LP * TGc, list_s = (1 + nws) * plane_s;
  const int stay_i = NL == 8 ? s : 0;
#define LVA_CRF_OF(i) (NL == 8 ? (i) : ((i) == 0 ? s : s - 4))

  float h[NL];
#pragma unroll
  for (int i = 0; i < NL; i++) {
    float sc = NEG;
    if (i == stay_i) sc = __uint_as_float(mine[0]);
    else if (has_src) sc = __uint_as_float(srcg[LVA_CRF_OF(i) * list_s]);
    h[i] = sc + wrow[LVA_CRF_OF(i)];
  }
  uint32_t ptrs = 0;  // 4-bit read position per list
  bool done = false;
#pragma unroll
  for (int l = 0; l < LP; l++) {
    out.sc[l] = NEG;
    if (!done && l < L) {
      bool accepted = false;
      while (!accepted) {
        float best = h[0];
#pragma unroll
        for (int i = 1; i < NL; i++) best = fmaxf(best, h[i]);
        if (best == NEG) { done = true; break; }
        uint32_t mask = 0;
#pragma unroll
        for (int i = 0; i < NL; i++) mask |= (h[i] == best) ? (1u << i) : 0u;
        if (mask & (mask - 1u)) return false;  // exact tie between two heads
        const int i = __ffs(mask) - 1;
        const int j = (ptrs >> (4 * i)) & 15;
        const bool is_stay = i == stay_i;
        const int kc = LVA_CRF_OF(i);
        const uint32_t *base = is_stay ? mine : srcg + kc * list_s;
        const int slot_s = is_stay ? NT : TGc, pl_s = is_stay ? LP * NT : plane_s;
        const int sh_e = is_stay ? 0 : sh;
        const uint32_t drop_e = is_stay ? 0u : drop;
        const bool top_word = is_stay || nws == NWD;  // does this list carry NWD words already?
        // candidate history
        uint32_t cand[NWD > 0 ? NWD : 1];
        if (NWD > 0) {
          uint32_t old[NWD > 0 ? NWD : 1];
#pragma unroll
          for (int w = 0; w < NWD; w++) {
            old[w] = 0;
            if (w < NWD - 1 || top_word) old[w] = base[(1 + w) * pl_s + j * slot_s];
          }
          cand[0] = (old[0] << sh_e) | drop_e;
#pragma unroll
          for (int w = 1; w < NWD; w++) cand[w] = __funnelshift_l(old[w - 1], old[w], sh_e);
        }
        // advance list i
        const int jn = j + 1;
        float nh = NEG;
        if (jn < L) {
          nh = __uint_as_float(base[jn * slot_s]) + wrow[kc];
          ptrs += 1u << (4 * i);
        }
#pragma unroll
        for (int i2 = 0; i2 < NL; i2++) h[i2] = (i2 == i) ? nh : h[i2];
        // de-duplicate against what is already in the new list (VCC:778-779)
        bool dup = false;
        if (NWD == 0) {
          dup = l > 0;
        } else {
#pragma unroll
          for (int a = 0; a < l; a++) {
            uint32_t diff = 0;
#pragma unroll
            for (int w = 0; w < NWD; w++) diff |= cand[w] ^ out.id[a][w];
            dup = dup || (diff == 0);
          }
        }
        if (!dup) {
          out.sc[l] = best;
#pragma unroll
          for (int w = 0; w < NWD; w++) out.id[l][w] = cand[w];
          accepted = true;
          if (NWD == 0) done = true;  // every further candidate carries the same (empty) history
        } else if (NWD == 0) {
          done = true;
          break;
        }
      }
    }
  }
#undef LVA_CRF_OF
  return true;
}

template <int LP, int NWD>
__global__ void __launch_bounds__(256) lva_step_fast(const StepArgs A) {
  constexpr int NT = 256, TG = 32;
  extern __shared__ __align__(128) uint32_t smem[];
  const CodeDev &K = c_code[A.code_slot];
  const ReadDev rd = A.reads[blockIdx.z];
  if (A.t >= rd.T) return;
  const uint32_t a = A.band[rd.band_off + A.t];
  uint32_t bend = a + A.dev2;
  if (bend > K.nrows) bend = K.nrows;
  if (bend > A.row_hi) bend = A.row_hi;
  const uint32_t p = (a > A.row_lo ? a : A.row_lo) + blockIdx.y;
  if (p >= bend) return;

  const RowDev R = K.row[p];
  const int tid = threadIdx.x;
  const uint32_t tile = blockIdx.x;
  const Target T = locate_target<TG>(K, R, p, tile, tid);

  uint64_t *bar = reinterpret_cast<uint64_t *>(smem);          // bytes 0..7
  float *post = reinterpret_cast<float *>(smem + 4);           // bytes 16..175
  uint32_t *src = smem + 64;                                   // byte 256
  const int nws = p > 0 ? K.row[p - 1].idwords : 0;
  uint32_t *own = src + 8 * (1 + nws) * LP * TG;
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (!__syncthreads_or(T.valid)) return;

  const int L = (int)K.L;
  uint32_t *prv = A.state + ((unsigned long long)blockIdx.z * 2 + (A.t & 1u)) * A.slot_words;
  uint32_t *cur = A.state + ((unsigned long long)blockIdx.z * 2 + ((A.t + 1u) & 1u)) * A.slot_words;

  // ---- stage ----
  if (tid == 0) {
    const uint32_t tile_bytes = p > 0 ? 8u * (1 + nws) * LP * TG * 4u : 0u;
    mbar_expect_tx(bar, tile_bytes + 160u);
    bulk_g2s(post, A.post + (rd.post_block + A.t) * 40, 160u, bar);
    if (p > 0) {
      const RowDev S = K.row[p - 1];
      bulk_g2s(src, prv + S.off + (unsigned long long)tile * S.tile_words, tile_bytes, bar);
    }
  }
  const unsigned long long own_base = R.off + (unsigned long long)(T.c >> 5) * R.tile_words + (T.c & 31u) +
                                      (unsigned long long)T.s * (1 + NWD) * LP * 32;
  uint32_t *mine = own + tid;
  if (T.valid) {
#pragma unroll
    for (int i = 0; i < (1 + NWD) * LP; i++) cp_async4(mine + i * NT, prv + own_base + (unsigned long long)i * 32);
  }
  cp_async_wait_all();
  mbar_wait(bar, 0);
  if (!T.valid) return;

  OutList<LP, NWD> out;
  bool in_regs = true;
  if (p == 0) {  // VCC:706-713
    const int row = T.s >= 4 ? 4 : T.s;
#pragma unroll
    for (int l = 0; l < LP; l++) out.sc[l] = neg_inf();
    out.sc[0] = __uint_as_float(mine[0]) + post[row * 8 + T.s];
  } else {
    const int row = T.s >= 4 ? 4 : T.s;
    const float *wrow = post + row * 8;
    const uint32_t *srcg = src + (T.src_g >= 0 ? T.src_g : 0);
    const bool has_src = T.src_g >= 0;
    bool ok;
    if (LP == 1) {
      // VCC:715-742: first strict maximum, own list first then predecessor crf ascending
      float best = neg_inf();
      int bk = -1;
      const float s0 = __uint_as_float(mine[0]) + wrow[T.s];
      if (s0 > best) { best = s0; bk = T.s; }
      const int list_s = (1 + nws) * TG;
      if (has_src) {
        const int k0 = T.s < 4 ? 0 : T.s - 4, k1 = T.s < 4 ? 7 : T.s - 4;
        for (int k = k0; k <= k1; k++) {
          if (k == T.s) continue;
          const float sc = __uint_as_float(srcg[k * list_s]) + wrow[k];
          if (sc > best) { best = sc; bk = k; }
        }
      }
      out.sc[0] = best;
      if (NWD > 0 && bk >= 0) {
        if (bk == T.s) {
#pragma unroll
          for (int w = 0; w < NWD; w++) out.id[0][w] = mine[(1 + w) * NT];
        } else {
          uint32_t prev = 0;
#pragma unroll
          for (int w = 0; w < NWD; w++) {
            uint32_t old = 0;
            if (w < nws) old = srcg[bk * list_s + (1 + w) * TG];
            out.id[0][w] = (old << R.shift) | (w == 0 ? T.drop : (prev >> (32 - R.shift)));
            prev = old;
          }
        }
      } else if (NWD > 0) {
#pragma unroll
        for (int w = 0; w < NWD; w++) out.id[0][w] = mine[(1 + w) * NT];  // metric -inf: history is don't-care
      }
      ok = true;
    } else if (T.s < 4) {
      ok = merge_lists<LP, NWD, 8>(out, srcg, mine, wrow, L, nws, R.shift, T.drop, T.s, has_src);
    } else {
      ok = merge_lists<LP, NWD, 2>(out, srcg, mine, wrow, L, nws, R.shift, T.drop, T.s, has_src);
    }
    if (!ok) {
      exact_merge<LP>(src, mine, post, L, nws, NWD, TG, NT, R.shift, T.drop, T.s, T.src_g);
      atomicAdd(A.counters, 1ull);
      in_regs = false;
    }
  }
  uint32_t *dst = cur + own_base;
  if (in_regs) {
#pragma unroll
    for (int l = 0; l < LP; l++) {
      if (l < L) {
        dst[l * 32] = __float_as_uint(out.sc[l]);
#pragma unroll
        for (int w = 0; w < NWD; w++) dst[((1 + w) * LP + l) * 32] = out.id[l][w];
      }
    }
  } else {
    for (int l = 0; l < L; l++)
      for (int w = 0; w < 1 + NWD; w++) dst[(w * LP + l) * 32] = mine[(w * LP + l) * NT];
  }
}

// ---------------------------------------------------------------------------------------------
// host-side dispatch
// ---------------------------------------------------------------------------------------------
inline size_t fast_smem_bytes(uint32_t LP, uint32_t nwd) {
  // worst case: predecessors carry as many words as the target row
  return 256 + ((size_t)8 * (1 + nwd) * LP * 32 + (size_t)(1 + nwd) * LP * 256) * 4;
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
  lva_step_fast<LP, NWD><<<grid, 256, fast_smem_bytes(LP, NWD), st>>>(a);
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
    dim3 grid(k.nconv / 32, q - p, active);
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
