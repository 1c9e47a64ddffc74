// C ABI of liblva_b200.so (include/lva_b200.h): handle management, batch scheduling, and the
// host-side tail of the decode (final std::sort + bit extraction, VCC:806-836, kept on the host
// so that tie order is libstdc++'s by construction -- SURVEY.md §8a row A7).
// VCC = /root/reference/viterbi/viterbi_convolutional_code.cpp.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <mutex>
#include <numeric>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/lva_b200.h"
#include "code_tables.hpp"
#include "lva_kernels.cuh"
#include "lva_fast.cuh"

using namespace lva;

namespace {

std::mutex g_slot_mutex;
bool g_slot_used[16][kCodeSlots];  // [device][constant-memory code slot]

struct PinnedBuf {
  void *p = nullptr;
  size_t cap = 0;
  bool ensure(size_t n) {
    if (n <= cap) return true;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    size_t want = std::max(n, (size_t)1 << 16);
    if (cudaMallocHost(&p, want) != cudaSuccess) { p = nullptr; return false; }
    cap = want;
    return true;
  }
  ~PinnedBuf() { if (p) cudaFreeHost(p); }
};
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  bool ensure(size_t n) {
    if (n <= cap) return true;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    if (cudaMalloc(&p, n) != cudaSuccess) { p = nullptr; return false; }
    cap = n;
    return true;
  }
  ~DevBuf() { if (p) cudaFree(p); }
};

}  // namespace

struct lva_handle {
  lva_params prm{};
  std::string marker;
  CodeSpec code;
  CodeDev dev_code{};
  int device = 0, code_slot = -1;
  uint32_t L = 1, LP = 1, max_dev = 0, nw_last = 0;
  uint32_t n_slots = 0;
  unsigned long long slot_words = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[8] = {};
  DevBuf d_state, d_reads, d_band, d_post, d_final, d_counters;
  CUtensorMap own_maps[kMaxIdWords + 1];  // per history width: view of d_state for the own-list TMA copies
  bool own_tma = false;
  PinnedBuf h_stage, h_final, h_post;
  std::string err;
  lva_timing timing{};
  // CRC request of the call in progress (lva_decode_batch_crc)
  int crc_pad = -1;
  uint32_t crc_num_oligos = 0;
  uint8_t *crc_ok_out = nullptr;
  uint16_t *crc_index_out = nullptr;
  int32_t *crc_choice_out = nullptr;
  bool survivor = false;     // list size 1: metrics + 4-bit decisions + traceback instead of histories
  DevBuf d_dec, d_trace;
  PinnedBuf h_trace;
  int gen_tg = 32;           // tile width of the generic kernel
  size_t gen_smem = 0;
  bool use_fast = false;
};

namespace {

#define CK(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                        \
      return e_ == cudaErrorMemoryAllocation ? LVA_ERR_OOM : LVA_ERR_CUDA;                \
    }                                                                                     \
  } while (0)

uint32_t next_pow2(uint32_t v) {
  uint32_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

// host mirror of emitted_base(): used once per handle to verify the butterfly structure the
// kernels rely on (every base has at most one conv predecessor)
uint32_t host_emitted_base(const CodeSpec &k, uint32_t blk, uint32_t c_src, uint32_t in1, uint32_t in2) {
  const uint32_t r = k.rc ? 1u : 0u;
  auto par = [](uint32_t v) { return (uint32_t)__builtin_parity(v); };
  const uint32_t w1 = c_src | (in1 << k.mem);
  const uint32_t o0 = par(w1 & k.G[0]) ^ r, o1 = par(w1 & k.G[1]) ^ r;
  if (blk == 0) return 2 * o0 + o1;
  const uint32_t w2 = (w1 >> 1) | (in2 << k.mem);
  const uint32_t o2 = par(w2 & k.G[0]) ^ r, o3 = par(w2 & k.G[1]) ^ r;
  if (blk == 1) return r ? 2 * o2 + o1 : 2 * o1 + o2;
  if (blk == 2) return r ? 2 * o3 + o0 : 2 * o0 + o3;
  return r ? 2 * o3 + o1 : 2 * o1 + o3;
}

bool butterfly_structure_ok(const CodeSpec &k) {
  for (uint32_t blk = 0; blk < 4; blk++) {
    bool used = false;
    for (uint32_t i = 0; i < k.punc_len; i++) used |= k.punc[i] == blk;
    if (!used) continue;
    const uint32_t sh = blk == 0 ? 1 : 2;
    for (uint32_t c = 0; c < k.nconv; c++) {
      const uint32_t run = c >> (k.mem - sh), lo = c & ((1u << (k.mem - sh)) - 1);
      uint32_t seen = 0;
      for (uint32_t xy = 0; xy < (1u << sh); xy++) {
        const uint32_t b = host_emitted_base(k, blk, (lo << sh) | xy, sh == 1 ? run : (run & 1), run >> 1);
        if (seen & (1u << b)) return false;
        seen |= 1u << b;
      }
    }
  }
  return true;
}

template <int LP, int TG>
cudaError_t launch_generic(const lva_handle *h, dim3 grid, const StepArgs &a) {
  lva_step_generic<LP, TG><<<grid, TG * 8, h->gen_smem, h->stream>>>(a);
  return cudaGetLastError();
}
template <int LP, int TG>
cudaError_t prep_generic(size_t smem) {
  return cudaFuncSetAttribute(lva_step_generic<LP, TG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

#define LVA_GENERIC_CASES(X) \
  X(1, 32) X(2, 32) X(4, 32) X(8, 32) X(16, 32) X(16, 16) X(32, 16) X(32, 8) X(64, 8) X(64, 4)

cudaError_t generic_prepare(const lva_handle *h) {
#define X(lp, tg) \
  if (h->LP == lp && h->gen_tg == tg) return prep_generic<lp, tg>(h->gen_smem);
  LVA_GENERIC_CASES(X)
#undef X
  return cudaErrorInvalidValue;
}
cudaError_t generic_launch(const lva_handle *h, dim3 grid, const StepArgs &a) {
#define X(lp, tg) \
  if (h->LP == lp && h->gen_tg == tg) return launch_generic<lp, tg>(h, grid, a);
  LVA_GENERIC_CASES(X)
#undef X
  return cudaErrorInvalidValue;
}

size_t generic_smem_bytes(uint32_t LP, uint32_t tg, uint32_t nw) {
  return 256 + ((size_t)8 * (1 + nw) * LP * tg + (size_t)(1 + nw) * LP * tg * 8) * 4;
}

struct FinalEntry {
  float score;
  uint32_t id[kMaxIdWords];
  uint32_t crcw;
};

}  // namespace

extern "C" {

const char *lva_version(void) { return "lva_b200 0.1 (sm_100a)"; }

// Tensor maps for the own-list copies of lva_step_fast: the state buffer as a 3-D tensor
// (32 conv columns | 8 crf lists, ROWS*128 bytes apart | 128-byte rows), box {8, 4, ROWS}; ROWS depends on
// the history width of the row class, hence one map per width.  cuTensorMapEncodeTiled comes from the
// driver through the runtime's entry-point query (no link against libcuda).  Failure -> cp.async path.
static bool make_own_maps(lva_handle *h) {
  typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
      qres != cudaDriverEntryPointSuccess)
    return false;
  const unsigned long long total_rows = (unsigned long long)h->n_slots * 2ull * h->slot_words / 32ull;
  if (total_rows == 0 || total_rows > 0xffffffffull) return false;
  for (uint32_t nw = 0; nw <= h->nw_last; nw++) {
    const cuuint32_t rows = (1u + nw) * h->LP;
    if (rows > 256) return false;
    const cuuint64_t dims[3] = {32, 8, total_rows};
    const cuuint64_t strides[2] = {(cuuint64_t)rows * 128ull, 128ull};  // bytes, dims 1 and 2
    const cuuint32_t box[3] = {8, 4, rows};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (((encode_fn)fn)(&h->own_maps[nw], CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, h->d_state.p, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return false;
  }
  return true;
}

lva_handle *lva_create(const lva_params *params, int *status, char *err, size_t err_len) {
  auto fail = [&](int st, const std::string &msg) -> lva_handle * {
    if (status) *status = st;
    if (err && err_len) snprintf(err, err_len, "%s", msg.c_str());
    return nullptr;
  };
  if (!params) return fail(LVA_ERR_BAD_ARG, "null params");
  lva_handle *h = new lva_handle();
  h->prm = *params;
  h->marker = params->sync_marker ? params->sync_marker : "";
  h->prm.sync_marker = nullptr;
  int st = build_code_spec(h->code, params->mem_conv, params->rate, params->msg_len, params->rc != 0,
                           h->marker, params->sync_period);
  if (st != kCodeOk) { delete h; return fail(st, code_status_text(st)); }
  const CodeSpec &k = h->code;
  if ((uint64_t)k.nrows * 8 * k.nconv >= ((uint64_t)1 << 32)) {
    delete h;
    return fail(LVA_ERR_TOO_MANY_STATES, "Too many states, can't fit in 32 bits");
  }
  if (params->msg_len > 256 || params->msg_len + params->mem_conv > 256) {
    delete h;
    return fail(LVA_ERR_MSG_TOO_LONG, "msg_len can't be above BITSET_SIZE");
  }
  if (params->list_size == 0 || params->list_size > LVA_MAX_LIST) {
    delete h;
    return fail(LVA_ERR_LIST_SIZE, "list_size must be in 1..64");
  }
  if (!butterfly_structure_ok(k)) {
    delete h;
    return fail(LVA_ERR_CODE_STRUCTURE, "code violates the one-predecessor-per-base structure");
  }
  h->L = params->list_size;
  h->LP = next_pow2(h->L);
  h->max_dev = params->max_deviation == LVA_DEFAULT_DEVIATION ? params->msg_len + params->mem_conv + 1
                                                              : params->max_deviation;
  h->nw_last = k.idwords[k.nrows - 1];
  // List size 1 needs no message identity at all (VCC:715-742 has no de-duplication): keep metrics
  // only, record one 4-bit decision per state and step, rebuild the message by traceback.
  h->survivor = h->L == 1 && !(params->flags & LVA_FLAG_EXACT_ONLY);
  if (h->survivor) {
    h->nw_last = 0;
    for (uint32_t p = 0; p < k.nrows; p++) h->code.idwords[p] = 0;
  }

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    delete h;
    return fail(LVA_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU path)");
  }
  h->device = params->device;
  if (h->device < 0 || h->device >= ndev || h->device >= 16) {
    delete h;
    return fail(LVA_ERR_NO_DEVICE, "bad device ordinal");
  }
  auto cuda_fail = [&](const char *what, cudaError_t e) -> lva_handle * {
    std::string m = std::string(what) + ": " + cudaGetErrorString(e);
    lva_destroy(h);
    return fail(e == cudaErrorMemoryAllocation ? LVA_ERR_OOM : LVA_ERR_CUDA, m);
  };
  cudaError_t e;
  if ((e = cudaSetDevice(h->device)) != cudaSuccess) return cuda_fail("cudaSetDevice", e);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, h->device)) != cudaSuccess) return cuda_fail("props", e);
  h->sm_count = prop.multiProcessorCount;
  if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess)
    return cuda_fail("cudaStreamCreate", e);
  for (auto &ev : h->ev)
    if ((e = cudaEventCreate(&ev)) != cudaSuccess) return cuda_fail("cudaEventCreate", e);

  // ---- device code tables (constant memory slot) ----
  {
    std::lock_guard<std::mutex> g(g_slot_mutex);
    for (int i = 0; i < kCodeSlots; i++)
      if (!g_slot_used[h->device][i]) { g_slot_used[h->device][i] = true; h->code_slot = i; break; }
  }
  if (h->code_slot < 0) { lva_destroy(h); return fail(LVA_ERR_BAD_ARG, "too many live handles on this device"); }
  CodeDev &d = h->dev_code;
  d.mem = k.mem; d.nconv = k.nconv; d.G0 = k.G[0]; d.G1 = k.G[1]; d.init = k.init; d.fin = k.fin;
  d.rc = k.rc ? 1 : 0; d.nrows = k.nrows; d.L = h->L; d.LP = h->LP; d.msg_len = k.msg_len; d.pad = 0;
  unsigned long long off = 0;
  for (uint32_t p = 0; p < k.nrows; p++) {
    RowDev &r = d.row[p];
    r.off = off;
    r.vmask = k.vmask[p]; r.vval = k.vval[p];
    const uint32_t idw = h->survivor ? 0u : k.idwords[p];
    r.tile_words = 8u * (1u + idw) * h->LP * 32u;
    r.block = k.block[p]; r.shift = k.shift[p]; r.idwords = (uint8_t)idw; r.dropmask = k.dropmask[p];
    r.inv_lin = 0; r.lin1 = 0; r.pad0 = r.pad1 = 0; r.pad2 = 0;
    if (p > 0) {  // linear map low predecessor bits -> emitted base (verified bijective by butterfly_structure_ok)
      const uint32_t b0 = host_emitted_base(k, k.block[p], 0, 0, 0);
      for (uint32_t xy = 0; xy < (1u << k.shift[p]); xy++) {
        const uint32_t t = host_emitted_base(k, k.block[p], xy, 0, 0) ^ b0;
        r.inv_lin |= (uint8_t)(xy << (2 * t));
        if (xy == 1) r.lin1 = (uint8_t)t;
      }
    }
    off += (unsigned long long)(k.nconv / kTileGroups) * r.tile_words;
  }
  h->slot_words = off;
  if ((e = cudaMemcpyToSymbol(c_code, &d, sizeof(CodeDev), sizeof(CodeDev) * h->code_slot)) != cudaSuccess)
    return cuda_fail("cudaMemcpyToSymbol", e);

  // ---- kernel selection ----
  h->gen_tg = 0;
  for (int tg : {32, 16, 8, 4}) {
    bool have = false;
#define X(lp, tgc) if (h->LP == lp && tg == tgc) have = true;
    LVA_GENERIC_CASES(X)
#undef X
    if (!have) continue;
    size_t sm = generic_smem_bytes(h->LP, tg, h->nw_last);
    if (sm <= (size_t)prop.sharedMemPerBlockOptin) { h->gen_tg = tg; h->gen_smem = sm; break; }
  }
  if (!h->gen_tg) { lva_destroy(h); return fail(LVA_ERR_LIST_SIZE, "list_size x msg_len too large for shared memory"); }
  if ((e = generic_prepare(h)) != cudaSuccess) return cuda_fail("cudaFuncSetAttribute(generic)", e);
  h->use_fast = !(params->flags & LVA_FLAG_EXACT_ONLY) && fast_supported(h->LP, h->nw_last);
  if (h->use_fast && (e = fast_prepare(h->LP, h->nw_last)) != cudaSuccess)
    return cuda_fail("cudaFuncSetAttribute(fast)", e);

  // ---- state memory: n_slots reads in flight, two parity buffers each ----
  size_t free_b = 0, total_b = 0;
  if ((e = cudaMemGetInfo(&free_b, &total_b)) != cudaSuccess) return cuda_fail("cudaMemGetInfo", e);
  const size_t per_slot = (size_t)h->slot_words * 2 * 4;
  size_t budget = (size_t)(free_b * 0.70);
  // default: enough reads in flight that one step launch is >= ~64 waves of CTAs (small codes are launch-bound otherwise), at least 64
  uint32_t slots = params->max_slots;
  if (!slots) {
    const uint64_t band_rows = std::min<uint64_t>(2ull * h->max_dev, k.nrows);
    const uint64_t ctas_per_read = std::max<uint64_t>(1, band_rows * 2 * (k.nconv / 32));
    slots = (uint32_t)std::min<uint64_t>(kMaxSlots, std::max<uint64_t>(64, (64ull * 4 * h->sm_count + ctas_per_read - 1) / ctas_per_read));
  }
  size_t per_slot_all = per_slot;
  if (h->survivor) {
    // list size 1: a step of one read is a few microseconds of work -> as many reads in flight as fit
    // next to their decision arrays (T x band rows x conv states words; T estimated at 5 blocks per row)
    const uint64_t band_rows = std::min<uint64_t>(std::max<uint64_t>(2ull * h->max_dev, 1), k.nrows);
    per_slot_all += (size_t)(5ull * k.nrows * band_rows * k.nconv * 4ull);
    if (!params->max_slots) slots = kMaxSlots;
  }
  if (slots > (uint32_t)kMaxSlots) slots = kMaxSlots;
  if ((size_t)slots * per_slot_all > budget) slots = (uint32_t)std::max<size_t>(1, budget / per_slot_all);
  if ((size_t)slots * per_slot > budget) slots = (uint32_t)(budget / per_slot);
  if (slots == 0) { lva_destroy(h); return fail(LVA_ERR_OOM, "not enough device memory for one read"); }
  h->n_slots = slots;
  if (!h->d_state.ensure((size_t)slots * per_slot)) { lva_destroy(h); return fail(LVA_ERR_OOM, "state allocation failed"); }
  if (!h->d_reads.ensure(sizeof(ReadDev) * slots) || !h->d_counters.ensure(64) ||
      !h->d_final.ensure((size_t)slots * 8 * h->LP * (2 + h->nw_last) * 4)) {
    lva_destroy(h);
    return fail(LVA_ERR_OOM, "allocation failed");
  }
  cudaMemset(h->d_counters.p, 0, 64);
  h->own_tma = h->use_fast && h->LP > 1 && getenv("LVA_NO_OWN_TMA") == nullptr && make_own_maps(h);
  if (status) *status = LVA_OK;
  return h;
}

void lva_destroy(lva_handle *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (auto &ev : h->ev) if (ev) cudaEventDestroy(ev);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->code_slot >= 0) {
    std::lock_guard<std::mutex> g(g_slot_mutex);
    g_slot_used[h->device][h->code_slot] = false;
  }
  delete h;
}

const char *lva_last_error(const lva_handle *h) { return h ? h->err.c_str() : "null handle"; }
uint32_t lva_nstate_pos(const lva_handle *h) { return h ? h->code.nrows : 0; }
uint32_t lva_slots(const lva_handle *h) { return h ? h->n_slots : 0; }

int lva_count_work(const lva_handle *h, uint32_t T, lva_work *out) {
  if (!h || !out) return LVA_ERR_BAD_ARG;
  const CodeSpec &k = h->code;
  out->state_updates = out->edge_acs = out->state_bytes = 0;
  for (uint32_t t = 0; t < T; t++) {
    const uint32_t a = band_start(t, T, k.nrows, h->max_dev), b = band_end(a, k.nrows, h->max_dev);
    for (uint32_t p = a; p < b; p++) {
      const double nst = std::ldexp(8.0, (int)k.mem - __builtin_popcount(k.vmask[p]));
      out->state_updates += nst;
      out->edge_acs += nst * (p == 0 ? 1.0 : (k.block[p] == 0 ? 3.0 : 5.0));
      // one read of the previous list + one write of the new list, LP slots of (1+nw) words
      out->state_bytes += nst * 2.0 * 4.0 * h->LP * (1 + k.idwords[p]) + (h->survivor ? nst * 0.5 : 0.0);
    }
  }
  return LVA_OK;
}

int lva_get_timing(const lva_handle *h, lva_timing *out) {
  if (!h || !out) return LVA_ERR_BAD_ARG;
  *out = h->timing;
  return LVA_OK;
}

int lva_check_params(uint32_t mem_conv, uint32_t rate, uint32_t msg_len) {
  CodeSpec k;
  return build_code_spec(k, mem_conv, rate, msg_len, false, "", 0);
}

int lva_get_code_tables(uint32_t mem_conv, uint32_t rate, uint32_t msg_len, int32_t rc, const char *sync_marker,
                        uint32_t sync_period, lva_code_tables *out) {
  if (!out) return LVA_ERR_BAD_ARG;
  CodeSpec k;
  int st = build_code_spec(k, mem_conv, rate, msg_len, rc != 0, sync_marker ? sync_marker : "", sync_period);
  if (st != kCodeOk) return st;
  memset(out, 0, sizeof(*out));
  out->G[0] = k.G[0]; out->G[1] = k.G[1];
  out->initial_state = k.init; out->final_state = k.fin;
  out->punc_len = k.punc_len; out->nstate_pos = k.nrows;
  for (uint32_t i = 0; i < k.punc_len; i++) out->punc[i] = k.punc[i];
  for (uint32_t p = 0; p < k.nrows; p++) {
    out->pos2msg[p] = k.bits_at[p];
    out->valid_mask[p] = k.vmask[p];
    out->valid_value[p] = k.vval[p];
    out->history_words[p] = k.idwords[p];
  }
  return LVA_OK;
}

void lva_band(uint32_t t, uint32_t nblk, uint32_t nstate_pos, uint32_t max_deviation, uint32_t *start,
              uint32_t *end) {
  const uint32_t a = band_start(t, nblk, nstate_pos, max_deviation);
  if (start) *start = a;
  if (end) *end = band_end(a, nstate_pos, max_deviation);
}

int lva_encode(uint32_t mem_conv, uint32_t rate, const uint8_t *msg, uint32_t msg_len, uint8_t *out_bits,
               uint32_t out_cap, uint32_t *n_out) {
  CodeSpec k;
  int st = build_code_spec(k, mem_conv, rate, msg_len, false, "", 0);
  if (st != kCodeOk) return st;
  std::vector<uint8_t> m(msg, msg + msg_len), enc;
  if (!conv_encode(k, m, enc)) return LVA_ERR_BAD_ARG;
  if (enc.size() > out_cap) return LVA_ERR_BAD_ARG;
  std::copy(enc.begin(), enc.end(), out_bits);
  if (n_out) *n_out = (uint32_t)enc.size();
  return LVA_OK;
}

static int decode_core(lva_handle *h, const float *d_post, const uint64_t *offs, uint32_t n_reads,
                       uint8_t *out_bits, float *out_scores, uint32_t *out_counts, int32_t *out_status) {
  const CodeSpec &k = h->code;
  const uint32_t L = h->L, LP = h->LP, msg_len = k.msg_len, planes = 2 + h->nw_last;  // metric, history, CRC word
  const float NEG = -std::numeric_limits<float>::infinity();

  std::vector<uint32_t> order;
  order.reserve(n_reads);
  for (uint32_t i = 0; i < n_reads; i++) {
    const uint64_t T = offs[i + 1] - offs[i];
    out_counts[i] = 0;
    if (T < (uint64_t)k.nrows + 1 || T > 0xFFFFFFFFull) {  // VCC:600-601
      if (out_status) out_status[i] = LVA_ERR_POST_TOO_SMALL;
      continue;
    }
    if (out_status) out_status[i] = LVA_OK;
    order.push_back(i);
  }
  std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
    return offs[x + 1] - offs[x] > offs[y + 1] - offs[y];
  });

  const uint32_t dev2 = 2u * h->max_dev;
  float acc_kernel = 0.f, acc_d2h = 0.f;
  std::vector<FinalEntry> fin;
  const uint32_t dec_rows = (uint32_t)std::min<uint64_t>((uint64_t)dev2 > 0 ? dev2 : 1, k.nrows);
  size_t chunk_cap = h->n_slots;
  if (h->survivor && !order.empty()) {  // decisions: T x band rows x conv states words per read in flight
    const uint64_t Tlongest = offs[order[0] + 1] - offs[order[0]];
    const uint64_t per_read = Tlongest * dec_rows * k.nconv * 4ull;
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    const uint64_t budget = (uint64_t)((free_b + h->d_dec.cap) * 0.8);
    chunk_cap = (size_t)std::max<uint64_t>(1, std::min<uint64_t>(h->n_slots, budget / std::max<uint64_t>(per_read, 1)));
  }
  for (size_t c0 = 0; c0 < order.size(); c0 += chunk_cap) {
    const uint32_t ns = (uint32_t)std::min<size_t>(chunk_cap, order.size() - c0);
    // ---- per-slot control + band tables ----
    size_t band_total = 0;
    for (uint32_t s = 0; s < ns; s++) band_total += offs[order[c0 + s] + 1] - offs[order[c0 + s]];
    const size_t stage_bytes = sizeof(ReadDev) * ns + band_total * sizeof(uint16_t);
    if (!h->h_stage.ensure(stage_bytes) || !h->d_band.ensure(band_total * sizeof(uint16_t) + 16)) {
      h->err = "staging allocation failed";
      return LVA_ERR_OOM;
    }
    ReadDev *rd = reinterpret_cast<ReadDev *>(h->h_stage.p);
    uint16_t *band = reinterpret_cast<uint16_t *>(rd + ns);
    std::vector<uint32_t> Ts(ns);
    size_t bo = 0;
    for (uint32_t s = 0; s < ns; s++) {
      const uint32_t i = order[c0 + s];
      const uint32_t T = (uint32_t)(offs[i + 1] - offs[i]);
      Ts[s] = T;
      rd[s].post_block = offs[i];
      rd[s].T = T;
      rd[s].band_off = (uint32_t)bo;
      for (uint32_t t = 0; t < T; t++) band[bo + t] = (uint16_t)band_start(t, T, k.nrows, h->max_dev);
      bo += T;
    }
    CK(cudaMemcpyAsync(h->d_reads.p, rd, sizeof(ReadDev) * ns, cudaMemcpyHostToDevice, h->stream));
    h->timing.h2d_bytes += sizeof(ReadDev) * ns;
    unsigned long long dec_slot_words = 0;
    if (h->survivor) {
      CK(cudaMemcpyAsync(h->d_band.p, band, band_total * sizeof(uint16_t), cudaMemcpyHostToDevice, h->stream));
      h->timing.h2d_bytes += band_total * sizeof(uint16_t);
      dec_slot_words = (unsigned long long)Ts[0] * dec_rows * k.nconv;
      if (!h->d_dec.ensure((size_t)dec_slot_words * ns * 4) || !h->d_trace.ensure((size_t)ns * 12 * 4) ||
          !h->h_trace.ensure((size_t)ns * 12 * 4)) {
        h->err = "decision memory allocation failed";
        return LVA_ERR_OOM;
      }
    }

    // ---- initialise the slots ----
    const unsigned long long words = (unsigned long long)ns * 2 * h->slot_words;
    lva_fill_neg_inf<<<h->sm_count * 8, 256, 0, h->stream>>>((uint32_t *)h->d_state.p, words);
    lva_seed_start<<<(ns * 8 + 127) / 128, 128, 0, h->stream>>>((uint32_t *)h->d_state.p, h->slot_words,
                                                                  h->code_slot, ns);
    h->timing.launches += 2;
    CK(cudaGetLastError());

    // ---- time steps ----
    CK(cudaEventRecord(h->ev[2], h->stream));
    StepArgs a{};
    a.state = (uint32_t *)h->d_state.p;
    a.slot_words = h->slot_words;
    a.post = d_post;
    a.counters = (unsigned long long *)h->d_counters.p;
    a.dev2 = dev2;
    a.code_slot = h->code_slot;
    a.dec = h->survivor ? (uint32_t *)h->d_dec.p : nullptr;
    a.dec_slot_words = dec_slot_words;
    a.dec_rows = dec_rows;
    a.own_tma = h->own_tma ? 1u : 0u;
    // developer switch for A/B runs: list size 1 through the tile-staging kernel lva_step_fast<1,0>
    static const bool no_l1_kernel = getenv("LVA_NO_L1_KERNEL") != nullptr;
    const uint32_t Tmax = Ts[0];
    uint32_t active = ns;
    for (uint32_t t = 0; t < Tmax; t++) {
      while (active > 0 && Ts[active - 1] <= t) active--;  // reads are sorted by T, longest first
      // rows touched this step: union over active reads of [a, b)
      uint32_t lo = 0xffffffffu, hi = 0;
      for (uint32_t s = 0; s < active; s++) {
        const uint32_t st = band[rd[s].band_off + t], en = band_end(st, k.nrows, h->max_dev);
        lo = std::min(lo, st);
        hi = std::max(hi, en);
      }
      if (hi <= lo) continue;
      a.t = t;
      for (uint32_t s = 0; s < active; s++) {
        a.band_a[s] = band[rd[s].band_off + t];
        a.post_blk[s] = (uint32_t)(rd[s].post_block + t);
      }
      if (h->use_fast && h->survivor && !no_l1_kernel) {
        L1Args l1{a.state, a.slot_words, a.post, a.dec, a.dec_slot_words, (const ReadDev *)h->d_reads.p,
                  (const uint16_t *)h->d_band.p, a.dec_rows, t, a.dev2, a.code_slot};
        cudaError_t e = l1_launch_step(h->stream, l1, k, active);
        if (e != cudaSuccess) { h->err = std::string("L=1 step launch: ") + cudaGetErrorString(e); return LVA_ERR_CUDA; }
        h->timing.step_launches++;
      } else if (h->use_fast) {
        cudaError_t e = fast_launch_step(h->LP, h->stream, a, k, lo, hi, active, &h->timing.step_launches, h->own_maps);
        if (e != cudaSuccess) { h->err = std::string("fast step launch: ") + cudaGetErrorString(e); return LVA_ERR_CUDA; }
      } else {
        a.row_lo = 0;
        a.row_hi = k.nrows;
        uint32_t span = 0;
        for (uint32_t s = 0; s < active; s++) {
          const uint32_t st = band[rd[s].band_off + t];
          span = std::max(span, band_end(st, k.nrows, h->max_dev) - st);
        }
        dim3 grid(k.nconv / h->gen_tg, span, active);
        cudaError_t e = generic_launch(h, grid, a);
        if (e != cudaSuccess) { h->err = std::string("generic step launch: ") + cudaGetErrorString(e); return LVA_ERR_CUDA; }
        h->timing.step_launches++;
      }
    }
    CK(cudaEventRecord(h->ev[3], h->stream));

    // ---- collect the candidate lists of the end state ----
    FinalArgs f{};
    f.state = (const uint32_t *)h->d_state.p;
    f.slot_words = h->slot_words;
    f.reads = (const ReadDev *)h->d_reads.p;
    f.out = (uint32_t *)h->d_final.p;
    f.code_slot = h->code_slot;
    f.n_slots = ns;
    f.crc_pad = h->crc_pad;
    const uint32_t nfin = ns * 8 * LP;
    lva_collect_final<<<(nfin + 127) / 128, 128, 0, h->stream>>>(f);
    h->timing.launches += 1;
    CK(cudaGetLastError());
    if (h->survivor) {
      TraceArgs tr{};
      tr.state = (const uint32_t *)h->d_state.p;
      tr.slot_words = h->slot_words;
      tr.dec = (const uint32_t *)h->d_dec.p;
      tr.dec_slot_words = dec_slot_words;
      tr.dec_rows = dec_rows;
      tr.reads = (const ReadDev *)h->d_reads.p;
      tr.band = (const uint16_t *)h->d_band.p;
      tr.out = (uint32_t *)h->d_trace.p;
      tr.code_slot = h->code_slot;
      tr.n_slots = ns;
      tr.crc_pad = h->crc_pad;
      lva_traceback<<<(ns + 63) / 64, 64, 0, h->stream>>>(tr);
      h->timing.launches += 1;
      CK(cudaGetLastError());
      CK(cudaMemcpyAsync(h->h_trace.p, h->d_trace.p, (size_t)ns * 12 * 4, cudaMemcpyDeviceToHost, h->stream));
      h->timing.d2h_bytes += (size_t)ns * 12 * 4;
    }
    const size_t fin_bytes = (size_t)nfin * planes * 4;
    if (!h->h_final.ensure(fin_bytes)) { h->err = "pinned allocation failed"; return LVA_ERR_OOM; }
    CK(cudaEventRecord(h->ev[4], h->stream));
    CK(cudaMemcpyAsync(h->h_final.p, h->d_final.p, fin_bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaEventRecord(h->ev[5], h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->timing.d2h_bytes += fin_bytes;
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]);
    acc_kernel += ms;
    cudaEventElapsedTime(&ms, h->ev[4], h->ev[5]);
    acc_d2h += ms;

    // ---- host tail: sort by metric, keep L, unpack bits (VCC:806-836) ----
    const uint32_t *fw = reinterpret_cast<const uint32_t *>(h->h_final.p);
    for (uint32_t s = 0; s < ns; s++) {
      const uint32_t i = order[c0 + s];
      fin.clear();
      if (h->survivor) {  // one entry at most: the traced-back best path
        const uint32_t *tw = reinterpret_cast<const uint32_t *>(h->h_trace.p) + (size_t)s * 12;
        if (tw[1] == 2u) { h->err = "traceback did not reach the start row"; return LVA_ERR_CUDA; }
        if (tw[1] == 1u) {
          FinalEntry fe;
          memcpy(&fe.score, tw, 4);
          for (uint32_t w = 0; w < 8; w++) fe.id[w] = tw[4 + w];
          fe.crcw = tw[2];
          fin.push_back(fe);
        }
      } else
      for (uint32_t crf = 0; crf < 8; crf++)
        for (uint32_t j = 0; j < L; j++) {
          const uint32_t *e = fw + ((size_t)(s * 8 + crf) * LP + j) * planes;
          float sc;
          memcpy(&sc, e, 4);
          if (sc != NEG) {
            FinalEntry fe;
            fe.score = sc;
            for (uint32_t w = 0; w < h->nw_last; w++) fe.id[w] = e[1 + w];
            fe.crcw = e[1 + h->nw_last];
            fin.push_back(fe);
          }
        }
      std::sort(fin.begin(), fin.end(),
                [](const FinalEntry &p1, const FinalEntry &p2) -> bool { return p1.score > p2.score; });
      if (fin.size() > L) fin.resize(L);
      out_counts[i] = (uint32_t)fin.size();
      if (h->crc_pad >= 0) {  // helper.decode_list_CRC_index: first entry whose CRC matches and index < num_oligos
        int32_t choice = -1;
        for (size_t e = 0; e < fin.size(); e++) {
          const bool ok = (fin[e].crcw & kCrcOk) != 0;
          const uint32_t idx = fin[e].crcw & 0xfffu;
          if (h->crc_ok_out) h->crc_ok_out[(size_t)i * L + e] = ok ? 1 : 0;
          if (h->crc_index_out) h->crc_index_out[(size_t)i * L + e] = (uint16_t)idx;
          if (choice < 0 && ok && idx < h->crc_num_oligos) choice = (int32_t)e;
        }
        if (h->crc_choice_out) h->crc_choice_out[i] = choice;
      }
      for (size_t e = 0; e < fin.size(); e++) {
        if (out_scores) out_scores[(size_t)i * L + e] = fin[e].score;
        uint8_t *dst = out_bits + ((size_t)i * L + e) * msg_len;
        for (uint32_t b = 0; b < msg_len; b++) {
          const uint32_t bit = msg_len - 1 - b;  // history bit of message position b
          const uint8_t v = (uint8_t)((fin[e].id[bit >> 5] >> (bit & 31)) & 1u);
          dst[k.rc ? msg_len - 1 - b : b] = v;
        }
      }
    }
  }
  h->timing.kernel_ms = acc_kernel;
  h->timing.d2h_ms = acc_d2h;
  h->timing.launches += h->timing.step_launches;
  return LVA_OK;
}

static int decode_entry(lva_handle *h, const float *post, bool post_on_device, const uint64_t *offs,
                        uint32_t n_reads, uint8_t *out_bits, float *out_scores, uint32_t *out_counts,
                        int32_t *out_status) {
  if (!h) return LVA_ERR_BAD_ARG;
  if (n_reads == 0) return LVA_OK;
  if (!post || !offs || !out_bits || !out_counts) { h->err = "null argument"; return LVA_ERR_BAD_ARG; }
  if (offs[n_reads] > 0xFFFF0000ull) { h->err = "batch too large (block index must fit 32 bits)"; return LVA_ERR_BAD_ARG; }
  CK(cudaSetDevice(h->device));
  h->timing = lva_timing{};
  CK(cudaMemsetAsync(h->d_counters.p, 0, 64, h->stream));
  CK(cudaEventRecord(h->ev[0], h->stream));
  const float *d_post = post;
  if (!post_on_device) {
    const uint64_t b0 = offs[0], b1 = offs[n_reads];
    const size_t bytes = (size_t)(b1 - b0) * 160;
    if (!h->d_post.ensure(bytes + 16)) { h->err = "posterior device allocation failed"; return LVA_ERR_OOM; }
    // Staged so that read i's block b sits at d_post + (block_offsets[i] + b) * 40, as with device input
    CK(cudaMemcpyAsync(h->d_post.p, post + b0 * 40, bytes, cudaMemcpyHostToDevice, h->stream));
    d_post = reinterpret_cast<const float *>(h->d_post.p) - b0 * 40;
    h->timing.h2d_bytes += bytes;
  }
  CK(cudaEventRecord(h->ev[1], h->stream));
  int st = decode_core(h, d_post, offs, n_reads, out_bits, out_scores, out_counts, out_status);
  if (st != LVA_OK) return st;
  CK(cudaEventRecord(h->ev[6], h->stream));
  CK(cudaEventSynchronize(h->ev[6]));
  cudaEventElapsedTime(&h->timing.total_ms, h->ev[0], h->ev[6]);
  cudaEventElapsedTime(&h->timing.h2d_ms, h->ev[0], h->ev[1]);
  unsigned long long fb = 0;
  CK(cudaMemcpy(&fb, h->d_counters.p, sizeof(fb), cudaMemcpyDeviceToHost));
  h->timing.exact_fallbacks = fb;
  return LVA_OK;
}

int lva_decode_batch(lva_handle *h, const float *post, const uint64_t *block_offsets, uint32_t n_reads,
                     uint8_t *out_bits, float *out_scores, uint32_t *out_counts, int32_t *out_status) {
  return decode_entry(h, post, false, block_offsets, n_reads, out_bits, out_scores, out_counts, out_status);
}
int lva_decode_batch_crc(lva_handle *h, const float *post, const uint64_t *block_offsets, uint32_t n_reads,
                         int32_t pad, uint32_t num_oligos, uint8_t *out_bits, float *out_scores,
                         uint32_t *out_counts, int32_t *out_status, uint8_t *out_crc_ok, uint16_t *out_index,
                         int32_t *out_choice) {
  if (!h) return LVA_ERR_BAD_ARG;
  h->crc_pad = pad ? 1 : 0;
  h->crc_num_oligos = num_oligos;
  h->crc_ok_out = out_crc_ok;
  h->crc_index_out = out_index;
  h->crc_choice_out = out_choice;
  if (out_choice) for (uint32_t i = 0; i < n_reads; i++) out_choice[i] = -1;
  const int st = decode_entry(h, post, false, block_offsets, n_reads, out_bits, out_scores, out_counts, out_status);
  h->crc_pad = -1;
  h->crc_ok_out = nullptr;
  h->crc_index_out = nullptr;
  h->crc_choice_out = nullptr;
  return st;
}

int lva_decode_batch_device(lva_handle *h, const float *post_dev, const uint64_t *block_offsets,
                            uint32_t n_reads, uint8_t *out_bits, float *out_scores, uint32_t *out_counts,
                            int32_t *out_status) {
  return decode_entry(h, post_dev, true, block_offsets, n_reads, out_bits, out_scores, out_counts, out_status);
}

}  // extern "C"
