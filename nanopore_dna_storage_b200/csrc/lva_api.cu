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
#include "lva_wide.cuh"

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
  // like ensure(), but the old allocation survives a failed attempt
  bool try_grow(size_t n) {
    if (n <= cap) return true;
    void *q = nullptr;
    if (cudaMalloc(&q, n) != cudaSuccess) { cudaGetLastError(); return false; }
    if (p) cudaFree(p);
    p = q;
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
  uint32_t L = 1, LP = 1, max_dev = 0;
  uint32_t n_slots = 0;
  // state geometry (lva_device.cuh): window rows, words per piece / row / parity buffer
  uint32_t window = 0, piece_words = 0, row_words = 0, nid_ring = 1;
  unsigned long long parity_words = 0, nid_parity_words = 0;
  uint32_t tab_slots = 0;      // interning table slots per read in flight (power of two)
  unsigned long long slot_words = 0;  // survivor mode: words of one parity buffer
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[8] = {};
  DevBuf d_state, d_nid, d_tab, d_reads, d_band, d_post, d_final, d_counters;
  FastMaps maps{};
  PinnedBuf h_stage, h_final, h_post;
  std::string err;
  lva_timing timing{};
  bool survivor = false;     // list size 1: metrics + 4-bit decisions + traceback instead of list entries
  DevBuf d_dec, d_trace;
  PinnedBuf h_trace;
  int gen_tg = 32;           // tile width of the generic kernel
  size_t gen_smem = 0;
  bool use_fast = false;
  bool use_wide = false;     // list sizes 17..64: lva_step_wide (lva_wide.cuh)
  bool all_one_bit = false;
  bool fused_band = true;    // VCC:677-679 as a fused multiply-subtract (LVA_FLAG_UNFUSED_BAND clears it)
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
  X(1, 32) X(2, 32) X(4, 32) X(8, 32) X(16, 32) X(16, 16) X(32, 16) X(32, 8) X(64, 8)

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

size_t generic_smem_bytes(uint32_t LP, uint32_t tg) {
  return 256 + ((size_t)24 * LP * tg + (size_t)3 * LP * tg * 8) * 4;
}

constexpr size_t kCounterBytes = (8 + kMaxSlots) * sizeof(unsigned long long);  // [0] fallbacks [1] overflow [2] structure check failures [8+z] keys of slot z

struct FinalEntry {
  float score;
  uint32_t id[kMaxIdWords];
  uint32_t crcw;
};

// CRC-8 / index request of one call (lva_decode_batch_crc)
struct CrcReq {
  int pad = -1;  // -1: none
  uint32_t num_oligos = 0;
  uint8_t *ok_out = nullptr;
  uint16_t *index_out = nullptr;
  int32_t *choice_out = nullptr;
};

__global__ void lva_seed_start_l1(uint32_t *state, unsigned long long slot_words, uint32_t code_slot, uint32_t n_slots) {
  const CodeDev &K = c_code[code_slot];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_slots * 8) return;
  state[(unsigned long long)(i >> 3) * 2 * slot_words + l1_word(K, 0, K.init, i & 7)] = 0u;  // 0.0f
}

typedef CUresult (*tmap_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Tensor map of lva_step_fast (lva_fast.cuh) for the own (stay) lists: the state buffer seen as
//   own: (8 columns | 8 crf lists | LP slots | 3 planes | pieces), box {8, 4 crf, LP, 3, 1 piece}
// (the predecessor tile and the nid planes are whole pieces: plain bulk copies, no map).
// cuTensorMapEncodeTiled comes from the driver through the runtime's entry-point query (no link against libcuda).
std::string make_maps(lva_handle *h) {
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
      qres != cudaDriverEntryPointSuccess)
    return "cuTensorMapEncodeTiled entry point not available";
  const cuuint64_t ppr = h->code.nconv / 8;
  const cuuint64_t pieces = (cuuint64_t)h->n_slots * 2ull * h->window * ppr;
  const cuuint64_t piece_bytes = (cuuint64_t)h->piece_words * 4ull, LP = h->LP;
  if (pieces > 0x7fffffffull) return "state buffer too large for 32-bit tensor coordinates";
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  auto enc = [&](CUtensorMap *m, cuuint32_t rank, void *base, const cuuint64_t *dims, const cuuint64_t *strides,
                 const cuuint32_t *box) {
    return ((tmap_encode_fn)fn)(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, rank, base, dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  {
    const cuuint64_t dims[5] = {8, 8, LP, 3, pieces}, strides[4] = {LP * 32, 32, 8 * LP * 32, piece_bytes};
    const cuuint32_t box[5] = {8, 4, (cuuint32_t)LP, 3, 1};
    CUresult r = enc(&h->maps.own, 5, h->d_state.p, dims, strides, box);
    if (r != CUDA_SUCCESS) return "cuTensorMapEncodeTiled(own) failed: " + std::to_string((int)r);
  }
  return "";
}

}  // namespace

extern "C" {

const char *lva_version(void) { return "lva_b200 0.2 (sm_100a)"; }

int lva_device_count(void) {
  int n = 0;
  return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
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
  // List size 1 needs no message identity at all (VCC:715-742 has no de-duplication): keep metrics
  // only, record one 4-bit decision per state and step, rebuild the message by traceback.
  h->survivor = h->L == 1 && !(params->flags & LVA_FLAG_EXACT_ONLY);
  h->fused_band = !(params->flags & LVA_FLAG_UNFUSED_BAND);
  h->all_one_bit = true;
  for (uint32_t r = 1; r < k.nrows; r++) h->all_one_bit = h->all_one_bit && k.shift[r] == 1;

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

  // ---- state geometry ----
  h->window = (uint32_t)std::min<uint64_t>(k.nrows, 2ull * h->max_dev + 1ull);
  h->piece_words = 192u * h->LP;
  h->row_words = (k.nconv / 8u) * h->piece_words;
  h->parity_words = (unsigned long long)h->window * h->row_words;
  h->slot_words = (unsigned long long)k.nrows * k.nconv * 8ull;  // survivor layout: every row, one word per state
  // nid ring: as many planes as boundary rows can be alive at once (rows [a-1, b) of a step), plus one
  h->nid_ring = 1;
  for (uint32_t p = 0; p < k.nrows; p++) {
    uint32_t n = 0;
    for (uint32_t q = p; q < k.nrows && q < p + h->window; q++) n += k.boundary[q];
    h->nid_ring = std::max(h->nid_ring, n + 1);
  }
  h->nid_parity_words = (unsigned long long)h->nid_ring * k.nconv * 8ull * h->LP;
  {
    // interning table: distinct (chain id, chunk) pairs per read.  Measured: up to 4.2 M keys per
    // C2 read (m=11, L=8, six chunks; hard reads need the most), 7.6 M at L=16, 1.2 M at m=14 with
    // three chunks -- the count follows the list size and the noise of the read, not the state count.
    // Default 8 M slots per 8 list slots; a call that fills a table beyond ~45 % grows them for the
    // next call, a probe chain that runs away raises a flag and the call is repeated with larger tables.
    uint64_t want = (1ull << 23) * std::max<uint64_t>(1, (h->LP + 7) / 8);
    if (const char *e2 = getenv("LVA_TAB_LOG2")) want = 1ull << std::min(28, std::max(10, atoi(e2)));
    want = std::min<uint64_t>(std::max<uint64_t>(want, 1ull << 16), 1ull << 27);
    h->tab_slots = next_pow2((uint32_t)want);
    if (k.nchunks == 0) h->tab_slots = 1024;  // short messages: nothing is ever interned
  }

  // ---- device code tables (constant memory slot) ----
  {
    std::lock_guard<std::mutex> g(g_slot_mutex);
    for (int i = 0; i < kCodeSlots; i++)
      if (!g_slot_used[h->device][i]) { g_slot_used[h->device][i] = true; h->code_slot = i; break; }
  }
  if (h->code_slot < 0) { lva_destroy(h); return fail(LVA_ERR_BAD_ARG, "too many live handles on this device"); }
  CodeDev &d = h->dev_code;
  d.mem = k.mem; d.nconv = k.nconv; d.G0 = k.G[0]; d.G1 = k.G[1]; d.init = k.init; d.fin = k.fin;
  d.rc = k.rc ? 1 : 0; d.nrows = k.nrows; d.L = h->L; d.LP = h->LP; d.msg_len = k.msg_len;
  d.window = h->window; d.nchunks = k.nchunks; d.last_plen = k.plen[k.nrows - 1]; d.nid_ring = h->nid_ring; d.pad = 0;
  for (int i = 0; i < 16; i++) d.chunk_len[i] = k.chunk_len[i];
  {
    uint32_t nb = 0;  // boundary rows seen so far
    uint8_t slot_of[kMaxRows] = {0};
    for (uint32_t p = 0; p < k.nrows; p++) {
      RowDev &r = d.row[p];
      memset(&r, 0, sizeof(r));
      r.vmask = k.vmask[p]; r.vval = k.vval[p];
      r.block = k.block[p]; r.shift = k.shift[p];
      r.app = k.app[p];
      r.boundary = k.boundary[p];
      if (r.boundary) { slot_of[p] = (uint8_t)(nb % h->nid_ring); nb++; }
      r.nid_slot = slot_of[p];
      r.post_boundary = p > 0 ? k.boundary[p - 1] : 0;
      r.src_nid_slot = p > 0 ? slot_of[p - 1] : 0;
      r.app_pos = (p > 0 && !r.post_boundary) ? k.plen[p - 1] : 0;
      r.nohist = (k.plen[p] == 0 && k.chunk_of[p] == 0) ? 1 : 0;
      r.xdup = (p == 0 || k.first_run[p - 1]) ? 1 : 0;  // the source row may hold one message in two crf lists
      if (p > 0) {  // linear map low predecessor bits -> emitted base (verified bijective by butterfly_structure_ok)
        const uint32_t b0 = host_emitted_base(k, k.block[p], 0, 0, 0);
        for (uint32_t xy = 0; xy < (1u << k.shift[p]); xy++) {
          const uint32_t t = host_emitted_base(k, k.block[p], xy, 0, 0) ^ b0;
          r.inv_lin |= (uint8_t)(xy << (2 * t));
          if (xy == 1) r.lin1 = (uint8_t)t;
        }
      }
    }
  }
  if ((e = cudaMemcpyToSymbol(c_code, &d, sizeof(CodeDev), sizeof(CodeDev) * h->code_slot)) != cudaSuccess)
    return cuda_fail("cudaMemcpyToSymbol", e);

  // ---- kernel selection ----
  h->gen_tg = 0;
  for (int tg : {32, 16, 8}) {
    bool have = false;
#define X(lp, tgc) if (h->LP == lp && tg == tgc) have = true;
    LVA_GENERIC_CASES(X)
#undef X
    if (!have) continue;
    size_t sm = generic_smem_bytes(h->LP, tg);
    if (sm <= (size_t)prop.sharedMemPerBlockOptin) { h->gen_tg = tg; h->gen_smem = sm; break; }
  }
  if (!h->gen_tg) { lva_destroy(h); return fail(LVA_ERR_LIST_SIZE, "list_size too large for shared memory"); }
  if ((e = generic_prepare(h)) != cudaSuccess) return cuda_fail("cudaFuncSetAttribute(generic)", e);
  h->use_fast = !(params->flags & LVA_FLAG_EXACT_ONLY) && (h->survivor || fast_supported(h->LP));
  h->use_wide = !(params->flags & LVA_FLAG_EXACT_ONLY) && !h->survivor && wide_supported(h->LP);
  if (const char *wv = getenv("LVA_WIDE"))  // developer A/B: list sizes 5..16 through lva_step_wide<8|16>
    if (atoi(wv) && !(params->flags & LVA_FLAG_EXACT_ONLY) && !h->survivor && (h->LP == 8 || h->LP == 16)) { h->use_wide = true; h->use_fast = false; }
  if (h->use_wide && (e = wide_prepare(h->LP)) != cudaSuccess) return cuda_fail("cudaFuncSetAttribute(wide)", e);
  if (h->use_fast && !h->survivor && (e = fast_prepare(h->LP)) != cudaSuccess)
    return cuda_fail("cudaFuncSetAttribute(fast)", e);

  // ---- state memory: n_slots reads in flight, two parity buffers each ----
  size_t free_b = 0, total_b = 0;
  if ((e = cudaMemGetInfo(&free_b, &total_b)) != cudaSuccess) return cuda_fail("cudaMemGetInfo", e);
  size_t per_slot = h->survivor ? (size_t)h->slot_words * 2 * 4
                                : (size_t)(h->parity_words + h->nid_parity_words) * 2 * 4 + (size_t)h->tab_slots * 8;
  size_t budget = (size_t)(free_b * 0.70);
  // default: enough reads in flight that one step launch is >= ~64 waves of CTAs (small codes are launch-bound otherwise), at least 64
  uint32_t slots = params->max_slots;
  if (!slots) {
    const uint64_t band_rows = std::min<uint64_t>(2ull * h->max_dev, k.nrows);
    const uint64_t ctas_per_read = std::max<uint64_t>(1, band_rows * (k.nconv / 32));
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
  const size_t state_bytes = h->survivor ? (size_t)slots * h->slot_words * 2 * 4 : (size_t)slots * h->parity_words * 2 * 4;
  if (!h->d_state.ensure(state_bytes)) { lva_destroy(h); return fail(LVA_ERR_OOM, "state allocation failed"); }
  if (!h->survivor && (!h->d_nid.ensure((size_t)slots * h->nid_parity_words * 2 * 4) ||
                       !h->d_tab.ensure((size_t)slots * h->tab_slots * 8))) {
    lva_destroy(h);
    return fail(LVA_ERR_OOM, "chain table allocation failed");
  }
  if (!h->d_reads.ensure(sizeof(ReadDev) * slots) || !h->d_counters.ensure(kCounterBytes) ||
      !h->d_final.ensure((size_t)slots * 8 * h->LP * kFinalWords * 4)) {
    lva_destroy(h);
    return fail(LVA_ERR_OOM, "allocation failed");
  }
  cudaMemset(h->d_counters.p, 0, kCounterBytes);
  if (h->use_fast && !h->survivor) {
    const std::string m = make_maps(h);
    if (!m.empty()) { lva_destroy(h); return fail(LVA_ERR_CUDA, m); }
  }
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
    const uint32_t a = band_start(t, T, k.nrows, h->max_dev, h->fused_band), b = band_end(a, k.nrows, h->max_dev);
    for (uint32_t p = a; p < b; p++) {
      const double nst = std::ldexp(8.0, (int)k.mem - __builtin_popcount(k.vmask[p]));
      out->state_updates += nst;
      out->edge_acs += nst * (p == 0 ? 1.0 : (k.block[p] == 0 ? 3.0 : 5.0));
      if (h->survivor) {
        out->state_bytes += nst * 8.5;  // metric read + metric written + 4-bit decision
      } else {
        // states some conv transition reaches (one-bit rows: two of the four bases): one read and one
        // write of the list, LP slots of 3 words, plus the published chain id on boundary rows
        const double live = p == 0 ? 8.0 : (k.shift[p] == 1 ? nst * 0.5 : nst);
        out->state_bytes += live * 4.0 * h->LP * (6.0 + (k.boundary[p] ? 1.0 : 0.0));
      }
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
    out->first_run_row[p] = k.first_run[p];
  }
  return LVA_OK;
}

void lva_band_ex(uint32_t t, uint32_t nblk, uint32_t nstate_pos, uint32_t max_deviation, uint32_t flags, uint32_t *start,
                 uint32_t *end) {
  const uint32_t a = band_start(t, nblk, nstate_pos, max_deviation, !(flags & LVA_FLAG_UNFUSED_BAND));
  if (start) *start = a;
  if (end) *end = band_end(a, nstate_pos, max_deviation);
}
void lva_band(uint32_t t, uint32_t nblk, uint32_t nstate_pos, uint32_t max_deviation, uint32_t *start,
              uint32_t *end) {
  lva_band_ex(t, nblk, nstate_pos, max_deviation, 0, start, end);
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

}  // extern "C"

// read i owns posterior blocks [starts[i], ends[i]) of d_post
static int decode_core(lva_handle *h, const float *d_post, const uint64_t *starts, const uint64_t *ends, uint32_t n_reads,
                       const CrcReq &crc, uint8_t *out_bits, float *out_scores, uint32_t *out_counts,
                       int32_t *out_status) {
  const CodeSpec &k = h->code;
  const uint32_t L = h->L, LP = h->LP, msg_len = k.msg_len;
  const float NEG = -std::numeric_limits<float>::infinity();

  std::vector<uint32_t> order;
  order.reserve(n_reads);
  for (uint32_t i = 0; i < n_reads; i++) {
    const uint64_t T = ends[i] >= starts[i] ? ends[i] - starts[i] : 0;
    out_counts[i] = 0;
    if (T < (uint64_t)k.nrows + 1 || T > 0xFFFFFFFFull) {  // VCC:600-601
      if (out_status) out_status[i] = LVA_ERR_POST_TOO_SMALL;
      continue;
    }
    if (out_status) out_status[i] = LVA_OK;
    order.push_back(i);
  }
  auto len_of = [&](uint32_t i) { return ends[i] - starts[i]; };
  std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return len_of(x) > len_of(y); });

  const uint32_t dev2 = 2u * h->max_dev;
  if (dev2 == 0) return LVA_OK;  // --max-deviation 0: the band [a, a + 0) is empty at every step, no state is ever updated, every list is empty (VCC:677-687)
  float acc_kernel = 0.f, acc_d2h = 0.f;
  std::vector<FinalEntry> fin;
  const uint32_t dec_rows = (uint32_t)std::min<uint64_t>((uint64_t)dev2 > 0 ? dev2 : 1, k.nrows);
  size_t chunk_cap = h->n_slots;
  if (h->survivor && !order.empty()) {  // decisions: T x band rows x conv states words per read in flight
    const uint64_t Tlongest = len_of(order[0]);
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
    for (uint32_t s = 0; s < ns; s++) band_total += len_of(order[c0 + s]);
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
      const uint32_t T = (uint32_t)len_of(i);
      Ts[s] = T;
      rd[s].post_block = starts[i];
      rd[s].T = T;
      rd[s].band_off = (uint32_t)bo;
      for (uint32_t t = 0; t < T; t++) band[bo + t] = (uint16_t)band_start(t, T, k.nrows, h->max_dev, h->fused_band);
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

    // ---- initialise the slots: every metric -inf, the start state seeded, chain tables empty ----
    if (h->survivor) {
      lva_fill_u32x4<<<h->sm_count * 8, 256, 0, h->stream>>>((uint32_t *)h->d_state.p, (unsigned long long)ns * 2 * h->slot_words, 0xff800000u);
      lva_seed_start_l1<<<(ns * 8 + 127) / 128, 128, 0, h->stream>>>((uint32_t *)h->d_state.p, h->slot_words, h->code_slot, ns);
    } else {
      lva_fill_u32x4<<<h->sm_count * 8, 256, 0, h->stream>>>((uint32_t *)h->d_state.p, (unsigned long long)ns * 2 * h->parity_words, 0xff800000u);
      lva_seed_start<<<(ns * 8 + 127) / 128, 128, 0, h->stream>>>((uint32_t *)h->d_state.p, h->parity_words, h->piece_words, h->code_slot, ns);
      if (k.nchunks > 0) CK(cudaMemsetAsync(h->d_tab.p, 0xFF, (size_t)ns * h->tab_slots * 8, h->stream));
    }
    h->timing.launches += 2;
    CK(cudaGetLastError());

    // ---- time steps ----
    CK(cudaEventRecord(h->ev[2], h->stream));
    StepArgs a{};
    a.state = (uint32_t *)h->d_state.p;
    a.nid = (uint32_t *)h->d_nid.p;
    a.tab = (unsigned long long *)h->d_tab.p;
    a.parity_words = h->parity_words;
    a.nid_parity_words = h->nid_parity_words;
    a.post = d_post;
    a.counters = (unsigned long long *)h->d_counters.p;
    a.row_words = h->row_words;
    a.piece_words = h->piece_words;
    a.tab_mask = h->tab_slots - 1;
    a.dev2 = dev2;
    a.code_slot = h->code_slot;
    a.one = 1u;
    { static const char *dbg_env = getenv("LVA_DBG"); a.dbg = dbg_env ? (uint32_t)atoi(dbg_env) : 0u; }
    ClearArgs clr{};
    clr.state = a.state;
    clr.parity_words = h->parity_words;
    clr.row_words = h->row_words;
    clr.piece_words = h->piece_words;
    clr.code_slot = h->code_slot;
    const uint32_t Tmax = Ts[0];
    uint32_t active = ns;
    for (uint32_t t = 0; t < Tmax; t++) {
      while (active > 0 && Ts[active - 1] <= t) active--;  // reads are sorted by T, longest first
      a.t = t;
      uint32_t span = 0;
      bool any_clear = false;
      if (h->survivor) {
        span = std::min<uint32_t>(dev2, k.nrows);  // lva_step_l1 reads its per-read control data from the device tables
      } else {
        for (uint32_t s = 0; s < active; s++) {
          const uint32_t st = band[rd[s].band_off + t], en = band_end(st, k.nrows, h->max_dev);
          span = std::max(span, en - st);
          a.band_a[s] = (uint16_t)st;
          a.band_slot[s] = (uint16_t)(st % h->window);
          a.post_blk[s] = (uint32_t)(rd[s].post_block + t);
          clr.row[s] = 0xFFFFu;
          if (t > 0 && en > 0) {
            // a row entering the band re-uses a window slot: wipe it first (lva_clear_rows); the band
            // end grows by at most one row per step (T >= nstate_pos + 1)
            const uint32_t en_prev = band_end(band[rd[s].band_off + t - 1], k.nrows, h->max_dev);
            if (en > en_prev && en - 1 >= h->window) { clr.row[s] = (uint16_t)(en - 1); any_clear = true; }
          }
        }
      }
      if (span == 0) continue;
      if (any_clear) {
        lva_clear_rows<<<dim3(std::max<uint32_t>(1, (k.nconv / 8) * (16 * LP) * 2 / 256), active), 256, 0, h->stream>>>(clr);
        h->timing.launches++;
      }
      if (h->use_fast && h->survivor) {
        L1Args l1{a.state, h->slot_words, a.post, (uint32_t *)h->d_dec.p, dec_slot_words, (const ReadDev *)h->d_reads.p,
                  (const uint16_t *)h->d_band.p, dec_rows, t, a.dev2, a.code_slot};
        cudaError_t e = l1_launch_step(h->stream, l1, k, active);
        if (e != cudaSuccess) { h->err = std::string("L=1 step launch: ") + cudaGetErrorString(e); return LVA_ERR_CUDA; }
      } else if (h->use_fast) {
        cudaError_t e = fast_launch_step(h->LP, h->stream, a, k, span, active, h->maps, h->all_one_bit);
        if (e != cudaSuccess) { h->err = std::string("fast step launch: ") + cudaGetErrorString(e); return LVA_ERR_CUDA; }
      } else if (h->use_wide) {
        cudaError_t e = wide_launch_step(h->LP, h->stream, a, k, span, active, h->all_one_bit);
        if (e != cudaSuccess) { h->err = std::string("wide step launch: ") + cudaGetErrorString(e); return LVA_ERR_CUDA; }
      } else {
        dim3 grid(k.nconv / h->gen_tg, span, active);
        cudaError_t e = generic_launch(h, grid, a);
        if (e != cudaSuccess) { h->err = std::string("generic step launch: ") + cudaGetErrorString(e); return LVA_ERR_CUDA; }
      }
      h->timing.step_launches++;
    }
    CK(cudaEventRecord(h->ev[3], h->stream));

    // ---- collect the candidate lists of the end state ----
    const uint32_t nfin = ns * 8 * LP;
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
      tr.crc_pad = crc.pad;
      lva_traceback<<<(ns + 63) / 64, 64, 0, h->stream>>>(tr);
      h->timing.launches += 1;
      CK(cudaGetLastError());
      CK(cudaEventRecord(h->ev[4], h->stream));
      CK(cudaMemcpyAsync(h->h_trace.p, h->d_trace.p, (size_t)ns * 12 * 4, cudaMemcpyDeviceToHost, h->stream));
      h->timing.d2h_bytes += (size_t)ns * 12 * 4;
    } else {
      FinalArgs f{};
      f.state = (const uint32_t *)h->d_state.p;
      f.tab = (const unsigned long long *)h->d_tab.p;
      f.parity_words = h->parity_words;
      f.reads = (const ReadDev *)h->d_reads.p;
      f.out = (uint32_t *)h->d_final.p;
      f.row_words = h->row_words;
      f.piece_words = h->piece_words;
      f.tab_slots = h->tab_slots;
      f.code_slot = h->code_slot;
      f.n_slots = ns;
      f.crc_pad = crc.pad;
      lva_collect_final<<<(nfin + 127) / 128, 128, 0, h->stream>>>(f);
      h->timing.launches += 1;
      CK(cudaGetLastError());
      const size_t fin_bytes = (size_t)nfin * kFinalWords * 4;
      if (!h->h_final.ensure(fin_bytes)) { h->err = "pinned allocation failed"; return LVA_ERR_OOM; }
      CK(cudaEventRecord(h->ev[4], h->stream));
      CK(cudaMemcpyAsync(h->h_final.p, h->d_final.p, fin_bytes, cudaMemcpyDeviceToHost, h->stream));
      h->timing.d2h_bytes += fin_bytes;
    }
    if (!h->survivor && k.nchunks > 0) {  // keys per chain table of this chunk of reads, then reset for the next
      static_assert(sizeof(unsigned long long) == 8, "");
      if (!h->h_post.ensure(ns * 8)) { h->err = "pinned allocation failed"; return LVA_ERR_OOM; }
      CK(cudaMemcpyAsync(h->h_post.p, (unsigned long long *)h->d_counters.p + 8, ns * 8, cudaMemcpyDeviceToHost, h->stream));
      CK(cudaMemsetAsync((unsigned long long *)h->d_counters.p + 8, 0, ns * 8, h->stream));
    }
    CK(cudaEventRecord(h->ev[5], h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (!h->survivor && k.nchunks > 0)
      for (uint32_t s = 0; s < ns; s++) {
        const unsigned long long n = reinterpret_cast<const unsigned long long *>(h->h_post.p)[s];
        h->timing.chain_inserts += n;
        h->timing.chain_keys_max = std::max<uint64_t>(h->timing.chain_keys_max, n);
      }
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
          const uint32_t *e = fw + ((size_t)(s * 8 + crf) * LP + j) * kFinalWords;
          float sc;
          memcpy(&sc, e, 4);
          if (sc != NEG) {
            FinalEntry fe;
            fe.score = sc;
            for (uint32_t w = 0; w < 8; w++) fe.id[w] = e[1 + w];
            fe.crcw = e[9];
            fin.push_back(fe);
          }
        }
      std::sort(fin.begin(), fin.end(),
                [](const FinalEntry &p1, const FinalEntry &p2) -> bool { return p1.score > p2.score; });
      if (fin.size() > L) fin.resize(L);
      out_counts[i] = (uint32_t)fin.size();
      if (crc.pad >= 0) {  // helper.decode_list_CRC_index: first entry whose CRC matches and index < num_oligos
        int32_t choice = -1;
        for (size_t e = 0; e < fin.size(); e++) {
          const bool ok = (fin[e].crcw & kCrcOk) != 0;
          const uint32_t idx = fin[e].crcw & 0xfffu;
          if (crc.ok_out) crc.ok_out[(size_t)i * L + e] = ok ? 1 : 0;
          if (crc.index_out) crc.index_out[(size_t)i * L + e] = (uint16_t)idx;
          if (choice < 0 && ok && idx < crc.num_oligos) choice = (int32_t)e;
        }
        if (crc.choice_out) crc.choice_out[i] = choice;
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
  h->timing.kernel_ms += acc_kernel;
  h->timing.d2h_ms += acc_d2h;
  return LVA_OK;
}

// `post` host or device memory; read i = blocks [starts[i], ends[i]).
static int decode_entry(lva_handle *h, const float *post, bool post_on_device, const uint64_t *starts,
                        const uint64_t *ends, uint32_t n_reads, const CrcReq &crc, uint8_t *out_bits, float *out_scores,
                        uint32_t *out_counts, int32_t *out_status) {
  if (!h) return LVA_ERR_BAD_ARG;
  if (n_reads == 0) return LVA_OK;
  if (!post || !starts || !ends || !out_bits || !out_counts) { h->err = "null argument"; return LVA_ERR_BAD_ARG; }
  uint64_t lo = ~0ull, hi = 0;
  for (uint32_t i = 0; i < n_reads; i++) {
    if (ends[i] < starts[i]) { h->err = "block range ends before it starts"; return LVA_ERR_BAD_ARG; }
    lo = std::min(lo, starts[i]);
    hi = std::max(hi, ends[i]);
  }
  if (hi > 0xFFFF0000ull) { h->err = "batch too large (block index must fit 32 bits)"; return LVA_ERR_BAD_ARG; }
  CK(cudaSetDevice(h->device));
  h->timing = lva_timing{};
  CK(cudaEventRecord(h->ev[0], h->stream));
  const float *d_post = post;
  if (!post_on_device) {
    const size_t bytes = (size_t)(hi - lo) * 160;
    if (!h->d_post.ensure(bytes + 16)) { h->err = "posterior device allocation failed"; return LVA_ERR_OOM; }
    // Staged so that block b of the caller's array sits at d_post + b * 40, as with device input
    CK(cudaMemcpyAsync(h->d_post.p, post + lo * 40, bytes, cudaMemcpyHostToDevice, h->stream));
    d_post = reinterpret_cast<const float *>(h->d_post.p) - lo * 40;
    h->timing.h2d_bytes += bytes;
  }
  CK(cudaEventRecord(h->ev[1], h->stream));
  const lva_timing before = h->timing;
  for (int attempt = 0;; attempt++) {
    {  // a repeated attempt starts its accounting afresh
      const uint64_t regrows = h->timing.table_regrows;
      h->timing = before;
      h->timing.table_regrows = regrows;
    }
    CK(cudaMemsetAsync(h->d_counters.p, 0, kCounterBytes, h->stream));
    h->timing.chain_keys_max = 0;
    int st = decode_core(h, d_post, starts, ends, n_reads, crc, out_bits, out_scores, out_counts, out_status);
    if (st != LVA_OK) return st;
    unsigned long long cnt[3] = {0, 0, 0};
    CK(cudaMemcpy(cnt, h->d_counters.p, sizeof(cnt), cudaMemcpyDeviceToHost));
    h->timing.exact_fallbacks = cnt[0];
    h->timing.structure_check_failures = cnt[2];
    h->timing.chain_table_slots = h->tab_slots;
    if (cnt[1] == 0) {
      // proactive growth for the next call: linear probing degrades quickly above half full
      if (!h->survivor && h->timing.chain_keys_max * 20ull > (unsigned long long)h->tab_slots * 9ull && h->tab_slots < (1u << 27) &&
          h->d_tab.try_grow((size_t)h->n_slots * h->tab_slots * 2 * 8))
        h->tab_slots *= 2;
      break;
    }
    // a chain table filled up: every result of this call is suspect.  Grow the tables and decode again.
    if (attempt == 3 || h->tab_slots >= (1u << 27)) { h->err = "chain table overflow"; return LVA_ERR_OOM; }
    h->timing.table_regrows++;
    if (!h->d_tab.try_grow((size_t)h->n_slots * h->tab_slots * 4 * 8)) { h->err = "chain table growth failed"; return LVA_ERR_OOM; }
    h->tab_slots *= 4;
  }
  CK(cudaEventRecord(h->ev[6], h->stream));
  CK(cudaEventSynchronize(h->ev[6]));
  cudaEventElapsedTime(&h->timing.total_ms, h->ev[0], h->ev[6]);
  cudaEventElapsedTime(&h->timing.h2d_ms, h->ev[0], h->ev[1]);
  h->timing.launches += h->timing.step_launches;
  return LVA_OK;
}

namespace {
struct OffsetRanges {  // cumulative offsets -> (starts, ends)
  const uint64_t *starts, *ends;
  explicit OffsetRanges(const uint64_t *offs) : starts(offs), ends(offs ? offs + 1 : nullptr) {}
};
}  // namespace

extern "C" {

int lva_decode_batch(lva_handle *h, const float *post, const uint64_t *block_offsets, uint32_t n_reads,
                     uint8_t *out_bits, float *out_scores, uint32_t *out_counts, int32_t *out_status) {
  OffsetRanges r(block_offsets);
  return decode_entry(h, post, false, r.starts, r.ends, n_reads, CrcReq{}, out_bits, out_scores, out_counts, out_status);
}

int lva_decode_batch_crc(lva_handle *h, const float *post, const uint64_t *block_offsets, uint32_t n_reads,
                         int32_t pad, uint32_t num_oligos, uint8_t *out_bits, float *out_scores,
                         uint32_t *out_counts, int32_t *out_status, uint8_t *out_crc_ok, uint16_t *out_index,
                         int32_t *out_choice) {
  CrcReq crc;
  crc.pad = pad ? 1 : 0;
  crc.num_oligos = num_oligos;
  crc.ok_out = out_crc_ok;
  crc.index_out = out_index;
  crc.choice_out = out_choice;
  if (out_choice) for (uint32_t i = 0; i < n_reads; i++) out_choice[i] = -1;
  OffsetRanges r(block_offsets);
  return decode_entry(h, post, false, r.starts, r.ends, n_reads, crc, out_bits, out_scores, out_counts, out_status);
}

int lva_decode_batch_device(lva_handle *h, const float *post_dev, const uint64_t *block_offsets,
                            uint32_t n_reads, uint8_t *out_bits, float *out_scores, uint32_t *out_counts,
                            int32_t *out_status) {
  OffsetRanges r(block_offsets);
  return decode_entry(h, post_dev, true, r.starts, r.ends, n_reads, CrcReq{}, out_bits, out_scores, out_counts, out_status);
}

int lva_decode_ranges(lva_handle *h, const float *post, int32_t post_on_device, const uint64_t *block_starts,
                      const uint64_t *block_ends, uint32_t n_reads, uint8_t *out_bits, float *out_scores,
                      uint32_t *out_counts, int32_t *out_status) {
  return decode_entry(h, post, post_on_device != 0, block_starts, block_ends, n_reads, CrcReq{}, out_bits, out_scores,
                      out_counts, out_status);
}

}  // extern "C"
