// Device-side data model shared by the trellis kernels.
//
// State layout in HBM ("tile-transposed"): a parity buffer holds, for every trellis row p
// (= nstate_pos index of the reference, VCC:74), C/32 tiles of 32 consecutive conv states.
// Inside a tile the 32 conv states are the FASTEST index:
//
//     word(p, c, s, plane, slot) = row[p].off + (c >> 5) * row[p].tile_words
//                                + ((s * (1 + nw_p) + plane) * LP + slot) * 32 + (c & 31)
//
// plane 0 = path metric (fp32 bits), planes 1..nw_p = message history words (the message bits
// that have already left the conv register, LSW first; the register itself is the state index).
// A tile is one contiguous chunk, so a CTA stages its 32 predecessor conv states with ONE bulk
// copy, shared-memory bank == conv state (conflict-free for any slot / plane / crf pattern), and
// a warp's 4-byte stores land as complete 32-byte sectors.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "code_tables.hpp"

namespace lva {

constexpr int kCodeSlots = 6;
constexpr int kTileGroups = 32;  // conv states per HBM tile

struct RowDev {
  unsigned long long off;  // word offset of the row inside a parity buffer
  uint32_t vmask, vval;    // conv state reachable iff (c & vmask) == vval   (VCC:944-978)
  uint32_t tile_words;     // 8 * (1 + idwords) * LP * 32
  uint8_t block, shift, idwords, dropmask;
  // The base a conv transition emits is XOR-linear in the predecessor's low `shift` bits xy:
  // base(c', xy) = base(c', 0) ^ lin[xy].  inv_lin packs, for t = 0..3, the xy with lin[xy] == t
  // (2 bits each; one-bit steps: only t in {0, lin[1]} exist) and lin1 = lin[1] for one-bit steps.
  uint8_t inv_lin, lin1, pad0, pad1;
  uint32_t pad2;
};

struct CodeDev {
  uint32_t mem, nconv, G0, G1, init, fin, rc, nrows, L, LP;
  uint32_t msg_len, pad;
  RowDev row[kMaxRows];
};

struct ReadDev {
  unsigned long long post_block;  // first posterior block of the read
  uint32_t T;                     // number of blocks
  uint32_t band_off;              // offset of its band-start table
};

constexpr int kMaxSlots = 512;    // reads in flight per handle (3 KB of kernel parameters)

// Everything a step launch needs travels in the kernel parameters (constant bank): a CTA can
// issue its tile copy without first chasing per-read control data through global memory.
struct StepArgs {
  uint32_t *state;                // [slot][parity][slot_words]
  unsigned long long slot_words;
  const float *post;
  unsigned long long *counters;   // [0] exact fallbacks
  uint32_t t, dev2, code_slot;
  uint32_t row_lo, row_hi;        // rows this launch is responsible for
  // survivor mode (list size 1): 4-bit decisions, 8 per conv state, [slot][t][row - a_t][c] words
  uint32_t *dec;                  // nullptr unless the handle runs in survivor mode
  unsigned long long dec_slot_words;
  uint32_t dec_rows;              // band rows per step in the decision array
  uint32_t own_tma;               // 1: the own lists of two-bit rows come through the tensor map passed beside this struct
  uint32_t post_blk[kMaxSlots];   // per active slot: index of the posterior block of this step
  uint16_t band_a[kMaxSlots];     // per active slot: first in-band row of this step (VCC:677-678)
};

struct FinalArgs {
  const uint32_t *state;
  unsigned long long slot_words;
  const ReadDev *reads;
  uint32_t *out;                  // [slot][8][LP][1 + nw_last + 1]: metric, history words, CRC word
  uint32_t code_slot, n_slots;
  int32_t crc_pad;                // -1: no CRC check; 0/1: helper.decode_list_CRC_index with pad = 0/1
};

struct TraceArgs {
  const uint32_t *state;          // parity buffers (metrics only in survivor mode)
  unsigned long long slot_words;
  const uint32_t *dec;
  unsigned long long dec_slot_words;
  uint32_t dec_rows;
  const ReadDev *reads;
  const uint16_t *band;           // band-start tables (ReadDev.band_off)
  uint32_t *out;                  // [slot][12]: metric, found flag, CRC word, pad, 8 message words
  uint32_t code_slot, n_slots;
  int32_t crc_pad;
};

// CRC word of a candidate: bit 31 = CRC-8 matches, bits 0..11 = oligo index (PRP inverted)
constexpr uint32_t kCrcOk = 0x80000000u;

__device__ __forceinline__ float neg_inf() { return __int_as_float(0xff800000); }

}  // namespace lva
