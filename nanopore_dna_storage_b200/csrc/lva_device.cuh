// Device-side data model shared by the trellis kernels.
//
// List entry = 3 words: path metric (fp32), chain id, partial chunk.
//   The reference carries a 256-bit message per entry and de-duplicates on it (VCC:603-615,
//   778-779).  Two entries that meet at one state share their last m bits (the conv register =
//   the state index), so only the message HISTORY -- the bits that already left the register --
//   tells them apart.  The history is kept in absolute coordinates (message bit b at string
//   position b) as completed chunks of <= 32 bits plus the partial chunk being filled.  The
//   completed chunks are named by ONE canonical 32-bit id: the slot of the key
//   (parent id, chunk bits) in a per-read open-addressing table (insert-or-get, slots never
//   change once set).  Canonical => (id, partial) equality <=> history equality, so the
//   de-duplication stays exact; the final message is read back by walking the chain.
//
// State layout in HBM ("octet pieces"): per read slot and parity buffer a circular window of W
// trellis rows (row p lives in window slot p % W; W = 2*max_deviation+1 or all rows); a row is
// nconv/8 pieces of 8 consecutive conv states:
//
//     word(p, c, plane, crf, slot) = rowbase + (c >> 3) * piece_words
//                                  + ((plane * 8 + crf) * LP + slot) * 8 + (c & 7)
//
// plane 0 = metric, 1 = chain id, 2 = partial chunk; piece_words = 3 * 8 * LP * 8.  The CTA that
// owns 32 predecessor conv states produces, per register run, 8 consecutive conv states = whole
// pieces, so every piece is written by one CTA and read as complete contiguous blocks (TMA
// tensor copies re-shape them into bank = conv-state rows of 32 words in shared memory).
// Boundary rows (code_tables.hpp) additionally publish the interned id of (id, partial) of every
// entry in a side ring ("nid", one plane per piece): the row above reads it in place of the id
// plane.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "code_tables.hpp"

namespace lva {

constexpr int kCodeSlots = 6;
constexpr int kTileGroups = 32;          // predecessor conv states per CTA
constexpr uint32_t kRootId = 0xFFFFFFFEu;  // chain id of "no completed chunk"
constexpr unsigned long long kEmptyKey = ~0ull;

struct RowDev {
  uint32_t vmask, vval;    // conv state reachable iff (c & vmask) == vval   (VCC:944-978)
  uint8_t block, shift;    // puncturing block entering the row, message bits it consumes
  uint8_t app;             // history bits a move INTO this row appends (0..2, the newest of the dropped bits)
  uint8_t app_pos;         // bit position inside the partial chunk where they go (0 when the source row is a boundary row)
  uint8_t post_boundary;   // 1: source row p-1 is a boundary row -> candidate id = its nid plane, partial = appended bits
  uint8_t boundary;        // 1: entries written to this row also publish their interned id
  uint8_t nid_slot;        // ring slot of this row's nid plane (boundary rows) ...
  uint8_t src_nid_slot;    // ... and of the source row's (post-boundary rows)
  uint8_t nohist;          // 1: no history bit yet: all candidates of a state carry the same (empty) message
  // The base a conv transition emits is XOR-linear in the predecessor's low `shift` bits xy:
  // base(c', xy) = base(c', 0) ^ lin[xy].  inv_lin packs, for t = 0..3, the xy with lin[xy] == t
  // (2 bits each; one-bit steps: only t in {0, lin[1]} exist) and lin1 = lin[1] for one-bit steps.
  uint8_t inv_lin, lin1;
  uint8_t xdup;            // 1: the source row p-1 may hold one message in two crf lists (first_run_rows): pop-by-pop merge
};

struct CodeDev {
  uint32_t mem, nconv, G0, G1, init, fin, rc, nrows, L, LP;
  uint32_t msg_len, window;        // window: rows per parity buffer (W)
  uint32_t nchunks, last_plen;     // completed chunks / partial bits of an entry of the last row
  uint32_t nid_ring, pad;          // nid planes per parity buffer
  uint8_t chunk_len[16];
  RowDev row[kMaxRows];
};

struct ReadDev {
  unsigned long long post_block;  // first posterior block of the read
  uint32_t T;                     // number of blocks
  uint32_t band_off;              // offset of its band-start table
};

constexpr int kMaxSlots = 512;    // reads in flight per handle

// Everything a step launch needs travels in the kernel parameters (constant bank): a CTA can
// issue its copies without first chasing per-read control data through global memory.
struct StepArgs {
  uint32_t *state;                // [slot][parity][W rows][pieces][piece_words]
  uint32_t *nid;                  // [slot][parity][nid_ring][pieces][8*LP*8]
  unsigned long long *tab;        // [slot][tab_slots] interning tables
  unsigned long long parity_words;  // words of one parity buffer (W * row_words)
  unsigned long long nid_parity_words;
  const float *post;
  unsigned long long *counters;   // [0] exact fallbacks, [1] chain table overflow flag, [2] structure check failures (LVA_DBG bit 3), [8 + slot] keys of that read's chain table
  uint32_t row_words, piece_words;
  uint32_t tab_mask;              // tab_slots - 1
  uint32_t t, dev2, code_slot;
  uint32_t one;                   // = 1, a value the compiler cannot see: cond_move() multiplies by it to turn a select into a predicated integer multiply-add (FMA pipe)
  uint32_t dbg;                   // developer switches (LVA_DBG), 0 in production: bit 1 data movement only, bit 2 pop-by-pop merge on every row, bit 3 check the two-stage merge's structural claim, bit 4 no stores, bit 5 wait for every copy before any work
  uint32_t post_blk[kMaxSlots];   // per active slot: index of the posterior block of this step
  uint16_t band_a[kMaxSlots];     // per active slot: first in-band row of this step (VCC:677-678)
  uint16_t band_slot[kMaxSlots];  // per active slot: band_a % window (the window slot of the first in-band row)
};

struct FinalArgs {
  const uint32_t *state;
  const unsigned long long *tab;
  unsigned long long parity_words;
  const ReadDev *reads;
  uint32_t *out;                  // [slot][8][LP][1 + 8 + 1]: metric, history words (bit k = message bit msg_len-1-k), CRC word
  uint32_t row_words, piece_words, tab_slots;
  uint32_t code_slot, n_slots;
  int32_t crc_pad;                // -1: no CRC check; 0/1: helper.decode_list_CRC_index with pad = 0/1
};

constexpr int kFinalWords = 10;   // words per collected candidate

// ---- survivor mode (list size 1): metrics only, 4-bit decisions, traceback ----
// layout: word(p, c, s) = p * nconv * 8 + (c >> 5) * 256 + s * 32 + (c & 31), all rows resident
struct TraceArgs {
  const uint32_t *state;          // parity buffers (metrics only)
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
