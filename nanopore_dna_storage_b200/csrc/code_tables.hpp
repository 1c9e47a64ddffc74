// Host-side code tables for the joint basecaller / convolutional-code trellis.
//
// Mirrors the parameter surface of the reference tool
// (/root/reference/viterbi/viterbi_convolutional_code.cpp, "VCC"):
//   set_conv_params  VCC:264-415   generator polynomials, start/end register, puncturing
//                                  building blocks, nstate_pos, st_pos2msg_pos, rc mirroring,
//                                  sync marker
//   is_valid_state   VCC:944-978   folded here into one (mask, value) pair per trellis row
//   band             VCC:677-679   st_pos_start / st_pos_end per time step
//   conv_encode      VCC:450-499
// plus the row geometry of the B200 state layout.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace lva {

constexpr int kMaxRows = 260;   // msg_len + mem <= 256 -> at most 257 rows
constexpr int kMaxIdWords = 8;  // 256 message bits

enum CodeStatus {
  kCodeOk = 0,
  kBadMem = 1,        // "Invalid mem_conv (allowed: 6, 8, 11, 14)"            VCC:290-292
  kBadRate = 2,       // "Invalid rate parameter (allowed: 1, 3, 5, 7)"         VCC:336-338
  kOddLength = 3,     // "Output length not even. Try padding ..."              VCC:353-357
  kMarkerTooLong = 4, // "Too large sync marker"                                VCC:391-394
  kPeriodTooShort = 5,// "Sync period shorter than sync marker length"          VCC:396-399
  kMarkerBadChar = 6, // "Invalid character in sync marker (only 0/1 allowed)"  VCC:409-411
};
const char *code_status_text(int st);

struct CodeSpec {
  uint32_t mem = 0, nconv = 0, G[2] = {0, 0}, init = 0, fin = 0;
  uint32_t rate = 1, msg_len = 0;
  bool rc = false;
  uint8_t punc[16] = {0};
  uint32_t punc_len = 0;
  uint32_t nrows = 0;               // nstate_pos
  uint32_t bits_at[kMaxRows] = {0}; // st_pos2msg_pos
  uint32_t marker_len = 0, marker_period = 0;
  uint8_t marker[256] = {0};

  // derived per trellis row p
  uint8_t block[kMaxRows] = {0};    // puncturing building block entering row p (p >= 1)
  uint8_t shift[kMaxRows] = {0};    // message bits consumed entering row p (0 for p = 0)
  uint32_t vmask[kMaxRows] = {0};   // conv state c is reachable at row p iff (c & vmask) == vval
  uint32_t vval[kMaxRows] = {0};
  uint8_t idwords[kMaxRows] = {0};  // 32-bit words of message history (what the reference's bitset would hold)
  uint8_t dropmask[kMaxRows] = {0}; // which of the `shift` bits leaving the register are message bits

  // Message identity of a list entry: the message bits that already left the conv
  // register are kept in ABSOLUTE coordinates (message bit b at string position b) as a chain of
  // completed chunks of <= 32 bits, named by one canonical 32-bit id, plus the partial chunk.
  uint8_t app[kMaxRows] = {0};      // message bits appended to the partial chunk by a move INTO row p (0..2)
  uint8_t plen[kMaxRows] = {0};     // bits in the partial chunk at row p (<= 32)
  uint8_t boundary[kMaxRows] = {0}; // 1: the partial chunk of row p is complete; a move out of it starts a new one
  uint8_t chunk_of[kMaxRows] = {0}; // number of completed chunks behind an entry of row p
  // first_run[p] = 1: some message prefix reaches row p with ALL its bases equal (row 0 counts).  All eight
  // crf states start at row 0 (VCC:657-663), so whether such a path sits in the flip or the flop state of
  // its base is open: the same message can be in two crf lists of one conv state.  Everywhere else the crf
  // state follows from the base sequence (flip on a new base, flip/flop alternating inside a run).
  uint8_t first_run[kMaxRows] = {0};
  uint32_t nchunks = 0;             // completed chunks behind an entry of the last row
  uint8_t chunk_len[16] = {0};      // bits of completed chunk k
};

int build_code_spec(CodeSpec &k, uint32_t mem, uint32_t rate, uint32_t msg_len, bool rc,
                    const std::string &marker, uint32_t period);

// base emitted by the conv transition out of c_src under puncturing block blk with input bits in1 (, in2)
// (VCC:440-448 and the block table of VCC:860-942); the kernels' emitted_base() is the same function
uint32_t host_emitted_base(const CodeSpec &k, uint32_t blk, uint32_t c_src, uint32_t in1, uint32_t in2);

void first_run_rows(CodeSpec &k);  // fills CodeSpec::first_run (called by build_code_spec)

// band start of time step t (VCC:677-679).  fused = true: t/nblk*nstate_pos - max_deviation evaluated
// as ONE fused multiply-subtract, which is what g++ -O3 emits for the reference on FMA hardware
// (install.sh:10 builds with -march=native); fused = false: product rounded first (a reference
// built with -ffp-contract=off, by clang, or for a CPU without FMA).  The two differ by one row for
// some (t, nblk), e.g. whenever nblk shares a factor with nstate_pos.
uint32_t band_start(uint32_t t, uint32_t nblk, uint32_t nrows, uint32_t max_deviation, bool fused = true);
inline uint32_t band_end(uint32_t a, uint32_t nrows, uint32_t max_deviation) {
  uint32_t e = a + 2u * max_deviation;  // uint32 arithmetic as in the reference
  return e < nrows ? e : nrows;
}

// conv_encode: returns false if the register does not end in the final state
bool conv_encode(const CodeSpec &k, const std::vector<uint8_t> &msg, std::vector<uint8_t> &out);

// per-read work in the reference's own unit: state updates and edge-ACS
struct WorkCount {
  double state_updates = 0, edge_acs = 0;
};
WorkCount count_work(const CodeSpec &k, uint32_t nblk, uint32_t max_deviation);

}  // namespace lva
