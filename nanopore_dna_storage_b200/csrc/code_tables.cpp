// See code_tables.hpp.  Reference file cited as VCC =
// /root/reference/viterbi/viterbi_convolutional_code.cpp.
#include "code_tables.hpp"

#include <algorithm>
#include <cmath>

namespace lva {

const char *code_status_text(int st) {
  switch (st) {
    case kCodeOk: return "";
    case kBadMem: return "Invalid mem_conv (allowed: 6, 8, 11, 14)\n";
    case kBadRate: return "Invalid rate parameter (allowed: 1, 3, 5, 7)\n";
    case kOddLength: return "Output length not even. Try padding with a single 0 at end.\n";
    case kMarkerTooLong: return "Too large sync marker\n";
    case kPeriodTooShort: return "Sync period shorter than sync marker length\n";
    case kMarkerBadChar: return "Invalid character in sync marker (only 0/1 allowed)\n";
    default: return "unknown code parameter error\n";
  }
}

static uint32_t mirror_bits(uint32_t v, uint32_t n) {
  uint32_t r = 0;
  for (uint32_t i = 0; i < n; i++) r |= ((v >> i) & 1u) << (n - 1 - i);
  return r;
}

int build_code_spec(CodeSpec &k, uint32_t mem, uint32_t rate, uint32_t msg_len, bool rc,
                    const std::string &marker, uint32_t period) {
  k = CodeSpec();
  k.mem = mem;
  k.rate = rate;
  k.msg_len = msg_len;
  k.rc = rc;
  struct Poly { uint32_t mem, g0, g1, start; };
  static const Poly polys[] = {{6, 0171, 0133, 0x25}, {8, 0515, 0677, 0x96},
                               {11, 05537, 06131, 0x4B1}, {14, 075063, 056711, 0x258D}};
  const Poly *poly = nullptr;
  for (const Poly &q : polys) if (q.mem == mem) poly = &q;
  if (!poly) return kBadMem;
  k.nconv = 1u << mem;
  k.G[0] = poly->g0;
  k.G[1] = poly->g1;
  k.init = poly->start;
  k.fin = mirror_bits(k.init, mem);

  // puncturing as a periodic sequence of per-base building blocks (VCC:296-339)
  static const struct { uint32_t rate, len; uint8_t pat[5]; } pats[] = {
      {1, 1, {0}}, {2, 3, {0, 2, 0}}, {3, 2, {0, 1}},
      {4, 5, {0, 3, 0, 2, 1}}, {5, 3, {0, 1, 2}}, {7, 4, {0, 3, 1, 1}}};
  bool found = false;
  for (const auto &q : pats)
    if (q.rate == rate) {
      k.punc_len = q.len;
      for (uint32_t i = 0; i < q.len; i++) k.punc[i] = q.pat[i];
      found = true;
    }
  if (!found) return kBadRate;

  // rows of the position trellis and the number of message bits behind each (VCC:344-357)
  const uint32_t total = msg_len + mem;
  uint32_t used = 0;
  k.nrows = 1;
  k.bits_at[0] = 0;
  while (used < total) {
    used += k.punc[(k.nrows - 1) % k.punc_len] == 0 ? 1 : 2;
    if (k.nrows >= (uint32_t)kMaxRows - 1) return kOddLength;
    k.bits_at[k.nrows++] = used;
  }
  if (used != total) return kOddLength;

  if (rc) {  // decode the reverse-complement strand: mirror everything (VCC:359-386)
    k.G[0] = mirror_bits(k.G[0], mem + 1);
    k.G[1] = mirror_bits(k.G[1], mem + 1);
    const uint32_t old_init = k.init;
    k.init = mirror_bits(k.fin, mem);
    k.fin = mirror_bits(old_init, mem);
    uint8_t fwd[16];
    std::copy(k.punc, k.punc + k.punc_len, fwd);
    const uint32_t last = (k.nrows - 2) % k.punc_len;  // block of the last forward base
    static const uint8_t swapped[4] = {0, 2, 1, 3};    // blocks 1 and 2 trade places when read backwards
    for (uint32_t i = 0; i < k.punc_len; i++)
      k.punc[i] = swapped[fwd[(k.punc_len - i + last) % k.punc_len]];
    std::reverse(k.bits_at, k.bits_at + k.nrows);
    for (uint32_t i = 0; i < k.nrows; i++) k.bits_at[i] = total - k.bits_at[i];
  }

  if (!marker.empty()) {  // VCC:388-414
    k.marker_len = (uint32_t)marker.size();
    if (k.marker_len >= 256) return kMarkerTooLong;
    k.marker_period = period;
    if (period < k.marker_len) return kPeriodTooShort;
    for (uint32_t i = 0; i < k.marker_len; i++) {
      if (marker[i] != '0' && marker[i] != '1') return kMarkerBadChar;
      k.marker[i] = marker[i] == '1';
    }
  }

  // ---- derived rows ----
  for (uint32_t p = 0; p < k.nrows; p++) {
    if (p > 0) {
      k.block[p] = k.punc[(p - 1) % k.punc_len];  // VCC:693-696
      k.shift[p] = k.block[p] == 0 ? 1 : 2;
    }
    // reachability of a register value at row p (VCC:944-978): every register bit that lies
    // before the message, after it, or inside a sync marker is pinned
    uint32_t mask = 0, val = 0;
    for (uint32_t sh = 0; sh < mem; sh++) {
      const int64_t q = (int64_t)k.bits_at[p] - 1 - (int64_t)sh;
      const int64_t qf = rc ? (int64_t)msg_len - 1 - q : q;
      const uint32_t bitpos = mem - 1 - sh;
      int pin = -1;
      if (q < 0) pin = (k.init >> (mem + q)) & 1u;
      else if (q >= (int64_t)msg_len) pin = (k.fin >> (q - msg_len)) & 1u;
      else if (k.marker_len > 0 && (uint64_t)qf % k.marker_period < k.marker_len)
        pin = k.marker[(uint64_t)qf % k.marker_period];
      if (pin >= 0) {
        mask |= 1u << bitpos;
        val |= (uint32_t)pin << bitpos;
      }
    }
    k.vmask[p] = mask;
    k.vval[p] = val;
    // message history kept per list entry: the bits that have already left the register
    const uint32_t hist = k.bits_at[p] > mem ? k.bits_at[p] - mem : 0;
    k.idwords[p] = (uint8_t)((hist + 31) / 32);
    const uint32_t real = std::min<uint32_t>(k.shift[p], hist);
    k.dropmask[p] = (uint8_t)((1u << real) - 1u);
    k.app[p] = (uint8_t)real;
  }
  // chunking of the message history: a move into row p appends app[p] bits to the partial chunk
  // of row p-1; when they no longer fit 32 bits, row p-1 is a boundary row (its partial chunk is
  // complete and gets interned) and row p starts a new chunk with the appended bits only
  k.nchunks = 0;
  for (uint32_t p = 1; p < k.nrows; p++) {
    if ((uint32_t)k.plen[p - 1] + k.app[p] > 32u) {
      k.boundary[p - 1] = 1;
      k.chunk_len[k.nchunks++] = k.plen[p - 1];
      k.plen[p] = k.app[p];
    } else {
      k.plen[p] = (uint8_t)(k.plen[p - 1] + k.app[p]);
    }
    k.chunk_of[p] = (uint8_t)k.nchunks;
  }
  first_run_rows(k);
  return kCodeOk;
}

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

// Rows a message prefix can reach with all its bases equal (CodeSpec::first_run): per base b, the set of
// conv states of row q reached from the start state through valid states by transitions that all emit b.
void first_run_rows(CodeSpec &k) {
  std::vector<uint8_t> cur[4], nxt[4];
  for (int b = 0; b < 4; b++) {
    cur[b].assign(k.nconv, 0);
    cur[b][k.init] = 1;
  }
  for (uint32_t p = 0; p < k.nrows; p++) k.first_run[p] = 0;
  k.first_run[0] = 1;
  for (uint32_t q = 1; q < k.nrows; q++) {
    const uint32_t sh = k.shift[q];
    bool any = false;
    for (uint32_t b = 0; b < 4; b++) {
      nxt[b].assign(k.nconv, 0);
      for (uint32_t c = 0; c < k.nconv; c++) {
        if (!cur[b][c]) continue;
        for (uint32_t in = 0; in < (1u << sh); in++) {
          const uint32_t in1 = in & 1u, in2 = in >> 1;
          uint32_t c2 = (c | (in1 << k.mem)) >> 1;
          if (sh == 2) c2 = (c2 | (in2 << k.mem)) >> 1;
          if ((c2 & k.vmask[q]) != k.vval[q]) continue;
          if (host_emitted_base(k, k.block[q], c, in1, in2) != b) continue;
          nxt[b][c2] = 1;
          any = true;
        }
      }
      cur[b].swap(nxt[b]);
    }
    if (!any) break;
    k.first_run[q] = 1;
  }
}

uint32_t band_start(uint32_t t, uint32_t nblk, uint32_t nrows, uint32_t max_deviation, bool fused) {
  const double frac = (double)t / (double)nblk;
  double v;
  if (fused) {
    v = std::fma(frac, (double)nrows, -(double)max_deviation);
  } else {
    volatile double prod = frac * (double)nrows;  // rounded product (volatile: no contraction by this build either)
    v = prod - (double)max_deviation;
  }
  const int64_t a = (int64_t)v;
  return a < 0 ? 0u : (uint32_t)a;
}

bool conv_encode(const CodeSpec &k, const std::vector<uint8_t> &msg, std::vector<uint8_t> &out) {
  const uint32_t total = (uint32_t)msg.size() + k.mem;
  std::vector<uint8_t> full;
  full.reserve(2 * total);
  uint32_t reg = k.init;
  for (uint32_t i = 0; i < total; i++) {
    const uint32_t bit = i < msg.size() ? (msg[i] ? 1u : 0u) : ((k.fin >> (i - msg.size())) & 1u);
    const uint32_t wide = reg | (bit << k.mem);
    full.push_back((uint8_t)(__builtin_parity(wide & k.G[0]) ^ (k.rc ? 1 : 0)));
    full.push_back((uint8_t)(__builtin_parity(wide & k.G[1]) ^ (k.rc ? 1 : 0)));
    reg = wide >> 1;
  }
  if (reg != k.fin) return false;
  out.clear();
  size_t i = 0;
  for (uint32_t p = 0; p + 1 < k.nrows; p++) {
    switch (k.punc[p % k.punc_len]) {  // which two of the (up to four) coded bits survive (VCC:473-494)
      case 0: out.push_back(full[i]);     out.push_back(full[i + 1]); i += 2; break;
      case 1: out.push_back(full[i + 1]); out.push_back(full[i + 2]); i += 4; break;
      case 2: out.push_back(full[i]);     out.push_back(full[i + 3]); i += 4; break;
      default: out.push_back(full[i + 1]); out.push_back(full[i + 3]); i += 4; break;
    }
  }
  return i == full.size();
}

WorkCount count_work(const CodeSpec &k, uint32_t nblk, uint32_t max_deviation) {
  WorkCount w;
  std::vector<double> per_row(k.nrows), edges(k.nrows);
  for (uint32_t p = 0; p < k.nrows; p++) {
    const int free_bits = (int)k.mem - __builtin_popcount(k.vmask[p]);
    per_row[p] = std::ldexp(8.0, free_bits);
    edges[p] = p == 0 ? 1.0 : (k.block[p] == 0 ? 3.0 : 5.0);
  }
  for (uint32_t t = 0; t < nblk; t++) {
    const uint32_t a = band_start(t, nblk, k.nrows, max_deviation);
    const uint32_t b = band_end(a, k.nrows, max_deviation);
    for (uint32_t p = a; p < b; p++) {
      w.state_updates += per_row[p];
      w.edge_acs += per_row[p] * edges[p];
    }
  }
  return w;
}

}  // namespace lva
