// TEST INFRASTRUCTURE ONLY.  CPU oracle for the list-Viterbi decode path.
//
// A from-the-spec restatement of the reference decoder
//   /root/reference/viterbi/viterbi_convolutional_code.cpp  ("VCC" below),
// written to be called in-process with arrays and to return the path metrics the
// reference computes but never emits.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline leg may load this library; the product (liblva_b200.so)
// never links, imports or calls it.
//
// Pinning: the reference holds no golden vectors for this path, so
// this oracle is pinned against the reference ITSELF, compiled unmodified into
// oracle/_ref/ (see Makefile, ref_shim.cpp): tests/test_oracle_vs_reference.py
// compares decoded lists byte-for-byte on a seeded hazard corpus, and
// tests/golden/*.json holds reference outputs that travel to the GPU box.
//
// Order-defining library routines (std::make_heap / pop_heap / push_heap / sort)
// are the same libstdc++ routines the reference calls (VCC:762-795, 818-821), so
// tie behaviour is identical by construction.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr int kCrf = 8;        // flip-flop states A+ C+ G+ T+ A- C- G- T-   (VCC:19-21)
constexpr int kMsgWords = 4;   // 256-bit message register                   (VCC:28-29)

struct Code {
  uint32_t m = 0, nconv = 0, G[2] = {0, 0}, init = 0, fin = 0;
  bool rc = false;
  uint8_t punc[16] = {0};
  uint32_t punc_len = 0;
  uint32_t P = 0;                // nstate_pos
  uint32_t pos2msg[260] = {0};   // st_pos2msg_pos
  uint32_t marker_len = 0, marker_period = 0;
  bool marker[256] = {false};
  uint32_t msg_len = 0;
};

uint32_t bitrev(uint32_t v, uint32_t n) {  // VCC:417-424
  uint32_t r = 0;
  for (uint32_t i = 0; i < n; i++) r = (r << 1) | ((v >> i) & 1u);
  return r;
}

// VCC:264-415.  Returns 0 ok; 1 bad m; 2 bad rate; 3 odd length; 4 marker too long;
// 5 period < marker; 6 bad marker char.
int build_code(Code &k, uint32_t m, uint32_t rate, uint32_t msg_len, bool rc,
               const std::string &marker, uint32_t period) {
  k = Code();
  k.m = m;
  k.nconv = 1u << m;
  k.rc = rc;
  k.msg_len = msg_len;
  switch (m) {                                           // VCC:269-293
    case 6:  k.G[0] = 0171;   k.G[1] = 0133;   k.init = 0x25;   break;
    case 8:  k.G[0] = 0515;   k.G[1] = 0677;   k.init = 0x96;   break;
    case 11: k.G[0] = 05537;  k.G[1] = 06131;  k.init = 0x4B1;  break;
    case 14: k.G[0] = 075063; k.G[1] = 056711; k.init = 0x258D; break;
    default: return 1;
  }
  k.fin = bitrev(k.init, m);                             // VCC:294
  static const uint8_t pat1[] = {0}, pat2[] = {0, 2, 0}, pat3[] = {0, 1},
                       pat4[] = {0, 3, 0, 2, 1}, pat5[] = {0, 1, 2}, pat7[] = {0, 3, 1, 1};
  const uint8_t *pat = nullptr;
  switch (rate) {                                        // VCC:299-339
    case 1: pat = pat1; k.punc_len = 1; break;
    case 2: pat = pat2; k.punc_len = 3; break;
    case 3: pat = pat3; k.punc_len = 2; break;
    case 4: pat = pat4; k.punc_len = 5; break;
    case 5: pat = pat5; k.punc_len = 3; break;
    case 7: pat = pat7; k.punc_len = 4; break;
    default: return 2;
  }
  for (uint32_t i = 0; i < k.punc_len; i++) k.punc[i] = pat[i];
  uint32_t consumed = 0;                                 // VCC:344-357
  k.P = 1;
  k.pos2msg[0] = 0;
  while (consumed < msg_len + m) {
    consumed += (k.punc[(k.P - 1) % k.punc_len] == 0) ? 1 : 2;
    if (k.P >= 259) return 3;
    k.pos2msg[k.P++] = consumed;
  }
  if (consumed != msg_len + m) return 3;
  if (rc) {                                              // VCC:359-386
    k.G[0] = bitrev(k.G[0], m + 1);
    k.G[1] = bitrev(k.G[1], m + 1);
    uint32_t t = bitrev(k.init, m);
    k.init = bitrev(k.fin, m);
    k.fin = t;
    uint8_t old[16];
    for (uint32_t i = 0; i < k.punc_len; i++) old[i] = k.punc[i];
    uint32_t end_idx = (k.P - 2) % k.punc_len;
    static const uint8_t mirror[4] = {0, 2, 1, 3};
    for (uint32_t i = 0; i < k.punc_len; i++)
      k.punc[i] = mirror[old[(k.punc_len - i + end_idx) % k.punc_len]];
    std::reverse(k.pos2msg, k.pos2msg + k.P);
    for (uint32_t i = 0; i < k.P; i++) k.pos2msg[i] = msg_len + m - k.pos2msg[i];
  }
  if (marker.empty()) return 0;                          // VCC:388-414
  k.marker_len = (uint32_t)marker.size();
  if (k.marker_len >= 256) return 4;
  k.marker_period = period;
  if (period < k.marker_len) return 5;
  for (uint32_t i = 0; i < k.marker_len; i++) {
    if (marker[i] == '0') k.marker[i] = false;
    else if (marker[i] == '1') k.marker[i] = true;
    else return 6;
  }
  return 0;
}

inline uint32_t step_back(const Code &k, uint32_t c, uint32_t bit) {      // VCC:433-438
  return ((c << 1) | bit) & (k.nconv - 1);
}
inline uint32_t step_fwd(const Code &k, uint32_t c, uint32_t bit) {       // VCC:426-431
  return (c | (bit << k.m)) >> 1;
}
inline uint32_t emit(const Code &k, int which, uint32_t c, uint32_t bit) {  // VCC:440-448
  return (uint32_t)__builtin_parity((c | (bit << k.m)) & k.G[which]) ^ (k.rc ? 1u : 0u);
}

bool state_ok(const Code &k, uint32_t msg_pos, uint32_t c) {               // VCC:944-978
  for (uint32_t sh = 0; sh < k.m; sh++) {
    int64_t q = (int64_t)msg_pos - 1 - (int64_t)sh;
    int64_t qf = k.rc ? (int64_t)k.msg_len - 1 - q : q;
    uint32_t bit = (c >> (k.m - 1 - sh)) & 1u;
    if (q < 0) {
      if (bit != ((k.init >> (k.m + q)) & 1u)) return false;
    } else if (q >= (int64_t)k.msg_len) {
      if (bit != ((k.fin >> (q - k.msg_len)) & 1u)) return false;
    } else if (k.marker_len > 0 && (uint64_t)qf % k.marker_period < k.marker_len) {
      // VCC:969 mixes int64 and uint32 (signed arithmetic); qf >= 0 on this branch.
      if (bit != (k.marker[(uint64_t)qf % k.marker_period] ? 1u : 0u)) return false;
    }
  }
  return true;
}

struct Edge {
  uint32_t conv;
  uint8_t crf, row, col, shift, bits;
};

// Predecessor list of (c2, s2) under puncturing block `blk`, in the reference's
// psidx order (VCC:860-942): stay first, then source crf ascending.
int edges_into(const Code &k, uint32_t c2, uint32_t s2, uint32_t blk, Edge *out) {
  int n = 0;
  const uint8_t row = (uint8_t)(s2 >= 4 ? 4 : s2);       // VCC:582-587
  const uint32_t top = (c2 >> (k.m - 1)) & 1u, top1 = (c2 >> (k.m - 2)) & 1u;
  out[n++] = Edge{c2, (uint8_t)s2, row, (uint8_t)s2, 0, 0};
  for (uint32_t s1 = 0; s1 < kCrf; s1++) {
    if (s2 >= 4 && !(s1 == s2 || s1 == s2 - 4)) continue;
    if (s1 == s2) continue;
    const uint32_t want = s2 % 4;
    if (blk == 0) {
      for (uint32_t b = 0; b < 2; b++) {
        uint32_t c1 = step_back(k, c2, b);
        if (2 * emit(k, 0, c1, top) + emit(k, 1, c1, top) == want)
          out[n++] = Edge{c1, (uint8_t)s1, row, (uint8_t)s1, 1, (uint8_t)top};
      }
    } else {
      for (uint32_t b = 0; b < 2; b++)
        for (uint32_t b1 = 0; b1 < 2; b1++) {
          uint32_t cm = step_back(k, c2, b), c1 = step_back(k, cm, b1);
          uint32_t o0 = emit(k, 0, c1, top1), o1 = emit(k, 1, c1, top1);
          uint32_t o2 = emit(k, 0, cm, top), o3 = emit(k, 1, cm, top);
          uint32_t base = 0;
          if (blk == 1) base = k.rc ? 2 * o2 + o1 : 2 * o1 + o2;
          else if (blk == 2) base = k.rc ? 2 * o3 + o0 : 2 * o0 + o3;
          else base = k.rc ? 2 * o3 + o1 : 2 * o1 + o3;
          if (base == want)
            out[n++] = Edge{c1, (uint8_t)s1, row, (uint8_t)s1, 2, (uint8_t)(2 * top1 + top)};
        }
    }
  }
  return n;
}

struct Msg {
  uint64_t w[kMsgWords];
  bool operator==(const Msg &o) const {
    return w[0] == o.w[0] && w[1] == o.w[1] && w[2] == o.w[2] && w[3] == o.w[3];
  }
};
inline Msg extend(const Msg &a, unsigned sh, unsigned bits) {  // (msg << sh) | bits on 256 bits
  if (sh == 0) return a;
  Msg r;
  for (int i = kMsgWords - 1; i > 0; i--) r.w[i] = (a.w[i] << sh) | (a.w[i - 1] >> (64 - sh));
  r.w[0] = (a.w[0] << sh) | bits;
  return r;
}

struct HeapItem {   // VCC:49-55: ordering looks at the score only
  float score;
  uint32_t psidx, slot;
  size_t src;
  bool operator<(const HeapItem &o) const { return score < o.score; }
};

struct Final {
  Msg msg;
  float score;
};

// optional statistics for kernel design (tests never enable them): pops per state update,
// split by target kind: [0] flip with predecessors, [1] flip without, [2] flop with, [3] flop without
bool g_stats_on = false;
unsigned long long g_pop_hist[4][80];
unsigned long long g_acc_hist[4][80];

}  // namespace

extern "C" {

struct lvo_params {
  uint32_t mem_conv, rate, msg_len, list_size, max_deviation;  // max_deviation 0xFFFFFFFF = default
  int32_t rc;
  const char *sync_marker;  // may be NULL / ""
  uint32_t sync_period;
  uint32_t num_thr;
  int32_t unfused_band;     // test knob: 1 = evaluate the band expression without FMA
};

// Band start for step t (VCC:677-679 as compiled by install.sh:10 on an FMA machine:
// the multiply and subtract are contracted).
int64_t lvo_band_start(uint32_t t, uint32_t nblk, uint32_t P, uint32_t dev, int unfused) {
  double q = (double)t / (double)nblk;
  double v;
  if (unfused) {
    volatile double prod = q * (double)P;
    v = prod - (double)dev;
  } else {
    v = std::fma(q, (double)P, -(double)dev);
  }
  int64_t a = (int64_t)v;
  return a < 0 ? 0 : a;
}

int lvo_tables(uint32_t m, uint32_t rate, uint32_t msg_len, int rc, uint32_t *G_out,
               uint32_t *init_final, uint8_t *punc, uint32_t *punc_len, uint32_t *P_out,
               uint32_t *pos2msg) {
  Code k;
  int st = build_code(k, m, rate, msg_len, rc != 0, "", 0);
  if (st) return st;
  G_out[0] = k.G[0];
  G_out[1] = k.G[1];
  init_final[0] = k.init;
  init_final[1] = k.fin;
  *punc_len = k.punc_len;
  for (uint32_t i = 0; i < k.punc_len; i++) punc[i] = k.punc[i];
  *P_out = k.P;
  for (uint32_t i = 0; i < k.P && i < 256; i++) pos2msg[i] = k.pos2msg[i];
  return 0;
}

// VCC:450-499.  out must hold 2*(msg_len+m) bytes.
int lvo_encode(uint32_t m, uint32_t rate, const uint8_t *msg, uint32_t msg_len, uint8_t *out,
               uint32_t *n_out) {
  Code k;
  int st = build_code(k, m, rate, msg_len, false, "", 0);
  if (st) return st;
  std::vector<uint8_t> full;
  full.reserve(2 * (msg_len + m));
  uint32_t c = k.init;
  for (uint32_t i = 0; i < msg_len + m; i++) {
    uint32_t bit = i < msg_len ? (msg[i] ? 1u : 0u) : ((k.fin >> (i - msg_len)) & 1u);
    full.push_back((uint8_t)emit(k, 0, c, bit));
    full.push_back((uint8_t)emit(k, 1, c, bit));
    c = step_fwd(k, c, bit);
  }
  if (c != k.fin) return 10;
  uint32_t n = 0;
  size_t i = 0;
  for (uint32_t p = 0; p + 1 < k.P; p++) {
    switch (k.punc[p % k.punc_len]) {
      case 0: out[n++] = full[i];     out[n++] = full[i + 1]; i += 2; break;
      case 1: out[n++] = full[i + 1]; out[n++] = full[i + 2]; i += 4; break;
      case 2: out[n++] = full[i];     out[n++] = full[i + 3]; i += 4; break;
      default: out[n++] = full[i + 1]; out[n++] = full[i + 3]; i += 4; break;
    }
  }
  if (i != full.size()) return 11;
  *n_out = n;
  return 0;
}

// Decode one read.  post = nblk x 40 floats in file order (VCC:553-575 reads the file
// straight into [5][8] blocks, so post[t][row][col] = post[t*40 + row*8 + col]).
// Returns 0 ok; 1..6 parameter errors (build_code); 20 "Too many states" (VCC:595-597);
// 21 "Too small post matrix" (VCC:600-601); 22 msg_len > 256 (VCC:604-605).
// out_bits: list_size rows of msg_len bytes; out_scores: list_size floats.
int lvo_decode(const lvo_params *prm, const float *post, uint32_t nblk, uint8_t *out_bits,
               float *out_scores, uint32_t *n_out) {
  Code k;
  int st = build_code(k, prm->mem_conv, prm->rate, prm->msg_len, prm->rc != 0,
                      prm->sync_marker ? prm->sync_marker : "", prm->sync_period);
  if (st) return st;
  const uint32_t L = prm->list_size, P = k.P, C = k.nconv, m = k.m, msg_len = prm->msg_len;
  const float NEG = -std::numeric_limits<float>::infinity();
  if ((uint64_t)P * kCrf * C >= ((uint64_t)1 << 32)) return 20;
  if (nblk < P + 1) return 21;
  if (msg_len > 256) return 22;
  uint32_t dev = prm->max_deviation == 0xFFFFFFFFu ? msg_len + m + 1 : prm->max_deviation;  // VCC:238-242
#ifdef _OPENMP
  int nthr = prm->num_thr ? (int)prm->num_thr : 1;
#endif

  const size_t nstate = (size_t)P * C * kCrf;
  std::vector<float> scoreA(nstate * L, NEG), scoreB(nstate * L, NEG);     // VCC:607-620
  std::vector<Msg> msgA(nstate * L, Msg{{0, 0, 0, 0}}), msgB(nstate * L, Msg{{0, 0, 0, 0}});
  float *cur_s = scoreA.data(), *prv_s = scoreB.data();
  Msg *cur_m = msgA.data(), *prv_m = msgB.data();

  std::vector<uint8_t> ok((size_t)P * C);                                   // VCC:624-630
#pragma omp parallel for num_threads(nthr)
  for (uint32_t p = 0; p < P; p++)
    for (uint32_t c = 0; c < C; c++) ok[(size_t)p * C + c] = state_ok(k, k.pos2msg[p], c);

  // predecessor tables per block type actually used (VCC:636-650)
  std::vector<Edge> etab[4];
  std::vector<uint8_t> ecnt[4];
  for (uint32_t blk = 0; blk < 4; blk++) {
    bool used = false;
    for (uint32_t i = 0; i < k.punc_len; i++) used |= (k.punc[i] == blk);
    if (!used) continue;
    etab[blk].resize((size_t)C * kCrf * 8);
    ecnt[blk].resize((size_t)C * kCrf);
#pragma omp parallel for num_threads(nthr)
    for (uint32_t c = 0; c < C; c++)
      for (uint32_t s = 0; s < kCrf; s++)
        ecnt[blk][(size_t)c * kCrf + s] =
            (uint8_t)edges_into(k, c, s, blk, &etab[blk][((size_t)c * kCrf + s) * 8]);
  }
  // block 0's table is also what position 0 indexes (VCC:693-705); make sure it exists
  if (etab[0].empty()) {
    etab[0].resize((size_t)C * kCrf * 8);
    ecnt[0].resize((size_t)C * kCrf);
    for (uint32_t c = 0; c < C; c++)
      for (uint32_t s = 0; s < kCrf; s++)
        ecnt[0][(size_t)c * kCrf + s] =
            (uint8_t)edges_into(k, c, s, 0, &etab[0][((size_t)c * kCrf + s) * 8]);
  }

  auto sidx = [&](uint32_t p, uint32_t c, uint32_t s) -> size_t {            // VCC:577-580
    return ((size_t)p * C + c) * kCrf + s;
  };
  for (uint32_t s = 0; s < kCrf; s++) cur_s[sidx(0, k.init, s) * L] = 0.0f;  // VCC:657-663

  for (uint32_t t = 0; t < nblk; t++) {                                      // VCC:667
    std::swap(cur_s, prv_s);
    std::swap(cur_m, prv_m);
    const float *w = post + (size_t)t * 40;
    int64_t a64 = lvo_band_start(t, nblk, P, dev, prm->unfused_band);
    const uint32_t a = (uint32_t)a64;
    const uint32_t b = std::min(a + 2u * dev, P);                            // VCC:679 (uint32 math)
#pragma omp parallel for schedule(dynamic) num_threads(nthr)
    for (uint32_t p = a; p < b; p++) {
      std::vector<HeapItem> heap;
      heap.reserve(16);
      const uint32_t blk = p ? k.punc[(p - 1) % k.punc_len] : 0;             // VCC:693-696
      for (uint32_t c = 0; c < C; c++) {
        if (!ok[(size_t)p * C + c]) continue;
        for (uint32_t s = 0; s < kCrf; s++) {
          const size_t x = sidx(p, c, s);
          const Edge *e = &etab[blk][((size_t)c * kCrf + s) * 8];
          const int ne = ecnt[blk][(size_t)c * kCrf + s];
          if (p == 0) {                                                      // VCC:706-713
            cur_m[x * L] = prv_m[x * L];
            cur_s[x * L] = prv_s[x * L] + w[e[0].row * 8 + e[0].col];
            for (uint32_t l = 1; l < L; l++) cur_s[x * L + l] = NEG;
          } else if (L == 1) {                                               // VCC:715-742
            float best = NEG;
            int bi = 0;
            size_t bsrc = 0;
            for (int i = 0; i < ne; i++) {
              size_t src = sidx(p - (i ? 1 : 0), e[i].conv, e[i].crf);
              float sc = prv_s[src] + w[e[i].row * 8 + e[i].col];
              if (sc > best) { best = sc; bi = i; bsrc = src; }
            }
            cur_s[x] = best;
            if (best != NEG) cur_m[x] = extend(prv_m[bsrc], e[bi].shift, e[bi].bits);
          } else {                                                           // VCC:743-800
            heap.clear();
            for (int i = 0; i < ne; i++) {
              size_t src = sidx(p - (i ? 1 : 0), e[i].conv, e[i].crf);
              if (prv_s[src * L] != NEG)
                heap.push_back(HeapItem{prv_s[src * L] + w[e[i].row * 8 + e[i].col],
                                        (uint32_t)i, 0, src});
            }
            std::make_heap(heap.begin(), heap.end());
            uint32_t l = 0;
            uint32_t npop = 0;
            while (!heap.empty() && l < L) {
              npop++;
              std::pop_heap(heap.begin(), heap.end());
              HeapItem h = heap.back();
              heap.pop_back();
              Msg mm = extend(prv_m[h.src * L + h.slot], e[h.psidx].shift, e[h.psidx].bits);
              if (std::find(cur_m + x * L, cur_m + x * L + l, mm) == cur_m + x * L + l) {
                cur_m[x * L + l] = mm;
                cur_s[x * L + l] = h.score;
                l++;
              }
              if (h.slot == L - 1) continue;
              float nxt = prv_s[h.src * L + h.slot + 1];
              if (nxt != NEG) {
                heap.push_back(HeapItem{nxt + w[e[h.psidx].row * 8 + e[h.psidx].col], h.psidx,
                                        h.slot + 1, h.src});
                std::push_heap(heap.begin(), heap.end());
              }
            }
            if (g_stats_on) {
              const int kind = (s >= 4 ? 2 : 0) + (ne > 1 ? 0 : 1);
#pragma omp atomic
              g_pop_hist[kind][npop < 79 ? npop : 79]++;
#pragma omp atomic
              g_acc_hist[kind][l < 79 ? l : 79]++;
            }
            for (; l < L; l++) cur_s[x * L + l] = NEG;
          }
        }
      }
    }
  }

  std::vector<Final> fin;                                                    // VCC:806-815
  for (uint32_t s = 0; s < kCrf; s++) {
    size_t x = sidx(P - 1, k.fin, s);
    for (uint32_t l = 0; l < L; l++)
      if (cur_s[x * L + l] != NEG) fin.push_back(Final{cur_m[x * L + l], cur_s[x * L + l]});
  }
  std::sort(fin.begin(), fin.end(),                                          // VCC:818-821
            [](const Final &p1, const Final &p2) -> bool { return p1.score > p2.score; });
  if (fin.size() > L) fin.resize(L);                                         // VCC:823-824
  *n_out = (uint32_t)fin.size();
  for (size_t i = 0; i < fin.size(); i++) {                                  // VCC:829-836
    out_scores[i] = fin[i].score;
    for (uint32_t j = 0; j < msg_len; j++) {
      uint32_t bit = msg_len + m - 1 - j;
      // std::bitset<256>::operator[] beyond 255 is undefined in the reference; callers
      // keep msg_len + m <= 256
      uint8_t v = bit < 256 ? (uint8_t)((fin[i].msg.w[bit >> 6] >> (bit & 63)) & 1u) : 0;
      out_bits[i * msg_len + (prm->rc ? msg_len - 1 - j : j)] = v;
    }
  }
  return 0;
}

void lvo_stats(int on, unsigned long long *pop_hist, unsigned long long *acc_hist) {
  if (pop_hist) memcpy(pop_hist, g_pop_hist, sizeof g_pop_hist);
  if (acc_hist) memcpy(acc_hist, g_acc_hist, sizeof g_acc_hist);
  if (on >= 0) {
    g_stats_on = on != 0;
    memset(g_pop_hist, 0, sizeof g_pop_hist);
    memset(g_acc_hist, 0, sizeof g_acc_hist);
  }
}

// --- libstdc++ heap primitives exposed for the device heap-emulation unit test ---------
// op: 0 = make_heap, 1 = pop_heap (moves max to the back), 2 = push_heap (last element)
void lvo_std_heap_op(int op, float *score, uint32_t *tag, uint32_t n) {
  std::vector<HeapItem> h(n);
  for (uint32_t i = 0; i < n; i++) h[i] = HeapItem{score[i], tag[i], 0, 0};
  if (op == 0) std::make_heap(h.begin(), h.end());
  else if (op == 1) std::pop_heap(h.begin(), h.end());
  else std::push_heap(h.begin(), h.end());
  for (uint32_t i = 0; i < n; i++) { score[i] = h[i].score; tag[i] = h[i].psidx; }
}

}  // extern "C"
