// viterbi_nanopore.out -- drop-in for the reference tool
//   /root/reference/viterbi/viterbi_convolutional_code.cpp  ("VCC": main at :137-262)
// Same options, same input/output files, same messages and exit codes; `decode` runs on the GPU
// through the C ABI (include/lva_b200.h), `encode` stays on the host.  There is no CPU decode.
//
// The reference parses argv with a vendored cxxopts 2.2.0; that header is third-party code and
// is NOT re-used here.  This file re-implements the subset of its behaviour that the tool and
// its callers (generate_decoded_lists.py:86-90, simulator.py:82-85, helper.py:305) can observe:
//   * "--name value", "--name=value", "-n value", grouped short flags; boolean flags take no value
//   * an argv element that is not option-shaped is ignored unless it starts with '-' (the callers
//     pass a literal '' when the read is not reverse-complemented)
//   * unknown option / missing argument / malformed number -> uncaught exception in the
//     reference, i.e. abort(); reproduced with the same what() text on stderr and SIGABRT
//   * integers: optional '-', optional 0x, range-checked against the option's type (uint8/uint32)
//
// Extensions (rejected by the reference, so they cannot change its meaning):
//   --batch-list FILE   decode many reads in one process: each line "infile outfile [rc]"
//   --post-dir DIR --out-prefix P   decode every DIR/*.post (name order) into P_<i>, the naming of
//                       the reference's driver loop (generate_decoded_lists.py:85); DIR/<name>.post.rc
//                       marks a reverse-complement read
//   --info-file FILE    one line per read of a batch: input, list file, orientation, blocks, entries, status
//   --device N          CUDA device ordinal (default 0)
//   --devices LIST      several GPUs of the box ("0-7", "0,2,3", "all"): one worker thread and one
//                       handle per GPU, reads dealt longest-first, forward / rc reads apart
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include <dirent.h>
#include <thread>

#include "../../include/lva_b200.h"

namespace {

const char *kHelp =
    "Viterbi decoder for nanopore dna storage codes\n"
    "Usage:\n"
    "  viterbi_nanopore [OPTION...]\n"
    "\n"
    "  -m, --mode arg           Mode: encode, decode\n"
    "  -i, --infile arg         Infile with message (encoding) or posterior matrix\n"
    "                           (decoding)\n"
    "  -o, --outfile arg        Outfile with encoded/decoded message (list)\n"
    "      --msg-len arg        Message length\n"
    "      --mem-conv arg       Code memory for convolutional code\n"
    "      --sync-marker arg    Sync marker for convolutional code decoding as\n"
    "                           string (e.g. 110) (default '') (default: )\n"
    "      --sync-period arg    Sync marker period for convolutional code decoding\n"
    "  -l, --list-size arg      List size for convolutional code decoding (default\n"
    "                           1) (default: 1)\n"
    "  -r, --rate arg           Rate of convolutional code: options 1 (1/2), 2\n"
    "                           (2/3), 3 (3/4), 4 (4/5), 5 (5/6), 7 (7/8) (default 1).\n"
    "                           Use standard puncturing patterns, expects\n"
    "                           appropriate padding (at most 1 bit needed) to make output\n"
    "                           length even. (default: 1)\n"
    "      --max-deviation arg  Max allowable deviation of st_pos around its\n"
    "                           expected value during decoding (tradeoff b/w speed and\n"
    "                           accuracy) (default: infinite)\n"
    "      --rc                 Reverse complement read (for decoding)\n"
    "  -t, --num-thr arg        Number of threads for convolutional code decoding\n"
    "                           (default 1) (default: 1)\n"
    "  -h, --help               Display this message\n";

// what the reference does on an uncaught exception: std::terminate -> abort (exit status -6)
[[noreturn]] void die_like_uncaught(const char *type, const std::string &what) {
  std::cout.flush();
  fprintf(stderr, "terminate called after throwing an instance of '%s'\n  what():  %s\n", type, what.c_str());
  std::abort();
}
const char *LQ = "\xE2\x80\x98", *RQ = "\xE2\x80\x99";  // cxxopts quotes option names with ‘ ’

enum Kind { kString, kU8, kU32, kBool };
struct Opt {
  const char *lname;
  char sname;
  Kind kind;
  bool has_default;
  const char *def;
};
const Opt kOpts[] = {
    {"mode", 'm', kString, false, ""},        {"infile", 'i', kString, false, ""},
    {"outfile", 'o', kString, false, ""},     {"msg-len", 0, kU32, false, ""},
    {"mem-conv", 0, kU8, false, ""},          {"sync-marker", 0, kString, true, ""},
    {"sync-period", 0, kU32, false, ""},      {"list-size", 'l', kU32, true, "1"},
    {"rate", 'r', kU8, true, "1"},            {"max-deviation", 0, kU32, false, ""},
    {"rc", 0, kBool, true, "false"},          {"num-thr", 't', kU32, true, "1"},
    {"help", 'h', kBool, true, "false"},
    // extensions
    {"batch-list", 0, kString, false, ""},    {"device", 0, kU32, true, "0"},
    {"devices", 0, kString, false, ""},       {"post-dir", 0, kString, false, ""},
    {"out-prefix", 0, kString, false, ""},    {"info-file", 0, kString, false, ""},
};

struct Parsed {
  std::map<std::string, int> count;
  std::map<std::string, std::string> text;
  std::map<std::string, uint64_t> num;
  int n(const std::string &k) const { auto it = count.find(k); return it == count.end() ? 0 : it->second; }
};

bool is_alnum(char c) { return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z'); }

// cxxopts integer_parser: "(-)?(0x)?([0-9a-zA-Z]+)|((0x)?0)"
uint64_t parse_unsigned(const std::string &t, uint64_t umax) {
  if (t == "0" || t == "0x0") return 0;
  size_t i = 0;
  bool neg = false;
  if (i < t.size() && t[i] == '-') { neg = true; i++; }
  int base = 10;
  if (t.compare(i, 2, "0x") == 0) { base = 16; i += 2; }
  if (i >= t.size()) die_like_uncaught("cxxopts::argument_incorrect_type",
                                      std::string("Argument ") + LQ + t + RQ + " failed to parse");
  uint64_t r = 0;
  for (; i < t.size(); i++) {
    const char c = t[i];
    uint64_t d;
    if (c >= '0' && c <= '9') d = c - '0';
    else if (base == 16 && c >= 'a' && c <= 'f') d = c - 'a' + 10;
    else if (base == 16 && c >= 'A' && c <= 'F') d = c - 'A' + 10;
    else die_like_uncaught("cxxopts::argument_incorrect_type",
                           std::string("Argument ") + LQ + t + RQ + " failed to parse");
    if (umax - d < r * base) die_like_uncaught("cxxopts::argument_incorrect_type",
                                              std::string("Argument ") + LQ + t + RQ + " failed to parse");
    r = r * base + d;
  }
  if (neg) die_like_uncaught("cxxopts::argument_incorrect_type",
                             std::string("Argument ") + LQ + t + RQ + " failed to parse");
  return r;
}

void store(Parsed &P, const Opt &o, const std::string &arg) {
  P.count[o.lname]++;
  P.text[o.lname] = arg;
  if (o.kind == kU8) P.num[o.lname] = parse_unsigned(arg, 0xFF);
  else if (o.kind == kU32) P.num[o.lname] = parse_unsigned(arg, 0xFFFFFFFFull);
  else if (o.kind == kBool) {
    // "(t|T)(rue)?" true, "((f|F)(alse)?)?" false
    if (arg == "t" || arg == "T" || arg == "true" || arg == "True") P.num[o.lname] = 1;
    else if (arg == "" || arg == "f" || arg == "F" || arg == "false" || arg == "False") P.num[o.lname] = 0;
    else die_like_uncaught("cxxopts::argument_incorrect_type",
                           std::string("Argument ") + LQ + arg + RQ + " failed to parse");
  }
}

const Opt *find_long(const std::string &n) {
  for (const Opt &o : kOpts) if (n == o.lname) return &o;
  return nullptr;
}
const Opt *find_short(char c) {
  for (const Opt &o : kOpts) if (o.sname && o.sname == c) return &o;
  return nullptr;
}

Parsed parse_args(int argc, char **argv) {
  Parsed P;
  for (int cur = 1; cur < argc; cur++) {
    const std::string a = argv[cur];
    if (a == "--") break;  // everything after is positional; none are declared, so it is ignored
    // "--([[:alnum:]][-_[:alnum:]]+)(=(.*))?"
    bool long_ok = false;
    std::string name, val;
    bool has_eq = false;
    if (a.size() >= 4 && a[0] == '-' && a[1] == '-' && is_alnum(a[2])) {
      size_t i = 3;
      while (i < a.size() && (is_alnum(a[i]) || a[i] == '-' || a[i] == '_')) i++;
      if (i >= 4 && (i == a.size() || a[i] == '=')) {
        long_ok = true;
        name = a.substr(2, i - 2);
        if (i < a.size()) { has_eq = true; val = a.substr(i + 1); }
      }
    }
    bool short_ok = false;
    if (!long_ok && a.size() >= 2 && a[0] == '-') {
      short_ok = true;
      for (size_t i = 1; i < a.size(); i++) short_ok = short_ok && is_alnum(a[i]);
    }
    if (!long_ok && !short_ok) {
      if (a.size() >= 2 && a[0] == '-')
        die_like_uncaught("cxxopts::option_syntax_exception",
                          std::string("Argument ") + LQ + a + RQ + " starts with a - but has incorrect syntax");
      continue;  // not option-shaped (e.g. the '' the Python callers pass): ignored
    }
    auto take_value = [&](const Opt &o, const std::string &nm) {
      if (o.kind == kBool) { store(P, o, "true"); return; }  // implicit value, never consumes argv
      if (cur + 1 >= argc)
        die_like_uncaught("cxxopts::missing_argument_exception",
                          std::string("Option ") + LQ + nm + RQ + " is missing an argument");
      store(P, o, argv[++cur]);
    };
    if (long_ok) {
      const Opt *o = find_long(name);
      if (!o) die_like_uncaught("cxxopts::option_not_exists_exception",
                                std::string("Option ") + LQ + name + RQ + " does not exist");
      if (has_eq) store(P, *o, val);
      else take_value(*o, name);
    } else {
      for (size_t i = 1; i < a.size(); i++) {
        const std::string nm(1, a[i]);
        const Opt *o = find_short(a[i]);
        if (!o) die_like_uncaught("cxxopts::option_not_exists_exception",
                                  std::string("Option ") + LQ + nm + RQ + " does not exist");
        if (i + 1 == a.size()) take_value(*o, nm);
        else if (o->kind == kBool) store(P, *o, "true");
        else die_like_uncaught("cxxopts::option_requires_argument_exception",
                               std::string("Option ") + LQ + nm + RQ + " requires an argument");
      }
    }
  }
  for (const Opt &o : kOpts)
    if (!P.n(o.lname) && o.has_default) {
      int keep = 0;
      store(P, o, o.def);
      P.count[o.lname] = keep;  // defaults do not count as "given"
    }
  return P;
}

// read_crf_post (VCC:553-575), including its behaviour on a trailing partial block: the last
// value read is repeated for the floats that are missing.
std::vector<float> read_posteriors(const std::string &path) {
  std::vector<float> f;
  std::ifstream fin(path, std::ios::binary);
  if (fin) {
    fin.seekg(0, std::ios::end);
    const std::streamoff sz = fin.tellg();
    fin.seekg(0, std::ios::beg);
    const size_t n = sz > 0 ? (size_t)sz / 4 : 0;
    f.resize(n);
    if (n) fin.read(reinterpret_cast<char *>(f.data()), (std::streamsize)(n * 4));
  }
  if (!f.empty() && f.size() % 40 != 0) {
    const float last = f.back();
    f.resize((f.size() + 39) / 40 * 40, last);
  }
  return f;
}

int usage_error(const char *msg) {
  std::cout << msg;
  std::cout << kHelp << "\n";
  return -1;
}

struct Job {
  std::string in, out;
  bool rc = false;
  uint64_t blocks = 0;   // posterior blocks in the file (work estimate)
  int status = 0;        // 0 ok, 1 too small a posterior, 2 decode failure
  uint32_t entries = 0;  // list entries written
};

lva_params decode_params(const Parsed &P, const std::string &marker, int rc, int device) {
  lva_params prm{};
  prm.mem_conv = (uint32_t)P.num.at("mem-conv");
  prm.rate = (uint32_t)P.num.at("rate");
  prm.msg_len = (uint32_t)P.num.at("msg-len");
  prm.list_size = (uint32_t)P.num.at("list-size");
  prm.max_deviation = P.n("max-deviation") ? (uint32_t)P.num.at("max-deviation") : LVA_DEFAULT_DEVIATION;
  prm.rc = rc;
  prm.sync_marker = marker.c_str();
  prm.sync_period = P.n("sync-period") ? (uint32_t)P.num.at("sync-period") : 0;
  prm.device = device;
  return prm;
}

// One GPU's share of a batch: the reads of one orientation, decoded in chunks of a few slot-fulls
// through ONE handle (tables, state memory and chain tables are set up once per handle).
// Posterior files are read by the worker itself, so the GPUs of a box also read in parallel.
bool decode_on_device(lva_handle *h, uint32_t msg_len, uint32_t L, std::vector<Job> &jobs, const std::vector<size_t> &mine,
                      std::string &err) {
  const size_t chunk_reads = std::max<size_t>(64, 4 * (size_t)lva_slots(h));
  const uint64_t chunk_blocks = (uint64_t)1 << 22;  // 640 MB of posteriors per call at most
  size_t at = 0;
  while (at < mine.size()) {
    std::vector<float> post;
    std::vector<uint64_t> offs(1, 0);
    size_t n = 0;
    while (at + n < mine.size() && n < chunk_reads && (n == 0 || offs.back() + jobs[mine[at + n]].blocks <= chunk_blocks)) {
      std::vector<float> f = read_posteriors(jobs[mine[at + n]].in);
      post.insert(post.end(), f.begin(), f.end());
      offs.push_back(post.size() / 40);
      n++;
    }
    std::vector<uint8_t> bits(n * L * msg_len);
    std::vector<uint32_t> counts(n);
    std::vector<int32_t> status(n);
    if (post.empty()) post.push_back(0.f);
    const int st = lva_decode_batch(h, post.data(), offs.data(), (uint32_t)n, bits.data(), nullptr, counts.data(), status.data());
    if (st != LVA_OK) {
      err = lva_last_error(h);
      for (size_t r = 0; r < n; r++) jobs[mine[at + r]].status = 2;
      return false;
    }
    for (size_t r = 0; r < n; r++) {
      Job &j = jobs[mine[at + r]];
      if (status[r] == LVA_ERR_POST_TOO_SMALL) { j.status = 1; continue; }
      std::ofstream fout(j.out);  // VCC:248-253
      for (uint32_t e = 0; e < counts[r]; e++) {
        const uint8_t *b = &bits[(r * L + e) * msg_len];
        std::string line(msg_len, '0');
        for (uint32_t k = 0; k < msg_len; k++) line[k] = b[k] ? '1' : '0';
        fout << line << "\n";
      }
      j.entries = counts[r];
    }
    at += n;
  }
  return true;
}

// Multi-GPU host of the batch extension: reads are independent, so they are
// dealt longest-first onto the least-loaded GPU, forward and reverse-complement reads apart (they
// need different code tables, i.e. different handles); one worker thread per GPU decodes its share
// and writes the list files by read; nothing is exchanged between GPUs.
int run_decode(const Parsed &P, std::vector<Job> &jobs, const std::vector<int> &devices) {
  const uint32_t msg_len = (uint32_t)P.num.at("msg-len"), L = (uint32_t)P.num.at("list-size");
  const std::string marker = P.text.at("sync-marker");
  for (Job &j : jobs) {
    std::ifstream f(j.in, std::ios::binary | std::ios::ate);
    j.blocks = f ? ((uint64_t)std::max<std::streamoff>(f.tellg(), 0) + 159) / 160 : 0;
  }
  const size_t ndev = devices.size();
  std::vector<std::vector<size_t>> share[2];  // [orientation][device] -> job indices
  for (int pass = 0; pass < 2; pass++) {
    share[pass].assign(ndev, {});
    std::vector<size_t> idx;
    for (size_t i = 0; i < jobs.size(); i++) if (jobs[i].rc == (pass == 1)) idx.push_back(i);
    std::stable_sort(idx.begin(), idx.end(), [&](size_t x, size_t y) { return jobs[x].blocks > jobs[y].blocks; });
    std::vector<uint64_t> load(ndev, 0);
    for (size_t i : idx) {
      const size_t d = (size_t)(std::min_element(load.begin(), load.end()) - load.begin());
      share[pass][d].push_back(i);
      load[d] += std::max<uint64_t>(jobs[i].blocks, 1);
    }
  }
  // The first handle is made on this thread so that parameter errors surface exactly like the
  // reference's set_conv_params failures (text + usage on stdout, return -1; VCC:211-214).
  lva_handle *first = nullptr;
  int first_pass = -1;
  size_t first_dev = 0;
  for (int pass = 0; pass < 2 && !first; pass++)
    for (size_t d = 0; d < ndev && !first; d++)
      if (!share[pass][d].empty()) {
        const lva_params prm = decode_params(P, marker, pass, devices[d]);
        int st = 0;
        char err[256] = {0};
        first = lva_create(&prm, &st, err, sizeof err);
        if (!first) {
          if (st >= LVA_ERR_MEM_CONV && st <= LVA_ERR_MARKER_CHAR) {
            std::cout << err;
            std::cout << kHelp << "\n";
            return -1;
          }
          if (st == LVA_ERR_TOO_MANY_STATES || st == LVA_ERR_MSG_TOO_LONG) die_like_uncaught("std::runtime_error", err);
          fprintf(stderr, "viterbi_nanopore: %s\n", err);
          return 1;
        }
        first_pass = pass;
        first_dev = d;
      }
  if (!first) return 0;  // nothing to decode
  std::vector<std::string> errors(ndev);
  auto worker = [&](size_t d) {
    for (int pass = 0; pass < 2; pass++) {
      if (share[pass][d].empty()) continue;
      lva_handle *h = (pass == first_pass && d == first_dev) ? first : nullptr;
      if (!h) {
        const lva_params prm = decode_params(P, marker, pass, devices[d]);
        int st = 0;
        char err[256] = {0};
        h = lva_create(&prm, &st, err, sizeof err);
        if (!h) {
          errors[d] = err;
          for (size_t i : share[pass][d]) jobs[i].status = 2;
          continue;
        }
      }
      std::string e;
      if (!decode_on_device(h, msg_len, L, jobs, share[pass][d], e)) errors[d] = e;
      lva_destroy(h);
    }
  };
  if (ndev == 1) {
    worker(0);
  } else {
    std::vector<std::thread> th;
    for (size_t d = 0; d < ndev; d++) th.emplace_back(worker, d);
    for (auto &t : th) t.join();
  }
  int rc_all = 0;
  for (size_t d = 0; d < ndev; d++)
    if (!errors[d].empty()) {
      fprintf(stderr, "viterbi_nanopore: device %d: decode failed: %s\n", devices[d], errors[d].c_str());
      rc_all = 1;
    }
  for (const Job &j : jobs)
    if (j.status == 1) {
      if (jobs.size() == 1) die_like_uncaught("std::runtime_error", "Too small post matrix");  // VCC:600-601
      fprintf(stderr, "viterbi_nanopore: %s: Too small post matrix\n", j.in.c_str());
      rc_all = 1;
    }
  return rc_all;
}

// "0-3", "0,2,5", "all" -> device ordinals (extension --devices)
std::vector<int> parse_devices(const std::string &spec) {
  std::vector<int> out;
  if (spec == "all") {
    for (int d = 0; d < lva_device_count(); d++) out.push_back(d);
    return out;
  }
  std::stringstream ss(spec);
  std::string tok;
  while (std::getline(ss, tok, ',')) {
    if (tok.empty()) continue;
    const size_t dash = tok.find('-');
    try {
      if (dash == std::string::npos) out.push_back(std::stoi(tok));
      else for (int d = std::stoi(tok.substr(0, dash)); d <= std::stoi(tok.substr(dash + 1)); d++) out.push_back(d);
    } catch (...) { out.clear(); return out; }
  }
  return out;
}

int run_encode(const Parsed &P) {
  const uint32_t msg_len = (uint32_t)P.num.at("msg-len");
  const uint32_t mem = (uint32_t)P.num.at("mem-conv"), rate = (uint32_t)P.num.at("rate");
  {  // parameter validation happens before the input is read (VCC:203-214)
    const int st = lva_check_params(mem, rate, msg_len);
    static const char *msgs[] = {"", "Invalid mem_conv (allowed: 6, 8, 11, 14)\n",
                                 "Invalid rate parameter (allowed: 1, 3, 5, 7)\n",
                                 "Output length not even. Try padding with a single 0 at end.\n"};
    if (st >= 1 && st <= 3) return usage_error(msgs[st]);
  }
  // read_bit_array (VCC:501-524): one message per '\n'-terminated line; an unterminated tail is dropped
  std::ifstream fin(P.text.at("infile"));
  std::vector<std::vector<uint8_t>> msgs;
  std::vector<uint8_t> cur;
  char ch;
  while (fin >> std::noskipws >> ch) {
    if (ch == '0') cur.push_back(0);
    else if (ch == '1') cur.push_back(1);
    else if (ch == '\n') { msgs.push_back(cur); cur.clear(); }
    else die_like_uncaught("std::runtime_error", "invalid character in input file");
  }
  std::vector<std::string> lines;
  for (auto &m : msgs) {
    if (m.size() != msg_len) {
      std::cout << "Message length does not match msg_len parameter.\n";
      return -1;
    }
    std::vector<uint8_t> enc(2 * (msg_len + mem) + 8);
    uint32_t n = 0;
    int st = lva_encode(mem, rate, m.data(), msg_len, enc.data(), (uint32_t)enc.size(), &n);
    if (st != LVA_OK) die_like_uncaught("std::runtime_error", "state after encoding does not match final_state_conv");
    static const char base[4] = {'A', 'C', 'G', 'T'};  // VCC:18, 540-551
    std::string s(n / 2, 'A');
    for (uint32_t i = 0; i < n / 2; i++) s[i] = base[2 * enc[2 * i] + enc[2 * i + 1]];
    lines.push_back(s);
  }
  std::ofstream fout(P.text.at("outfile"));
  for (auto &s : lines) fout << s << "\n";
  return 0;
}

}  // namespace

int main(int argc, char **argv) {
  const Parsed P = parse_args(argc, argv);
  if (P.n("help")) {
    std::cout << kHelp << "\n";
    return 0;
  }
  const bool dir_mode = P.n("post-dir") > 0;
  const bool batch = P.n("batch-list") > 0 || dir_mode;
  if (!P.n("mode") || (!batch && (!P.n("infile") || !P.n("outfile"))) || (dir_mode && !P.n("out-prefix")))
    return usage_error("Invalid options.\n");
  const std::string mode = P.text.at("mode");
  if (mode != "encode" && mode != "decode") return usage_error("Invalid mode.\n");
  if (!P.n("mem-conv")) return usage_error("Memory of convolutional code not specified.\n");
  if (!P.n("msg-len")) return usage_error("msg-len not specified.\n");
  const bool rc_flag = mode == "decode" && P.num.at("rc") != 0;
  if (rc_flag) std::cout << "Reverse complement flag detected.\n";
  if (P.text.at("sync-marker") != "" && !P.n("sync-period"))
    die_like_uncaught("std::domain_error", "No value");  // result["sync-period"].as<uint32_t>() on an absent option
  if (mode == "encode") {
    if (batch) return usage_error("Invalid options.\n");
    return run_encode(P);
  }
  std::vector<Job> jobs;
  if (dir_mode) {
    // every *.post of a directory, in name order, to <out-prefix>_<i> -- the naming of the reference's
    // driver loop (generate_decoded_lists.py:85); a file <name>.rc beside a posterior marks a
    // reverse-complement read
    std::vector<std::string> names;
    if (DIR *d = opendir(P.text.at("post-dir").c_str())) {
      while (dirent *e = readdir(d)) {
        const std::string n = e->d_name;
        if (n.size() > 5 && n.compare(n.size() - 5, 5, ".post") == 0) names.push_back(n);
      }
      closedir(d);
    } else {
      fprintf(stderr, "viterbi_nanopore: cannot open directory %s\n", P.text.at("post-dir").c_str());
      return 1;
    }
    std::sort(names.begin(), names.end());
    for (size_t i = 0; i < names.size(); i++) {
      Job j;
      j.in = P.text.at("post-dir") + "/" + names[i];
      j.out = P.text.at("out-prefix") + "_" + std::to_string(i);
      j.rc = rc_flag || std::ifstream(j.in + ".rc").good();
      jobs.push_back(j);
    }
  } else if (batch) {
    std::ifstream fl(P.text.at("batch-list"));
    std::string line;
    while (std::getline(fl, line)) {
      std::istringstream ss(line);
      Job j;
      std::string rcs;
      if (!(ss >> j.in >> j.out)) continue;
      j.rc = rc_flag;
      if (ss >> rcs) j.rc = (rcs == "rc" || rcs == "1" || rcs == "--rc");
      jobs.push_back(j);
    }
  } else {
    Job j;
    j.in = P.text.at("infile");
    j.out = P.text.at("outfile");
    j.rc = rc_flag;
    jobs.push_back(j);
  }
  std::vector<int> devices;
  if (P.n("devices")) {
    devices = parse_devices(P.text.at("devices"));
    if (devices.empty()) { fprintf(stderr, "viterbi_nanopore: bad --devices list\n"); return 1; }
  } else {
    devices.push_back((int)P.num.at("device"));
  }
  const int rc = run_decode(P, jobs, devices);
  if (P.n("info-file")) {  // one line per read: input, list file, orientation, blocks, entries, status
    std::ofstream fi(P.text.at("info-file"));
    for (const Job &j : jobs)
      fi << j.in << "\t" << j.out << "\t" << (j.rc ? "rc" : "fwd") << "\t" << j.blocks << "\t" << j.entries << "\t"
         << (j.status == 0 ? "ok" : j.status == 1 ? "too_small" : "failed") << "\n";
  }
  return rc;
}
