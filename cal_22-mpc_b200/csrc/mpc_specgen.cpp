// Config compiler for the specialised (thread-per-block) MPC kernel.
//
// Turns a flattened config (mpc_config_pod, i.e. what VPC::parseConfig reads, reference VPC.cpp:72-330) into the
// source of a straight-line schedule over the hand-written primitives of mpc_spec.cuh, with every table folded into
// immediates:
//   * predictor byte gathers (PredictorModule.cpp:37-173) and the root-first residue order (ResidueModule.cpp:26-39)
//     become PRMTs with constant selectors or plain register renaming,
//   * DiffTable bytes / WeightTable shift distances become immediate operands,
//   * the scan permutation (ScanModule.cpp:14-20) is recognised as column-major (a scan row = two residue bytes) or
//     plane-major (a scan row = one bit plane of 16 bytes); column-major modules are scored lazily, row by row, and
//     stop at their first non-zero row (VPC.cpp:378-387).
// The same generator serves the ahead-of-time build (tools/specgen -> csrc/spec/spec_<name>.cu, compiled by nvcc)
// and the run-time path (mpc_jit.cpp: NVRTC at mpc_create for configs that were not compiled in).
// Eligible: lineSize 32 / 64 / 128 and every PredComp scan table column-major or plane-major; everything else stays on the
// generic warp-per-block kernel.
#include "mpc_specgen.h"
#include "mpc_layout.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <set>
#include <string>
#include <vector>

namespace mpc {
namespace {

// Line geometry of the config being compiled (set by build_modules): L bytes = W words per line, R = L / 2 scan rows of 16 bits,
// NCH = L / 16 column chunks per bit plane.  A thread always holds 128 bytes = 128 / L consecutive lines (mpc_spec.cuh).
constexpr int kMaxL = 128;
thread_local int L = 128;
thread_local int W = 32;
inline int R() { return L / 2; }
inline int NCH() { return L / 16; }

std::string fmt(const char* f, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, f);
  int n = vsnprintf(buf, sizeof(buf), f, ap);
  va_end(ap);
  if (n < (int)sizeof(buf)) return std::string(buf, (size_t)std::max(n, 0));
  std::string big((size_t)n + 1, '\0');
  va_start(ap, f);
  vsnprintf(&big[0], big.size(), f, ap);
  va_end(ap);
  big.resize((size_t)n);
  return big;
}

std::string join(const std::vector<std::string>& v, const char* sep) {
  std::string s;
  for (size_t i = 0; i < v.size(); i++) {
    if (i) s += sep;
    s += v[i];
  }
  return s;
}

typedef std::vector<std::string> Lines;

// ---- byte-gather planning ---------------------------------------------------------------------------------------
// Expression for a 32-bit word whose byte q is byte srcs[q] of the 128-byte array `arr` (available as arr[0..31]),
// or zero when srcs[q] < 0.
std::string gather_expr(const std::string& arr, const int (&srcs)[4]) {
  bool any = false;
  for (int s : srcs) any |= (s >= 0);
  if (!any) return "0u";
  std::vector<int> words;
  for (int s : srcs)
    if (s >= 0 && std::find(words.begin(), words.end(), s / 4) == words.end()) words.push_back(s / 4);
  uint32_t mask = 0;
  for (int q = 0; q < 4; q++)
    if (srcs[q] >= 0) mask |= 0xFFu << (8 * q);
  auto w = [&](int i) { return fmt("%s[%d]", arr.c_str(), i); };
  auto finish = [&](const std::string& e) { return mask == 0xFFFFFFFFu ? e : fmt("(%s & 0x%08xu)", e.c_str(), mask); };
  bool identity = (words.size() == 1);
  for (int q = 0; q < 4 && identity; q++)
    if (srcs[q] >= 0 && srcs[q] % 4 != q) identity = false;
  if (identity) return finish(w(words[0]));
  if (words.size() <= 2) {
    const int a = words[0], b = words.size() == 2 ? words[1] : words[0];
    unsigned sel = 0;
    for (int q = 0; q < 4; q++) {
      unsigned nib = 0;
      if (srcs[q] >= 0) nib = (unsigned)(srcs[q] % 4) + (srcs[q] / 4 == a ? 0u : 4u);
      sel |= nib << (4 * q);
    }
    return finish(fmt("prmt(%s, %s, 0x%04xu)", w(a).c_str(), w(b).c_str(), sel));
  }
  // three or four source words: two partial gathers merged by a third PRMT
  std::vector<int> first(words.begin(), words.begin() + 2), second(words.begin() + 2, words.end());
  auto in = [](const std::vector<int>& ws, int word) { return std::find(ws.begin(), ws.end(), word) != ws.end(); };
  auto partial = [&](const std::vector<int>& ws) {
    const int a = ws[0], b = ws.size() == 2 ? ws[1] : ws[0];
    unsigned sel = 0;
    for (int q = 0; q < 4; q++) {
      unsigned nib = 0;
      if (srcs[q] >= 0 && in(ws, srcs[q] / 4)) nib = (unsigned)(srcs[q] % 4) + (srcs[q] / 4 == a ? 0u : 4u);
      sel |= nib << (4 * q);
    }
    if (ws.size() == 1 && sel == 0x3210u) return w(a);
    return fmt("prmt(%s, %s, 0x%04xu)", w(a).c_str(), w(b).c_str(), sel);
  };
  unsigned sel = 0;
  for (int q = 0; q < 4; q++) {
    const unsigned nib = (srcs[q] < 0 || in(first, srcs[q] / 4)) ? (unsigned)q : 4u + (unsigned)q;
    sel |= nib << (4 * q);
  }
  return finish(fmt("prmt(%s, %s, 0x%04xu)", partial(first).c_str(), partial(second).c_str(), sel));
}

// Canonical layout of a plane-major scan (mpc_spec.cuh, pm_classify): c[16 h + k], byte lane q = k-th column (scan order) of column
// chunk 4 h + q.  One c word at a time that is a gather from four residue words = three PRMTs.  But columns a whole number of
// words apart give c words that are the SAME byte position of the same four residue words (A, B, C, D) -- the rows of a 4 x 4 byte
// transpose, which takes two PRMT stages: (A, B) and (C, D) interleaved in their low / high halves, then the pairs combined; 8 PRMTs
// for 4 words instead of 12, 7 instead of 9 for three of them (P6: 68 instead of 96 per block).  Words that do not fit keep the gather.
// `transposes` = false: gathers only (the per-module winner passes of configs that mix scan families run at their register limit: measured
// 1-2 % slower on E5 with the transposes' temporaries).
void emit_pm_canonical(const std::vector<int>& cols, const char* arr, Lines& out, bool transposes = true) {
  const int nh = (NCH() + 3) / 4;
  std::vector<std::vector<int>> srcs((size_t)16 * nh, std::vector<int>(4, -1));
  for (int h = 0; h < nh; h++)
    for (int k = 0; k < 16; k++)
      for (int q = 0; q < 4; q++) srcs[16 * h + k][q] = (4 * h + q) < NCH() ? cols[16 * (4 * h + q) + k] : -1;  // lines of 32 bytes: two chunks, two zero lanes
  const char* env = getenv("MPC_SPEC_TRANSPOSE4");
  const bool on = transposes && !(env && env[0] == '0');
  std::map<std::vector<int>, std::map<int, int>> groups;  // (A, B, C, D) -> byte position -> c index
  if (on)
    for (int i = 0; i < 16 * nh; i++) {
      const std::vector<int>& s4 = srcs[i];
      if (s4[0] < 0 || s4[1] < 0 || s4[2] < 0 || s4[3] < 0) continue;
      const int b = s4[0] % 4;
      if (s4[1] % 4 != b || s4[2] % 4 != b || s4[3] % 4 != b) continue;
      std::vector<int> words = {s4[0] / 4, s4[1] / 4, s4[2] / 4, s4[3] / 4};
      if (std::set<int>(words.begin(), words.end()).size() != 4) continue;
      if (groups[words].count(b)) continue;  // a duplicate column: the gather below serves it
      groups[words][b] = i;
    }
  std::set<int> done;
  int gi = 0;
  for (auto& g : groups) {
    const std::map<int, int>& by_b = g.second;
    const bool lo = by_b.count(0) || by_b.count(1), hi = by_b.count(2) || by_b.count(3);
    if ((int)by_b.size() + 2 * ((lo ? 1 : 0) + (hi ? 1 : 0)) >= 3 * (int)by_b.size()) continue;  // no cheaper than the gathers
    const std::vector<int>& w = g.first;
    if (lo) out.push_back(fmt("  const uint32_t tl%d = prmt(%s[%d], %s[%d], 0x5140u), ul%d = prmt(%s[%d], %s[%d], 0x5140u);", gi, arr, w[0], arr, w[1], gi, arr, w[2], arr, w[3]));
    if (hi) out.push_back(fmt("  const uint32_t th%d = prmt(%s[%d], %s[%d], 0x7362u), uh%d = prmt(%s[%d], %s[%d], 0x7362u);", gi, arr, w[0], arr, w[1], gi, arr, w[2], arr, w[3]));
    for (auto& bi : by_b) {
      const int b = bi.first;
      out.push_back(fmt("  c[%d] = prmt(t%c%d, u%c%d, 0x%04xu);", bi.second, b < 2 ? 'l' : 'h', gi, b < 2 ? 'l' : 'h', gi, (b & 1) ? 0x7632u : 0x5410u));
      done.insert(bi.second);
    }
    gi++;
  }
  for (int i = 0; i < 16 * nh; i++) {
    if (done.count(i)) continue;
    const int s4[4] = {srcs[i][0], srcs[i][1], srcs[i][2], srcs[i][3]};
    out.push_back(fmt("  c[%d] = %s;", i, gather_expr(arr, s4).c_str()));
  }
}

int count_prmts(const std::string& e);

// ---- module model --------------------------------------------------------------------------------------------------
struct Module {
  int idx = 0;
  int predictor = 0, root = 0;
  bool cxor = false;
  int xsrc[kMaxL], psrc[kMaxL], pval[kMaxL];
  enum Op { kNone, kAdd, kShift } op = kNone;
  int root_pred = 0;
  enum Family { kCm, kPm, kBg } family = kCm;
  std::vector<int> cols;  // scan order (cm: one entry per column group; pm: the 128 columns of a plane group)
  std::vector<int> bits;  // bg (any other table): scan position i reads bit bits[i] of the XOR-ed residue line seen as W little-endian words
  int rho[8] = {0, 1, 2, 3, 4, 5, 6, 7};  // pm: plane scanned by plane group p

  bool rho_identity() const {
    for (int p = 0; p < 8; p++)
      if (rho[p] != p) return false;
    return true;
  }

  bool init(int index, const mpc_module_pod& m, std::string* why) {
    idx = index;
    predictor = m.predictor;
    root = m.root;
    cxor = m.consecutive_xor != 0;
    if (predictor == MPC_PRED_CONSEC && root != 0) { *why = "Consecutive predictor with root != 0"; return false; }
    int tperm[kMaxL], t = 0;
    for (int plane = 3; plane >= 0; plane--)
      for (int i = plane; i < L; i += 4) tperm[t++] = i;
    // source tables in residue order (same construction as build_generic_tables, mpc_generic.cu)
    for (int j = 0; j < L; j++) {
      const int i = (j == 0) ? root : (j <= root ? j - 1 : j);
      xsrc[j] = i;
      int ps = root, pv = 0;
      switch (predictor) {
        case MPC_PRED_ONE: ps = root; break;
        case MPC_PRED_CONSEC: ps = tperm[i > 0 ? i - 1 : 0]; break;
        case MPC_PRED_DIFF: ps = m.base[i]; pv = m.diff[i]; break;
        case MPC_PRED_WEIGHT: ps = m.base[i]; pv = (i != root) ? (int)m.shift[i] : 0; break;
        default: *why = "unknown predictor"; return false;
      }
      if (j == 0) { ps = root; pv = 0; }
      if (ps < 0 || ps >= L) { *why = "base index outside the line"; return false; }
      psrc[j] = ps;
      pval[j] = pv;
    }
    op = predictor == MPC_PRED_DIFF ? kAdd : (predictor == MPC_PRED_WEIGHT ? kShift : kNone);
    root_pred = (predictor == MPC_PRED_CONSEC) ? tperm[root] : root;
    // ---- scan family ----
    const int T = m.table_size;
    const uint8_t* rows = m.scan_row;
    const uint8_t* cs = m.scan_col;
    bool cm = (T % 8 == 0);
    for (int i = 0; i < T && cm; i++) cm = (rows[i] == i % 8) && (cs[i] == cs[i - i % 8]);
    if (cm) {
      family = kCm;
      for (int i = 0; i < T; i += 8) cols.push_back(cs[i]);
      return true;
    }
    bool pm = (T == 8 * L);
    bool seen[8] = {false};
    for (int g = 0; g < 8 && pm; g++) {
      const int r = rows[g * L];
      for (int i = 0; i < L && pm; i++) pm = (rows[g * L + i] == r) && (cs[g * L + i] == cs[i]);
      if (pm) { pm = !seen[r]; seen[r] = true; rho[g] = r; }
    }
    if (pm) {
      family = kPm;
      cols.assign(cs, cs + L);
      return true;
    }
    // Any other table (ScanModule.cpp:6-22 takes arbitrary (row, column) pairs, duplicates and short tables included): the
    // "bit-gather" family -- every scan row is assembled from its 16 source bits with shift-and-mask pairs, bits that share a
    // source word and a shift distance in one pair.  MPC_SPEC_BITGATHER=0 leaves such configs to the generic kernel.
    {
      const char* e = getenv("MPC_SPEC_BITGATHER");
      if (e && e[0] == '0') {
        *why = fmt("module %d: scan table is neither column-major nor plane-major", idx);
        return false;
      }
    }
    if (T < 0 || T > 8 * L) { *why = fmt("module %d: scan table larger than the line", idx); return false; }
    family = kBg;
    for (int i = 0; i < T; i++) {
      if (rows[i] > 7 || cs[i] >= L) { *why = fmt("module %d: scan entry outside the line", idx); return false; }
      bits.push_back(32 * (cs[i] / 4) + 8 * (cs[i] % 4) + 7 - rows[i]);
    }
    return true;
  }

  // ---- expressions ----
  std::string pred_expr(int w) const {
    const int srcs[4] = {psrc[4 * w], psrc[4 * w + 1], psrc[4 * w + 2], psrc[4 * w + 3]};
    const std::string base = gather_expr("x", srcs);
    const int* vals = &pval[4 * w];
    if (op == kAdd) {
      uint32_t d = 0;
      for (int q = 0; q < 4; q++) d |= (uint32_t)(vals[q] & 0xFF) << (8 * q);
      return d == 0 ? base : fmt("add_u8x4(%s, 0x%08xu)", base.c_str(), d);
    }
    if (op == kShift) {
      std::map<int, std::vector<int>> groups;
      for (int q = 0; q < 4; q++) groups[vals[q]].push_back(q);
      if (groups.size() == 1 && groups.begin()->first == 0) return base;
      std::vector<std::string> terms;
      for (auto& kv : groups) {
        const int s = kv.first;
        if (std::abs(s) >= 8) continue;
        uint32_t m = 0;
        if (s >= 0) {
          for (int q : kv.second) m |= (uint32_t)((0xFF << s) & 0xFF) << (8 * q);
          terms.push_back(s ? fmt("((p << %d) & 0x%08xu)", s, m) : fmt("(p & 0x%08xu)", m));
        } else {
          for (int q : kv.second) m |= (uint32_t)(0xFF >> -s) << (8 * q);
          terms.push_back(fmt("(mpcdev::shr_fma(p, %d) & 0x%08xu)", -s, m));  // the shift on the FMA pipe (IMAD.HI): the kernels are ALU-pipe bound
        }
      }
      if (terms.empty()) return "0u";
      return fmt("shiftmix(%s, [](uint32_t p) { return %s; })", base.c_str(), join(terms, " | ").c_str());
    }
    return base;
  }

  // Residue word w of a shifting (Weight) predictor in 16-bit lanes (mpc_device.cuh: lanes_even / lanes_odd / lanes_merge): the line
  // word is split into its even and odd bytes and every predicted byte -- byte sb of line word sw, shifted by s -- is subtracted in
  // place by one multiply-add on the FMA pipe whose multiplier -2^k is the shift and the byte position at once.  How the value gets
  // ready for that, cheapest first:
  //   (A) top byte of its word, right shift: x[sw] >> (24 - s) is exact (IMAD.HI), no ALU-pipe instruction;
  //   (B) target byte 2 or 3, left shift: x[sw] >> 8 sb has the byte at the bottom and only junk ABOVE it, which the multiply moves
  //       past the target byte (byte 3 of the even half is discarded, anything beyond bit 31 is gone), no ALU-pipe instruction;
  //   (C) otherwise x[sw] & mask keeps the bits of the byte that survive the shift (one LOP3); terms of one half that share the
  //       source word and the net shift 8 (q - sb) + s share the mask and the multiply-add; a net RIGHT shift is an IMAD.HI first.
  // P6's WeightBasePredictor (previous byte, shifts 0 / -1 / +1 / -1): 4 ALU-pipe + 5 FMA-pipe instructions per word against
  // 8 + 3 for gather, shift masks and sub_u8x4 -- the plane-major kernels are ALU-pipe bound.  Empty when the form does not apply.
  std::string lanes_residue_expr(int w) const {
    // Shifting predictors only.  Plain byte gathers were tried in this form as well (ConsecutiveBasePredictor's byte-plane order: four
    // source words per predicted word, most of them top bytes, 3 ALU + 8 FMA instead of 7 + 1 per word): fewer ALU-pipe instructions but
    // more instructions in all, and P6 lost 3-8 % on every data class (profiles/r02_p6_lanes.txt) -- the trade pays only while the
    // instruction count does not grow.
    if (op != kShift) return std::string();
    const char* env = getenv("MPC_SPEC_WLANES");
    if (env && env[0] == '0') return std::string();
    for (int q = 0; q < 4; q++)
      if (xsrc[4 * w + q] != 4 * w + q) return std::string();
    struct Masked { int sw, e; uint32_t mask; };
    std::string half[2] = {fmt("mpcdev::lanes_even(x[%d])", w), fmt("mpcdev::lanes_odd(x[%d])", w)};
    std::vector<Masked> masked[2];
    for (int q = 0; q < 4; q++) {
      if (w == 0 && q == 0) continue;  // the root byte: the caller puts line[root] there (ResidueModule.cpp:26-27)
      const int ps = psrc[4 * w + q], s = pval[4 * w + q], sw = ps / 4, sb = ps % 4, h = q & 1;
      if (std::abs(s) >= 8) continue;  // predicted byte 0
      if (sb == 3 && s <= 0) {         // (A)
        half[h] = fmt("mpcdev::mad_fma(mpcdev::shr_fma(x[%d], %d), 0x%08xu, %s)", sw, 24 - s, 0u - (1u << (8 * q)), half[h].c_str());
      } else if (q >= 2 && s >= 0) {   // (B)
        const std::string v = sb == 0 ? fmt("x[%d]", sw) : fmt("mpcdev::shr_fma(x[%d], %d)", sw, 8 * sb);
        half[h] = fmt("mpcdev::mad_fma(%s, 0x%08xu, %s)", v.c_str(), 0u - (1u << (8 * q + s)), half[h].c_str());
      } else {                         // (C)
        const int e = 8 * (q - sb) + s;
        const uint32_t keep = s < 0 ? (uint32_t)((0xff << -s) & 0xff) : (uint32_t)(0xff >> s);
        bool merged = false;
        for (auto& t : masked[h])
          if (t.sw == sw && t.e == e) { t.mask |= keep << (8 * sb); merged = true; }
        if (!merged) masked[h].push_back({sw, e, keep << (8 * sb)});
      }
    }
    for (int h = 0; h < 2; h++)
      for (auto& t : masked[h]) {
        if (t.e >= 0)
          half[h] = fmt("mpcdev::mad_fma((x[%d] & 0x%08xu), 0x%08xu, %s)", t.sw, t.mask, 0u - (1u << t.e), half[h].c_str());
        else
          half[h] = fmt("mpcdev::mad_fma(mpcdev::shr_fma((x[%d] & 0x%08xu), %d), 0xffffffffu, %s)", t.sw, t.mask, -t.e, half[h].c_str());
      }
    // ALU-pipe instructions per word: two splits, one merge and the masks here; the gather's PRMTs, one LOP3 per shift distance and
    // the four of sub_u8x4 in the plain form
    const int alu_lanes = 3 + (int)(masked[0].size() + masked[1].size());
    const int psrcs[4] = {psrc[4 * w], psrc[4 * w + 1], psrc[4 * w + 2], psrc[4 * w + 3]};
    int alu_plain = count_prmts(gather_expr("x", psrcs)) + 4;
    std::set<int> dist;
    for (int q = 0; q < 4; q++) dist.insert(pval[4 * w + q]);
    if (!(dist.size() == 1 && *dist.begin() == 0)) alu_plain += (int)dist.size();
    if (alu_lanes >= alu_plain) return std::string();
    return fmt("mpcdev::lanes_merge(%s, %s)", half[0].c_str(), half[1].c_str());
  }

  // Predicted word w as ONE byte gather over at most two line words, no arithmetic: *wa, *wb = source words, bytes[q] = (source
  // word, byte) of lane q.  False for predictors that add or shift, or gather from three or four words.
  bool pred_sel_form(int w, std::vector<int>* words, int (&src)[4]) const {
    for (int q = 0; q < 4; q++) {
      if (op == kShift && pval[4 * w + q] != 0) return false;
      if (op == kAdd && (pval[4 * w + q] & 0xff) != 0) return false;
      src[q] = psrc[4 * w + q];
    }
    words->clear();
    for (int q = 0; q < 4; q++)
      if (std::find(words->begin(), words->end(), src[q] / 4) == words->end()) words->push_back(src[q] / 4);
    return words->size() <= 2;
  }

  std::string root_byte() const {
    const int rw = root / 4, rb = root % 4;
    return rb == 0 ? fmt("(x[%d] & 0xffu)", rw) : fmt("((x[%d] >> %d) & 0xffu)", rw, 8 * rb);
  }

  static bool plain_word(const std::string& e, int* index) {
    if (e.size() < 4 || e.compare(0, 2, "x[") != 0 || e.back() != ']') return false;
    for (size_t i = 2; i + 1 < e.size(); i++)
      if (e[i] < '0' || e[i] > '9') return false;
    *index = atoi(e.c_str() + 2);
    return true;
  }

  // statement defining `const uint32_t <name>` = residue word w (before the XOR stage)
  std::string residue_stmts(int w, const std::string& name, bool shared_low = false, const std::string& pred_override = std::string()) const {
    const int srcs[4] = {xsrc[4 * w], xsrc[4 * w + 1], xsrc[4 * w + 2], xsrc[4 * w + 3]};
    const std::string xe = gather_expr("x", srcs), pe = pred_override.empty() ? pred_expr(w) : pred_override;
    std::string e;
    int a = 0, b = 0;
    const std::string lanes = pred_override.empty() ? lanes_residue_expr(w) : std::string();
    if (!lanes.empty())
      e = lanes;
    else if (shared_low && plain_word(xe, &a) && plain_word(pe, &b))
      e = fmt("sub_u8x4_shared(x[%d], x[%d], ah[%d], ah[%d])", a, b, a, b);
    else
      e = fmt("sub_u8x4(%s, %s)", xe.c_str(), pe.c_str());
    if (w == 0) e = fmt("((%s & 0xffffff00u) | %s)", e.c_str(), root_byte().c_str());
    return fmt("const uint32_t %s = %s;", name.c_str(), e.c_str());
  }

  // line word ^ predicted word for residue word w; a zero byte <=> that residue byte is zero
  std::string eq_expr(int w) const {
    const int srcs[4] = {xsrc[4 * w], xsrc[4 * w + 1], xsrc[4 * w + 2], xsrc[4 * w + 3]};
    std::string e = fmt("(%s ^ %s)", gather_expr("x", srcs).c_str(), pred_expr(w).c_str());
    if (w == 0) e = fmt("((%s & 0xffffff00u) | %s)", e.c_str(), root_byte().c_str());  // ResidueModule.cpp:26-27
    return e;
  }

  std::string g_from_r(int w, const std::string& r) const {
    if (cxor) return fmt("xc(%s, 0x%08xu)", r.c_str(), w == 0 ? 0x7f7f7f00u : 0x7f7f7f7fu);
    return fmt("xf(%s, 0x%08xu)", r.c_str(), w == 0 ? 0x7f7f7f00u : 0x7f7f7f7fu);
  }
};

// ---- row layout of column-major modules -----------------------------------------------------------------------------
// The encoder reads 32 words of two 16-bit scan rows each.  Legacy layout: word j = rows 2j, 2j+1 -- one byte from each
// of up to four residue words, i.e. three PRMTs per word.  Paired layout: word j = rows (ra[j], rb[j]) chosen so that
// the four bytes come from at most two residue words (ONE PRMT); the non-zero flags of the rows are then accumulated per
// "run" (words whose ra and rb both advance by one) and shifted into place in the 64-bit zero-row mask at the end.
struct RowLayout {
  bool paired = false;
  int ra[kMaxL / 4], rb[kMaxL / 4];      // rows in the low / high halfword of word j
  int grp[kMaxL / 4], slot[kMaxL / 4];   // flag accumulator of word j and its bit position in it
  struct Group { int A, B, len; };
  std::vector<Group> groups;
};

int count_prmts(const std::string& e) {
  int n = 0;
  for (size_t pos = e.find("prmt("); pos != std::string::npos; pos = e.find("prmt(", pos + 1)) n++;
  return n;
}

// srcs of encoder word holding rows (a, b) of a module with scan columns `cols` (resized to L, -1 = no column)
void row_pair_srcs(const std::vector<int>& cols, int a, int b, int (&srcs)[4]) {
  srcs[0] = cols[2 * a + 1]; srcs[1] = cols[2 * a]; srcs[2] = cols[2 * b + 1]; srcs[3] = cols[2 * b];
}

RowLayout plan_row_layout(const std::vector<std::vector<int>>& cm_cols) {
  RowLayout lay;
  for (int j = 0; j < W; j++) { lay.ra[j] = 2 * j; lay.rb[j] = 2 * j + 1; lay.grp[j] = 0; lay.slot[j] = 0; }
  if (cm_cols.empty()) return lay;
  const std::vector<int>& cols = cm_cols[0];
  auto words_of = [&](int r) {
    std::set<int> w;
    for (int c : {cols[2 * r], cols[2 * r + 1]})
      if (c >= 0) w.insert(c / 4);
    return w;
  };
  auto fits = [&](int a, int b) {
    std::set<int> u = words_of(a), v = words_of(b);
    u.insert(v.begin(), v.end());
    return u.size() <= 2;
  };
  // rows that read the same pair of residue words are paired first (in row order), then any two rows that fit one PRMT
  std::map<std::set<int>, std::vector<int>> buckets;
  for (int r = 0; r < R(); r++) buckets[words_of(r)].push_back(r);
  std::vector<std::pair<int, int>> pairs;
  std::vector<int> left;
  for (auto& kv : buckets) {
    auto& v = kv.second;
    size_t i = 0;
    if (kv.first.size() == 2)
      for (; i + 1 < v.size(); i += 2) pairs.push_back({v[i], v[i + 1]});
    for (; i < v.size(); i++) left.push_back(v[i]);
  }
  std::sort(left.begin(), left.end());
  std::vector<bool> used(left.size(), false);
  for (size_t i = 0; i < left.size(); i++) {
    if (used[i]) continue;
    for (size_t k = i + 1; k < left.size(); k++)
      if (!used[k] && fits(left[i], left[k])) { pairs.push_back({left[i], left[k]}); used[i] = used[k] = true; break; }
  }
  std::vector<int> rest;
  for (size_t i = 0; i < left.size(); i++)
    if (!used[i]) rest.push_back(left[i]);
  for (size_t i = 0; i + 1 < rest.size(); i += 2) pairs.push_back({rest[i], rest[i + 1]});
  if (pairs.size() != (size_t)W) return lay;
  std::sort(pairs.begin(), pairs.end());
  // PRMT count of both layouts over every column-major module of the config
  int cost_legacy = 0, cost_paired = 0;
  for (auto& mc : cm_cols)
    for (int j = 0; j < W; j++) {
      int s0[4], s1[4];
      row_pair_srcs(mc, 2 * j, 2 * j + 1, s0);
      row_pair_srcs(mc, pairs[j].first, pairs[j].second, s1);
      cost_legacy += count_prmts(gather_expr("g", s0));
      cost_paired += count_prmts(gather_expr("g", s1));
    }
  // runs: consecutive words whose two rows both advance by one share a flag accumulator (16 flags per halfword)
  RowLayout p;
  p.paired = true;
  for (int j = 0; j < W; j++) {
    p.ra[j] = pairs[j].first;
    p.rb[j] = pairs[j].second;
    if (j > 0 && p.ra[j] == p.ra[j - 1] + 1 && p.rb[j] == p.rb[j - 1] + 1 && p.groups.back().len < 16) {
      p.groups.back().len++;
    } else {
      p.groups.push_back({p.ra[j], p.rb[j], 1});
    }
    p.grp[j] = (int)p.groups.size() - 1;
    p.slot[j] = p.groups.back().len - 1;
  }
  const char* e = getenv("MPC_SPEC_PAIRED");
  if (e && e[0] == '0') return lay;
  if (p.groups.size() > 8 || cost_paired + 4 * (int)p.groups.size() >= cost_legacy) return lay;
  return p;
}

// encode_rows: the row classifier for the paired layout (row-cost table in shared memory), emitted per config
void emit_encode_rows(const RowLayout& lay, Lines& out);
std::string bg_row_pair_expr(const Module& m, int ra, int rb);

// full_<m>: all 32 residue words -> residue sums, canonical row layout c[32].
// lut_xor != 0 (column-major modules only): the row-cost table is indexed with the residue bytes BEFORE the XOR stage --
// the stage is a per-byte bijection (Gray code / conditional complement), so it is folded into the table and disappears
// from the kernel; only the root byte, which the XOR stage skips (XORModule.cpp:12), is run through the inverse map so
// that the table maps it back to itself.
// ---- selector groups: ONE winner pass for several modules in warps whose lanes picked different winners -------------------
// Modules whose predicted words are plain byte gathers over the SAME pair of line words differ only in the PRMT selector
// (F4: "previous byte" = prmt(x[w-1], x[w], 0x6543) and "previous word" = x[w-1] = prmt(x[w-1], x[w], 0x3210)).  With the
// selector in a register that every lane sets from its own winner, one pass serves all lanes of the group: a warp whose
// lanes picked modules of one group runs one residue pass instead of one per module.  (Warps whose lanes agree keep the
// specialised per-module pass: its shared-operand subtract is one instruction cheaper per word.)
struct SelGroup {
  std::vector<int> members;                       // indices into the module list
  int wa[kMaxL / 4], wb[kMaxL / 4];               // operand words of word w
  int cls[kMaxL / 4];                             // selector class of word w
  std::vector<std::vector<unsigned>> cls_sel;     // per class: selector of every member
};

bool plan_sel_group(const std::vector<Module>& mods, const std::vector<int>& members, SelGroup* g) {
  g->members = members;
  g->cls_sel.clear();
  const Module& m0 = mods[(size_t)members[0]];
  for (int mi : members) {
    const Module& m = mods[(size_t)mi];
    if (m.family != Module::kCm || m.predictor == MPC_PRED_CONSEC || m.root != m0.root || m.cols != m0.cols || m.cxor != m0.cxor) return false;
    for (int j = 0; j < L; j++)
      if (m.xsrc[j] != m0.xsrc[j]) return false;
  }
  for (int w = 0; w < W; w++) {
    std::vector<int> uni;
    std::vector<std::vector<int>> srcs;
    for (int mi : members) {
      std::vector<int> ws;
      int src[4];
      if (!mods[(size_t)mi].pred_sel_form(w, &ws, src)) return false;
      for (int x : ws)
        if (std::find(uni.begin(), uni.end(), x) == uni.end()) uni.push_back(x);
      srcs.push_back(std::vector<int>(src, src + 4));
    }
    if (uni.size() > 2) return false;
    g->wa[w] = uni[0];
    g->wb[w] = uni.size() == 2 ? uni[1] : uni[0];
    std::vector<unsigned> sels;
    for (auto& sv : srcs) {
      unsigned sel = 0;
      for (int q = 0; q < 4; q++) sel |= ((unsigned)(sv[(size_t)q] % 4) + (sv[(size_t)q] / 4 == g->wa[w] ? 0u : 4u)) << (4 * q);
      sels.push_back(sel);
    }
    auto it = std::find(g->cls_sel.begin(), g->cls_sel.end(), sels);
    if (it == g->cls_sel.end()) { g->cls_sel.push_back(sels); it = g->cls_sel.end() - 1; }
    g->cls[w] = (int)(it - g->cls_sel.begin());
  }
  return g->cls_sel.size() <= 4;  // selector registers per lane
}

void emit_full(const Module& m, Lines& out, int lut_xor, const RowLayout& lay, const SelGroup* grp = nullptr, int grp_id = 0) {
  if (grp) {
    std::string params;
    for (size_t c = 0; c < grp->cls_sel.size(); c++) params += fmt(", uint32_t s%d", (int)c);
    out.push_back(fmt("__device__ __forceinline__ void full_g%d(const uint32_t (&x)[32], uint32_t (&c)[32], uint32_t& sa, uint32_t& sq%s) {", grp_id, params.c_str()));
  } else
  out.push_back(fmt("__device__ __forceinline__ void full_%d(const uint32_t (&x)[32], uint32_t (&c)[32], uint32_t& sa, uint32_t& sq) {", m.idx));
  out.push_back("  uint32_t g[32];");
  out.push_back("  uint32_t sa0 = 0, sa1 = 0, sa2 = 0, sa3 = 0, sq0 = 0, sq1 = 0, sq2 = 0, sq3 = 0;  // four short chains instead of one long one");
  out.push_back("  uint32_t ah[32];");
  out.push_back("#pragma unroll");
  out.push_back(fmt("  for (int i = 0; i < %d; i++) ah[i] = x[i] | 0x80808080u;  // only the words a plain-copy predictor uses survive", W));
  for (int w = 0; w < W; w++) {
    if (grp) out.push_back("  { " + m.residue_stmts(w, "r", false, fmt("prmt(x[%d], x[%d], s%d)", grp->wa[w], grp->wb[w], grp->cls[w])));
    else
    out.push_back("  { " + m.residue_stmts(w, "r", true));
    if (w == 0) {
      // MAE/MSE run over all line positions (ResidueModule.cpp:43-73): the residue line holds the root byte itself at
      // position 0, the statistics hold line[root] - predicted[root] instead
      const int rw = m.root / 4, rb = m.root % 4, pw = m.root_pred / 4, pb = m.root_pred % 4;
      if (m.root_pred == m.root)
        out.push_back("    const uint32_t rs = r & 0xffffff00u;");
      else
        out.push_back(fmt("    const uint32_t rs = (r & 0xffffff00u) | (((x[%d] >> %d) - (x[%d] >> %d)) & 0xffu);", rw, 8 * rb, pw, 8 * pb));
      out.push_back("    sa0 = mpcdev::sum_u8x4_acc(rs, sa0); sq0 = __dp4a(rs, rs, sq0);");
    } else {
      out.push_back(fmt("    sa%d = mpcdev::sum_u8x4_acc(r, sa%d); sq%d = __dp4a(r, r, sq%d);", w % 4, w % 4, w % 4, w % 4));
    }
    if (lut_xor && m.family == Module::kCm) {
      if (w == 0)
        out.push_back(fmt("    g[0] = (r & 0xffffff00u) | %s(r & 0xffu); }", lut_xor == 1 ? "inv_gray8" : "inv_first8"));
      else
        out.push_back(fmt("    g[%d] = r; }", w));
    } else {
      out.push_back(fmt("    g[%d] = %s; }", w, m.g_from_r(w, "r").c_str()));
    }
  }
  out.push_back("  sa = (sa0 + sa1) + (sa2 + sa3); sq = (sq0 + sq1) + (sq2 + sq3);");
  if (m.family == Module::kCm) {
    std::vector<int> cols = m.cols;
    cols.resize(L, -1);
    for (int j = 0; j < W; j++) {
      int srcs[4];
      row_pair_srcs(cols, lay.ra[j], lay.rb[j], srcs);  // legacy layout: rows 2j, 2j+1
      out.push_back(fmt("  c[%d] = %s;", j, gather_expr("g", srcs).c_str()));
    }
  } else if (m.family == Module::kBg) {
    for (int j = 0; j < W; j++) out.push_back(fmt("  c[%d] = %s;", j, bg_row_pair_expr(m, lay.ra[j], lay.rb[j]).c_str()));
  } else {
    emit_pm_canonical(m.cols, "g", out, false);
  }
  out.push_back("}");
}

void emit_encode_rows(const RowLayout& lay, Lines& out) {
  out.push_back("// Row classifier for the paired row layout: word j holds scan rows ra (low halfword) and rb (high halfword); each row");
  out.push_back("// costs one table load; the non-zero flags of a run of words land in one accumulator and are shifted into the 64-bit");
  out.push_back("// row mask at the end (zero runs: FPCModule.cpp:27-45, 69-79).");
  out.push_back("__device__ __forceinline__ uint32_t encode_rows(const uint32_t (&c)[32], const uint8_t* __restrict__ lut) {");
  out.push_back("  uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;");
  std::vector<std::string> ns;
  for (size_t g = 0; g < lay.groups.size(); g++) ns.push_back(fmt("n%d = 0", (int)g));
  out.push_back("  uint32_t " + join(ns, ", ") + ";");
  for (int j = 0; j < W; j++)
    out.push_back(fmt("  row_pair_step(c[%d], lut, a%d, a%d, n%d, 0x%xu);  // rows %d, %d", j, (j & 1) * 2, (j & 1) * 2 + 1, lay.grp[j],
                      1u << lay.slot[j], lay.ra[j], lay.rb[j]));
  out.push_back("  uint64_t nzm = 0;");
  for (size_t g = 0; g < lay.groups.size(); g++)
    out.push_back(fmt("  nzm |= ((uint64_t)(n%d & 0xffffu) << %d) | ((uint64_t)(n%d >> 16) << %d);", (int)g, lay.groups[g].A, (int)g, lay.groups[g].B));
  if (R() < 64) out.push_back(fmt("  return (a0 + a1) + (a2 + a3) + mpcdev::zero_run_cost(~nzm & 0x%llxull);  // %d rows", (1ull << R()) - 1ull, R()));
  else out.push_back("  return (a0 + a1) + (a2 + a3) + mpcdev::zero_run_cost(~nzm);");
  out.push_back("}");
  out.push_back("");
}

// Leading zero rows of a column-major module.  A scan row is two bytes of the XOR-ed residue line, and such a byte is
// zero exactly when the line byte equals its prediction (both XOR variants map 0 -> 0 only), so the score needs no
// subtraction: e<w> = line word ^ predicted word, tested under the byte masks of the row.
void emit_score_cm(const Module& m, Lines& out) {
  out.push_back(fmt("__device__ __forceinline__ uint32_t score_%d(const uint32_t (&x)[32]) {", m.idx));
  std::set<int> have;
  std::vector<int> cols = m.cols;
  cols.resize(L, -1);
  const int nrows = R();
  std::vector<std::map<int, uint32_t>> tests(nrows);
  for (int k = 0; k < nrows; k++)
    for (int cidx : {cols[2 * k], cols[2 * k + 1]})
      if (cidx >= 0) tests[k][cidx / 4] |= 0xFFu << (8 * (cidx % 4));
  for (int k = 0; k < nrows; k += 2) {
    std::vector<int> group;
    for (int kk : {k, k + 1})
      if (kk < nrows && !tests[kk].empty()) group.push_back(kk);
    for (int kk : group)
      for (auto& wm : tests[kk])
        if (!have.count(wm.first)) {
          out.push_back(fmt("  const uint32_t e%d = %s;", wm.first, m.eq_expr(wm.first).c_str()));
          have.insert(wm.first);
        }
    std::vector<std::pair<int, std::string>> exprs;
    for (int kk : group) {
      std::vector<std::pair<uint32_t, std::vector<std::string>>> terms;  // grouped by mask, first-appearance order
      for (auto& wm : tests[kk]) {
        auto it = std::find_if(terms.begin(), terms.end(), [&](auto& t) { return t.first == wm.second; });
        if (it == terms.end()) { terms.push_back({wm.second, {}}); it = terms.end() - 1; }
        it->second.push_back(fmt("e%d", wm.first));
      }
      std::vector<std::string> parts;
      for (auto& t : terms)
        parts.push_back(t.first != 0xFFFFFFFFu ? fmt("((%s) & 0x%08xu)", join(t.second, " | ").c_str(), t.first)
                                                : fmt("(%s)", join(t.second, " | ").c_str()));
      exprs.push_back({kk, join(parts, " | ")});
    }
    if (exprs.size() == 2)
      out.push_back(fmt("  { const uint32_t t0 = %s, t1 = %s; if ((t0 | t1) != 0u) return t0 ? %du : %du; }", exprs[0].second.c_str(),
                        exprs[1].second.c_str(), exprs[0].first, exprs[1].first));
    else if (exprs.size() == 1)
      out.push_back(fmt("  if ((%s) != 0u) return %du;", exprs[0].second.c_str(), exprs[0].first));
  }
  out.push_back(fmt("  return %du;", nrows));
  out.push_back("}");
}

// ---- bit-gather family: arbitrary scan tables ------------------------------------------------------------------------------
// source bits of scan row j grouped by source word: word -> mask (for the zero test of the scoring pass)
std::map<int, uint32_t> bg_row_masks(const Module& m, int j) {
  std::map<int, uint32_t> by_word;
  for (int q = 0; q < 16; q++) {
    const size_t i = (size_t)16 * j + q;
    if (i < m.bits.size()) by_word[m.bits[i] / 32] |= 1u << (m.bits[i] % 32);
  }
  return by_word;
}

// expression for the encoder word holding rows ra (low halfword) and rb (high halfword), scan position 0 in bit 15 of a halfword
std::string bg_row_pair_expr(const Module& m, int ra, int rb) {
  std::map<std::pair<int, int>, uint32_t> groups;  // (source word, right-shift distance) -> mask of target bits
  for (int half = 0; half < 2; half++) {
    const int row = half ? rb : ra;
    for (int q = 0; q < 16; q++) {
      const size_t i = (size_t)16 * row + q;
      if (i >= m.bits.size()) continue;
      const int t = 16 * half + 15 - q, b = m.bits[i] % 32;
      groups[{m.bits[i] / 32, b - t}] |= 1u << t;
    }
  }
  if (groups.empty()) return "0u";
  std::vector<std::string> terms;
  for (auto& kv : groups) {
    const int w = kv.first.first, sh = kv.first.second;
    if (sh == 0) terms.push_back(fmt("(g[%d] & 0x%08xu)", w, kv.second));
    else if (sh > 0) terms.push_back(fmt("((g[%d] >> %d) & 0x%08xu)", w, sh, kv.second));
    else terms.push_back(fmt("((g[%d] << %d) & 0x%08xu)", w, -sh, kv.second));
  }
  return join(terms, " | ");
}

// Leading zero rows of a bit-gather module: the XOR-ed residue words a row reads are computed when first needed, a row is zero
// iff its source bits are (one AND-OR per source word), and the walk stops at the first non-zero row (VPC.cpp:378-387).
void emit_score_bg(const Module& m, Lines& out) {
  out.push_back(fmt("__device__ __forceinline__ uint32_t score_%d(const uint32_t (&x)[32]) {", m.idx));
  out.push_back("  uint32_t g[32];");
  std::set<int> have;
  for (int j = 0; j < R(); j++) {
    const auto masks = bg_row_masks(m, j);
    if (masks.empty()) continue;  // past the end of a short table: always zero
    std::vector<std::string> terms;
    for (auto& wm : masks) {
      if (!have.count(wm.first)) {
        out.push_back("  { " + m.residue_stmts(wm.first, "r") + fmt(" g[%d] = %s; }", wm.first, m.g_from_r(wm.first, "r").c_str()));
        have.insert(wm.first);
      }
      terms.push_back(wm.second == 0xFFFFFFFFu ? fmt("g[%d]", wm.first) : fmt("(g[%d] & 0x%08xu)", wm.first, wm.second));
    }
    out.push_back(fmt("  if ((%s) != 0u) return %du;", join(terms, " | ").c_str(), j));
  }
  out.push_back(fmt("  return %du;", R()));
  out.push_back("}");
}

// 8-bit value whose bit 7-p is bit 7-rho[p] of v (planes reordered into scan order)
std::string plane_order_expr(const std::string& v, const Module& m) {
  if (m.rho_identity()) return fmt("(%s & 0xffu)", v.c_str());
  std::vector<std::string> terms;
  for (int p = 0; p < 8; p++) {
    const int src = 7 - m.rho[p], dst = 7 - p;
    if (src >= dst) terms.push_back(fmt("(((%s) >> %d) & 0x%02xu)", v.c_str(), src - dst, 1 << dst));
    else terms.push_back(fmt("(((%s) << %d) & 0x%02xu)", v.c_str(), dst - src, 1 << dst));
  }
  return "(" + join(terms, " | ") + ")";
}

void emit_score_pm(const Module& m, Lines& out) {
  out.push_back(fmt("__device__ __forceinline__ uint32_t score_%d(const uint32_t (&x)[32]) {", m.idx));
  out.push_back("  uint32_t g[32];");
  const bool early = (m.rho[0] == 0);  // plane 0 (bit 7) is untouched by the XOR stage: test it on the residues alone
  const int nch = NCH();
  std::vector<std::map<int, uint32_t>> chunks(nch);
  for (int j = 0; j < nch; j++)
    for (int i = 0; i < 16; i++) {
      const int cidx = m.cols[16 * j + i];
      chunks[j][cidx / 4] |= 0xFFu << (8 * (cidx % 4));
    }
  auto chunk_or = [&](int j, const char* arr) {
    std::vector<std::string> terms;
    for (auto& wm : chunks[j])
      terms.push_back(wm.second != 0xFFFFFFFFu ? fmt("(%s[%d] & 0x%08xu)", arr, wm.first, wm.second) : fmt("%s[%d]", arr, wm.first));
    return join(terms, " | ");
  };
  if (early) {
    // Rows 0..7 are plane 0 of the eight column chunks, and plane 0 (bit 7) is untouched by the XOR stage: walk the
    // chunks in scan order, computing only the residue words a chunk needs, and stop at the first set bit -- a module
    // that loses on its first rows costs a handful of words instead of the whole line.
    std::set<int> have;
    for (int j = 0; j < nch; j++) {
      for (auto& wm : chunks[j])
        if (!have.count(wm.first)) {
          out.push_back("  { " + m.residue_stmts(wm.first, "r") + fmt(" g[%d] = r; }", wm.first));
          have.insert(wm.first);
        }
      out.push_back(fmt("  if ((%s) & 0x80808080u) return %du;", chunk_or(j, "g").c_str(), j));
    }
    for (int w = 0; w < W; w++)
      if (!have.count(w)) out.push_back("  { " + m.residue_stmts(w, "r") + fmt(" g[%d] = r; }", w));
  } else {
    for (int w = 0; w < W; w++) out.push_back("  { " + m.residue_stmts(w, "r") + fmt(" g[%d] = r; }", w));
  }
  if (!m.rho_identity()) {
    for (int w = 0; w < W; w++) out.push_back(fmt("  g[%d] = %s;", w, m.g_from_r(w, fmt("g[%d]", w)).c_str()));
  } else {
    out.push_back("  // planes are scanned MSB first: while every higher plane is zero, plane b of the XOR-ed residue equals");
    out.push_back("  // plane b of the residue itself (XORModule.cpp:9-20), so the leading-zero count needs no XOR stage");
  }
  out.push_back("  uint32_t f[8];");
  for (int j = 0; j < nch; j++)
    out.push_back(fmt("  { uint32_t o = %s; o |= o >> 16; o |= o >> 8; f[%d] = %s; }", chunk_or(j, "g").c_str(), j, plane_order_expr("o", m).c_str()));
  out.push_back(fmt("  return pm_leading_zero_rows<%d>(f);", nch));
  out.push_back("}");
}

// after transpose8x8 byte c holds plane 7-c; output byte p must hold plane rho[p]
std::pair<unsigned, unsigned> pm_selectors(const Module& m) {
  unsigned s0 = 0, s1 = 0;
  for (int p = 0; p < 4; p++) {
    s0 |= (unsigned)(7 - m.rho[p]) << (4 * p);
    s1 |= (unsigned)(7 - m.rho[4 + p]) << (4 * p);
  }
  return {s0, s1};
}

// ---- "pm2": configs whose modules are all plane-major with ONE scan family (same column order, planes MSB first) -------------
// The first generation of the plane-major path emitted, per module, a scoring function (all 32 residue words) and a winner
// pass (residues again + statistics + XOR stage + byte transposition): ~70 KiB of straight-line SASS that the instruction
// cache cannot hold (ncu: a quarter to almost half of the stall samples were instruction fetches), and lanes of a warp that
// picked different winners each ran a whole winner pass.  Here a module contributes ONE piece of code, its residue pass
// res_<m> (x -> r[32], before the XOR stage), used both for scoring (it then stops at the first column chunk whose plane 0
// is set) and for the winner; everything else -- leading-zero-row count, MAE/MSE sums, XOR stage, byte transposition,
// bit-sliced row classifier -- is emitted once and shared by all modules and all lanes.
std::vector<std::map<int, uint32_t>> pm_chunks(const Module& m) {
  std::vector<std::map<int, uint32_t>> chunks(NCH());
  for (int j = 0; j < NCH(); j++)
    for (int i = 0; i < 16; i++) {
      const int cidx = m.cols[16 * j + i];
      chunks[j][cidx / 4] |= 0xFFu << (8 * (cidx % 4));
    }
  return chunks;
}

std::string pm_chunk_or(const std::vector<std::map<int, uint32_t>>& chunks, int j, const char* arr) {
  // words under the same byte mask are OR-ed first, so that a chunk costs one LOP3 per two words
  std::vector<std::pair<uint32_t, std::vector<std::string>>> terms;
  for (auto& wm : chunks[j]) {
    auto it = std::find_if(terms.begin(), terms.end(), [&](auto& t) { return t.first == wm.second; });
    if (it == terms.end()) { terms.push_back({wm.second, {}}); it = terms.end() - 1; }
    it->second.push_back(fmt("%s[%d]", arr, wm.first));
  }
  std::vector<std::string> parts;
  for (auto& t : terms)
    parts.push_back(t.first != 0xFFFFFFFFu ? fmt("((%s) & 0x%08xu)", join(t.second, " | ").c_str(), t.first) : fmt("(%s)", join(t.second, " | ").c_str()));
  return join(parts, " | ");
}

void emit_res_pm2(const Module& m, Lines& out) {
  out.push_back(fmt("__device__ __forceinline__ uint32_t res_%d(const uint32_t (&x)[32], uint32_t (&r)[32], bool scoring) {", m.idx));
  const char* eah = getenv("MPC_SPEC_PM2_AH");
  const bool use_ah = !(eah && eah[0] == '0');
  if (use_ah) {
    out.push_back("  uint32_t ah[32];");
    out.push_back("#pragma unroll");
    out.push_back(fmt("  for (int i = 0; i < %d; i++) ah[i] = x[i] | 0x80808080u;  // only the words a plain-copy predictor uses survive", W));
  }
  const auto chunks = pm_chunks(m);
  std::set<int> have;
  for (int j = 0; j < NCH(); j++) {
    for (auto& wm : chunks[j])
      if (!have.count(wm.first)) {
        out.push_back("  { " + m.residue_stmts(wm.first, "rr", use_ah) + fmt(" r[%d] = rr; }", wm.first));
        have.insert(wm.first);
      }
    // rows 0..7 are plane 0 (bit 7, untouched by the XOR stage) of the column chunks in scan order
    out.push_back(fmt("  if (scoring && ((%s) & 0x80808080u)) return %du;", pm_chunk_or(chunks, j, "r").c_str(), j));
  }
  for (int w = 0; w < W; w++)
    if (!have.count(w)) out.push_back("  { " + m.residue_stmts(w, "rr", use_ah) + fmt(" r[%d] = rr; }", w));
  out.push_back(fmt("  return %du;  // plane 0 is zero in every chunk (or not scoring): all residue words are in r", NCH()));
  out.push_back("}");
}

void emit_pm2_shared(const std::vector<Module>& mods, Lines& out) {
  const Module& m0 = mods[0];
  const auto chunks = pm_chunks(m0);
  out.push_back("// leading zero rows from complete residues: planes are scanned MSB first, and while every higher plane is zero, plane b of");
  out.push_back("// the XOR-ed residue equals plane b of the residue itself (XORModule.cpp:9-20), so the count needs no XOR stage");
  out.push_back("__device__ __forceinline__ uint32_t pm2_lz(const uint32_t (&r)[32]) {");
  out.push_back("  uint32_t f[8];");
  for (int j = 0; j < NCH(); j++)
    out.push_back(fmt("  { uint32_t o = %s; o |= o >> 16; o |= o >> 8; f[%d] = o & 0xffu; }", pm_chunk_or(chunks, j, "r").c_str(), j));
  out.push_back(fmt("  return pm_leading_zero_rows<%d>(f);", NCH()));
  out.push_back("}");
  out.push_back("");
  const auto sel = pm_selectors(m0);
  out.push_back("// residue sums (VPC.cpp:417-443), XOR stage (XORModule.cpp:5-23), scan (ScanModule.cpp:6-22) as a byte transposition into the");
  out.push_back("// canonical layout of pm_classify, common encoder (FPCModule.cpp:19-158) -- one copy for every module");
  out.push_back("__device__ __forceinline__ uint32_t pm2_tail(int best, const uint32_t (&x)[32], uint32_t (&r)[32], uint32_t& sa, uint32_t& sq) {");
  // MAE/MSE run over all line positions: position `root` holds line[root] - predicted[root] there, the residue line holds the root byte
  out.push_back("  uint32_t rootfix = 0u;");
  for (auto& m : mods)
    if (m.root_pred != m.root)
      out.push_back(fmt("  if (best == %d) rootfix = ((x[%d] >> %d) - (x[%d] >> %d)) & 0xffu;", m.idx, m.root / 4, 8 * (m.root % 4), m.root_pred / 4, 8 * (m.root_pred % 4)));
  out.push_back("  uint32_t sa0 = 0, sa1 = 0, sa2 = 0, sa3 = 0, sq0 = 0, sq1 = 0, sq2 = 0, sq3 = 0;");
  out.push_back("  { const uint32_t rs = (r[0] & 0xffffff00u) | rootfix; sa0 = mpcdev::sum_u8x4_acc(rs, sa0); sq0 = __dp4a(rs, rs, sq0); }");
  out.push_back("#pragma unroll");
  out.push_back(fmt("  for (int w = 1; w < %d; w++) {", W));
  out.push_back("    if ((w & 3) == 0) { sa0 = mpcdev::sum_u8x4_acc(r[w], sa0); sq0 = __dp4a(r[w], r[w], sq0); }");
  out.push_back("    else if ((w & 3) == 1) { sa1 = mpcdev::sum_u8x4_acc(r[w], sa1); sq1 = __dp4a(r[w], r[w], sq1); }");
  out.push_back("    else if ((w & 3) == 2) { sa2 = mpcdev::sum_u8x4_acc(r[w], sa2); sq2 = __dp4a(r[w], r[w], sq2); }");
  out.push_back("    else { sa3 = mpcdev::sum_u8x4_acc(r[w], sa3); sq3 = __dp4a(r[w], r[w], sq3); }");
  out.push_back("  }");
  out.push_back("  sa = (sa0 + sa1) + (sa2 + sa3); sq = (sq0 + sq1) + (sq2 + sq3);");
  bool any_c = false, any_f = false;
  std::vector<std::string> cmods;
  for (auto& m : mods) { if (m.cxor) { any_c = true; cmods.push_back(fmt("best == %d", m.idx)); } else any_f = true; }
  const std::string xc_loop_s = fmt("    r[0] = xc(r[0], 0x7f7f7f00u);\n#pragma unroll\n    for (int w = 1; w < %d; w++) r[w] = xc(r[w], 0x7f7f7f7fu);", W);
  const std::string xf_loop_s = fmt("    r[0] = xf(r[0], 0x7f7f7f00u);\n#pragma unroll\n    for (int w = 1; w < %d; w++) r[w] = xf(r[w], 0x7f7f7f7fu);", W);
  const char* xc_loop = xc_loop_s.c_str();
  const char* xf_loop = xf_loop_s.c_str();
  if (any_c && any_f) {
    out.push_back("  if (" + join(cmods, " || ") + ") {  // consecutive XOR");
    out.push_back(xc_loop);
    out.push_back("  } else {  // first-plane XOR");
    out.push_back(xf_loop);
    out.push_back("  }");
  } else {
    out.push_back("  {");
    out.push_back(any_c ? xc_loop : xf_loop);
    out.push_back("  }");
  }
  out.push_back("  uint32_t c[32];");
  emit_pm_canonical(m0.cols, "r", out);
  out.push_back(fmt("  return encode_pm<%d, 0x%04xu, 0x%04xu>(c);", NCH(), sel.first, sel.second));
  out.push_back("}");
  out.push_back("");
}

// Cfg::select_encode for pm2 configs: the per-thread state machine (module k, scoring or winner pass) around ONE switch
// over the residue passes.  VPC.cpp:372-395: most leading zero rows wins, ties go to the later module -- so the last module
// wins unscored while no earlier module has a zero row.
void emit_pm2_select_encode(const std::vector<Module>& mods, Lines& out) {
  const int first = mods.front().idx, last = mods.back().idx;
  out.push_back("  // preset >= 0: the winner is known (a block taken back from the warp's deferral buffer): winner pass only.  defer(best) may take");
  out.push_back("  // a block whose winner is not the last module off this lane's hands (mpc_spec.cuh: it is finished later in a full batch of such");
  out.push_back("  // blocks); the lane then returns kDeferred.");
  out.push_back("  static constexpr uint32_t kDeferred = 0xffffffffu;");
  out.push_back("  template <class Defer>");
  out.push_back("  __device__ static __forceinline__ uint32_t select_encode(uint32_t (&x)[32], int& best, uint32_t& sa, uint32_t& sq, unsigned lanes, int preset, Defer&& defer) {");
  out.push_back("    uint32_t r[32];");
  out.push_back("    uint32_t bestz = 0u;");
  out.push_back("    bool deferred = false;");
  out.push_back(fmt("    int k = %d;", first));
  out.push_back(fmt("    bool scoring = %s;", first == last ? "false" : "true"));
  if (first == last) out.push_back(fmt("    best = %d;", last));
  out.push_back("    if (preset >= 0) { best = preset; k = preset; scoring = false; }");
  out.push_back("    for (;;) {");
  out.push_back("      // the block does not change between iterations, so the compiler would hoist every predictor gather of every module out of");
  out.push_back("      // the loop (and spill them); an empty asm per word makes the block opaque at the top of each iteration (no instruction)");
  out.push_back("#pragma unroll");
  out.push_back(fmt("      for (int i = 0; i < %d; i++) asm volatile(\"\" : \"+r\"(x[i]));", W));
  out.push_back(fmt("      uint32_t ze = %du;", NCH()));
  out.push_back("      switch (k) {");
  for (auto& m : mods) out.push_back(fmt("        case %d: ze = res_%d(x, r, scoring); break;", m.idx, m.idx));
  out.push_back("        default: break;");
  out.push_back("      }");
  out.push_back("      if (!scoring) break;");
  out.push_back(fmt("      const uint32_t z = (ze == %du) ? pm2_lz(r) : ze;", NCH()));
  out.push_back("      if (bestz <= z) { best = k; bestz = z; }");
  out.push_back(fmt("      if (k == %d) {  // every module is scored: r holds the winner's residues only if the last module won with a complete pass", last));
  out.push_back("        scoring = false;");
  out.push_back(fmt("        if (best == %d && ze == %du) break;", last, NCH()));
  out.push_back(fmt("        if (best != %d && defer(best)) { deferred = true; break; }", last));
  out.push_back("        k = best;");
  out.push_back("        continue;");
  out.push_back("      }");
  out.push_back("      k++;");
  out.push_back(fmt("      if (k == %d && bestz == 0u) { best = %d; scoring = false; }  // the last module wins every tie", last, last));
  out.push_back("    }");
  out.push_back("    __syncwarp(lanes);  // one pass over the shared tail for the whole warp, whatever modules its lanes picked");
  out.push_back("    if (deferred) return kDeferred;");
  out.push_back("    return pm2_tail(best, x, r, sa, sq);");
  out.push_back("  }");
}

template <class T>
std::string int_array(const T* vals, int count, int width) {
  std::string s = "{";
  for (int i = 0; i < width; i++) {
    if (i) s += ",";
    s += std::to_string(i < count ? (int)vals[i] : 0);
  }
  return s + "}";
}

std::string pod_initializer(const mpc_config_pod& c) {
  std::vector<std::string> mods;
  for (int i = 0; i < MPC_MAX_MODULES; i++) {
    if (i >= c.num_modules) { mods.push_back("{}"); continue; }
    const mpc_module_pod& m = c.modules[i];
    if (m.kind == MPC_MOD_ALLZERO) mods.push_back("{MPC_MOD_ALLZERO}");
    else if (m.kind == MPC_MOD_ALLWORDSAME) mods.push_back("{MPC_MOD_ALLWORDSAME}");
    else
      mods.push_back(fmt("{MPC_MOD_PREDCOMP, %d, %d, %d, %d, ", m.predictor, m.root, m.consecutive_xor, m.table_size) +
                     int_array(m.base, L, kMaxL) + ", " + int_array(m.diff, L, kMaxL) + ", " + int_array(m.shift, L, kMaxL) + ", " +
                     int_array(m.scan_row, m.table_size, 8 * kMaxL) + ", " + int_array(m.scan_col, m.table_size, 8 * kMaxL) + "}");
  }
  return fmt("{%d, %d, %d, %d, ", c.line_size, c.num_modules, c.has_wordsame, c.first_predcomp) +
         int_array(c.enc_bits, c.num_modules + 1, MPC_MAX_MODULES + 1) + ", {" + join(mods, ",\n   ") + "}}";
}

bool build_modules(const mpc_config_pod& cfg, std::vector<Module>* mods, std::string* why) {
  if (cfg.line_size != 32 && cfg.line_size != 64 && cfg.line_size != 128) { *why = "lineSize is not 32, 64 or 128"; return false; }
  L = cfg.line_size;
  W = L / 4;
  if (cfg.num_modules < 1 || cfg.num_modules > MPC_MAX_MODULES) { *why = "num_modules out of range"; return false; }
  for (int i = cfg.first_predcomp; i < cfg.num_modules; i++) {
    Module m;
    if (!m.init(i, cfg.modules[i], why)) return false;
    mods->push_back(m);
  }
  return true;
}

}  // namespace

// Test hook (tests/test_specgen_residues.py): the residue statements of every PredComp module as plain C++ functions over the
// host forms of the primitives, next to the module's source tables, so that the generated arithmetic is checked on the CPU
// against the per-byte definition (PredictorModule.cpp:37-173, ResidueModule.cpp:12-41) before it ever runs on a GPU.
std::string generate_residue_probe(const mpc_config_pod& cfg, std::string* why) {
  std::vector<Module> mods;
  std::string w;
  if (!build_modules(cfg, &mods, &w)) {
    if (why) *why = w;
    return std::string();
  }
  Lines out;
  out.push_back(fmt("static const int kProbeL = %d, kProbeModules = %d;", L, (int)mods.size()));
  std::vector<std::string> fn, xs, ps, pv, ops, roots;
  for (auto& m : mods) {
    out.push_back(fmt("static void probe_res_%d(const uint32_t* x, uint32_t* r) {", m.idx));
    out.push_back("  uint32_t ah[32];");
    out.push_back(fmt("  for (int i = 0; i < %d; i++) ah[i] = x[i] | 0x80808080u;", W));
    out.push_back("  (void)ah;");
    for (int k = 0; k < W; k++) out.push_back("  { " + m.residue_stmts(k, "rr", true) + fmt(" r[%d] = rr; }", k));
    out.push_back("}");
    std::vector<std::string> a, b, c;
    for (int j = 0; j < L; j++) { a.push_back(fmt("%d", m.xsrc[j])); b.push_back(fmt("%d", m.psrc[j])); c.push_back(fmt("%d", m.pval[j])); }
    xs.push_back("{" + join(a, ",") + "}");
    ps.push_back("{" + join(b, ",") + "}");
    pv.push_back("{" + join(c, ",") + "}");
    fn.push_back(fmt("probe_res_%d", m.idx));
    ops.push_back(fmt("%d", (int)m.op));
    roots.push_back(fmt("%d", m.root));
  }
  // the byte transposition of the first plane-major module into the canonical layout of pm_classify (emit_pm_canonical)
  const Module* pm = nullptr;
  for (auto& m : mods)
    if (!pm && m.family == Module::kPm) pm = &m;
  out.push_back(fmt("static const int kProbeCanonWords = %d;", pm ? 16 * ((NCH() + 3) / 4) : 0));
  out.push_back("static void probe_canon(const uint32_t* r, uint32_t* c) {");
  if (pm) emit_pm_canonical(pm->cols, "r", out);
  out.push_back("  (void)r; (void)c;");
  out.push_back("}");
  {
    std::vector<std::string> cc;
    for (int i = 0; i < L; i++) cc.push_back(fmt("%d", pm ? pm->cols[i] : 0));
    out.push_back(fmt("static const int kProbeCols[%d] = {%s};", L, join(cc, ",").c_str()));
  }
  out.push_back(fmt("static void (*const kProbeFn[])(const uint32_t*, uint32_t*) = {%s};", join(fn, ", ").c_str()));
  out.push_back(fmt("static const int kProbeX[][%d] = {%s};", L, join(xs, ",\n  ").c_str()));
  out.push_back(fmt("static const int kProbeP[][%d] = {%s};", L, join(ps, ",\n  ").c_str()));
  out.push_back(fmt("static const int kProbeV[][%d] = {%s};", L, join(pv, ",\n  ").c_str()));
  out.push_back(fmt("static const int kProbeOp[] = {%s};  // 0 copy, 1 add, 2 shift", join(ops, ", ").c_str()));
  out.push_back(fmt("static const int kProbeRoot[] = {%s};", join(roots, ", ").c_str()));
  std::string text;
  for (auto& l : out) { text += l; text += "\n"; }
  return text;
}

bool spec_eligible(const mpc_config_pod& cfg, std::string* why) {
  std::vector<Module> mods;
  std::string w;
  const bool ok = build_modules(cfg, &mods, &w);
  if (why) *why = w;
  return ok;
}

SpecTraits spec_traits(const mpc_config_pod& cfg) {
  SpecTraits t;
  std::vector<Module> mods;
  std::string why;
  if (!build_modules(cfg, &mods, &why)) return t;
  bool has_pm = false, has_cm = false, all_c = true, any_c = false;
  // pm2 (see emit_res_pm2): every module plane-major, planes MSB first, one column order
  t.pm2 = !mods.empty();
  for (auto& m : mods) t.pm2 = t.pm2 && m.family == Module::kPm && m.rho_identity() && m.cols == mods[0].cols;
  {
    const char* e2 = getenv("MPC_SPEC_PM2");
    if (e2 && e2[0] == '0') t.pm2 = false;
  }
  bool has_bg = false;
  {
    const char* e3 = getenv("MPC_SPEC_DEFER");
    // Measured on B200 (profiles/r02_defer.txt): it pays only where winners other than the last module are rare (P6 smooth +5 %,
    // hash-mixed +4 %) and costs 5-10 % where they are not (ramp, sparse) or never occur (random: code size), so it is OFF unless
    // MPC_SPEC_DEFER=1 asks.
    t.defer = t.pm2 && cfg.line_size == 128 && mods.size() >= 2 && (e3 && e3[0] == '1');
  }
  for (auto& m : mods) {
    if (m.family == Module::kPm) has_pm = true;
    else { has_cm = true; all_c = all_c && m.cxor; any_c = any_c || m.cxor; }
    if (m.family == Module::kBg) has_bg = true;
  }
  const char* e;
  t.eligible = true;
  t.line_size = cfg.line_size;
  t.use_lut = has_cm && !((e = getenv("MPC_SPEC_LUT")) && e[0] == '0');
  t.lut_xor = 0;
  // (bit-gather rows take single bits of the XOR-ed residues: the stage cannot be folded into the table)
  if (t.use_lut && !has_bg && !((e = getenv("MPC_SPEC_LUTXOR")) && e[0] == '0')) t.lut_xor = all_c ? 1 : (!any_c ? 2 : 0);
  // column-major only: ONE CTA of 20 warps per SM at 96 registers (the next allocation step down, 80, spills), one
  // 4 KiB tile stage per warp; measured on B200: 16 warps x 2 stages 4.11 TB/s, 18 x 1 4.16, 20 x 1 4.26, 21/22/24 x 1 (80
  // registers) 3.86-3.93 on the headline workload
  t.warps = (t.use_lut && !has_pm && !has_bg) ? 20 : 16;
  // configs with plane-major modules: ONE CTA of 16 warps at 128 registers (measured against 8 warps at 160 registers, the
  // former choice: P6 smooth 1.42 -> 1.79 TB/s, random 2.23 -> 2.72, E5 mixed 2.08 -> 2.82; 20 warps at 96 registers spill)
  t.stages = 1;
  t.min_ctas = 1;
  if ((e = getenv("MPC_SPEC_WARPS")) && atoi(e) > 0) { t.warps = atoi(e); t.min_ctas = t.warps >= 16 ? 1 : t.min_ctas; }  // tuning overrides
  if ((e = getenv("MPC_SPEC_MIN_CTAS")) && atoi(e) > 0) t.min_ctas = atoi(e);
  e = getenv("MPC_SPEC_SKIP");
  t.skip_zero_groups = e ? (e[0] != '0') : !t.use_lut;
  if ((e = getenv("MPC_SPEC_STAGES")) && (atoi(e) == 1 || atoi(e) == 2)) t.stages = atoi(e);
  // tile loader: MPC_SPEC_TMA=1 selects one TMA tensor copy per warp and tile (cp.async.bulk.tensor 2D, SWIZZLE_128B = the layout the
  // per-thread reads want, completion on a per-warp mbarrier) instead of 8 cp.async per lane.  Bit-exact, but measured slower on B200
  // (F4 smooth 4 300 -> 4 020 GB/s, random 4 035 -> 3 785; 2.5 points of it are the fence.proxy.async that must order the warp's reads
  // of the stage before the next copy), so cp.async stays the default.
  t.tma = t.stages == 1 && cfg.line_size == 128 && (e = getenv("MPC_SPEC_TMA")) && e[0] == '1';
  t.smem_bytes = (size_t)t.warps * t.stages * 4096 + (t.tma ? (size_t)((t.warps * 8 + 15) / 16) * 16 : 0) + 4 * ((((size_t)(cfg.num_modules + 1) * (8 * 128 + 32) + 4 * (cfg.num_modules + 1)) + 3) & ~(size_t)3) +
                 (t.use_lut ? (size_t)kRowLutBytes : 0) + (t.defer ? (size_t)t.warps * kDeferBytesPerWarp + 128 : 0);
  // Regrouping queues (mpc_spec.cuh): one queue per PredComp module in the shared memory that is left (227 KiB per CTA on sm_100,
  // 1 KiB of it reserved by the driver); an entry is the block (128 B) + its index + a flag word.  At least two batches of 32 per
  // queue, else the kernel runs without them (MPC_SPEC_REGROUP=0 switches them off for A/B runs).
  {
    const int nq = (int)mods.size();
    const size_t limit = 227 * 1024 - 1024;
    int cap = 0;
    if (nq >= 2 && t.smem_bytes + 64 < limit) {
      cap = (int)((limit - t.smem_bytes - (size_t)nq * 16 - 16 - 128) / ((size_t)nq * 136));
      cap = cap / 32 * 32;
      if (cap > 128) cap = 128;
      if (cap < 64) cap = 0;
    }
    // Measured on B200 (profiles/r02_regroup.txt): the queues cost more than they save -- the part of the winner pass that
    // divergent lanes do not already share (the per-module residue pass) is ~260 warp instructions per mixed tile, the queue
    // traffic ~300, and warps that run different modules' passes at the same time overflow the 32 KiB instruction cache
    // (F4 mixed 2 524 -> 1 697 GB/s, homogeneous classes -10 %).  They are therefore OFF unless MPC_SPEC_REGROUP=1 asks.
    if (!((e = getenv("MPC_SPEC_REGROUP")) && e[0] == '1') || cfg.line_size != 128) cap = 0;
    if ((e = getenv("MPC_SPEC_QUEUE_CAP")) && atoi(e) > 0 && atoi(e) % 32 == 0 && atoi(e) <= cap) cap = atoi(e);
    t.queue_cap = cap;
    if (cap > 0) t.smem_bytes += (size_t)nq * cap * 136 + (size_t)nq * 16 + 16 + 128;
    // With regrouping a warp's lanes mostly run ONE module's winner pass, so the row classifier no longer has to be shared
    // across modules behind a reconvergence point: each module's pass runs straight into its own copy of the classifier and the
    // 32 row words never have to be live all at once next to the block (no spills at 96 registers; MPC_SPEC_FUSED=0/1 overrides).
    t.fused_encode = cap > 0 && !has_pm;
    if ((e = getenv("MPC_SPEC_FUSED"))) t.fused_encode = e[0] == '1';
    // Column-major configs get BOTH forms: a warp whose stage-3 lanes all picked the same module (homogeneous data -- the
    // common case) runs that module's fused pass, any other warp the per-module residue passes followed by the ONE shared
    // classifier (measured: fused +1.5 % on smooth / random data, shared +25 % on finely mixed data; MPC_SPEC_ADAPTIVE=0/1).
    t.adaptive_encode = !t.fused_encode && !has_pm && !has_bg && t.use_lut;
    if ((e = getenv("MPC_SPEC_ADAPTIVE"))) t.adaptive_encode = e[0] == '1' && !t.fused_encode;
  }
  return t;
}

std::string generate_spec_source(const mpc_config_pod& cfg, const std::string& name, bool jit, std::string* why) {
  std::vector<Module> mods;
  std::string w;
  if (!build_modules(cfg, &mods, &w)) {
    if (why) *why = w;
    return std::string();
  }
  const SpecTraits t = spec_traits(cfg);
  const int n = cfg.num_modules, first = cfg.first_predcomp;
  Lines out;
  if (jit) {
    out.push_back("// Generated at run time by libmpc_b200 (mpc_specgen.cpp) for a config without a compiled-in specialisation.");
  } else {
    out.push_back(fmt("// AUTO-GENERATED by tools/specgen (csrc/mpc_specgen.cpp) from configs/%s.json -- do not edit.", name.c_str()));
    out.push_back("// Straight-line schedule of the MPC per-block path for this config over the primitives of mpc_spec.cuh.");
  }
  out.push_back(jit ? "#include \"mpc_spec.cuh\"" : "#include \"../mpc_spec.cuh\"");
  if (!jit) out.push_back("#include \"../mpc_spec.h\"");
  out.push_back("");
  out.push_back("namespace mpc {");
  out.push_back(fmt("namespace spec_%s {", name.c_str()));
  out.push_back("using namespace mpc::spec;");
  out.push_back("template <class F> __device__ __forceinline__ uint32_t shiftmix(uint32_t p, F f) { return f(p); }");
  out.push_back("");
  static const char* kPredNames[] = {"OneBasePredictor", "ConsecutiveBasePredictor", "DiffBasePredictor", "WeightBasePredictor"};
  RowLayout lay;
  {
    std::vector<std::vector<int>> cm_cols;
    bool any_bg = false;
    for (auto& m : mods) {
      if (m.family == Module::kCm) { cm_cols.push_back(m.cols); cm_cols.back().resize(L, -1); }
      any_bg = any_bg || m.family == Module::kBg;
    }
    lay = plan_row_layout((t.use_lut && !any_bg) ? cm_cols : std::vector<std::vector<int>>());
  }
  if (lay.paired) emit_encode_rows(lay, out);
  for (auto& m : mods) {
    out.push_back(fmt("// ---- module %d: %s, root %d, %s XOR, %s-major scan ----", m.idx, kPredNames[m.predictor], m.root,
                      m.cxor ? "consecutive" : "first-plane", m.family == Module::kCm ? "column" : (m.family == Module::kPm ? "plane" : "bit-gather (arbitrary table), row")));
    if (t.pm2) {
      emit_res_pm2(m, out);
    } else {
      if (m.family == Module::kCm) emit_score_cm(m, out); else if (m.family == Module::kPm) emit_score_pm(m, out); else emit_score_bg(m, out);
      emit_full(m, out, t.lut_xor, lay);
    }
    out.push_back("");
  }
  if (t.pm2) emit_pm2_shared(mods, out);
  // selector groups (see plan_sel_group): merged winner passes for warps whose lanes picked different modules of a group
  std::vector<SelGroup> groups;
  std::vector<int> group_of(mods.size(), -1);
  {
    const char* eg = getenv("MPC_SPEC_SELGROUP");
    if (!t.pm2 && !t.fused_encode && !(eg && eg[0] == '0')) {
      for (size_t i = 0; i < mods.size(); i++) {
        if (group_of[i] >= 0) continue;
        std::vector<int> members{(int)i};
        SelGroup g, tryg;
        for (size_t j = i + 1; j < mods.size(); j++) {
          if (group_of[j] >= 0) continue;
          std::vector<int> cand = members;
          cand.push_back((int)j);
          if (plan_sel_group(mods, cand, &tryg)) { members = cand; g = tryg; }
        }
        if (members.size() >= 2) {
          for (int mi : members) group_of[(size_t)mi] = (int)groups.size();
          groups.push_back(g);
        }
      }
    }
    for (size_t gi = 0; gi < groups.size(); gi++) {
      std::vector<std::string> names;
      for (int mi : groups[gi].members) names.push_back(std::to_string(mods[(size_t)mi].idx));
      out.push_back(fmt("// ---- selector group %d: modules %s share one winner pass in warps whose lanes picked different winners ----", (int)gi, join(names, ", ").c_str()));
      emit_full(mods[(size_t)groups[gi].members[0]], out, t.lut_xor, lay, &groups[gi], (int)gi);
      out.push_back("");
    }
  }
  out.push_back("struct Cfg {");
  out.push_back(fmt("  static constexpr int kLineBytes = %d;  // a thread holds 128 bytes = 128 / kLineBytes consecutive lines", L));
  out.push_back(fmt("  static constexpr int kWords = %d;", W));
  out.push_back(fmt("  static constexpr int kNumModules = %d;", n));
  out.push_back(fmt("  static constexpr int kFirst = %d;", first));
  out.push_back(fmt("  static constexpr bool kHasWordSame = %s;", cfg.has_wordsame ? "true" : "false"));
  // register budget: plane-major modules keep the line, its residues and the transposed rows live at once;
  // column-major only: 20 warps in ONE CTA per SM (96 registers), one 4 KiB tile stage per warp, so that the 64 KiB row-cost table, the
  // tile stages and the histogram fit the 227 KiB of shared memory
  out.push_back(fmt("  static constexpr int kWarps = %d;", t.warps));
  out.push_back(fmt("  static constexpr int kStages = %d;  // shared-memory tile stages per warp", t.stages));
  out.push_back(fmt("  static constexpr bool kTma = %s;  // tiles arrive by TMA (cp.async.bulk.tensor, one per warp and tile) instead of cp.async", t.tma ? "true" : "false"));
  out.push_back(fmt("  static constexpr bool kUseLut = %s;  // shared-memory row-cost table (column-major modules)", t.use_lut ? "true" : "false"));
  out.push_back(fmt("  static constexpr bool kSkipZeroGroups = %s;  // branch around groups of eight zero rows in the encoder", t.skip_zero_groups ? "true" : "false"));
  out.push_back(fmt("  static constexpr int kLutXor = %d;  // 0: table indexed by scan rows; 1 / 2: XOR stage (consecutive / first-plane) folded into the table",
                    t.use_lut ? t.lut_xor : 0));
  out.push_back(fmt("  static constexpr int kMinCtasPerSm = %d;  // __launch_bounds__: register budget 65536 / (threads * CTAs)", t.min_ctas));
  out.push_back(fmt("  static constexpr int kQueueCap = %d;  // entries per regrouping queue (one queue per PredComp module); 0 = no regrouping", t.queue_cap));
  out.push_back(fmt("  static constexpr bool kPm2 = %s;  // one residue pass per module + shared scoring / statistics / classifier (select_encode)", t.pm2 ? "true" : "false"));
  out.push_back(fmt("  static constexpr bool kDefer = %s;  // per-warp deferral buffer for blocks whose winner is not the last module (pm2, 128-byte lines)", t.defer ? "true" : "false"));
  out.push_back("  __device__ static __forceinline__ uint32_t enc(int k) {  // encoding bits of cluster k-1, VPC.cpp:102-117");
  out.push_back("    switch (k) {");
  for (int k = 0; k <= n; k++) out.push_back(fmt("      case %d: return %du;", k, cfg.enc_bits[k]));
  out.push_back("    }");
  out.push_back("    return 0u;");
  out.push_back("  }");
  if (t.pm2) {
    emit_pm2_select_encode(mods, out);
    out.push_back("  __device__ static __forceinline__ void select(const uint32_t (&)[32], int&, uint32_t&, unsigned) {}");
    out.push_back("  static constexpr bool kAdaptiveEncode = false;");
    out.push_back("  __device__ static __forceinline__ uint32_t encode(int, const uint32_t (&)[32], uint32_t&, uint32_t&, unsigned, const uint8_t*, bool) { return 0u; }");
    out.push_back("};");
    out.push_back("");
  } else {
  out.push_back("  static constexpr uint32_t kDeferred = 0xffffffffu;");
  out.push_back("  template <class Defer>");
  out.push_back("  __device__ static __forceinline__ uint32_t select_encode(uint32_t (&)[32], int&, uint32_t&, uint32_t&, unsigned, int, Defer&&) { return 0u; }");
  out.push_back("  // VPC.cpp:372-395: most leading zero rows wins, ties go to the later module");
  out.push_back("  __device__ static __forceinline__ void select(const uint32_t (&x)[32], int& best, uint32_t& bestz, unsigned lanes) {");
  out.push_back("    uint32_t z;");
  for (size_t i = 0; i + 1 < mods.size(); i++)
    out.push_back(fmt("    z = score_%d(x); if (bestz <= z) { best = %d; bestz = z; }", mods[i].idx, mods[i].idx));
  if (!mods.empty()) {
    const int li = mods.back().idx;
    out.push_back("    // the last module wins every tie, so while no earlier module has a zero row it wins unscored");
    out.push_back(fmt("    if (bestz == 0u) { best = %d; } else { z = score_%d(x); if (bestz <= z) { best = %d; bestz = z; } }", li, li, li));
    out.push_back("    __syncwarp(lanes);  // reconverge before the encoder: both sides of the branch share it");
  } else {
    out.push_back("    (void)z; (void)x; (void)best; (void)bestz; (void)lanes;");
  }
  out.push_back("  }");
  out.push_back("  // residue sums (VPC.cpp:417-443) + common encoder (FPCModule.cpp:19-85) of the chosen module");
  out.push_back(fmt("  static constexpr bool kAdaptiveEncode = %s;  // homogeneous warps run a fused per-module pass, others the shared classifier", t.adaptive_encode ? "true" : "false"));
  out.push_back("  __device__ static __forceinline__ uint32_t encode(int best, const uint32_t (&x)[32], uint32_t& sa, uint32_t& sq, unsigned lanes, const uint8_t* lut, bool uniform) {");
  out.push_back("    uint32_t c[32];");
  // encoder families in use: column-major, or plane-major with a given pair of plane selectors
  std::vector<std::pair<int, std::pair<unsigned, unsigned>>> fams;
  auto fam_of = [&](const Module& m) {
    return std::make_pair(m.family == Module::kPm ? 1 : 0, m.family == Module::kPm ? pm_selectors(m) : std::make_pair(0u, 0u));
  };
  for (auto& m : mods)
    if (std::find(fams.begin(), fams.end(), fam_of(m)) == fams.end()) fams.push_back(fam_of(m));
  std::sort(fams.begin(), fams.end());
  auto call = [&](const std::pair<int, std::pair<unsigned, unsigned>>& f) {
    return f.first == 0 ? (lay.paired ? std::string("encode_rows(c, lut)") : fmt("encode_cm<%d, kUseLut, kSkipZeroGroups>(c, lut)", W))
                        : fmt("encode_pm<%d, 0x%04xu, 0x%04xu>(c)", NCH(), f.second.first, f.second.second);
  };
  if (t.adaptive_encode) {
    out.push_back("    if (uniform) {  // every lane of the warp runs the same module: straight from its residue pass into the classifier");
    out.push_back("      switch (best) {");
    for (auto& m : mods) {
      const int fid = (int)(std::find(fams.begin(), fams.end(), fam_of(m)) - fams.begin());
      out.push_back(fmt("        case %d: full_%d(x, c, sa, sq); return %s;", m.idx, m.idx, call(fams[(size_t)fid]).c_str()));
    }
    out.push_back("        default: return 0u;");
    out.push_back("      }");
    out.push_back("    }");
  } else {
    out.push_back("    (void)uniform;");
  }
  out.push_back("    int fam = 0;");
  std::string else_prefix = "    ";
  for (size_t gi = 0; gi < groups.size(); gi++) {
    const SelGroup& g = groups[gi];
    std::vector<std::string> tests;
    for (int mi : g.members) tests.push_back(fmt("best == %d", mods[(size_t)mi].idx));
    out.push_back(else_prefix + "if (" + join(tests, " || ") + ") {  // one pass for the group, every lane with its own selectors");
    std::string args;
    for (size_t c = 0; c < g.cls_sel.size(); c++) {
      std::string e = fmt("0x%04xu", g.cls_sel[c].back());
      for (int k = (int)g.members.size() - 2; k >= 0; k--) e = fmt("(best == %d ? 0x%04xu : %s)", mods[(size_t)g.members[(size_t)k]].idx, g.cls_sel[c][(size_t)k], e.c_str());
      out.push_back(fmt("      const uint32_t s%d = %s;", (int)c, e.c_str()));
      args += fmt(", s%d", (int)c);
    }
    const int fid = (int)(std::find(fams.begin(), fams.end(), fam_of(mods[(size_t)g.members[0]])) - fams.begin());
    out.push_back(fmt("      full_g%d(x, c, sa, sq%s); fam = %d;", (int)gi, args.c_str(), fid));
    out.push_back("    } else");
    else_prefix = "    ";
  }
  out.push_back("    switch (best) {");
  for (size_t i = 0; i < mods.size(); i++) {
    const Module& m = mods[i];
    if (group_of[i] >= 0) continue;
    const int fid = (int)(std::find(fams.begin(), fams.end(), fam_of(m)) - fams.begin());
    if (t.fused_encode) out.push_back(fmt("      case %d: full_%d(x, c, sa, sq); return %s;", m.idx, m.idx, call(fams[(size_t)fid]).c_str()));
    else out.push_back(fmt("      case %d: full_%d(x, c, sa, sq); fam = %d; break;", m.idx, m.idx, fid));
  }
  out.push_back("      default: break;");
  out.push_back("    }");
  if (t.fused_encode) {
    out.push_back("    (void)fam; (void)lanes; (void)lut;");
    out.push_back("    return 0u;");
  } else {
  out.push_back("    __syncwarp(lanes);  // the row classifier below is shared by all modules: run it once per warp");
  if (fams.size() == 1) {
    out.push_back("    (void)fam;");
    out.push_back("    return " + call(fams[0]) + ";");
  } else {
    for (size_t i = 0; i < fams.size(); i++) out.push_back(fmt("    if (fam == %d) return %s;", (int)i, call(fams[i]).c_str()));
    if (fams.empty()) out.push_back("    (void)fam; (void)c; (void)x; (void)sa; (void)sq; (void)lut;");
    out.push_back("    return 0u;");
  }
  }
  out.push_back("  }");
  out.push_back("};");
  out.push_back("");
  }  // !pm2
  if (jit) {
    out.push_back(fmt("}  // namespace spec_%s", name.c_str()));
    out.push_back("}  // namespace mpc");
    out.push_back(fmt("extern \"C\" __global__ void __launch_bounds__(mpc::spec_%s::Cfg::kWarps * 32, mpc::spec_%s::Cfg::kMinCtasPerSm)", name.c_str(), name.c_str()));
    out.push_back("mpc_jit_kernel(const uint4* __restrict__ lines, unsigned long long n_blocks, unsigned short* __restrict__ packed,");
    out.push_back("               unsigned long long* __restrict__ stats, const uint4* __restrict__ row_lut, unsigned int* __restrict__ sched,");
    out.push_back("               unsigned int static_rounds, const __grid_constant__ mpc::TileTmap tmap) {");
    out.push_back(fmt("  mpc::spec::spec_kernel_body<mpc::spec_%s::Cfg>(lines, n_blocks, packed, stats, row_lut, sched, static_rounds, &tmap);", name.c_str()));
    out.push_back("}");
  } else {
    out.push_back("static const mpc_config_pod kPod =");
    out.push_back("  " + pod_initializer(cfg) + ";");
    out.push_back("");
    out.push_back("static bool matches(const mpc_config_pod& cfg) { return spec_pod_equal(cfg, kPod); }");
    out.push_back("static cudaError_t launch(const mpc_config_pod&, const uint8_t* d_lines, uint64_t n_blocks, uint16_t* d_packed,");
    out.push_back("                          uint64_t* d_stats, const uint8_t* d_row_lut, uint32_t* d_sched, int sm_count, cudaStream_t stream) {");
    out.push_back("  return launch_spec<Cfg>(d_lines, n_blocks, d_packed, d_stats, d_row_lut, d_sched, sm_count, stream);");
    out.push_back("}");
    out.push_back(fmt("}  // namespace spec_%s", name.c_str()));
    out.push_back(fmt("extern const SpecKernel kSpec_%s = {\"%s\", spec_%s::matches, spec_%s::launch, spec_%s::Cfg::kLutXor};", name.c_str(),
                      name.c_str(), name.c_str(), name.c_str(), name.c_str()));
    out.push_back("}  // namespace mpc");
  }
  std::string text;
  for (auto& l : out) { text += l; text += "\n"; }
  return text;
}

}  // namespace mpc
