// TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// Thin extern "C" wrapper around the UNMODIFIED reference classes, compiled together with
// the reference's own sources where they lie under /root/reference (see build_ref.sh).
// It lets the tests drive comp::Compressor::CompressLine (Compressor.h:28) one line at a
// time, exactly as compressLines does (main.cpp:229-244), and read back the per-line
// return value, the cluster each line was accounted to (VPC.h:49-60) and the final stats.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "compressor/VPC.h"
#include "compressor/FPC.h"
#include "compressor/BDI.h"
#include "compressor/BPC.h"
#include "compressor/CPACK.h"
#include "compressor/SC2.h"
#include "compressor/Pattern.h"

namespace {
struct Handle {
  comp::Compressor* c = nullptr;
  std::string alg;
  unsigned lineSize = 0;
  int numModules = 0;
};
}  // namespace

extern "C" {

// alg: "VPC" (cfg = json path, line_size ignored), "BDI"/"FPC"/"BPC"/"CPACK" (cfg unused),
// "SC2" (sampling = number of sampling lines, as main.cpp:108-114 computes it).
void* ref_create(const char* alg, const char* cfg, unsigned line_size, unsigned long long sampling) {
  Handle* h = new Handle;
  h->alg = alg;
  h->lineSize = line_size;
  std::string a(alg);
  if (a == "VPC") {
    comp::VPC* v = new comp::VPC(std::string(cfg));
    h->c = v;
    h->lineSize = (unsigned)v->GetCachelineSize();
    h->numModules = v->GetNumModules();
  } else if (a == "BDI") h->c = new comp::BDI(line_size);
  else if (a == "FPC") h->c = new comp::FPC(line_size);
  else if (a == "BPC") h->c = new comp::BPC(line_size);
  else if (a == "CPACK") h->c = new comp::CPACK(line_size);
  else if (a == "SC2") h->c = new comp::SC2(line_size, sampling);
  else if (a == "PATTERN") h->c = new comp::Pattern(line_size);
  else { delete h; return nullptr; }
  return h;
}

unsigned ref_line_size(void* hv) { return ((Handle*)hv)->lineSize; }
int ref_num_modules(void* hv) { return ((Handle*)hv)->numModules; }

// Push n lines of `line_size` bytes through CompressLine.  sizes[i] = return value.
// sels[i] (VPC only, may be null) = cluster whose count moved (-1 = uncompressed).
void ref_compress(void* hv, const uint8_t* lines, unsigned long long n, unsigned line_size,
                  uint32_t* sizes, int32_t* sels) {
  Handle* h = (Handle*)hv;
  std::vector<uint8_t> line(line_size);
  comp::VPCResult* vr = (h->alg == "VPC") ? static_cast<comp::VPCResult*>(h->c->GetResult()) : nullptr;
  std::vector<uint64_t> before;
  for (unsigned long long i = 0; i < n; i++) {
    line.assign(lines + i * line_size, lines + (i + 1) * line_size);
    if (vr && sels) {
      before.clear();
      for (int m = -1; m < h->numModules; m++) before.push_back(vr->m_ClusterStats[m].count);
    }
    unsigned s = h->c->CompressLine(line);
    if (sizes) sizes[i] = s;
    if (vr && sels) {
      int sel = -2;
      for (int m = -1; m < h->numModules; m++)
        if (vr->m_ClusterStats[m].count != before[m + 1]) sel = m;
      sels[i] = sel;
    }
  }
}

// totals[0]=OriginalSize totals[1]=CompressedSize ; *ratio = CompRatio
void ref_totals(void* hv, uint64_t* totals, double* ratio) {
  comp::CompResult* r = ((Handle*)hv)->c->GetResult();
  totals[0] = r->OriginalSize;
  totals[1] = r->CompressedSize;
  *ratio = r->CompRatio;
}

// VPC only.  For cluster m = -1..N-1 (row m+1): stat[row*4+0..3] = count, orig, comp, numLines(MAE/MSE);
// fl[row*3+0..2] = compRatio, MAE, MSE ; hist[row*hist_bins + s] = compSizeHistogram[s] (s < hist_bins).
void ref_vpc_stats(void* hv, uint64_t* stat, double* fl, uint64_t* hist, unsigned hist_bins) {
  Handle* h = (Handle*)hv;
  comp::VPCResult* vr = static_cast<comp::VPCResult*>(h->c->GetResult());
  for (int m = -1; m < h->numModules; m++) {
    int row = m + 1;
    comp::ClusterStat& cs = vr->m_ClusterStats[m];
    stat[row * 4 + 0] = cs.count;
    stat[row * 4 + 1] = cs.originalSize;
    stat[row * 4 + 2] = cs.compressedSize;
    stat[row * 4 + 3] = vr->m_NumLines[m];
    fl[row * 3 + 0] = cs.compRatio;
    fl[row * 3 + 1] = vr->m_MAE[m];
    fl[row * 3 + 2] = vr->m_MSE[m];
    if (hist)
      for (auto& kv : cs.compSizeHistogram)
        if (kv.first >= 0 && (unsigned)kv.first < hist_bins) hist[(size_t)row * hist_bins + kv.first] = kv.second;
  }
}

// BDI: 9 counters (BDI.h:29-33); FPC: TotalWords + 8 (FPC.h:31-37); BPC: TotalWords + 7 (BPC.h:29-33)
int ref_counts(void* hv, uint64_t* out, int cap) {
  Handle* h = (Handle*)hv;
  int n = 0;
  if (h->alg == "BDI") {
    auto* r = static_cast<comp::BDIResult*>(h->c->GetResult());
    for (auto v : r->Counts) if (n < cap) out[n++] = v;
  } else if (h->alg == "FPC") {
    auto* r = static_cast<comp::FPCResult*>(h->c->GetResult());
    if (n < cap) out[n++] = r->TotalWords;
    for (auto v : r->Counts) if (n < cap) out[n++] = v;
  } else if (h->alg == "BPC") {
    auto* r = static_cast<comp::BPCResult*>(h->c->GetResult());
    if (n < cap) out[n++] = r->TotalWords;
    for (auto v : r->Counts) if (n < cap) out[n++] = v;
  }
  return n;
}

// PATTERN only: the counters of comp::PatternResult (Pattern.h:224-231) in the layout of orc_pattern_run
// ([0] Z [1] R [2] T [3] U [4] Total [5..10] Implicit [11..16] Explicit [17..272] SymbolCounts [273..528] ...Except).
void ref_pattern_stats(void* hv, uint64_t* out) {
  auto* r = static_cast<comp::PatternResult*>(((Handle*)hv)->c->GetResult());
  memset(out, 0, 529 * sizeof(uint64_t));
  out[0] = r->Z; out[1] = r->R; out[2] = r->T; out[3] = r->U; out[4] = r->Total;
  for (int i = 0; i < 6; i++) { out[5 + i] = r->ImplicitCounts[i]; out[11 + i] = r->ExplicitCounts[i]; }
  for (auto& kv : r->SymbolCounts) out[17 + kv.first] = kv.second;
  for (auto& kv : r->SymbolCountsExceptAllZerosAllWordSame) out[273 + kv.first] = kv.second;
}
void ref_pattern_print(void* hv, const char* workload, const char* path) {
  static_cast<comp::PatternResult*>(((Handle*)hv)->c->GetResult())->Print(workload, path);
}

// Run Print/PrintDetail exactly as main.cpp:160-165 does for VPC.
void ref_vpc_print(void* hv, const char* workload, const char* path, const char* detail_path) {
  comp::VPCResult* vr = static_cast<comp::VPCResult*>(((Handle*)hv)->c->GetResult());
  vr->Print(workload, path);
  vr->PrintDetail(workload, detail_path);
}

}  // extern "C"
