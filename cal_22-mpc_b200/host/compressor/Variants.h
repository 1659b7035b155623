// Host-side mirrors of the reference's secondary compressors (BASELINE config #5): comp::BDI, comp::FPC,
// comp::BPC (GPU, one block per thread), comp::SC2 (GPU histogram + host tree + GPU lookup) and comp::CPACK
// (host, sequential: its dictionary persists across lines).  Each forwards to the C ABI and fills a result
// struct whose Print writes the same CSV the reference writes
// (reference src/compressor/BDI.h:35-84, FPC.h:39-87, BPC.h:35-84, CPACK.h:45-92, CompResult.h:37-74).
#ifndef MPCB_VARIANTS_H_
#define MPCB_VARIANTS_H_

#include <string>
#include <vector>

#include "Compressor.h"
#include "mpc_capi.h"

namespace comp {

// CSV with a "Total Words" column (FPC/BPC/CPACK) or without (BDI), then one column per counter
struct CountedResult : public CompResult {
  CountedResult(unsigned lineSize, std::vector<std::string> columns, bool hasTotalWords)
      : CompResult(lineSize), TotalWords(0), Counts(columns.size(), 0), m_Columns(std::move(columns)), m_HasTotal(hasTotalWords) {}
  void Print(std::string workloadName = "", std::string filePath = "") override;
  uint64_t TotalWords;
  std::vector<uint64_t> Counts;

 private:
  std::vector<std::string> m_Columns;
  bool m_HasTotal;
};

class VariantCompressor : public Compressor {
 public:
  // alg: "BDI", "FPC", "BPC", "CPACK", "SC2"; loaderRows = trace::Loader::GetNumLines() (SC2 sampling, main.cpp:108-114)
  VariantCompressor(const std::string& alg, unsigned lineSize, unsigned long long loaderRows);
  ~VariantCompressor() override { delete m_Stat; }
  unsigned CompressLine(std::vector<uint8_t>& dataLine) override;
  void CompressBatch(const uint8_t* lines, uint64_t nLines) override;
  CompResult* GetResult() override;
  double KernelMs() const { return m_KernelMs; }

 private:
  void accumulate(const mpc_variant_stats& s);
  std::string m_Alg;
  unsigned m_LineSize;
  unsigned long long m_Sampling;
  std::vector<uint8_t> m_Pending;  // SC2 / CPACK depend on the whole stream: lines handed over one by one wait here
  double m_KernelMs = 0;
};

}  // namespace comp
#endif
