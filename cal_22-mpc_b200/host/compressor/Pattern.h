// Host-side mirror of the reference's data-pattern analysis (reference src/compressor/Pattern.h:34-272): comp::Pattern
// forwards to the C ABI (mpc_pattern_run_host: GPU kernel + hash sort; the temporal-locality cache falls back to an
// in-order host pass only when it would evict), comp::PatternResult writes the same CSV (Pattern.h:155-222).
// Like the reference's, it never calls Update, so the ratio the CLI prints is 0.
#ifndef MPCB_PATTERN_H_
#define MPCB_PATTERN_H_

#include <map>
#include <string>
#include <vector>

#include "Compressor.h"
#include "mpc_capi.h"

namespace comp {

struct PatternResult : public CompResult {
  explicit PatternResult(unsigned lineSize)
      : CompResult(lineSize), ImplicitCounts(6, 0), ExplicitCounts(6, 0), Z(0), R(0), T(0), U(0), Total(0) {}
  static double ComputeEntropy(const std::map<uint8_t, uint64_t>& symbolCounts);  // Pattern.h:127-153
  void Print(std::string workloadName = "", std::string filePath = "") override;
  std::vector<uint64_t> ImplicitCounts, ExplicitCounts;
  std::map<uint8_t, uint64_t> SymbolCounts, SymbolCountsExceptAllZerosAllWordSame;  // only symbols that occurred
  uint64_t Z, R, T, U;  // Zeros, Repeated, TemporalLocality, NotDefined (bytes)
  uint64_t Total;
};

class Pattern : public Compressor {
 public:
  explicit Pattern(unsigned lineSize);
  ~Pattern() override { delete m_Stat; }
  unsigned CompressLine(std::vector<uint8_t>& dataLine) override;  // queued: the cache makes the result stream-dependent
  void CompressBatch(const uint8_t* lines, uint64_t nLines) override;
  CompResult* GetResult() override;
  double KernelMs() const { return m_KernelMs; }
  bool TemporalOnHost() const { return m_TemporalOnHost; }

 private:
  unsigned m_LineSize;
  std::vector<uint8_t> m_Pending;
  double m_KernelMs = 0;
  bool m_TemporalOnHost = false;
};

}  // namespace comp
#endif
