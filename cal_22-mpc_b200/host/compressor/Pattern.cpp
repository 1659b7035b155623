#include "Pattern.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace comp {

// Pattern.h:127-153: probabilities in symbol order, entropy = sum of -p * log2(p) over the symbols that occurred
double PatternResult::ComputeEntropy(const std::map<uint8_t, uint64_t>& symbolCounts) {
  uint64_t sum = 0;
  for (auto& kv : symbolCounts) sum += kv.second;
  double entropy = 0;
  for (auto& kv : symbolCounts) {
    const double probability = (double)kv.second / (double)sum;
    entropy += -probability * log2(probability);
  }
  return entropy;
}

void PatternResult::Print(std::string workloadName, std::string filePath) {
  std::ofstream file;
  std::ostream* os = &std::cout;
  if (filePath != "") {
    if (!isFileExists(filePath)) {
      file.open(filePath);
      if (!file.is_open()) {
        std::cout << "File is not open: \"" << filePath << "\"" << std::endl;
        exit(1);
      }
      file << "Workload,Entropy [b/B],Entropy except AllZeros AllWordSame [b/B],Zeros [B],Repeated Line [B],Temporal Locality [B],"
           << "B8D1-Implicit [B],B8D1-Explicit [B],B8D2-Implicit [B],B8D2-Explicit [B],B8D4-Implicit [B],B8D4-Explicit [B],"
           << "B4D1-Implicit [B],B4D1-Explicit [B],B4D2-Implicit [B],B4D2-Explicit [B],B2D1-Implicit [B],B2D1-Explicit [B],"
           << "Undefined [B],Total Size [B]," << std::endl;
      file.close();
    }
    file.open(filePath, std::ios_base::app);
    os = &file;
  }
  *os << workloadName << "," << formatDouble(ComputeEntropy(SymbolCounts)) << ","
      << formatDouble(ComputeEntropy(SymbolCountsExceptAllZerosAllWordSame)) << "," << Z << "," << R << "," << T << ",";
  for (int i = 0; i < 6; i++) *os << ImplicitCounts[i] << "," << ExplicitCounts[i] << ",";
  *os << U << "," << Total << "," << std::endl;
}

Pattern::Pattern(unsigned lineSize) : m_LineSize(lineSize) {
  m_Stat = new PatternResult(lineSize);
  m_Stat->CompressorName = "Pattern Checker";
}

unsigned Pattern::CompressLine(std::vector<uint8_t>& dataLine) {
  m_Pending.insert(m_Pending.end(), dataLine.begin(), dataLine.end());
  return 0;
}

void Pattern::CompressBatch(const uint8_t* lines, uint64_t nLines) {
  m_Pending.insert(m_Pending.end(), lines, lines + nLines * m_LineSize);
}

CompResult* Pattern::GetResult() {
  if (!m_Pending.empty()) {
    mpc_pattern_stats s;
    float ms = 0;
    if (mpc_pattern_run_host(0, m_Pending.data(), m_Pending.size() / m_LineSize, m_LineSize, 0, nullptr, &s, &ms) != MPC_OK) {
      printf("PATTERN: %s\n", mpc_pattern_error());
      exit(1);
    }
    PatternResult* r = static_cast<PatternResult*>(m_Stat);
    r->Z += s.zeros_bytes; r->R += s.repeated_bytes; r->T += s.temporal_bytes; r->U += s.undefined_bytes; r->Total += s.total_bytes;
    for (int i = 0; i < 6; i++) { r->ImplicitCounts[i] += s.implicit_bytes[i]; r->ExplicitCounts[i] += s.explicit_bytes[i]; }
    for (int b = 0; b < 256; b++) {
      if (s.symbol_counts[b]) r->SymbolCounts[(uint8_t)b] += s.symbol_counts[b];
      if (s.symbol_counts_nontrivial[b]) r->SymbolCountsExceptAllZerosAllWordSame[(uint8_t)b] += s.symbol_counts_nontrivial[b];
    }
    m_KernelMs += ms;
    m_TemporalOnHost = s.temporal_path != 0;
    m_Pending.clear();
    m_Pending.shrink_to_fit();
  }
  return m_Stat;
}

}  // namespace comp
