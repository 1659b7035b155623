#include "Variants.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace comp {

void CountedResult::Print(std::string workloadName, std::string filePath) {
  std::ofstream file;
  std::ostream* os = &std::cout;
  if (filePath != "") {
    if (!isFileExists(filePath)) {
      file.open(filePath);
      if (!file.is_open()) {
        std::cout << "File is not open: \"" << filePath << "\"" << std::endl;
        exit(1);
      }
      file << "Workload,Original Size,Compressed Size,Compression Ratio,";
      if (m_HasTotal) file << "Total Words,";
      for (auto& c : m_Columns) file << c << ",";
      file << std::endl;
      file.close();
    }
    file.open(filePath, std::ios_base::app);
    os = &file;
  }
  *os << workloadName << "," << OriginalSize << "," << CompressedSize << "," << formatDouble(CompRatio) << ",";
  if (m_HasTotal) *os << TotalWords << ",";
  for (uint64_t c : Counts) *os << c << ",";
  *os << std::endl;
}

static std::vector<std::string> numbered(const char* prefix, int n) {
  std::vector<std::string> v;
  for (int i = 0; i < n; i++) v.push_back(std::string(prefix) + std::to_string(i));
  return v;
}

VariantCompressor::VariantCompressor(const std::string& alg, unsigned lineSize, unsigned long long loaderRows)
    : m_Alg(alg), m_LineSize(lineSize) {
  // main.cpp:108-114
  m_Sampling = std::max<unsigned long long>(10000, std::min<unsigned long long>(loaderRows / 100, 1000000));
  if (alg == "BDI") {
    m_Stat = new CountedResult(lineSize, {"Zeros", "Repeated", "B8D1", "B8D2", "B8D4", "B4D1", "B4D2", "B2D1", "Uncompressed"}, false);
    m_Stat->CompressorName = "Base-Delta Immediate";
  } else if (alg == "FPC") {
    m_Stat = new CountedResult(lineSize, numbered("Prefix", 8), true);
    m_Stat->CompressorName = "Frequent Pattern Compression";
  } else if (alg == "BPC") {
    m_Stat = new CountedResult(lineSize, numbered("Pattern", 7), true);
    m_Stat->CompressorName = "Bit-Plane Compression";
  } else if (alg == "CPACK") {
    m_Stat = new CountedResult(lineSize, numbered("Pattern", 6), true);
    m_Stat->CompressorName = "C-Pack";
  } else {  // SC2 reports through the base CompResult (SC2.h:96-104)
    m_Stat = new CompResult(lineSize);
    m_Stat->CompressorName = "SC2-Huffman";
  }
}

void VariantCompressor::accumulate(const mpc_variant_stats& s) {
  m_Stat->OriginalSize += s.original_bits;
  m_Stat->CompressedSize += s.compressed_bits;
  m_Stat->CompRatio = m_Stat->CompressedSize ? (double)m_Stat->OriginalSize / (double)m_Stat->CompressedSize : 0;
  CountedResult* cr = dynamic_cast<CountedResult*>(m_Stat);
  if (!cr) return;
  if (m_Alg == "BDI") {
    for (int i = 0; i < 9; i++) cr->Counts[i] += s.counts[i];
  } else if (m_Alg == "FPC") {
    for (int i = 0; i < 8; i++) { cr->Counts[i] += s.counts[i]; cr->TotalWords += s.counts[i]; }
  } else if (m_Alg == "BPC") {
    for (int i = 0; i < 7; i++) cr->Counts[i] += s.counts[i];
    cr->TotalWords += s.counts[7];
  } else if (m_Alg == "CPACK") {
    // library order = m_PatternLength order (zzzz, xxxx, mmmm, mmxx, zzzx, mmmx); CSV order = CPACKPattern (CPACK.h:18-26)
    static const int kFromLib[6] = {0, 4, 2, 5, 3, 1};
    for (int i = 0; i < 6; i++) { cr->Counts[i] += s.counts[kFromLib[i]]; cr->TotalWords += s.counts[kFromLib[i]]; }
  }
}

unsigned VariantCompressor::CompressLine(std::vector<uint8_t>& dataLine) {
  if (m_Alg == "SC2" || m_Alg == "CPACK") {  // stream-dependent: costed when the stream is complete (GetResult)
    m_Pending.insert(m_Pending.end(), dataLine.begin(), dataLine.end());
    return 0;
  }
  uint16_t size = 0;
  mpc_variant_stats s;
  const int alg = m_Alg == "BDI" ? MPC_ALG_BDI : m_Alg == "FPC" ? MPC_ALG_FPC : MPC_ALG_BPC;
  if (mpc_variant_run_host(alg, 0, dataLine.data(), 1, m_LineSize, &size, &s, nullptr) != MPC_OK) {
    printf("%s: %s\n", m_Alg.c_str(), mpc_variant_error());
    exit(1);
  }
  accumulate(s);
  return size;
}

void VariantCompressor::CompressBatch(const uint8_t* lines, uint64_t nLines) {
  if (m_Alg == "SC2" || m_Alg == "CPACK") {
    m_Pending.insert(m_Pending.end(), lines, lines + nLines * m_LineSize);
    return;
  }
  mpc_variant_stats s;
  float ms = 0;
  const int alg = m_Alg == "BDI" ? MPC_ALG_BDI : m_Alg == "FPC" ? MPC_ALG_FPC : MPC_ALG_BPC;
  if (mpc_variant_run_host(alg, 0, lines, nLines, m_LineSize, nullptr, &s, &ms) != MPC_OK) {
    printf("%s: %s\n", m_Alg.c_str(), mpc_variant_error());
    exit(1);
  }
  m_KernelMs += ms;
  accumulate(s);
}

CompResult* VariantCompressor::GetResult() {
  if (!m_Pending.empty()) {
    mpc_variant_stats s;
    float ms = 0;
    const uint64_t n = m_Pending.size() / m_LineSize;
    int rc = (m_Alg == "SC2") ? mpc_sc2_run_host(0, m_Pending.data(), n, m_LineSize, m_Sampling, nullptr, &s, &ms)
                              : mpc_cpack_run_host(m_Pending.data(), n, m_LineSize, nullptr, &s);
    if (rc != MPC_OK) {
      printf("%s: %s\n", m_Alg.c_str(), m_Alg == "SC2" ? mpc_sc2_error() : "failed");
      exit(1);
    }
    m_KernelMs += ms;
    accumulate(s);
    m_Pending.clear();
    m_Pending.shrink_to_fit();
  }
  return m_Stat;
}

}  // namespace comp
