#include "VPC.h"

#include <cstdio>
#include <cstdlib>
#include <thread>

namespace comp {

// ---- VPCResult ----------------------------------------------------------------------------------------------
void VPCResult::Fill(const mpc_stats_pod& s) {
  OriginalSize = s.original_bits;
  CompressedSize = s.compressed_bits;
  // the reference recomputes the ratio on every Update (CompResult.h:34); with no line it stays 0
  CompRatio = s.blocks ? (double)OriginalSize / (double)CompressedSize : 0;
  const uint64_t lineBits = (uint64_t)LineSize * BYTE;
  for (int i = -1; i < m_NumModules; i++) {
    const int k = i + 1;
    ClusterStat& cs = m_ClusterStats[i];
    cs.count = s.count[k];
    cs.originalSize = s.count[k] * lineBits;
    cs.compressedSize = s.comp_bits[k];
    cs.compRatio = cs.count ? (double)cs.originalSize / (double)cs.compressedSize : 0;  // VPC.h:55
    for (int b = 0; b < MPC_HIST_BINS; b++)
      if (s.hist[k][b] || b < COMPSIZELIMIT) cs.compSizeHistogram[b] = s.hist[k][b];
    // running means of per-line MAE / MSE (VPC.h:62-76); exact for power-of-two line sizes when evaluated as
    // (sum / L) / lines -- SURVEY.md section 7
    m_NumLines[i] = s.res_lines[k];
    m_MAE[i] = s.res_lines[k] ? ((double)s.res_abs[k] / (double)LineSize) / (double)s.res_lines[k] : 0;
    m_MSE[i] = s.res_lines[k] ? ((double)s.res_sq[k] / (double)LineSize) / (double)s.res_lines[k] : 0;
  }
}

static std::ostream* openAppend(std::ofstream& file, const std::string& path, const std::string& header) {
  if (path == "") return &std::cout;
  if (!isFileExists(path)) {
    file.open(path);
    if (!file.is_open()) {
      std::cout << "File is not open: \"" << path << "\"" << std::endl;  // VPC.h:92-97
      exit(1);
    }
    file << header;
    file.close();
  }
  file.open(path, std::ios_base::app);
  return &file;
}

void VPCResult::Print(std::string workloadName, std::string filePath) {
  std::string header = "workload,total,,,";
  for (int i = -1; i < m_NumModules; i++) header += std::to_string(i) + ",,,";
  header += "\n,original_size,compressed_size,compression_ratio,count,";  // the extra "count," column is the reference's
  for (int i = -1; i < m_NumModules; i++) header += "original_size,compressed_size,compression_ratio,";
  header += "\n";
  std::ofstream file;
  std::ostream& os = *openAppend(file, filePath, header);
  os << workloadName << "," << OriginalSize << "," << CompressedSize << "," << formatDouble(CompRatio) << ",";
  for (int i = -1; i < m_NumModules; i++) {
    ClusterStat& cs = m_ClusterStats[i];
    os << cs.originalSize << "," << cs.compressedSize << "," << formatDouble(cs.compRatio) << ",";
  }
  os << std::endl;
}

void VPCResult::PrintDetail(std::string workloadName, std::string filePath) {
  std::string header = "workload,";
  for (int i = -1; i < m_NumModules; i++) header += std::to_string(i) + ",,";
  for (int i = 0; i < m_NumModules; i++) header += std::to_string(i) + "," + std::string(COMPSIZELIMIT - 1, ',');
  header += "\n,";
  for (int i = -1; i < m_NumModules; i++) header += "mae,mse,";
  for (int i = 0; i < m_NumModules; i++)
    for (int j = 0; j < COMPSIZELIMIT; j++) header += std::to_string(j) + ",";
  header += "\n";
  std::ofstream file;
  std::ostream& os = *openAppend(file, filePath, header);
  os << workloadName << ",";
  for (int i = -1; i < m_NumModules; i++) os << formatDouble(m_MAE[i]) << "," << formatDouble(m_MSE[i]) << ",";
  for (int i = 0; i < m_NumModules; i++) {
    ClusterStat& cs = m_ClusterStats[i];
    for (int j = 0; j < COMPSIZELIMIT; j++) os << cs.compSizeHistogram[j] << ",";  // sizes >= 288 are not printed
  }
  os << std::endl;
}

// ---- VPC ------------------------------------------------------------------------------------------------------
void VPC::die(const char* what, mpc_ctx* ctx) {
  printf("%s: %s\n", what, ctx ? mpc_last_error(ctx) : mpc_global_error());
  exit(1);
}

VPC::VPC(std::string configPath, int numGpus, int kernel) {
  char err[1024] = {0};
  int rc = mpc_config_from_json_file(configPath.c_str(), &m_Cfg, err, sizeof(err));
  if (rc != MPC_OK) {  // the reference prints the message and exits 1 (VPC.cpp:77-81, 90-95, 180-184, 295-299, 320-324)
    printf("%s\n", err);
    exit(1);
  }
  if (numGpus < 1) numGpus = 1;
  for (int d = 0; d < numGpus; d++) {
    mpc_ctx* c = nullptr;
    if (mpc_create(&m_Cfg, d, &c) != MPC_OK) die("mpc_create", nullptr);
    if (kernel && mpc_set_kernel(c, kernel) != MPC_OK) die("mpc_set_kernel", c);
    if (mpc_prepare_host(c) != MPC_OK) die("mpc_prepare_host", c);  // pin the staging ring while the loader parses its header
    m_Ctx.push_back(c);
  }
  // communicators are created once per compressor, not per GetResult()
  if (numGpus > 1 && mpc_comm_init_all(m_Ctx.data(), numGpus) != MPC_OK) die("mpc_comm_init_all", m_Ctx[0]);
  m_Stat = new VPCResult((unsigned)m_Cfg.line_size, m_Cfg.num_modules);
  m_Stat->CompressorName = "Contrastive Clustering Compressor";  // VPC.h:248
}

VPC::~VPC() {
  for (mpc_ctx* c : m_Ctx) mpc_destroy(c);
  delete m_Stat;
}

const char* VPC::KernelName() const { return mpc_kernel_name(m_Ctx[0]); }

unsigned VPC::CompressLine(std::vector<uint8_t>& dataLine) {
  if ((int)dataLine.size() != m_Cfg.line_size) {
    printf("CompressLine: line of %zu bytes, config lineSize is %d\n", dataLine.size(), m_Cfg.line_size);
    exit(1);
  }
  uint16_t packed = 0;
  if (mpc_submit_host(m_Ctx[0], dataLine.data(), 1, &packed) != MPC_OK || mpc_sync(m_Ctx[0]) != MPC_OK)
    die("CompressLine", m_Ctx[0]);
  return packed & 0x7FFu;
}

void VPC::CompressBatch(const uint8_t* lines, uint64_t nLines) {
  const size_t G = m_Ctx.size();
  const uint64_t per = (nLines + G - 1) / G;  // GPU g takes [g*per, min(n, (g+1)*per))
  std::vector<std::thread> th;
  for (size_t g = 0; g < G; g++) {
    const uint64_t lo = std::min<uint64_t>(nLines, g * per), hi = std::min<uint64_t>(nLines, (g + 1) * per);
    if (hi == lo) continue;
    auto work = [this, g, lo, hi, lines]() {
      if (mpc_submit_host(m_Ctx[g], lines + lo * (uint64_t)m_Cfg.line_size, hi - lo, nullptr) != MPC_OK ||
          mpc_sync(m_Ctx[g]) != MPC_OK)
        die("CompressBatch", m_Ctx[g]);
    };
    if (G == 1) work(); else th.emplace_back(work);
  }
  for (auto& t : th) t.join();
  collectTiming();
}

void VPC::collectTiming() {
  const size_t G = m_Ctx.size();
  for (mpc_ctx* c : m_Ctx) {
    float ms = 0;
    int launches = 0;
    mpc_last_timing(c, &ms, &launches);
    if (ms > m_KernelMs || G == 1) m_KernelMs = (G == 1) ? m_KernelMs + ms : ms;
  }
}

// The file-backed hand-over: every GPU reads its contiguous shard of the byte range itself (mpc_submit_file), so the dump is
// neither mapped nor copied through an intermediate buffer on its way to the pinned staging ring.
bool VPC::CompressFile(int fd, uint64_t offset, uint64_t nLines, bool directIo) {
  const size_t G = m_Ctx.size();
  const uint64_t per = (nLines + G - 1) / G;
  std::vector<std::thread> th;
  for (size_t g = 0; g < G; g++) {
    const uint64_t lo = std::min<uint64_t>(nLines, g * per), hi = std::min<uint64_t>(nLines, (g + 1) * per);
    if (hi == lo) continue;
    auto work = [this, g, lo, hi, fd, offset, directIo]() {
      if (mpc_submit_file(m_Ctx[g], fd, offset + lo * (uint64_t)m_Cfg.line_size, hi - lo, nullptr, directIo ? 1 : 0) != MPC_OK ||
          mpc_sync(m_Ctx[g]) != MPC_OK)
        die("CompressFile", m_Ctx[g]);
    };
    if (G == 1) work(); else th.emplace_back(work);
  }
  for (auto& t : th) t.join();
  collectTiming();
  return true;
}

CompResult* VPC::GetResult() {
  // one exchange for the whole run: the library all-reduces (sum, uint64) a copy of the statistics
  // vectors over the communicators created in the constructor; with one GPU this is mpc_finish
  mpc_stats_pod* pod = new mpc_stats_pod;
  if (mpc_finish_allreduce(m_Ctx.data(), (int)m_Ctx.size(), pod) != MPC_OK) die("mpc_finish_allreduce", m_Ctx[0]);
  static_cast<VPCResult*>(m_Stat)->Fill(*pod);
  delete pod;
  return m_Stat;
}

}  // namespace comp
