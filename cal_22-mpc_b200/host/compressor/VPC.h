// comp::VPC (= MPC) and comp::VPCResult -- host-side mirror of reference src/compressor/VPC.h.
// The compressor owns one libmpc_b200 context per GPU; CompressLine / CompressBatch forward to the C ABI,
// GetResult() pulls the statistics vector back (all-reduced over NCCL when several GPUs are used) and
// fills the same fields the reference's VPCResult carries, so Print / PrintDetail write identical CSV files.
#ifndef MPCB_VPC_H_
#define MPCB_VPC_H_

#include <map>
#include <string>
#include <vector>

#include "Compressor.h"
#include "mpc_capi.h"

namespace comp {

struct ClusterStat {  // VPC.h:16-32
  uint64_t count = 0;
  uint64_t originalSize = 0;
  uint64_t compressedSize = 0;
  double compRatio = 0;
  std::map<int, uint64_t> compSizeHistogram;
};

struct VPCResult : public CompResult {
  VPCResult(unsigned lineSize, int numModules) : CompResult(lineSize), m_NumModules(numModules) {
    for (int i = -1; i < numModules; i++) {  // VPC.h:213-229
      ClusterStat st;
      for (int s = 0; s < COMPSIZELIMIT; s++) st.compSizeHistogram[s] = 0;
      m_ClusterStats[i] = st;
      m_MAE[i] = 0;
      m_MSE[i] = 0;
      m_NumLines[i] = 0;
    }
  }
  void Fill(const mpc_stats_pod& s);
  void Print(std::string workloadName = "", std::string filePath = "") override;        // VPC.h:78-135
  void PrintDetail(std::string workloadName = "", std::string filePath = "") override;  // VPC.h:137-211

  std::map<int, ClusterStat> m_ClusterStats;
  std::map<int, double> m_MAE, m_MSE;
  std::map<int, uint64_t> m_NumLines;
  int m_NumModules;
};

class VPC : public Compressor {
 public:
  // numGpus > 1: the batch is sharded contiguously over devices 0..numGpus-1
  explicit VPC(std::string configPath, int numGpus = 1, int kernel = 0);
  ~VPC() override;
  int GetCachelineSize() { return m_Cfg.line_size; }
  int GetNumModules() { return m_Cfg.num_modules; }
  int GetNumClusters() { return m_Cfg.num_modules + 1; }
  unsigned CompressLine(std::vector<uint8_t>& dataLine) override;       // VPC.cpp:22-25
  void CompressBatch(const uint8_t* lines, uint64_t nLines) override;  // replaces main.cpp:237-243
  bool CompressFile(int fd, uint64_t offset, uint64_t nLines, bool directIo) override;  // pread -> pinned ring -> H2D -> kernel
  CompResult* GetResult() override;
  double KernelMs() const { return m_KernelMs; }
  const char* KernelName() const;

 private:
  void die(const char* what, mpc_ctx* ctx);
  void collectTiming();
  mpc_config_pod m_Cfg;
  std::vector<mpc_ctx*> m_Ctx;
  double m_KernelMs = 0;
};

}  // namespace comp
#endif
