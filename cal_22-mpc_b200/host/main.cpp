// compressor -- drop-in for the reference CLI (reference src/main.cpp): same flags, same stdout line, same
// output file names and CSV bytes, with the per-line loop (main.cpp:229-244) replaced by chunked batches on
// the GPU.  `bin/run` invokes this binary exactly as it invokes the reference's.
//
//   compressor -a VPC -i FILE.npy -c CFG.json [-o OUTDIR] [--gpus N] [--kernel 0|1|2] [--time]
#include <cassert>
#include <chrono>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "compressor/Pattern.h"
#include "compressor/VPC.h"
#include "compressor/Variants.h"
#include "loader/LoaderGPGPU.h"
#include "loader/LoaderNPY.h"
#include "utils.h"

static const char* kHelp =
    "\nUsage:\n  Compressor [OPTION...]\n\n"
    "  -a, --algorithm arg  Compression algorithm [VPC/FPC/BDI/BPC/CPACK/SC2/PATTERN/VIEWER]. Default=VPC\n"
    "  -i, --input arg      Input path. Supported extensions: .npy (memory dump), .log (GPGPU-Sim trace)\n"
    "  -c, --config arg     Config file path (.json).\n"
    "  -o, --output arg     Output directory path\n"
    "      --gpus arg       Number of GPUs to shard the dump over (default 1)\n"
    "      --kernel arg     0 auto, 1 generic warp-per-block kernel, 2 specialised kernel\n"
    "      --time           Print device and wall time\n"
    "  -h, --help           Print usage\n";

comp::CompResult* compressLines(comp::Compressor* compressor, trace::Loader* loader, unsigned lineSize);
void viewLines(trace::Loader* loader);

int main(int argc, char** argv) {
  std::string algorithm = "VPC", tracePath, configPath, outputDirPath;
  bool haveInput = false, haveConfig = false, help = false, timing = false;
  int gpus = 1, kernel = 0;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    auto val = [&](std::string& dst) { if (i + 1 < argc) dst = argv[++i]; };
    if (a == "-a" || a == "--algorithm") val(algorithm);
    else if (a == "-i" || a == "--input") { val(tracePath); haveInput = true; }
    else if (a == "-c" || a == "--config") { val(configPath); haveConfig = true; }
    else if (a == "-o" || a == "--output") val(outputDirPath);
    else if (a == "--gpus") { std::string v; val(v); gpus = atoi(v.c_str()); }
    else if (a == "--kernel") { std::string v; val(v); kernel = atoi(v.c_str()); }
    else if (a == "--time") timing = true;
    else if (a == "-h" || a == "--help") help = true;
  }
  if (!haveInput) help = true;                        // main.cpp:50-53
  if (algorithm == "VPC" && !haveConfig) help = true;  // main.cpp:54-58
  if (help) {
    std::cout << kHelp << std::endl;  // main.cpp:66-71: help exits 0
    return 0;
  }

  trace::Loader* loader = nullptr;
  if (endsWith(tracePath, ".npy")) {
    trace::LoaderNPY* l = new trace::LoaderNPY(tracePath);
    if (!l->Error().empty()) { printf("%s\n", l->Error().c_str()); return 1; }
    loader = l;
  } else if (endsWith(tracePath, ".log")) {
    trace::gpgpusim::LoaderGPGPU* l = new trace::gpgpusim::LoaderGPGPU(tracePath);  // main.cpp:76-77
    if (!l->Error().empty()) { printf("%s\n", l->Error().c_str()); return 1; }
    loader = l;
  } else {
    // .txt (AXI channel dumps of the reference's apsim loader) is outside the GPU hot path
    printf("Unsupported extension: \"%s\" (this build reads .npy memory dumps and .log GPGPU-Sim traces)\n", tracePath.c_str());
    return 1;
  }
  const unsigned lineSize = loader->GetCachelineSize();  // main.cpp:86

  comp::Compressor* compressor = nullptr;
  if (algorithm == "VPC") {
    comp::VPC* vpc = new comp::VPC(configPath, gpus, kernel);  // main.cpp:88-91
    // The reference ignores the loader's line size for VPC and strides by the config's lineSize (VPC.cpp:101): with a
    // mismatch it reads past the loader's buffer (undefined behaviour).  Refused here, like the other UB configs.
    if ((unsigned)vpc->GetCachelineSize() != lineSize) {
      printf("Line size mismatch: \"%s\" holds %u-byte lines, config \"%s\" has lineSize %d\n", tracePath.c_str(), lineSize,
             configPath.c_str(), vpc->GetCachelineSize());
      return 1;
    }
    compressor = vpc;
  } else if (algorithm == "BDI" || algorithm == "FPC" || algorithm == "BPC" || algorithm == "CPACK" || algorithm == "SC2") {
    compressor = new comp::VariantCompressor(algorithm, lineSize, loader->GetNumLines());  // main.cpp:92-116
  } else if (algorithm == "PATTERN") {
    compressor = new comp::Pattern(lineSize);  // main.cpp:117-120
  } else if (algorithm == "VIEWER") {
    viewLines(loader);  // main.cpp:121-125: host only, no compressor
    return 0;
  } else {
    printf("Invalid name of algorithm: \"%s\" (this build implements VPC, BDI, FPC, BPC, CPACK, SC2, PATTERN, VIEWER)\n", algorithm.c_str());
    return 1;
  }

  // results file names, main.cpp:129-136
  std::string saveFileName = (algorithm == "VPC") ? parseConfig(configPath) : algorithm;
  std::string compOutputSavePath = outputDirPath + "/" + saveFileName + "_results.csv";
  std::string compDetailedOutputSavePath = outputDirPath + "/" + saveFileName + "_results_detail.csv";

  auto t0 = std::chrono::steady_clock::now();
  comp::CompResult* compStat = compressLines(compressor, loader, lineSize);
  auto t1 = std::chrono::steady_clock::now();

  // workload name = <parent dir>_<file stem>, main.cpp:142-157
  std::vector<std::string> parts = splitString(tracePath, "/");
  std::string benchmarkName = parts.size() >= 2 ? parts[parts.size() - 2] : "";
  std::string appName = parts.back();
  replaceAll(appName, ".log", "");
  replaceAll(appName, ".npy", "");
  replaceAll(appName, ".txt", "");
  std::string workloadName = benchmarkName + "_" + appName;
  std::cout << "comp.ratio: " << formatDouble(compStat->CompRatio) << std::endl;  // main.cpp:158
  if (timing && algorithm == "VPC") {
    comp::VPC* v = static_cast<comp::VPC*>(compressor);
    double wall = std::chrono::duration<double>(t1 - t0).count();
    fprintf(stderr, "kernel %s: device %.3f ms, wall %.3f s, %llu blocks of %u B\n", v->KernelName(), v->KernelMs(), wall,
            (unsigned long long)(compStat->OriginalSize / (8ull * lineSize)), lineSize);
  }
  compStat->Print(workloadName, compOutputSavePath);
  compStat->PrintDetail(workloadName, compDetailedOutputSavePath);

  delete loader;
  delete compressor;
  return 0;
}

// main.cpp:208-248 with the GetCacheline/CompressLine pair replaced by GetChunk/CompressBatch
comp::CompResult* compressLines(comp::Compressor* compressor, trace::Loader* loader, unsigned lineSize) {
  uint64_t n = 0;
  // 1. file-backed loaders and compressors that read files themselves: descriptor + byte range, no mapping, no copy
  {
    const char* mode = getenv("MPC_IO");  // "mmap": skip this path; "direct": O_DIRECT reads (cold dumps, no page cache)
    const bool direct = mode && std::string(mode) == "direct";
    if (!(mode && std::string(mode) == "mmap")) {
      uint64_t off = 0;
      const int fd = loader->GetFile(&off, &n, direct);
      if (fd >= 0) {
        if (n == 0 || compressor->CompressFile(fd, off, n, direct)) return compressor->GetResult();
        loader->Reset();  // the compressor does not read files: start over with the in-memory forms
      }
    }
  }
  const uint8_t* all = loader->GetAll(&n);
  if (all) {
    if (n) compressor->CompressBatch(all, n);  // the mapped file is handed over whole; the library chunks it
  } else {
    const uint64_t chunkLines = (256ull << 20) / lineSize;
    std::vector<uint8_t> buf(chunkLines * lineSize);
    while ((n = loader->GetChunk(buf.data(), chunkLines)) != 0) compressor->CompressBatch(buf.data(), n);
  }
  return compressor->GetResult();
}

// main.cpp:250-300: hex dump of the lines compressLines would see; GPGPU-Sim records carry an R:/W: prefix
void viewLines(trace::Loader* loader) {
  if (dynamic_cast<trace::gpgpusim::LoaderGPGPU*>(loader) != nullptr) {
    trace::gpgpusim::MemReqGPU_t req;
    while (1) {
      loader->GetCacheline(&req);
      if (req.isEnd) break;
      if (!(req.reqType == trace::gpgpusim::GLOBAL_ACC_R || req.reqType == trace::gpgpusim::GLOBAL_ACC_W)) continue;
      printf("%s: ", req.reqType == trace::gpgpusim::GLOBAL_ACC_R ? "R" : "W");
      for (size_t i = 0; i < req.data.size(); i++) printf("%02x ", req.data[i]);
      printf("\n");
    }
  } else {
    trace::MemReq_t req;
    while (1) {
      loader->GetCacheline(&req);
      if (req.isEnd) break;
      for (size_t i = 0; i < req.data.size(); i++) printf("%02x ", req.data[i]);
      printf("\n");
    }
  }
}
