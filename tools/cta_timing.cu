// Debug harness (GPU box): per-CTA timeline of the specialised F4 kernel -- when each CTA starts, leaves the prologue,
// leaves the tile loop and ends -- to separate launch ramp, prologue, tail imbalance and statistics flush.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DMPC_CTA_TIMING -Iinclude -Ical_22-mpc_b200/csrc \
//        tools/cta_timing.cu cal_22-mpc_b200/csrc/mpc_spec_registry.o cal_22-mpc_b200/csrc/mpc_synth.o -o tools/cta_timing
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../cal_22-mpc_b200/csrc/spec/spec_F4.cu"
#include "../cal_22-mpc_b200/csrc/mpc_internal.h"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

int main(int argc, char** argv) {
  const uint64_t n = argc > 1 ? strtoull(argv[1], 0, 10) : (1ull << 23);
  const int kind = argc > 2 ? atoi(argv[2]) : 2;
  uint8_t* d; uint64_t* st; uint8_t* lut; uint32_t* sched;
  CK(cudaMalloc(&d, n * 128)); CK(cudaMalloc(&st, mpc::kStatsWords * 8)); CK(cudaMalloc(&lut, 65536));
  CK(cudaMemset(st, 0, mpc::kStatsWords * 8));
  CK(cudaMalloc(&sched, 128)); CK(cudaMemset(sched, 0, 128));
  std::vector<uint8_t> h(65536);
  mpc::build_row_cost_lut(h.data(), mpc::spec_F4::Cfg::kLutXor);
  CK(cudaMemcpy(lut, h.data(), 65536, cudaMemcpyHostToDevice));
  CK(mpc::launch_synth(d, 0, n, n, kind, 2024, 0));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  mpc_config_pod* pod = new mpc_config_pod();
  for (int it = 0; it < 4; it++) {
    cudaEventRecord(e0);
    CK(mpc::kSpec_F4.launch(*pod, d, n, nullptr, st, lut, sched, p.multiProcessorCount, 0));
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<unsigned long long> t(4 * 1024);
    CK(cudaMemcpyFromSymbol(t.data(), mpc::spec::g_cta_t, sizeof(unsigned long long) * 4 * 1024));
    const int G = p.multiProcessorCount;
    unsigned long long t0 = ~0ull, tend = 0;
    for (int b = 0; b < G; b++) { t0 = std::min(t0, t[4 * b]); tend = std::max(tend, t[4 * b + 3]); }
    std::vector<double> s0, s1, s2, s3;
    for (int b = 0; b < G; b++) { s0.push_back((t[4*b]-t0)*1e-3); s1.push_back((t[4*b+1]-t0)*1e-3); s2.push_back((t[4*b+2]-t0)*1e-3); s3.push_back((t[4*b+3]-t0)*1e-3); }
    auto stat = [&](const char* name, std::vector<double> v) { std::sort(v.begin(), v.end());
      printf("  %-10s min %8.2f  p10 %8.2f  med %8.2f  p90 %8.2f  max %8.2f us\n", name, v[0], v[v.size()/10], v[v.size()/2], v[v.size()*9/10], v.back()); };
    printf("iter %d: event %.2f us, first start -> last end %.2f us, %.1f GB/s\n", it, ms * 1e3, (tend - t0) * 1e-3, n * 128 / (ms * 1e-3) / 1e9);
    stat("start", s0); stat("prologue", s1); stat("loop end", s2); stat("cta end", s3);
  }
  return 0;
}
namespace mpc {  // the registry object file references the list; this harness links one kernel only
const SpecKernel* const kSpecKernels[] = {&kSpec_F4};
const int kNumSpecKernels = 1;
}
