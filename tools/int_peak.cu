// Integer-pipe micro-benchmark for roofline B ("micro-benchmark LOP3/IADD3 issue rate").
// Each kernel runs 8 independent dependency chains per thread of ONE instruction kind (inline PTX, so ptxas cannot fold
// them), 4 warps per scheduler resident (enough to cover the 4-cycle dependent-issue latency with 8 chains), and reports
// thread-level operations per second over the whole GPU and warp instructions per clock per SM.  "mix" alternates an
// ALU-pipe op (LOP3) with an FMA-pipe op (IMAD): the two pipes issue from one scheduler port, so the mix shows the
// combined ceiling the MPC kernels (LOP3/PRMT/VIMNMX on ALU, IMAD/IDP on FMA) can reach.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/int_peak tools/int_peak.cu ; prints one JSON line
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int kChains = 8;
constexpr int kInner = 64;     // instructions per chain per loop iteration
constexpr int kThreads = 512;  // 16 warps = 4 per scheduler

enum { OP_LOP3 = 0, OP_IADD3 = 1, OP_IMAD = 2, OP_PRMT = 3, OP_SHF = 4, OP_VIMNMX = 5, OP_IDP = 6, OP_MIX = 7, OP_COUNT = 8 };

template <int OP>
__device__ __forceinline__ void step(unsigned& a, unsigned b, unsigned c, int i) {
  if (OP == OP_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
  // ptxas fuses two dependent adds / mins into ONE three-input IADD3 / VIMNMX3: written as the pair, counted as one instruction
  else if (OP == OP_IADD3) asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(a) : "r"(b), "r"(c));
  else if (OP == OP_IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
  else if (OP == OP_PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x2103;" : "+r"(a) : "r"(b));
  else if (OP == OP_SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a) : "r"(b));
  else if (OP == OP_VIMNMX) asm volatile("min.u16x2 %0, %0, %1;\n\tmin.u16x2 %0, %0, %2;" : "+r"(a) : "r"(b), "r"(c));
  else if (OP == OP_IDP) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(a) : "r"(b), "r"(c));
  else {  // OP_MIX
    if (i & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
  }
}

template <int OP>
__global__ void __launch_bounds__(kThreads) peak_kernel(unsigned* out, int iters, unsigned seed) {
  unsigned v[kChains];
#pragma unroll
  for (int k = 0; k < kChains; k++) v[k] = seed + threadIdx.x * 31u + k;
  const unsigned b = seed | 3u, c = seed * 7u + 1u;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < kInner; i++) {
#pragma unroll
      for (int k = 0; k < kChains; k++) step<OP>(v[k], b, c, i);
    }
  }
  unsigned s = 0;
#pragma unroll
  for (int k = 0; k < kChains; k++) s ^= v[k];
  if (s == 0x12345u) out[blockIdx.x * kThreads + threadIdx.x] = s;  // never true in practice: keeps the chains alive
}

template <int OP>
double run(unsigned* d_out, int sms, int iters, double* ms_out) {
  const int grid = sms * 1;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  peak_kernel<OP><<<grid, kThreads>>>(d_out, iters / 8, 12345u);  // warm-up
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    CK(cudaEventRecord(e0));
    peak_kernel<OP><<<grid, kThreads>>>(d_out, iters, 12345u + rep);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  *ms_out = best;
  const double thread_ops = (double)grid * kThreads * (double)iters * kInner * kChains;
  return thread_ops / (best * 1e-3);
}

int main(int argc, char** argv) {
  int dev = argc > 1 ? atoi(argv[1]) : 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, dev));
  int clock_khz = 0;
  CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, dev));
  unsigned* d_out;
  CK(cudaMalloc(&d_out, (size_t)p.multiProcessorCount * kThreads * sizeof(unsigned)));
  const int iters = 4000;
  const char* names[OP_COUNT] = {"lop3", "iadd3", "imad", "prmt", "shf", "vimnmx_u16x2", "idp4a", "mix_lop3_imad"};
  double ops[OP_COUNT], ms[OP_COUNT];
  ops[0] = run<OP_LOP3>(d_out, p.multiProcessorCount, iters, &ms[0]);
  ops[1] = run<OP_IADD3>(d_out, p.multiProcessorCount, iters, &ms[1]);
  ops[2] = run<OP_IMAD>(d_out, p.multiProcessorCount, iters, &ms[2]);
  ops[3] = run<OP_PRMT>(d_out, p.multiProcessorCount, iters, &ms[3]);
  ops[4] = run<OP_SHF>(d_out, p.multiProcessorCount, iters, &ms[4]);
  ops[5] = run<OP_VIMNMX>(d_out, p.multiProcessorCount, iters, &ms[5]);
  ops[6] = run<OP_IDP>(d_out, p.multiProcessorCount, iters, &ms[6]);
  ops[7] = run<OP_MIX>(d_out, p.multiProcessorCount, iters, &ms[7]);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_mhz_nominal\": %.0f, \"threads_per_sm\": %d", p.name, p.multiProcessorCount,
         clock_khz / 1000.0, kThreads);
  for (int i = 0; i < OP_COUNT; i++) {
    // warp instructions per clock per SM at the nominal clock (the clock under load is sampled by bench.py)
    const double wipc = ops[i] / 32.0 / p.multiProcessorCount / (clock_khz * 1e3);
    printf(", \"%s\": {\"tops\": %.3f, \"warp_inst_per_clk_per_sm\": %.3f, \"ms\": %.3f}", names[i], ops[i] / 1e12, wipc, ms[i]);
  }
  printf("}\n");
  cudaFree(d_out);
  return 0;
}
