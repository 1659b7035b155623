// NVRTC path of the specialised kernel: mpc_specgen.cpp writes the schedule for the context's config, NVRTC compiles it
// for sm_100a against the same headers the ahead-of-time build uses (embedded in the library as strings), the cubin
// is loaded through the driver API (entry points fetched with cudaGetDriverEntryPoint, so the library does not link
// libcuda and still loads on a machine without a driver).
#include "mpc_jit.h"

#include <cuda.h>
#include <nvrtc.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
#include <fstream>
#include <mutex>

#include "mpc_embedded_headers.inc"

namespace mpc {

struct JitKernel {
  CUmodule module = nullptr;
  CUfunction fn = nullptr;
  SpecTraits traits;
  int per_sm = 1;
};

namespace {

struct Driver {
  CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
  CUresult (*ModuleUnload)(CUmodule) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
  CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t) = nullptr;
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
  CUresult (*LaunchKernelEx)(const CUlaunchConfig*, CUfunction, void**, void**) = nullptr;  // optional: programmatic dependent launch
  CUresult (*TensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill) = nullptr;
  bool ok = false;
};

template <class F>
bool entry(const char* name, F* out) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || !p) return false;
  *out = reinterpret_cast<F>(p);
  return true;
}

Driver& driver() {
  static Driver d;
  static std::once_flag once;
  std::call_once(once, [] {
    d.ok = entry("cuModuleLoadData", &d.ModuleLoadData) && entry("cuModuleUnload", &d.ModuleUnload) &&
           entry("cuModuleGetFunction", &d.ModuleGetFunction) && entry("cuFuncSetAttribute", &d.FuncSetAttribute) &&
           entry("cuOccupancyMaxActiveBlocksPerMultiprocessor", &d.OccupancyMaxActiveBlocksPerMultiprocessor) &&
           entry("cuLaunchKernel", &d.LaunchKernel);
    entry("cuLaunchKernelEx", &d.LaunchKernelEx);
    entry("cuTensorMapEncodeTiled", &d.TensorMapEncodeTiled);
  });
  return d;
}

uint64_t fnv1a(const std::string& s) {
  uint64_t h = 1469598103934665603ull;
  for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
  return h;
}

}  // namespace

int jit_compile(const mpc_config_pod& cfg, std::vector<char>* cubin, SpecTraits* traits, std::string* log) {
  std::string why;
  const std::string src = generate_spec_source(cfg, "jit", true, &why);
  if (src.empty()) {
    if (log) *log = "config is not eligible for the specialised kernel: " + why;
    return MPC_E_CONFIG;
  }
  if (traits) *traits = spec_traits(cfg);
  // disk cache (optional): key = hash of the generated source and of the headers it is compiled against
  std::string cache_path;
  if (const char* dir = getenv("MPC_JIT_CACHE_DIR")) {
    char name[64];
    int nv_major = 0, nv_minor = 0;
    nvrtcVersion(&nv_major, &nv_minor);  // a cubin built by another compiler version is another entry
    snprintf(name, sizeof(name), "/mpc_b200_%016llx.cubin",
             (unsigned long long)(fnv1a(src) ^ (fnv1a(kHdr_mpc_spec_cuh) * 3) ^ (fnv1a(kHdr_mpc_device_cuh) * 5) ^
                                  (fnv1a(kHdr_mpc_layout_h) * 7) ^ ((uint64_t)(nv_major * 1000 + nv_minor) * 0x9E3779B97F4A7C15ull)));
    cache_path = std::string(dir) + name;
    std::ifstream in(cache_path, std::ios::binary);
    if (in) {
      cubin->assign(std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>());
      // entries are published with rename() (below), so a file that exists is complete; the ELF magic guards against
      // anything else that may sit under this name
      if (cubin->size() > 64 && memcmp(cubin->data(), "\x7f" "ELF", 4) == 0) return MPC_OK;
      cubin->clear();
    }
  }
  const char* header_names[] = {"mpc_spec.cuh", "mpc_device.cuh", "mpc_layout.h"};
  const char* header_src[] = {kHdr_mpc_spec_cuh, kHdr_mpc_device_cuh, kHdr_mpc_layout_h};
  nvrtcProgram prog;
  if (nvrtcCreateProgram(&prog, src.c_str(), "mpc_jit.cu", 3, header_src, header_names) != NVRTC_SUCCESS) {
    if (log) *log = "nvrtcCreateProgram failed";
    return MPC_E_STATE;
  }
  const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo"};
  const nvrtcResult rc = nvrtcCompileProgram(prog, 3, opts);
  size_t log_size = 0;
  nvrtcGetProgramLogSize(prog, &log_size);
  if (log && log_size > 1) {
    log->resize(log_size);
    nvrtcGetProgramLog(prog, &(*log)[0]);
  }
  if (rc != NVRTC_SUCCESS) {
    nvrtcDestroyProgram(&prog);
    if (log) *log = std::string("NVRTC: ") + nvrtcGetErrorString(rc) + "\n" + *log;
    return MPC_E_STATE;
  }
  size_t n = 0;
  nvrtcGetCUBINSize(prog, &n);
  cubin->resize(n);
  nvrtcGetCUBIN(prog, cubin->data());
  nvrtcDestroyProgram(&prog);
  if (!cache_path.empty()) {
    // several ranks may build the same config at once: write to a private name, then publish atomically
    const std::string tmp = cache_path + ".tmp." + std::to_string((long long)getpid());
    bool ok = false;
    {
      std::ofstream out(tmp, std::ios::binary);
      out.write(cubin->data(), (std::streamsize)cubin->size());
      ok = out.good();
    }
    if (!ok || rename(tmp.c_str(), cache_path.c_str()) != 0) remove(tmp.c_str());
  }
  return MPC_OK;
}

JitKernel* jit_create(const mpc_config_pod& cfg, std::string* log) {
  Driver& d = driver();
  if (!d.ok) {
    if (log) *log = "CUDA driver entry points unavailable";
    return nullptr;
  }
  std::vector<char> cubin;
  JitKernel* k = new JitKernel;
  if (jit_compile(cfg, &cubin, &k->traits, log) != MPC_OK) { delete k; return nullptr; }
  cudaFree(0);  // make sure the primary context is current
  CUresult r = d.ModuleLoadData(&k->module, cubin.data());
  if (r == CUDA_SUCCESS) r = d.ModuleGetFunction(&k->fn, k->module, "mpc_jit_kernel");
  if (r == CUDA_SUCCESS) r = d.FuncSetAttribute(k->fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)k->traits.smem_bytes);
  if (r == CUDA_SUCCESS) r = d.OccupancyMaxActiveBlocksPerMultiprocessor(&k->per_sm, k->fn, k->traits.warps * 32, k->traits.smem_bytes);
  if (r != CUDA_SUCCESS) {
    if (log) *log = "loading the NVRTC-built kernel failed (CUresult " + std::to_string((int)r) + ")";
    jit_destroy(k);
    return nullptr;
  }
  if (k->per_sm < 1) k->per_sm = 1;
  return k;
}

void jit_destroy(JitKernel* k) {
  if (!k) return;
  if (k->module && driver().ok) driver().ModuleUnload(k->module);
  delete k;
}

cudaError_t make_tile_tmap(TileTmap* out, const uint8_t* d_lines, uint64_t n_blocks) {
  static_assert(sizeof(TileTmap) == sizeof(CUtensorMap), "TileTmap must be a CUtensorMap");
  if (!driver().TensorMapEncodeTiled) return cudaErrorNotSupported;
  const cuuint64_t dims[2] = {128, n_blocks};          // innermost first: 128 bytes per block, n_blocks rows
  const cuuint64_t strides[1] = {128};                 // bytes between rows
  const cuuint32_t box[2] = {128, 32};                 // one tile: 32 blocks
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = driver().TensorMapEncodeTiled(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_UINT8, 2,
                                                   const_cast<uint8_t*>(d_lines), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

int jit_lut_xor(const JitKernel* k) { return k->traits.use_lut ? k->traits.lut_xor : 0; }

cudaError_t jit_launch(JitKernel* k, const uint8_t* d_lines, uint64_t n_blocks, uint16_t* d_packed, uint64_t* d_stats,
                       const uint8_t* d_row_lut, uint32_t* d_sched, int sm_count, cudaStream_t stream) {
  if (n_blocks == 0) return cudaSuccess;
  if (n_blocks > 0xffffff00ull) return cudaErrorInvalidValue;  // 32-bit block indices in the kernel: 512 GiB per launch
  const uint64_t tile_lines = 32ull * (uint64_t)(128 / k->traits.line_size);
  const uint64_t n_tiles = (n_blocks + tile_lines - 1) / tile_lines;
  uint64_t grid = (uint64_t)sm_count * k->per_sm;
  const uint64_t want = (n_tiles + k->traits.warps - 1) / k->traits.warps;
  if (grid > want) grid = want;
  unsigned long long n = n_blocks;
  const char* e = getenv("MPC_SCHED_STATIC_EIGHTHS");
  int eighths = e ? atoi(e) : 7;
  eighths = eighths < 0 ? 0 : (eighths > 8 ? 8 : eighths);
  const uint64_t total_warps = grid * (uint64_t)k->traits.warps;
  unsigned int static_rounds = eighths == 8 ? (unsigned int)((n_tiles + total_warps - 1) / total_warps)
                                            : (unsigned int)((n_tiles / total_warps) * (uint64_t)eighths / 8);
  TileTmap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (k->traits.tma) {
    if (n_blocks >= (1ull << 31)) return cudaErrorInvalidValue;  // tensor coordinates are signed 32-bit
    const cudaError_t te = make_tile_tmap(&tmap, d_lines, n_blocks);
    if (te != cudaSuccess) return te;
  }
  void* args[] = {(void*)&d_lines, (void*)&n, (void*)&d_packed, (void*)&d_stats, (void*)&d_row_lut, (void*)&d_sched, (void*)&static_rounds,
                  (void*)&tmap};
  CUresult r;
  const char* pdl = getenv("MPC_PDL");
  if (driver().LaunchKernelEx && !(pdl && pdl[0] == '0')) {
    // same launch attribute as launch_spec (mpc_spec.cuh): the prologue may overlap the previous kernel's tail
    CUlaunchConfig lc;
    memset(&lc, 0, sizeof(lc));
    lc.gridDimX = (unsigned)grid; lc.gridDimY = 1; lc.gridDimZ = 1;
    lc.blockDimX = (unsigned)(k->traits.warps * 32); lc.blockDimY = 1; lc.blockDimZ = 1;
    lc.sharedMemBytes = (unsigned)k->traits.smem_bytes;
    lc.hStream = (CUstream)stream;
    CUlaunchAttribute attr;
    memset(&attr, 0, sizeof(attr));
    attr.id = CU_LAUNCH_ATTRIBUTE_PROGRAMMATIC_STREAM_SERIALIZATION;
    attr.value.programmaticStreamSerializationAllowed = 1;
    lc.attrs = &attr;
    lc.numAttrs = 1;
    r = driver().LaunchKernelEx(&lc, k->fn, args, nullptr);
  } else {
    r = driver().LaunchKernel(k->fn, (unsigned)grid, 1, 1, (unsigned)(k->traits.warps * 32), 1, 1,
                              (unsigned)k->traits.smem_bytes, (CUstream)stream, args, nullptr);
  }
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorLaunchFailure;
}

}  // namespace mpc
