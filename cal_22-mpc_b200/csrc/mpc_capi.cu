// C-ABI of libmpc_b200 (include/mpc_capi.h): context, streams, the chunked host->device loader and
// the statistics vector.  Replaces the per-line driver loop of the reference
// (src/main.cpp:229-244: GetCacheline -> CompressLine -> VPCResult::Update) with batched launches.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types only: the library resolves the NCCL entry points with dlopen when a communicator is first used
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "mpc_capi.h"
#include "mpc_internal.h"
#include "mpc_jit.h"
#include "mpc_spec.h"

namespace {

thread_local std::string g_error;

constexpr int kHostStages = 3;  // pinned/device staging ring of mpc_submit_host / mpc_submit_file: fill, copy and compute overlap
constexpr size_t kStagePad = 4096;  // room for the unaligned head of an O_DIRECT read

struct Stage {             // one slot of the staging ring used by mpc_submit_host / mpc_submit_file
  cudaStream_t stream = nullptr;
  uint8_t* h_pinned = nullptr;
  uint8_t* d_lines = nullptr;
  uint16_t* d_packed = nullptr;
  uint16_t* h_packed = nullptr;  // pinned
  cudaEvent_t k_start = nullptr, k_stop = nullptr, done = nullptr;
  bool busy = false;
  // pending result copy (pinned -> user buffer) performed on the host once `done` fires
  uint16_t* user_packed = nullptr;
  uint64_t user_count = 0;
};

}  // namespace

struct mpc_ctx {
  mpc_config_pod cfg;
  int device = 0;
  int sm_count = 148;
  mpc::GenericParams gparams;
  mpc::GenericModule* d_gmods = nullptr;
  const mpc::SpecKernel* spec = nullptr;  // specialised kernel matching cfg, if any
  uint8_t* d_row_lut = nullptr;           // row-cost table used by the specialised kernels
  mpc::JitKernel* jit = nullptr;          // specialised kernel built at run time (NVRTC) when none is compiled in
  int kernel_choice = 0;
  uint64_t* d_stats = nullptr;
  uint32_t* d_sched = nullptr;            // tile-scheduler words of the specialised kernels: 1 + kHostStages slots (device stream, staging ring) x 32 words
  cudaStream_t own_stream = nullptr;      // created by mpc_create
  cudaStream_t stream = nullptr;          // stream used by mpc_submit_device / synth / stats (own or caller's)
  cudaEvent_t ev_start = nullptr, ev_stop = nullptr, ev_order = nullptr;
  Stage stage[kHostStages];
  uint64_t chunk_blocks = 0;
  float last_ms = 0.f;
  int last_launches = 0;
  bool last_timing_pending = false;       // ev_start/ev_stop recorded but not read yet
  bool timing_enabled = true;             // record the two events around mpc_submit_device launches
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending_host_events;
  std::string error;
  std::string kernel_name;
  std::string jit_note;  // why no specialised kernel exists for this context, if that is the case
  // multi-GPU: the job's one collective is an all-reduce of the statistics vector
  ncclComm_t comm = nullptr;
  bool comm_owned = false;                // created by mpc_comm_init_rank / mpc_comm_init_all: destroyed with the context
  uint64_t* d_reduced = nullptr;          // all-reduced copy of d_stats (d_stats itself keeps accumulating locally)
  bool stages_ready = false;              // every buffer of both mpc_submit_host stages exists
};

namespace {

int fail(mpc_ctx* ctx, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->error = buf;
  g_error = buf;
  return code;
}

#define MPC_CUDA(ctx, call)                                                                           \
  do {                                                                                                \
    cudaError_t e__ = (call);                                                                         \
    if (e__ != cudaSuccess) return fail(ctx, MPC_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__)); \
  } while (0)

bool use_spec(const mpc_ctx* ctx) {
  if (ctx->kernel_choice == 1) return false;
  return ctx->spec != nullptr || ctx->jit != nullptr;
}

void refresh_kernel_name(mpc_ctx* ctx) {
  if (!use_spec(ctx)) ctx->kernel_name = "generic_warp";
  else ctx->kernel_name = std::string("spec_thread:") + (ctx->spec ? ctx->spec->name : "jit");
}

// `slot` names the scheduler words the launch uses: launches that may overlap (the two stages of mpc_submit_host) must
// not share them; launches on one stream run one after the other and the kernel leaves its slot zeroed.
int launch(mpc_ctx* ctx, const uint8_t* d_lines, uint64_t n, uint16_t* d_packed, cudaStream_t s, int slot) {
  cudaError_t e;
  uint32_t* sched = ctx->d_sched + 32 * slot;
  if (use_spec(ctx) && ctx->spec)
    e = ctx->spec->launch(ctx->cfg, d_lines, n, d_packed, ctx->d_stats, ctx->d_row_lut, sched, ctx->sm_count, s);
  else if (use_spec(ctx))
    e = mpc::jit_launch(ctx->jit, d_lines, n, d_packed, ctx->d_stats, ctx->d_row_lut, sched, ctx->sm_count, s);
  else
    e = mpc::launch_generic(ctx->gparams, ctx->d_gmods, d_lines, n, d_packed, ctx->d_stats, ctx->sm_count, s);
  if (e != cudaSuccess) return fail(ctx, MPC_E_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
  return MPC_OK;
}

int drain_stage(mpc_ctx* ctx, Stage& st) {
  if (!st.busy) return MPC_OK;
  MPC_CUDA(ctx, cudaEventSynchronize(st.done));
  float ms = 0.f;
  MPC_CUDA(ctx, cudaEventElapsedTime(&ms, st.k_start, st.k_stop));
  ctx->last_ms += ms;
  if (st.user_packed) memcpy(st.user_packed, st.h_packed, st.user_count * sizeof(uint16_t));
  st.user_packed = nullptr;
  st.busy = false;
  return MPC_OK;
}

// Host worker pool for the staging copies (pageable -> pinned memcpy, page cache -> pinned pread).  One core moves 5-10 GB/s, PCIe Gen5
// wants ~55 GB/s, so every 32 MiB chunk is split over a dozen threads -- and creating those threads per chunk cost as much as the copy
// itself (a 1 GiB file is 32 chunks x 12 threads).  The pool is created on first use and lives as long as the process; run() hands the
// slices out, the caller works along, concurrent callers (one host thread per GPU in the CLI) take turns.
class HostPool {
 public:
  static HostPool& get() {
    static HostPool pool;
    return pool;
  }
  size_t width() const { return workers_.size() + 1; }
  void run(size_t n, const std::function<void(size_t)>& fn) {
    if (n == 0) return;
    if (n == 1 || workers_.empty()) { for (size_t i = 0; i < n; i++) fn(i); return; }
    std::lock_guard<std::mutex> turn(run_m_);
    {
      std::lock_guard<std::mutex> lk(m_);
      job_ = &fn; n_ = n; next_ = 0; pending_ = n; gen_++;
    }
    cv_.notify_all();
    work();
    std::unique_lock<std::mutex> lk(m_);
    done_cv_.wait(lk, [&] { return pending_ == 0; });
    job_ = nullptr;
  }

 private:
  HostPool() {
    const unsigned hw = std::thread::hardware_concurrency();
    // all cores: measured on the pool's 16-core hosts, pread from the page cache scales 5.9 (1 thread) -> 40 (12) -> 46.5 GB/s (16)
    const size_t nt = hw >= 4 ? (hw > 32 ? 31 : hw - 1) : 0;
    for (size_t i = 0; i < nt; i++) workers_.emplace_back([this] { loop(); });
  }
  ~HostPool() {
    { std::lock_guard<std::mutex> lk(m_); stop_ = true; }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  void work() {  // take slices until none is left
    for (;;) {
      const std::function<void(size_t)>* job;
      size_t i;
      {
        std::lock_guard<std::mutex> lk(m_);
        if (!job_ || next_ >= n_) return;
        job = job_;
        i = next_++;
      }
      (*job)(i);
      bool last;
      { std::lock_guard<std::mutex> lk(m_); last = --pending_ == 0; }
      if (last) done_cv_.notify_all();
    }
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return stop_ || gen_ != seen; });
        if (stop_) return;
        seen = gen_;
      }
      work();
    }
  }
  std::vector<std::thread> workers_;
  std::mutex m_, run_m_;
  std::condition_variable cv_, done_cv_;
  const std::function<void(size_t)>* job_ = nullptr;
  size_t n_ = 0, next_ = 0, pending_ = 0;
  uint64_t gen_ = 0;
  bool stop_ = false;
};

// Pageable -> pinned staging copy on the pool.
void parallel_memcpy(uint8_t* dst, const uint8_t* src, size_t bytes) {
  HostPool& pool = HostPool::get();
  size_t nt = bytes < (8u << 20) ? 1 : pool.width();
  if (nt <= 1) { memcpy(dst, src, bytes); return; }
  const size_t per = ((bytes / nt) + 4095) & ~(size_t)4095;
  nt = (bytes + per - 1) / per;
  pool.run(nt, [=](size_t t) {
    const size_t lo = t * per, hi = (lo + per < bytes) ? lo + per : bytes;
    memcpy(dst + lo, src + lo, hi - lo);
  });
}

void free_stages(mpc_ctx* ctx) {
  for (Stage& st : ctx->stage) {
    if (st.stream) cudaStreamDestroy(st.stream);
    if (st.h_pinned) cudaFreeHost(st.h_pinned);
    if (st.d_lines) cudaFree(st.d_lines);
    if (st.d_packed) cudaFree(st.d_packed);
    if (st.h_packed) cudaFreeHost(st.h_packed);
    if (st.k_start) cudaEventDestroy(st.k_start);
    if (st.k_stop) cudaEventDestroy(st.k_stop);
    if (st.done) cudaEventDestroy(st.done);
    st = Stage();
  }
  ctx->stages_ready = false;
}

// Page cache (or, with O_DIRECT, the device) -> pinned staging on several host threads: pread copies in the kernel without
// the page faults of a mapping; one core moves ~5-10 GB/s, PCIe Gen5 wants ~55 GB/s.  Returns false on a short read / error.
bool parallel_pread(int fd, uint8_t* dst, size_t bytes, uint64_t file_off, size_t align) {
  HostPool& pool = HostPool::get();
  size_t nt = bytes < (8u << 20) ? 1 : pool.width();
  size_t per = ((bytes / nt) + 4095) & ~(size_t)4095;
  if (per < align) per = align;
  if (per == 0) per = 4096;
  nt = (bytes + per - 1) / per;
  std::vector<int> ok(nt ? nt : 1, 1);
  int* okp = ok.data();
  pool.run(nt, [=](size_t t) {
    const size_t lo = t * per, hi = (lo + per < bytes) ? lo + per : bytes;
    size_t done = lo;
    while (done < hi) {
      const ssize_t r = pread(fd, dst + done, hi - done, (off_t)(file_off + done));
      if (r <= 0) { okp[t] = 0; return; }
      done += (size_t)r;
    }
  });
  for (int v : ok) if (!v) return false;
  return true;
}

int ensure_stages(mpc_ctx* ctx) {
  if (ctx->stages_ready) return MPC_OK;
  free_stages(ctx);  // a previous attempt may have failed half way: start from nothing
  ctx->chunk_blocks = (32ull << 20) / (uint64_t)ctx->cfg.line_size;  // 32 MiB per chunk: full PCIe rate, short pipeline fill and drain
  const size_t bytes = (size_t)ctx->chunk_blocks * ctx->cfg.line_size;
  auto alloc = [&]() -> cudaError_t {
    cudaError_t e;
    for (Stage& st : ctx->stage) {
      if ((e = cudaStreamCreateWithFlags(&st.stream, cudaStreamNonBlocking)) != cudaSuccess) return e;
      if ((e = cudaMallocHost(&st.h_pinned, bytes + kStagePad)) != cudaSuccess) return e;
      if ((e = cudaMalloc(&st.d_lines, bytes)) != cudaSuccess) return e;
      if ((e = cudaMalloc(&st.d_packed, ctx->chunk_blocks * sizeof(uint16_t))) != cudaSuccess) return e;
      if ((e = cudaMallocHost(&st.h_packed, ctx->chunk_blocks * sizeof(uint16_t))) != cudaSuccess) return e;
      if ((e = cudaEventCreate(&st.k_start)) != cudaSuccess) return e;
      if ((e = cudaEventCreate(&st.k_stop)) != cudaSuccess) return e;
      if ((e = cudaEventCreateWithFlags(&st.done, cudaEventDisableTiming)) != cudaSuccess) return e;
    }
    return cudaSuccess;
  };
  const cudaError_t e = alloc();
  if (e != cudaSuccess) {
    free_stages(ctx);
    return fail(ctx, MPC_E_CUDA, "allocating the host-submit stages failed: %s", cudaGetErrorString(e));
  }
  ctx->stages_ready = true;
  // The first launch of a kernel loads its code onto the device (CUDA loads modules lazily: ~5 ms for the specialised kernels,
  // measured as the difference to CUDA_MODULE_LOADING=EAGER).  Pay that here, with the other one-off set-up costs, instead of in the
  // first chunk of the first submission: one tile of zeroes into a scratch statistics vector that is thrown away.
  {
    uint64_t* scratch = nullptr;
    Stage& st = ctx->stage[0];
    if (cudaMalloc(&scratch, mpc::kStatsWords * sizeof(uint64_t)) == cudaSuccess) {
      cudaMemsetAsync(scratch, 0, mpc::kStatsWords * sizeof(uint64_t), st.stream);
      cudaMemsetAsync(st.d_lines, 0, 4096, st.stream);
      std::swap(ctx->d_stats, scratch);
      const int rc = launch(ctx, st.d_lines, 4096 / (uint64_t)ctx->cfg.line_size, nullptr, st.stream, 1);
      std::swap(ctx->d_stats, scratch);
      cudaStreamSynchronize(st.stream);
      cudaFree(scratch);
      if (rc != MPC_OK) { free_stages(ctx); return rc; }
    }
  }
  return MPC_OK;
}

// ---- NCCL, resolved at run time: a single-GPU user never loads it, and under torchrun the process-wide libnccl.so.2
// (torch's) is the one that gets used --------------------------------------------------------------------------------
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names)
      if ((api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL)) != nullptr) break;
    if (!api.handle) { api.error = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
#define MPC_NCCL_SYM(field, name)                                                   \
  *(void**)(&api.field) = dlsym(api.handle, name);                                  \
  if (!api.field) { api.error = std::string("libnccl lacks ") + name; return; }
    MPC_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    MPC_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    MPC_NCCL_SYM(CommInitAll, "ncclCommInitAll")
    MPC_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    MPC_NCCL_SYM(CommCount, "ncclCommCount")
    MPC_NCCL_SYM(AllReduce, "ncclAllReduce")
    MPC_NCCL_SYM(GroupStart, "ncclGroupStart")
    MPC_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    MPC_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef MPC_NCCL_SYM
  });
  return &api;
}

#define MPC_NCCL(ctx, api, call)                                                                                 \
  do {                                                                                                           \
    ncclResult_t r__ = (call);                                                                                   \
    if (r__ != ncclSuccess) return fail(ctx, MPC_E_CUDA, "%s failed: %s", #call, (api)->GetErrorString(r__));    \
  } while (0)

// NCCL prints its version banner on stdout when NCCL_DEBUG=VERSION/INFO is set (this image sets it); the CLI's stdout
// carries exactly what the reference prints, so the banner is sent to stderr while a communicator initialises.
struct StdoutToStderr {
  int saved = -1;
  StdoutToStderr() { fflush(stdout); saved = dup(1); if (saved >= 0) dup2(2, 1); }
  ~StdoutToStderr() { if (saved >= 0) { fflush(stdout); dup2(saved, 1); close(saved); } }
};

int ensure_reduced(mpc_ctx* ctx) {
  if (ctx->d_reduced) return MPC_OK;
  MPC_CUDA(ctx, cudaSetDevice(ctx->device));
  MPC_CUDA(ctx, cudaMalloc(&ctx->d_reduced, mpc::kStatsWords * sizeof(uint64_t)));
  return MPC_OK;
}

}  // namespace

extern "C" {

const char* mpc_version(void) { return "mpc_b200 0.1 (sm_100a)"; }
const char* mpc_global_error(void) { return g_error.c_str(); }
const char* mpc_last_error(const mpc_ctx* ctx) { return ctx ? ctx->error.c_str() : g_error.c_str(); }

int mpc_create(const mpc_config_pod* cfg, int device, mpc_ctx** out) {
  if (!cfg || !out) return fail(nullptr, MPC_E_ARG, "mpc_create: null argument");
  char err[512] = {0};
  int rc = mpc_config_validate(cfg, err, sizeof(err));
  if (rc != MPC_OK) return fail(nullptr, rc, "%s", err);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, MPC_E_CUDA, "no CUDA device available (%s); libmpc_b200 has no CPU path",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(nullptr, MPC_E_ARG, "device %d out of range (have %d)", device, ndev);
  mpc_ctx* ctx = new mpc_ctx;
  ctx->cfg = *cfg;
  ctx->device = device;
#define MPC_CREATE_CUDA(call)                                                                 \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      int rc__ = fail(nullptr, MPC_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__));  \
      mpc_destroy(ctx);                                                                       \
      return rc__;                                                                            \
    }                                                                                         \
  } while (0)
  MPC_CREATE_CUDA(cudaSetDevice(device));
  MPC_CREATE_CUDA(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device));
  MPC_CREATE_CUDA(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
  ctx->stream = ctx->own_stream;
  MPC_CREATE_CUDA(cudaEventCreate(&ctx->ev_start));
  MPC_CREATE_CUDA(cudaEventCreate(&ctx->ev_stop));
  MPC_CREATE_CUDA(cudaEventCreateWithFlags(&ctx->ev_order, cudaEventDisableTiming));
  MPC_CREATE_CUDA(cudaMalloc(&ctx->d_stats, mpc::kStatsWords * sizeof(uint64_t)));
  MPC_CREATE_CUDA(cudaMemsetAsync(ctx->d_stats, 0, mpc::kStatsWords * sizeof(uint64_t), ctx->stream));
  MPC_CREATE_CUDA(cudaMalloc(&ctx->d_sched, (1 + kHostStages) * 32 * sizeof(uint32_t)));
  MPC_CREATE_CUDA(cudaMemsetAsync(ctx->d_sched, 0, (1 + kHostStages) * 32 * sizeof(uint32_t), ctx->stream));
  std::vector<mpc::GenericModule> gm((size_t)MPC_MAX_MODULES);
  mpc::build_generic_tables(ctx->cfg, &ctx->gparams, gm.data());
  const size_t gbytes = sizeof(mpc::GenericModule) * (size_t)(ctx->gparams.num_predcomp > 0 ? ctx->gparams.num_predcomp : 1);
  MPC_CREATE_CUDA(cudaMalloc(&ctx->d_gmods, gbytes));
  MPC_CREATE_CUDA(cudaMemcpyAsync(ctx->d_gmods, gm.data(), gbytes, cudaMemcpyHostToDevice, ctx->stream));
  ctx->spec = mpc::find_spec_kernel(ctx->cfg);
  if (!ctx->spec) {
    // no compiled-in specialisation: build one now with NVRTC when the config is eligible (MPC_JIT=0 disables);
    // a config that is not eligible, or a failed build, leaves the generic kernel in charge and says why
    const char* env = getenv("MPC_JIT");
    std::string why;
    if (env && env[0] == '0') ctx->jit_note = "run-time specialisation disabled (MPC_JIT=0)";
    else if (!mpc::spec_eligible(ctx->cfg, &why)) ctx->jit_note = "not eligible for the specialised kernel: " + why;
    else {
      std::string log;
      ctx->jit = mpc::jit_create(ctx->cfg, &log);
      if (!ctx->jit) ctx->jit_note = "run-time specialisation failed: " + log;
    }
  }
  {
    std::vector<uint8_t> lut(mpc::kRowLutBytes);
    mpc::build_row_cost_lut(lut.data(), ctx->spec ? ctx->spec->lut_xor : (ctx->jit ? mpc::jit_lut_xor(ctx->jit) : 0));
    MPC_CREATE_CUDA(cudaMalloc(&ctx->d_row_lut, lut.size()));
    MPC_CREATE_CUDA(cudaMemcpy(ctx->d_row_lut, lut.data(), lut.size(), cudaMemcpyHostToDevice));
  }
  MPC_CREATE_CUDA(cudaStreamSynchronize(ctx->stream));
#undef MPC_CREATE_CUDA
  refresh_kernel_name(ctx);
  *out = ctx;
  return MPC_OK;
}

void mpc_destroy(mpc_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  if (ctx->comm && ctx->comm_owned) {
    NcclApi* api = nccl_api();
    if (api->CommDestroy) api->CommDestroy(ctx->comm);
  }
  if (ctx->d_reduced) cudaFree(ctx->d_reduced);
  free_stages(ctx);
  if (ctx->d_gmods) cudaFree(ctx->d_gmods);
  if (ctx->d_row_lut) cudaFree(ctx->d_row_lut);
  if (ctx->jit) mpc::jit_destroy(ctx->jit);
  if (ctx->d_stats) cudaFree(ctx->d_stats);
  if (ctx->d_sched) cudaFree(ctx->d_sched);
  if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
  if (ctx->ev_stop) cudaEventDestroy(ctx->ev_stop);
  if (ctx->ev_order) cudaEventDestroy(ctx->ev_order);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

int mpc_set_kernel(mpc_ctx* ctx, int which) {
  if (!ctx) return MPC_E_ARG;
  if (which < 0 || which > 2) return fail(ctx, MPC_E_ARG, "mpc_set_kernel: which = %d", which);
  if (which == 2 && !ctx->spec && !ctx->jit)
    return fail(ctx, MPC_E_STATE, "no specialised kernel for the given config (%s)", ctx->jit_note.c_str());
  ctx->kernel_choice = which;
  refresh_kernel_name(ctx);
  return MPC_OK;
}

const char* mpc_kernel_name(const mpc_ctx* ctx) { return ctx ? ctx->kernel_name.c_str() : ""; }

int mpc_jit_compile_check(const mpc_config_pod* cfg, char* log, size_t log_len, size_t* cubin_bytes) {
  if (!cfg) return MPC_E_ARG;
  std::vector<char> cubin;
  std::string l;
  mpc::SpecTraits t;
  const int rc = mpc::jit_compile(*cfg, &cubin, &t, &l);
  if (log && log_len) snprintf(log, log_len, "%s", l.c_str());
  if (cubin_bytes) *cubin_bytes = cubin.size();
  return rc;
}

int mpc_enable_timing(mpc_ctx* ctx, int enabled) {
  if (!ctx) return MPC_E_ARG;
  ctx->timing_enabled = enabled != 0;
  return MPC_OK;
}

int mpc_set_stream(mpc_ctx* ctx, void* cuda_stream) {
  if (!ctx) return MPC_E_ARG;
  int rc = mpc_sync(ctx);
  if (rc != MPC_OK) return rc;
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return MPC_OK;
}

int mpc_submit_device(mpc_ctx* ctx, const uint8_t* d_lines, uint64_t n_blocks, uint16_t* d_packed) {
  if (!ctx) return MPC_E_ARG;
  if (n_blocks && !d_lines) return fail(ctx, MPC_E_ARG, "mpc_submit_device: null lines");
  if ((uintptr_t)d_lines & 15) return fail(ctx, MPC_E_ARG, "mpc_submit_device: lines must be 16-byte aligned");
  MPC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (ctx->timing_enabled) MPC_CUDA(ctx, cudaEventRecord(ctx->ev_start, ctx->stream));
  int rc = n_blocks ? launch(ctx, d_lines, n_blocks, d_packed, ctx->stream, 0) : MPC_OK;
  if (rc != MPC_OK) return rc;
  if (ctx->timing_enabled) MPC_CUDA(ctx, cudaEventRecord(ctx->ev_stop, ctx->stream));
  ctx->last_timing_pending = ctx->timing_enabled;
  if (!ctx->timing_enabled) ctx->last_ms = 0.f;
  ctx->last_launches = n_blocks ? 1 : 0;
  return MPC_OK;
}

int mpc_submit_host(mpc_ctx* ctx, const uint8_t* h_lines, uint64_t n_blocks, uint16_t* h_packed) {
  if (!ctx) return MPC_E_ARG;
  if (n_blocks && !h_lines) return fail(ctx, MPC_E_ARG, "mpc_submit_host: null lines");
  MPC_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = ensure_stages(ctx);
  if (rc != MPC_OK) return rc;
  // Statistics memset / earlier device submits are ordered on ctx->stream; make the stage streams see them.
  MPC_CUDA(ctx, cudaEventRecord(ctx->ev_order, ctx->stream));
  for (Stage& st : ctx->stage) MPC_CUDA(ctx, cudaStreamWaitEvent(st.stream, ctx->ev_order, 0));
  ctx->last_timing_pending = false;
  ctx->last_ms = 0.f;
  ctx->last_launches = 0;
  const uint64_t L = (uint64_t)ctx->cfg.line_size;
  cudaPointerAttributes attr;
  bool pinned_src = false;
  if (n_blocks && cudaPointerGetAttributes(&attr, h_lines) == cudaSuccess) pinned_src = (attr.type == cudaMemoryTypeHost);
  cudaGetLastError();
  uint64_t done_blocks = 0;
  int which = 0;
  while (done_blocks < n_blocks) {
    Stage& st = ctx->stage[which];
    rc = drain_stage(ctx, st);
    if (rc != MPC_OK) return rc;
    const uint64_t nb = (n_blocks - done_blocks < ctx->chunk_blocks) ? (n_blocks - done_blocks) : ctx->chunk_blocks;
    const uint8_t* src = h_lines + done_blocks * L;
    if (pinned_src) {
      MPC_CUDA(ctx, cudaMemcpyAsync(st.d_lines, src, nb * L, cudaMemcpyHostToDevice, st.stream));
    } else {
      parallel_memcpy(st.h_pinned, src, nb * L);  // LoaderNPY.cpp:24-26 copied one line at a time; this is the chunked form
      MPC_CUDA(ctx, cudaMemcpyAsync(st.d_lines, st.h_pinned, nb * L, cudaMemcpyHostToDevice, st.stream));
    }
    MPC_CUDA(ctx, cudaEventRecord(st.k_start, st.stream));
    rc = launch(ctx, st.d_lines, nb, h_packed ? st.d_packed : nullptr, st.stream, 1 + which);
    if (rc != MPC_OK) return rc;
    MPC_CUDA(ctx, cudaEventRecord(st.k_stop, st.stream));
    ctx->last_launches++;
    if (h_packed) {
      MPC_CUDA(ctx, cudaMemcpyAsync(st.h_packed, st.d_packed, nb * sizeof(uint16_t), cudaMemcpyDeviceToHost, st.stream));
      st.user_packed = h_packed + done_blocks;
      st.user_count = nb;
    }
    MPC_CUDA(ctx, cudaEventRecord(st.done, st.stream));
    st.busy = true;
    done_blocks += nb;
    which = (which + 1) % kHostStages;
  }
  return MPC_OK;
}

int mpc_submit_file(mpc_ctx* ctx, int fd, uint64_t file_offset, uint64_t n_blocks, uint16_t* h_packed, int direct_io) {
  if (!ctx) return MPC_E_ARG;
  if (fd < 0) return fail(ctx, MPC_E_ARG, "mpc_submit_file: bad file descriptor");
  MPC_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = ensure_stages(ctx);
  if (rc != MPC_OK) return rc;
  MPC_CUDA(ctx, cudaEventRecord(ctx->ev_order, ctx->stream));
  for (Stage& st : ctx->stage) MPC_CUDA(ctx, cudaStreamWaitEvent(st.stream, ctx->ev_order, 0));
  ctx->last_timing_pending = false;
  ctx->last_ms = 0.f;
  ctx->last_launches = 0;
  const uint64_t L = (uint64_t)ctx->cfg.line_size;
  uint64_t done_blocks = 0;
  int which = 0;
  const bool trace = getenv("MPC_TRACE_IO") != nullptr;  // stderr: where the host thread's time goes
  double t_read = 0, t_wait = 0;
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  while (done_blocks < n_blocks) {
    Stage& st = ctx->stage[which];
    const double tw0 = now();
    rc = drain_stage(ctx, st);
    if (rc != MPC_OK) return rc;
    t_wait += now() - tw0;
    const uint64_t nb = (n_blocks - done_blocks < ctx->chunk_blocks) ? (n_blocks - done_blocks) : ctx->chunk_blocks;
    const uint64_t off = file_offset + done_blocks * L;
    const double tr0 = now();
    // O_DIRECT wants file offset, length and buffer aligned: read from the aligned offset below `off` (the stage has the room)
    // and copy to the device from where the chunk's first byte landed
    const uint64_t head = direct_io ? (off & (kStagePad - 1)) : 0;
    size_t want = (size_t)(nb * L + head);
    if (direct_io) want = (want + kStagePad - 1) & ~(kStagePad - 1);
    // the tail of the last chunk may reach past the end of the file: a short read there is not an error
    if (!parallel_pread(fd, st.h_pinned, want, off - head, direct_io ? kStagePad : 1)) {
      struct stat sb;
      if (fstat(fd, &sb) != 0 || (uint64_t)sb.st_size < off + nb * L)
        return fail(ctx, MPC_E_IO, "mpc_submit_file: short read at offset %llu", (unsigned long long)off);
    }
    t_read += now() - tr0;
    MPC_CUDA(ctx, cudaMemcpyAsync(st.d_lines, st.h_pinned + head, nb * L, cudaMemcpyHostToDevice, st.stream));
    MPC_CUDA(ctx, cudaEventRecord(st.k_start, st.stream));
    rc = launch(ctx, st.d_lines, nb, h_packed ? st.d_packed : nullptr, st.stream, 1 + which);
    if (rc != MPC_OK) return rc;
    MPC_CUDA(ctx, cudaEventRecord(st.k_stop, st.stream));
    ctx->last_launches++;
    if (h_packed) {
      MPC_CUDA(ctx, cudaMemcpyAsync(st.h_packed, st.d_packed, nb * sizeof(uint16_t), cudaMemcpyDeviceToHost, st.stream));
      st.user_packed = h_packed + done_blocks;
      st.user_count = nb;
    }
    MPC_CUDA(ctx, cudaEventRecord(st.done, st.stream));
    st.busy = true;
    done_blocks += nb;
    which = (which + 1) % kHostStages;
  }
  if (trace)
    fprintf(stderr, "mpc_submit_file: %.1f MiB, pread %.1f ms (%.1f GB/s), waiting for a free slot %.1f ms\n", n_blocks * L / 1048576.0,
            1e3 * t_read, t_read > 0 ? n_blocks * L / t_read / 1e9 : 0.0, 1e3 * t_wait);
  return MPC_OK;
}

int mpc_prepare_host(mpc_ctx* ctx) {
  if (!ctx) return MPC_E_ARG;
  MPC_CUDA(ctx, cudaSetDevice(ctx->device));
  return ensure_stages(ctx);
}

int mpc_sync(mpc_ctx* ctx) {
  if (!ctx) return MPC_E_ARG;
  MPC_CUDA(ctx, cudaSetDevice(ctx->device));
  for (Stage& st : ctx->stage) {
    int rc = drain_stage(ctx, st);
    if (rc != MPC_OK) return rc;
  }
  MPC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->last_timing_pending) {
    MPC_CUDA(ctx, cudaEventElapsedTime(&ctx->last_ms, ctx->ev_start, ctx->ev_stop));
    ctx->last_timing_pending = false;
  }
  return MPC_OK;
}

int mpc_stats_device_ptr(mpc_ctx* ctx, uint64_t** d_stats, size_t* n_words) {
  if (!ctx || !d_stats || !n_words) return MPC_E_ARG;
  *d_stats = ctx->d_stats;
  *n_words = mpc::kStatsWords;
  return MPC_OK;
}

int mpc_stats_expand(const mpc_config_pod* cfg, const uint64_t* w, size_t n_words, mpc_stats_pod* out) {
  if (!cfg || !w || !out || n_words != mpc::kStatsWords) return MPC_E_ARG;
  memset(out, 0, sizeof(*out));
  const uint64_t line_bits = 8ull * (uint64_t)cfg->line_size;
  for (int k = 0; k <= cfg->num_modules; k++) {
    uint64_t cnt = 0, comp = 0;
    for (int s = 0; s < MPC_HIST_BINS; s++) {
      uint64_t h = w[mpc::kHistOff + (size_t)k * mpc::kHB + s];
      out->hist[k][s] = h;
      cnt += h;
      comp += h * (uint64_t)s;
    }
    out->count[k] = cnt;           // ClusterStat::count, VPC.h:56
    out->comp_bits[k] = comp;      // ClusterStat::compressedSize, VPC.h:54
    out->res_abs[k] = w[mpc::kResAbsOff + k];
    out->res_sq[k] = w[mpc::kResSqOff + k];
    // lines that reached checkOtherPatterns: cluster -1 and every PredComp cluster (VPC.cpp:410-412)
    out->res_lines[k] = (k == 0 || k - 1 >= cfg->first_predcomp) ? cnt : 0;
    out->blocks += cnt;
    out->compressed_bits += comp;  // CompResult::CompressedSize, CompResult.h:33
  }
  out->original_bits = out->blocks * line_bits;  // CompResult::OriginalSize, CompResult.h:32
  return MPC_OK;
}

int mpc_finish(mpc_ctx* ctx, mpc_stats_pod* out) {
  if (!ctx || !out) return MPC_E_ARG;
  int rc = mpc_sync(ctx);
  if (rc != MPC_OK) return rc;
  std::vector<uint64_t> host(mpc::kStatsWords);
  MPC_CUDA(ctx, cudaMemcpy(host.data(), ctx->d_stats, mpc::kStatsWords * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return mpc_stats_expand(&ctx->cfg, host.data(), host.size(), out);
}

int mpc_reset(mpc_ctx* ctx) {
  if (!ctx) return MPC_E_ARG;
  int rc = mpc_sync(ctx);
  if (rc != MPC_OK) return rc;
  MPC_CUDA(ctx, cudaMemsetAsync(ctx->d_stats, 0, mpc::kStatsWords * sizeof(uint64_t), ctx->stream));
  MPC_CUDA(ctx, cudaMemsetAsync(ctx->d_sched, 0, (1 + kHostStages) * 32 * sizeof(uint32_t), ctx->stream));
  MPC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MPC_OK;
}

int mpc_last_timing(mpc_ctx* ctx, float* kernel_ms, int* launches) {
  if (!ctx) return MPC_E_ARG;
  int rc = mpc_sync(ctx);
  if (rc != MPC_OK) return rc;
  if (kernel_ms) *kernel_ms = ctx->last_ms;
  if (launches) *launches = ctx->last_launches;
  return MPC_OK;
}


/* ---- multi-GPU: the statistics all-reduce behind the ABI ------------------------------------------------------- */

int mpc_comm_unique_id(void* uid, size_t uid_len) {
  if (!uid || uid_len < sizeof(ncclUniqueId)) return fail(nullptr, MPC_E_ARG, "mpc_comm_unique_id: need %zu bytes", sizeof(ncclUniqueId));
  NcclApi* api = nccl_api();
  if (!api->error.empty()) return fail(nullptr, MPC_E_STATE, "%s", api->error.c_str());
  ncclUniqueId id;
  MPC_NCCL(nullptr, api, api->GetUniqueId(&id));
  memcpy(uid, &id, sizeof(id));
  return MPC_OK;
}

int mpc_comm_init_rank(mpc_ctx* ctx, const void* uid, size_t uid_len, int nranks, int rank) {
  if (!ctx || !uid || uid_len < sizeof(ncclUniqueId) || nranks < 1 || rank < 0 || rank >= nranks)
    return fail(ctx, MPC_E_ARG, "mpc_comm_init_rank: bad argument");
  if (ctx->comm) return fail(ctx, MPC_E_STATE, "mpc_comm_init_rank: the context already has a communicator");
  NcclApi* api = nccl_api();
  if (!api->error.empty()) return fail(ctx, MPC_E_STATE, "%s", api->error.c_str());
  MPC_CUDA(ctx, cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, uid, sizeof(id));
  {
    StdoutToStderr quiet;
    MPC_NCCL(ctx, api, api->CommInitRank(&ctx->comm, nranks, id, rank));
  }
  ctx->comm_owned = true;
  return ensure_reduced(ctx);
}

int mpc_comm_init_all(mpc_ctx** ctxs, int n) {
  if (!ctxs || n < 1) return fail(nullptr, MPC_E_ARG, "mpc_comm_init_all: bad argument");
  for (int i = 0; i < n; i++)
    if (!ctxs[i] || ctxs[i]->comm) return fail(ctxs[i], MPC_E_STATE, "mpc_comm_init_all: context %d is null or already has a communicator", i);
  NcclApi* api = nccl_api();
  if (!api->error.empty()) return fail(ctxs[0], MPC_E_STATE, "%s", api->error.c_str());
  std::vector<ncclComm_t> comms((size_t)n);
  std::vector<int> devs((size_t)n);
  for (int i = 0; i < n; i++) devs[(size_t)i] = ctxs[i]->device;
  {
    StdoutToStderr quiet;
    MPC_NCCL(ctxs[0], api, api->CommInitAll(comms.data(), n, devs.data()));
  }
  for (int i = 0; i < n; i++) {
    ctxs[i]->comm = comms[(size_t)i];
    ctxs[i]->comm_owned = true;
    const int rc = ensure_reduced(ctxs[i]);
    if (rc != MPC_OK) return rc;
  }
  return MPC_OK;
}

int mpc_attach_comm(mpc_ctx* ctx, void* nccl_comm) {
  if (!ctx) return MPC_E_ARG;
  if (ctx->comm && ctx->comm_owned) return fail(ctx, MPC_E_STATE, "mpc_attach_comm: the context owns a communicator already");
  NcclApi* api = nccl_api();
  if (!api->error.empty()) return fail(ctx, MPC_E_STATE, "%s", api->error.c_str());
  ctx->comm = (ncclComm_t)nccl_comm;
  ctx->comm_owned = false;
  return nccl_comm ? ensure_reduced(ctx) : MPC_OK;
}

int mpc_allreduce_stats(mpc_ctx** ctxs, int n) {
  if (!ctxs || n < 1) return fail(nullptr, MPC_E_ARG, "mpc_allreduce_stats: bad argument");
  NcclApi* api = nullptr;
  for (int i = 0; i < n; i++) {
    mpc_ctx* ctx = ctxs[i];
    if (!ctx) return fail(nullptr, MPC_E_ARG, "mpc_allreduce_stats: null context");
    int rc = ensure_reduced(ctx);
    if (rc != MPC_OK) return rc;
    MPC_CUDA(ctx, cudaSetDevice(ctx->device));
    // chunks still in flight on the mpc_submit_host stage streams must land in d_stats before it is read
    for (Stage& st : ctx->stage)
      if (st.busy) MPC_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, st.done, 0));
    MPC_CUDA(ctx, cudaMemcpyAsync(ctx->d_reduced, ctx->d_stats, mpc::kStatsWords * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
    if (ctx->comm && !api) {
      api = nccl_api();
      if (!api->error.empty()) return fail(ctx, MPC_E_STATE, "%s", api->error.c_str());
    }
  }
  if (!api) return MPC_OK;  // no communicator anywhere: one GPU, the copy is the reduction
  MPC_NCCL(ctxs[0], api, api->GroupStart());
  for (int i = 0; i < n; i++) {
    mpc_ctx* ctx = ctxs[i];
    if (!ctx->comm) { api->GroupEnd(); return fail(ctx, MPC_E_STATE, "mpc_allreduce_stats: context %d has no communicator", i); }
    cudaSetDevice(ctx->device);
    const ncclResult_t r = api->AllReduce(ctx->d_reduced, ctx->d_reduced, mpc::kStatsWords, ncclUint64, ncclSum, ctx->comm, ctx->stream);
    if (r != ncclSuccess) { api->GroupEnd(); return fail(ctx, MPC_E_CUDA, "ncclAllReduce failed: %s", api->GetErrorString(r)); }
  }
  MPC_NCCL(ctxs[0], api, api->GroupEnd());
  return MPC_OK;
}

int mpc_reduced_stats(mpc_ctx* ctx, mpc_stats_pod* out) {
  if (!ctx || !out) return MPC_E_ARG;
  if (!ctx->d_reduced) return fail(ctx, MPC_E_STATE, "mpc_reduced_stats: no all-reduce has been issued on this context");
  MPC_CUDA(ctx, cudaSetDevice(ctx->device));
  MPC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  std::vector<uint64_t> host(mpc::kStatsWords);
  MPC_CUDA(ctx, cudaMemcpy(host.data(), ctx->d_reduced, mpc::kStatsWords * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return mpc_stats_expand(&ctx->cfg, host.data(), host.size(), out);
}

int mpc_reduced_device_ptr(mpc_ctx* ctx, uint64_t** d_reduced, size_t* n_words) {
  if (!ctx || !d_reduced || !n_words) return MPC_E_ARG;
  int rc = ensure_reduced(ctx);
  if (rc != MPC_OK) return rc;
  *d_reduced = ctx->d_reduced;
  *n_words = mpc::kStatsWords;
  return MPC_OK;
}

int mpc_finish_allreduce(mpc_ctx** ctxs, int n, mpc_stats_pod* out) {
  if (!ctxs || n < 1 || !out) return fail(nullptr, MPC_E_ARG, "mpc_finish_allreduce: bad argument");
  for (int i = 0; i < n; i++) {
    int rc = mpc_sync(ctxs[i]);
    if (rc != MPC_OK) return rc;
  }
  int rc = mpc_allreduce_stats(ctxs, n);
  if (rc != MPC_OK) return rc;
  for (int i = 1; i < n; i++) {
    MPC_CUDA(ctxs[i], cudaSetDevice(ctxs[i]->device));
    MPC_CUDA(ctxs[i], cudaStreamSynchronize(ctxs[i]->stream));
  }
  return mpc_reduced_stats(ctxs[0], out);
}

int mpc_synth_device(mpc_ctx* ctx, uint8_t* d_lines, uint64_t first_block, uint64_t n_blocks, uint64_t total_blocks,
                     int kind, uint64_t seed) {
  if (!ctx || (n_blocks && !d_lines)) return MPC_E_ARG;
  if (kind < MPC_SYN_ZERO || kind > MPC_SYN_MIXED_REGIONS) return fail(ctx, MPC_E_ARG, "unknown synthetic kind %d", kind);
  if (ctx->cfg.line_size != 128) return fail(ctx, MPC_E_ARG, "synthetic dumps are defined for 128-byte blocks");
  MPC_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaError_t e = mpc::launch_synth(d_lines, first_block, n_blocks, total_blocks, kind, seed, ctx->stream);
  if (e != cudaSuccess) return fail(ctx, MPC_E_CUDA, "synth launch failed: %s", cudaGetErrorString(e));
  return MPC_OK;
}

}  // extern "C"
