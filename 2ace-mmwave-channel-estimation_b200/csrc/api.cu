// libtwoace C ABI (include/twoace.h): host orchestration of the sm_100a kernels.
// No CPU fallback exists: every entry point launches CUDA kernels or fails with TWOACE_E_CUDA.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/twoace.h"
#include "solve_kernels.cuh"
#include "big_stage.cuh"
#include "phaselift.cuh"
#include "metrics.cuh"
#include "synth.cuh"
#include "minl2.cuh"
#include <thread>

using namespace twoace;

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

struct twoace_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  int64_t launches = 0;
  int num_sms = 0;
  int chunk = 8192;
  DevBuf arena, ws, taskbuf;
  cd* cb_rm = nullptr;   // row-major codebook
  int cb_rows = 0, cb_n = 0;
  uint32_t* cb_codes = nullptr;   // 2-bit codes of the codebook (n == 256 and quantised) or nullptr
  double cb_mag = 0.0;
  int opt_fast = 1;      // 1: use the shared-memory cluster kernel when a launch is eligible
  int opt_fast_cs = 2;   // cluster size for the r = 20 stages (2 or 4)
  int opt_dedup_nuclear = 0;   // 1: do not re-execute the (bit-identical) rank-one rerun of inferLowRank_Nuclear
  int opt_cache_sinv = 1;   // 1: the stages of a trial share one (I + A A')^-1 (computed by the first of them)
  int opt_spectral_jacobi = 0;   // 1: full Jacobi eigendecomposition in the spectral initialisation (round-1 path)
  int opt_tensor = 1;    // 1: exact int8 tensor-core (tcgen05) A-products in the cluster kernel, 0: FP64 SIMT products
  int64_t fast_launches = 0, tc_launches = 0;
  double* trace_user = nullptr;   // optional residual-trace sink of the next solves (twoace_set_trace)
  int trace_mem = 0;
  size_t trace_cap = 0;
  bool timing = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> stage_events;
  std::vector<std::string> stage_labels;   // one per stage_events entry (TWOACE_TRACE_LAUNCHES=1 prints them)
  // The kernel groups of one InferADMM stage launch (different cluster shapes / kernels for different m) run
  // concurrently: the first on the main stream, the others on side streams, each with its own workspace.
  static constexpr int NSIDE = 4;
  cudaStream_t side_stream[NSIDE] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[NSIDE] = {nullptr, nullptr, nullptr, nullptr};
  DevBuf ws_side[NSIDE];
  bool in_side = false;                    // inside a concurrent launch group: per-kernel events are not summed
  DevBuf counters;                         // task-queue counters of the cluster kernels (ring of NCOUNTER ints)
  int counter_cursor = 0;
  int opt_overlap = 1;                     // 1: overlap the general-kernel group of a stage with its cluster-kernel groups
  std::vector<char> stage_side;            // per stage_events entry: 1 = ran on the side stream (not summed)
  std::vector<std::pair<void*, size_t>> stage_cache;   // free device buffers of the host-pointer staging (pointer, capacity)
  std::vector<twoace_ctx*> peers;          // twoace_create_multi: the contexts of the other GPUs (this one is device 0 of the set)
};

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      char buf_[512];                                                                         \
      snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),     \
               __FILE__, __LINE__);                                                           \
      ctx->err = buf_;                                                                        \
      return TWOACE_E_CUDA;                                                                   \
    }                                                                                         \
  } while (0)

#define FAIL(code, ...)                                  \
  do {                                                   \
    char buf_[512];                                      \
    snprintf(buf_, sizeof buf_, __VA_ARGS__);            \
    ctx->err = buf_;                                     \
    return code;                                         \
  } while (0)

static int ensure(twoace_ctx* ctx, DevBuf& b, size_t bytes) {
  if (bytes <= b.cap) return 0;
  if (b.p) {
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
  }
  size_t want = bytes + bytes / 8 + 4096;
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    FAIL(TWOACE_E_NOMEM, "cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
  }
  b.cap = want;
  return 0;
}

struct Bump {
  size_t off = 0;
  size_t take(size_t bytes) {
    size_t o = (off + 255) / 256 * 256;
    off = o + bytes;
    return o;
  }
};

extern "C" void twoace_default_params(twoace_params* p) {
  p->lambda = 0.0; p->r = 20; p->mu0 = 1e-3; p->rho = 1.03; p->cc_frac = 0.95;
  p->tol_rel = 1e-4; p->tol_abs = 1e-8; p->maxiter = 500;
}

extern "C" int twoace_version(void) { return 100; }

extern "C" int twoace_create(int device, twoace_ctx** out) {
  if (!out) return TWOACE_E_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    (void)cudaGetLastError();
    return TWOACE_E_CUDA;
  }
  twoace_ctx* ctx = new twoace_ctx();
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    (void)cudaGetLastError();
    delete ctx;
    return TWOACE_E_CUDA;
  }
  bool side_ok = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) == cudaSuccess;
  for (int k = 0; k < twoace_ctx::NSIDE && side_ok; ++k)
    side_ok = cudaStreamCreateWithFlags(&ctx->side_stream[k], cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&ctx->ev_join[k], cudaEventDisableTiming) == cudaSuccess;
  if (!side_ok) {
    (void)cudaGetLastError();
    ctx->side_stream[0] = nullptr;     // overlap disabled; everything runs on the main stream
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  ctx->num_sms = prop.multiProcessorCount;
  *out = ctx;
  return TWOACE_OK;
}

extern "C" void twoace_destroy(twoace_ctx* ctx) {
  if (!ctx) return;
  for (twoace_ctx* p : ctx->peers) twoace_destroy(p);
  ctx->peers.clear();
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->arena.p) cudaFree(ctx->arena.p);
  if (ctx->ws.p) cudaFree(ctx->ws.p);
  if (ctx->taskbuf.p) cudaFree(ctx->taskbuf.p);
  if (ctx->cb_rm) cudaFree(ctx->cb_rm);
  if (ctx->cb_codes) cudaFree(ctx->cb_codes);
  for (auto& pr : ctx->stage_cache) cudaFree(pr.first);
  for (int k = 0; k < twoace_ctx::NSIDE; ++k) {
    if (ctx->side_stream[k]) { cudaStreamSynchronize(ctx->side_stream[k]); cudaStreamDestroy(ctx->side_stream[k]); }
    if (ctx->ev_join[k]) cudaEventDestroy(ctx->ev_join[k]);
    if (ctx->ws_side[k].p) cudaFree(ctx->ws_side[k].p);
  }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->counters.p) cudaFree(ctx->counters.p);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char* twoace_last_error(const twoace_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" void* twoace_stream(twoace_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" int64_t twoace_launch_count(const twoace_ctx* ctx) {
  if (!ctx) return 0;
  int64_t n = ctx->launches;
  for (const twoace_ctx* p : ctx->peers) n += p->launches;
  return n;
}
extern "C" int twoace_synchronize(twoace_ctx* ctx) {
  if (!ctx) return TWOACE_E_INVALID;
  for (twoace_ctx* p : ctx->peers) { int rc = twoace_synchronize(p); if (rc) { ctx->err = p->err; return rc; } }
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  return TWOACE_OK;
}

// ---- several GPUs behind one context (instances are independent, each GPU owns a contiguous slice)
extern "C" int twoace_create_multi(const int* devices, int n_dev, twoace_ctx** out) {
  if (!out) return TWOACE_E_INVALID;
  *out = nullptr;
  if (!devices || n_dev < 1) return TWOACE_E_INVALID;
  // (a device may be listed more than once: it then runs that many independent pipelines, whose kernels interleave)
  twoace_ctx* c = nullptr;
  int rc = twoace_create(devices[0], &c);
  if (rc) return rc;
  for (int i = 1; i < n_dev; ++i) {
    twoace_ctx* p = nullptr;
    rc = twoace_create(devices[i], &p);
    if (rc) { twoace_destroy(c); return rc; }
    c->peers.push_back(p);
  }
  *out = c;
  return TWOACE_OK;
}

extern "C" int twoace_device_count(const twoace_ctx* ctx) { return ctx ? 1 + (int)ctx->peers.size() : 0; }

// Contiguous slices of a ragged batch, balanced by the row counts (the cost of a solve grows with m).
struct BatchSlice { int b0, b1; size_t rows0; };   // instances [b0, b1), rows0 = sum of m over [0, b0)
static std::vector<BatchSlice> split_batch(const int32_t* m, int nb, int parts) {
  std::vector<size_t> pre(nb + 1, 0);
  for (int b = 0; b < nb; ++b) pre[b + 1] = pre[b] + (m ? (size_t)std::max(1, m[b]) : 1);
  std::vector<BatchSlice> out;
  int b0 = 0;
  for (int k = 0; k < parts; ++k) {
    int b1 = nb;
    if (k + 1 < parts) {
      const size_t want = pre[nb] * (size_t)(k + 1) / parts;
      b1 = (int)(std::lower_bound(pre.begin(), pre.end(), want) - pre.begin());
      b1 = std::min(std::max(b1, b0), nb);
    }
    size_t rows0 = 0;
    if (m) for (int b = 0; b < b0; ++b) rows0 += m[b];
    out.push_back({b0, b1, rows0});
    b0 = b1;
  }
  return out;
}

// Run f(context of GPU k, slice k) on one host thread per GPU; the first failure is reported through `ctx`.
template <class F>
static int run_on_all(twoace_ctx* ctx, const std::vector<BatchSlice>& sl, F f) {
  const int parts = (int)sl.size();
  std::vector<int> rcs(parts, 0);
  std::vector<std::thread> th;
  auto sub = [&](int k) { return k == 0 ? ctx : ctx->peers[k - 1]; };
  for (int k = 1; k < parts; ++k)
    th.emplace_back([&, k]() { rcs[k] = sl[k].b1 > sl[k].b0 ? f(sub(k), sl[k]) : 0; });
  rcs[0] = sl[0].b1 > sl[0].b0 ? f(ctx, sl[0]) : 0;
  for (auto& t : th) t.join();
  for (int k = 1; k < parts; ++k)
    if (rcs[k]) { ctx->err = "GPU " + std::to_string(sub(k)->device) + ": " + sub(k)->err; return rcs[k]; }
  return rcs[0];
}

static int multi_guard(twoace_ctx* ctx, int mem) {
  if (mem != TWOACE_MEM_HOST) FAIL(TWOACE_E_INVALID, "a multi-GPU context takes host buffers only (mem = TWOACE_MEM_HOST)");
  if (ctx->trace_user) FAIL(TWOACE_E_UNSUPPORTED, "the residual trace is not available on a multi-GPU context");
  return 0;
}

// ------------------------------------------------------------------------------------------
template <class T>
static int upload_tasks(twoace_ctx* ctx, const std::vector<T>& v, size_t& cursor, const T** dptr) {
  // tasks are appended to ctx->taskbuf at `cursor`; the caller reserved enough space up front
  size_t o = (cursor + 255) / 256 * 256;
  size_t bytes = v.size() * sizeof(T);
  if (o + bytes > ctx->taskbuf.cap) FAIL(TWOACE_E_NOMEM, "internal: task buffer too small");
  CK(cudaMemcpyAsync((char*)ctx->taskbuf.p + o, v.data(), bytes, cudaMemcpyHostToDevice, ctx->stream));
  *dptr = (const T*)((char*)ctx->taskbuf.p + o);
  cursor = o + bytes;
  return 0;
}

static int stage_grid(twoace_ctx* ctx, size_t smem, int ntasks, int* grid) {
  int occ = 0;
  CK(gen_kernel_set_smem(smem));
  CK(gen_kernel_occupancy(&occ, smem));
  if (occ < 1) FAIL(TWOACE_E_UNSUPPORTED, "stage kernel does not fit: %zu bytes of shared memory", smem);
  *grid = std::max(1, std::min(ntasks, occ * ctx->num_sms));
  return 0;
}

constexpr size_t SMEM_LIMIT = (size_t)227 * 1024;
constexpr int NCOUNTER = 256;

static void push_stage_event(twoace_ctx* ctx, cudaEvent_t e0, cudaEvent_t e1, const char* label, bool group = false) {
  ctx->stage_events.emplace_back(e0, e1);
  std::string lb(label);
  if (ctx->in_side && !group) lb += "  (inside the concurrent group below: not summed)";
  ctx->stage_labels.emplace_back(lb);
  ctx->stage_side.resize(ctx->stage_events.size(), 0);
  ctx->stage_side.back() = (ctx->in_side && !group) ? 1 : 0;
}

// A zeroed task-queue counter for the next cluster-kernel launch on the context's current stream.
static int next_counter(twoace_ctx* ctx, int** out) {
  int rc = ensure(ctx, ctx->counters, NCOUNTER * sizeof(int));
  if (rc) return rc;
  int* p = (int*)ctx->counters.p + (ctx->counter_cursor++ % NCOUNTER);
  CK(cudaMemsetAsync(p, 0, sizeof(int), ctx->stream));
  *out = p;
  return 0;
}

template <int RL, int CS, bool TC>
static int launch_fast_t(twoace_ctx* ctx, const StageTask* dt, int ntasks, const DevParams& prm, FastDims fd,
                         bool* launched) {
  *launched = false;
  auto kern = fast_stage_kernel<RL, CS, TC>;
  size_t smem = fast_smem_bytes<RL>(fd);
  if (smem > SMEM_LIMIT) return 0;
  // the tensor-core kernel allocates all 512 tensor-memory columns: keep it alone on its SM
  if (TC) smem = std::max(smem, (size_t)117 * 1024);
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(NT, 1, 1);
  cfg.gridDim = dim3((unsigned)(CS * ntasks), 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int maxcl = 0;
  CK(cudaOccupancyMaxActiveClusters(&maxcl, kern, &cfg));
  if (maxcl < 1) return 0;
  const int ncl = std::max(1, std::min(ntasks, maxcl));
  fd.ws_stride = (fast_ws_elems(fd) + 15) / 16 * 16;
  int rc = ensure(ctx, ctx->ws, (size_t)ncl * fd.ws_stride * sizeof(cd));
  if (rc) return rc;
  cfg.gridDim = dim3((unsigned)(ncl * CS), 1, 1);
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (ctx->timing) {
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, ctx->stream));
  }
  int* counter = nullptr;
  rc = next_counter(ctx, &counter);
  if (rc) return rc;
  CK(cudaLaunchKernelEx(&cfg, kern, dt, ntasks, prm, fd, (cd*)ctx->ws.p, counter));
  if (ctx->timing) {
    CK(cudaEventRecord(e1, ctx->stream));
    char lb[160];
    snprintf(lb, sizeof lb, "fast_stage_kernel<%d,%d,%s> tasks %d clusters %d maxm %d smem %zu nslot %d n1 %d", RL, CS,
             TC ? "tc" : "simt", ntasks, ncl, fd.maxm, smem, fd.tc.nslot, fd.tc.n1);
    push_stage_event(ctx, e0, e1, lb);
  }
  ctx->launches++;
  ctx->fast_launches++;
  if (TC) ctx->tc_launches++;
  *launched = true;
  return 0;
}

// 256 < m <= 1024 on the chunked cluster kernel (big_stage.cuh)
static int launch_big(twoace_ctx* ctx, const StageTask* dt, int ntasks, const DevParams& prm, FastDims fd, bool* launched) {
  *launched = false;
  auto kern = big_stage_kernel;
  const size_t smem = std::max(fast_smem_bytes<BIG_RL>(fd), (size_t)117 * 1024);
  if (smem > SMEM_LIMIT) return 0;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(NT, 1, 1);
  cfg.gridDim = dim3((unsigned)(BIG_CS * ntasks), 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = BIG_CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int maxcl = 0;
  CK(cudaOccupancyMaxActiveClusters(&maxcl, kern, &cfg));
  if (maxcl < 1) return 0;
  const int ncl = std::max(1, std::min(ntasks, maxcl));
  fd.ws_stride = (big_ws_elems(fd.mfull) + 15) / 16 * 16;
  int rc = ensure(ctx, ctx->ws, (size_t)ncl * fd.ws_stride * sizeof(cd));
  if (rc) return rc;
  cfg.gridDim = dim3((unsigned)(ncl * BIG_CS), 1, 1);
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (ctx->timing) {
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, ctx->stream));
  }
  int* counter = nullptr;
  rc = next_counter(ctx, &counter);
  if (rc) return rc;
  CK(cudaLaunchKernelEx(&cfg, kern, dt, ntasks, prm, fd, (cd*)ctx->ws.p, counter));
  if (ctx->timing) {
    CK(cudaEventRecord(e1, ctx->stream));
    char lb[160];
    snprintf(lb, sizeof lb, "big_stage_kernel tasks %d clusters %d mfull %d smem %zu nslot %d n1 %d", ntasks, ncl, fd.mfull,
             smem, fd.tc.nslot, fd.tc.n1);
    push_stage_event(ctx, e0, e1, lb);
  }
  ctx->launches++;
  ctx->fast_launches++;
  ctx->tc_launches++;
  *launched = true;
  return 0;
}

// r = 1 stages with 256 < m <= 1024 (the refinement on all rows): one CTA per task, two per SM
static int launch_big1(twoace_ctx* ctx, const StageTask* dt, int ntasks, const DevParams& prm, FastDims fd, bool* launched) {
  *launched = false;
  auto kern = big1_stage_kernel;
  const size_t smem = fast_smem_bytes<1>(fd);
  if (smem > SMEM_LIMIT) return 0;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem));
  if (occ < 1) return 0;
  const int grid = std::max(1, std::min(ntasks, occ * ctx->num_sms));
  fd.ws_stride = (big1_ws_elems() + 15) / 16 * 16;
  int rc = ensure(ctx, ctx->ws, (size_t)grid * fd.ws_stride * sizeof(cd));
  if (rc) return rc;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (ctx->timing) {
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, ctx->stream));
  }
  kern<<<grid, NT, smem, ctx->stream>>>(dt, ntasks, prm, fd, (cd*)ctx->ws.p);
  CK(cudaGetLastError());
  if (ctx->timing) {
    CK(cudaEventRecord(e1, ctx->stream));
    char lb[160];
    snprintf(lb, sizeof lb, "big1_stage_kernel tasks %d grid %d (x%d per SM) maxm %d smem %zu", ntasks, grid, occ, fd.maxm, smem);
    push_stage_event(ctx, e0, e1, lb);
  }
  ctx->launches++;
  ctx->fast_launches++;
  *launched = true;
  return 0;
}

static bool big1_eligible(const twoace_ctx* ctx, const StageTask& t, int n, int tx, int rx) {
  return ctx->opt_fast && t.codes != nullptr && t.cscale != nullptr && n == FN && tx == FTX && rx == FTX &&
         t.m > 256 && t.m <= BIG_MMAX && t.r == 1 && !t.nuclear && (t.rank_one == 0 || t.rank_one == 1);
}

static bool big_eligible(const twoace_ctx* ctx, const StageTask& t, int n, int tx, int rx) {
  return ctx->opt_fast && ctx->opt_tensor && t.codes != nullptr && t.cscale != nullptr && n == FN && tx == FTX && rx == FTX &&
         t.m > 256 && t.m <= BIG_MMAX && t.r == BIG_R && !t.nuclear && (t.rank_one == 0 || t.rank_one == 1);
}

static bool fast_eligible(const twoace_ctx* ctx, const StageTask& t, int n, int tx, int rx) {
  return ctx->opt_fast && t.codes != nullptr && t.cscale != nullptr && n == FN && tx == FTX && rx == FTX &&
         t.m <= 256 && (t.r == 20 || t.r == 1) && (!t.nuclear || t.r == 1 || t.m >= 26) &&
         (t.rank_one == 0 || t.rank_one == 1);   // (the profiles of inferLowRank.m / V2.m: general kernel)
}

static int launch_stage_general(twoace_ctx* ctx, const std::vector<StageTask>& tasks, const DevParams& prm, int n,
                                int tx, int rx, size_t& cursor) {
  if (tasks.empty()) return 0;
  StageDims dm;
  dm.n = n; dm.tx = tx; dm.rx = rx; dm.maxm = 1; dm.maxr = 1; dm.dmax = 1;
  bool nuc = false;
  for (const StageTask& t : tasks) {
    dm.maxm = std::max(dm.maxm, t.m);
    dm.maxr = std::max(dm.maxr, t.r);
    dm.dmax = std::max(dm.dmax, use_woodbury(t.m, n) ? t.m : n);
    nuc = nuc || t.nuclear;
  }
  const StageTask* dt = nullptr;
  int rc = upload_tasks(ctx, tasks, cursor, &dt);
  if (rc) return rc;
  dm.ds = nuc ? std::max(tx, dm.maxr) : tx;
  if (dm.ds > SMALL_DMAX) FAIL(TWOACE_E_UNSUPPORTED, "tx (or r for the nuclear variant) > %d", SMALL_DMAX);
  dm.ws_stride = (stage_ws_elems(dm) + 15) / 16 * 16;
  // tasks that carry 2-bit codes of A keep them in shared memory when the rows of the largest task fit
  dm.wpr = 0;
  bool any_codes = false;
  for (const StageTask& t : tasks) any_codes = any_codes || (t.codes != nullptr && t.cscale != nullptr);
  if (any_codes && n % 16 == 0) {
    dm.wpr = n / 16;
    if (stage_smem_bytes(dm) > SMEM_LIMIT) dm.wpr = 0;
  }
  const size_t smem = stage_smem_bytes(dm);
  int grid = 0;
  rc = stage_grid(ctx, smem, (int)tasks.size(), &grid);
  if (rc) return rc;
  rc = ensure(ctx, ctx->ws, (size_t)grid * dm.ws_stride * sizeof(cd));
  if (rc) return rc;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (ctx->timing) {
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, ctx->stream));
  }
  CK(gen_kernel_launch(grid, smem, ctx->stream, dt, (int)tasks.size(), prm, dm, (cd*)ctx->ws.p));
  if (ctx->timing) {
    CK(cudaEventRecord(e1, ctx->stream));
    char lb[160];
    snprintf(lb, sizeof lb, "admm_stage_kernel tasks %zu grid %d maxm %d maxr %d%s", tasks.size(), grid, dm.maxm, dm.maxr,
             dm.wpr ? " codes" : "");
    push_stage_event(ctx, e0, e1, lb);
  }
  ctx->launches++;
  return 0;
}

// InferADMM of inferMinL2.m (no low-rank variable): general kernel with a per-CTA global workspace
static int launch_minl2(twoace_ctx* ctx, const std::vector<StageTask>& tasks, const DevParams& prm, int n, size_t& cursor) {
  if (tasks.empty()) return 0;
  Minl2Dims dm = {};
  dm.n = n; dm.maxm = 1; dm.maxr = 1;
  for (const StageTask& t : tasks) { dm.maxm = std::max(dm.maxm, t.m); dm.maxr = std::max(dm.maxr, t.r); }
  dm.ws_stride = (minl2_ws_elems(dm) + 15) / 16 * 16;
  const size_t smem = minl2_smem_bytes(dm);
  if (smem > SMEM_LIMIT) FAIL(TWOACE_E_UNSUPPORTED, "inferMinL2 stage kernel does not fit: %zu bytes of shared memory", smem);
  int occ = 0;
  CK(cudaFuncSetAttribute(minl2_stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, minl2_stage_kernel, NT, smem));
  if (occ < 1) FAIL(TWOACE_E_UNSUPPORTED, "inferMinL2 stage kernel does not fit: %zu bytes of shared memory", smem);
  const int grid = std::max(1, std::min((int)tasks.size(), occ * ctx->num_sms));
  int rc = ensure(ctx, ctx->ws, (size_t)grid * dm.ws_stride * sizeof(cd));
  if (rc) return rc;
  const StageTask* dt = nullptr;
  rc = upload_tasks(ctx, tasks, cursor, &dt);
  if (rc) return rc;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (ctx->timing) {
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, ctx->stream));
  }
  minl2_stage_kernel<<<grid, NT, smem, ctx->stream>>>(dt, (int)tasks.size(), prm, dm, (cd*)ctx->ws.p);
  CK(cudaGetLastError());
  if (ctx->timing) {
    CK(cudaEventRecord(e1, ctx->stream));
    char lb[160];
    snprintf(lb, sizeof lb, "minl2_stage_kernel tasks %d grid %d maxm %d maxr %d", (int)tasks.size(), grid, dm.maxm, dm.maxr);
    push_stage_event(ctx, e0, e1, lb);
  }
  ctx->launches++;
  return 0;
}

// Shared-memory geometry of the cluster kernel for an r = 20 task with m rows; false when it does not fit.
template <int RL>
static bool fast_dims(int m, bool nuc, bool tc, FastDims* out) {
  FastDims f = {};
  f.maxm = m; f.mw = (m + 15) / 16; f.r = 20; f.ws_stride = 0;
  f.nuclear = nuc ? 1 : 0; f.ds = nuc ? 20 : 16;
  if (tc) { if (!fast_tc_layout<RL>(f, SMEM_LIMIT)) return false; }
  else if (fast_smem_bytes<RL>(f) > SMEM_LIMIT) return false;
  if (out) *out = f;
  return true;
}

// One InferADMM launch.  Tasks that qualify for the shared-memory cluster kernel (16x16, quantised A,
// r in {20,1}, m <= 256) are split off and grouped by kernel configuration; the configuration of a task is
// decided by ITS OWN m (which layout fits in shared memory), never by its batch mates.
static int launch_stage(twoace_ctx* ctx, const std::vector<StageTask>& tasks, const DevParams& prm, int n,
                        int tx, int rx, size_t& cursor) {
  if (tasks.empty()) return 0;
  if (tasks.front().nuclear == 2) return launch_minl2(ctx, tasks, prm, n, cursor);     // a launch is all of one kind
  // groups: 0 = <10,2> tensor-core, 1 = <5,4> tensor-core, 2 = <10,2> SIMT, 3 = <5,4> SIMT, 4 = r = 1
  std::vector<StageTask> grp[5], gen, big, big1;
  bool nuc = false;
  for (const StageTask& t : tasks) nuc = nuc || t.nuclear;     // a launch is all-nuclear or all-V4
  const bool tc = ctx->opt_tensor != 0;
  for (const StageTask& t : tasks) {
    if (big_eligible(ctx, t, n, tx, rx)) { big.push_back(t); continue; }
    if (big1_eligible(ctx, t, n, tx, rx)) { big1.push_back(t); continue; }
    if (!fast_eligible(ctx, t, n, tx, rx)) { gen.push_back(t); continue; }
    if (t.r == 1) { grp[4].push_back(t); continue; }
    const bool want2 = ctx->opt_fast_cs == 2;
    const bool f2t = tc && fast_dims<10>(t.m, nuc, true, nullptr), f4t = tc && fast_dims<5>(t.m, nuc, true, nullptr);
    const bool f2 = fast_dims<10>(t.m, nuc, false, nullptr), f4 = fast_dims<5>(t.m, nuc, false, nullptr);
    if (want2 && f2t) grp[0].push_back(t);
    else if (f4t) grp[1].push_back(t);
    else if (f2t) grp[0].push_back(t);
    else if (want2 && f2) grp[2].push_back(t);
    else if (f4) grp[3].push_back(t);
    else if (f2) grp[2].push_back(t);
    else gen.push_back(t);
  }
  // The groups of a mixed launch are independent (different instances): they run concurrently -- the first on the
  // main stream, the others on side streams with their own workspaces -- and the cluster kernels take their tasks
  // from a device counter, so a group that gets SMs late simply processes what is left.  One group leaves SMs idle
  // (132 of 148 for clusters of 4, far fewer for a handful of tiny problems); together they fill the machine.
  int ngroups = (big.empty() ? 0 : 1) + (big1.empty() ? 0 : 1) + (gen.empty() ? 0 : 1);
  for (int g = 0; g < 5; ++g) ngroups += grp[g].empty() ? 0 : 1;
  const bool concurrent = ctx->opt_overlap && ctx->side_stream[0] != nullptr && ngroups >= 2;
  int slot = 0;
  bool side_used[twoace_ctx::NSIDE] = {false, false, false, false};
  cudaEvent_t g0 = nullptr, g1 = nullptr;
  if (concurrent) {
    if (ctx->timing) {
      CK(cudaEventCreate(&g0));
      CK(cudaEventCreate(&g1));
      CK(cudaEventRecord(g0, ctx->stream));
    }
    CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
    ctx->in_side = true;
  }
  auto run_group = [&](auto&& fn) -> int {
    const int my = slot++;
    if (!concurrent || my == 0) return fn();
    const int k = (my - 1) % twoace_ctx::NSIDE;
    if (!side_used[k]) {
      if (cudaStreamWaitEvent(ctx->side_stream[k], ctx->ev_fork, 0) != cudaSuccess) { ctx->err = "cudaStreamWaitEvent failed"; return TWOACE_E_CUDA; }
      side_used[k] = true;
    }
    std::swap(ctx->stream, ctx->side_stream[k]);
    std::swap(ctx->ws, ctx->ws_side[k]);
    const int rc = fn();
    std::swap(ctx->ws, ctx->ws_side[k]);
    std::swap(ctx->stream, ctx->side_stream[k]);
    return rc;
  };
  auto finish = [&](int rc) -> int {
    if (!concurrent) return rc;
    ctx->in_side = false;
    for (int k = 0; k < twoace_ctx::NSIDE; ++k) {
      if (!side_used[k]) continue;
      if (cudaEventRecord(ctx->ev_join[k], ctx->side_stream[k]) != cudaSuccess ||
          cudaStreamWaitEvent(ctx->stream, ctx->ev_join[k], 0) != cudaSuccess) {
        if (!rc) { ctx->err = "joining the side streams failed"; rc = TWOACE_E_CUDA; }
      }
    }
    if (g0) {
      if (!rc && cudaEventRecord(g1, ctx->stream) == cudaSuccess) {
        char lb[96];
        snprintf(lb, sizeof lb, "launch group: the %d kernels above, concurrently", ngroups);
        push_stage_event(ctx, g0, g1, lb, true);
      } else {
        cudaEventDestroy(g0);
        cudaEventDestroy(g1);
      }
    }
    return rc;
  };
  int rc = 0;
  if (!big.empty()) {
    rc = run_group([&]() -> int {
      std::stable_sort(big.begin(), big.end(), [](const StageTask& a, const StageTask& b) { return a.m > b.m; });
      FastDims fd = {};
      fd.maxm = BIG_CH; fd.mw = BIG_CH / 16; fd.r = BIG_R; fd.ws_stride = 0; fd.nuclear = 0; fd.ds = 16;
      fd.mfull = big.front().m;
      if (!fast_tc_layout<BIG_RL>(fd, SMEM_LIMIT)) FAIL(TWOACE_E_CUDA, "internal: chunked cluster kernel layout (mfull %d)", fd.mfull);
      const StageTask* dt = nullptr;
      int r2 = upload_tasks(ctx, big, cursor, &dt);
      if (r2) return r2;
      bool launched = false;
      r2 = launch_big(ctx, dt, (int)big.size(), prm, fd, &launched);
      if (r2) return r2;
      if (!launched) FAIL(TWOACE_E_CUDA, "chunked cluster kernel launch configuration rejected (mfull %d)", fd.mfull);
      return 0;
    });
    if (rc) return finish(rc);
  }
  const int order[5] = {1, 3, 0, 2, 4};     // clusters of 4 first: they need four free SMs of one GPC
  for (int gi = 0; gi < 5; ++gi) {
    const int g = order[gi];
    std::vector<StageTask>& ft = grp[g];
    if (ft.empty()) continue;
    rc = run_group([&]() -> int {
      // longest tasks first: the queue hands them out in this order, which balances the tail of the launch
      std::stable_sort(ft.begin(), ft.end(), [](const StageTask& a, const StageTask& b) { return a.m > b.m; });
      int maxm = 1;
      for (const StageTask& t : ft) maxm = std::max(maxm, t.m);
      FastDims fd = {};
      bool ok = true;
      if (g == 4) {
        fd.maxm = maxm; fd.mw = (maxm + 15) / 16; fd.r = 1; fd.ws_stride = 0; fd.nuclear = nuc ? 1 : 0; fd.ds = 16;
      } else if (g == 0 || g == 2) ok = fast_dims<10>(maxm, nuc, g == 0, &fd);
      else ok = fast_dims<5>(maxm, nuc, g == 1, &fd);
      if (!ok) FAIL(TWOACE_E_CUDA, "internal: cluster kernel layout (group %d, maxm %d)", g, maxm);
      const StageTask* dt = nullptr;
      int r2 = upload_tasks(ctx, ft, cursor, &dt);
      if (r2) return r2;
      bool launched = false;
      switch (g) {
        case 0: r2 = launch_fast_t<10, 2, true>(ctx, dt, (int)ft.size(), prm, fd, &launched); break;
        case 1: r2 = launch_fast_t<5, 4, true>(ctx, dt, (int)ft.size(), prm, fd, &launched); break;
        case 2: r2 = launch_fast_t<10, 2, false>(ctx, dt, (int)ft.size(), prm, fd, &launched); break;
        case 3: r2 = launch_fast_t<5, 4, false>(ctx, dt, (int)ft.size(), prm, fd, &launched); break;
        default: r2 = launch_fast_t<1, 1, false>(ctx, dt, (int)ft.size(), prm, fd, &launched); break;
      }
      if (r2) return r2;
      if (!launched) FAIL(TWOACE_E_CUDA, "cluster kernel launch configuration rejected (group %d, maxm %d)", g, fd.maxm);
      return 0;
    });
    if (rc) return finish(rc);
  }
  if (!big1.empty()) {
    rc = run_group([&]() -> int {
      std::stable_sort(big1.begin(), big1.end(), [](const StageTask& a, const StageTask& b) { return a.m > b.m; });
      FastDims fd = {};
      fd.maxm = big1.front().m; fd.mw = (fd.maxm + 15) / 16; fd.r = 1; fd.ds = 16; fd.lean = 1;
      const StageTask* dt = nullptr;
      int r2 = upload_tasks(ctx, big1, cursor, &dt);
      if (r2) return r2;
      bool launched = false;
      r2 = launch_big1(ctx, dt, (int)big1.size(), prm, fd, &launched);
      if (r2) return r2;
      if (!launched) FAIL(TWOACE_E_CUDA, "r = 1 large-m kernel launch configuration rejected (maxm %d)", fd.maxm);
      return 0;
    });
    if (rc) return finish(rc);
  }
  if (!gen.empty()) rc = run_group([&]() -> int { return launch_stage_general(ctx, gen, prm, n, tx, rx, cursor); });
  return finish(rc);
}

static int launch_spectral(twoace_ctx* ctx, const std::vector<SpecTask>& tasks, int n, size_t& cursor) {
  if (tasks.empty()) return 0;
  SpecDims dm = {};
  dm.n = n; dm.maxm = 1; dm.dmax = 1; dm.force_jacobi = ctx->opt_spectral_jacobi;
  for (const SpecTask& t : tasks) {
    dm.maxm = std::max(dm.maxm, t.m);
    dm.dmax = std::max(dm.dmax, t.m <= n ? t.m : n);
  }
  dm.ws_stride = (spec_ws_elems(dm) + 15) / 16 * 16;
  const size_t smem = spec_smem_bytes(dm);
  int occ = 0;
  CK(cudaFuncSetAttribute(spectral_init_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spectral_init_kernel, NT, smem));
  if (occ < 1) FAIL(TWOACE_E_UNSUPPORTED, "spectral kernel does not fit: %zu bytes of shared memory", smem);
  const int grid = std::max(1, std::min((int)tasks.size(), occ * ctx->num_sms));
  int rc = ensure(ctx, ctx->ws, (size_t)grid * dm.ws_stride * sizeof(cd));
  if (rc) return rc;
  const SpecTask* dt = nullptr;
  rc = upload_tasks(ctx, tasks, cursor, &dt);
  if (rc) return rc;
  spectral_init_kernel<<<grid, NT, smem, ctx->stream>>>(dt, (int)tasks.size(), dm, (cd*)ctx->ws.p);
  CK(cudaGetLastError());
  ctx->launches++;
  return 0;
}

static int launch_ortho(twoace_ctx* ctx, const std::vector<OrthoTask>& tasks, int n, size_t& cursor) {
  if (tasks.empty()) return 0;
  int rmax = 1;
  for (const OrthoTask& t : tasks) rmax = std::max(rmax, t.r);
  const size_t smem = ortho_smem_bytes(rmax);
  CK(cudaFuncSetAttribute(ortho_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = std::max(1, std::min((int)tasks.size(), 2 * ctx->num_sms));
  const OrthoTask* dt = nullptr;
  int rc = upload_tasks(ctx, tasks, cursor, &dt);
  if (rc) return rc;
  ortho_kernel<<<grid, NT, smem, ctx->stream>>>(dt, (int)tasks.size(), n, rmax);
  CK(cudaGetLastError());
  ctx->launches++;
  return 0;
}

// info words (twoace.h) from the control blocks and the stage bookkeeping
__global__ void info_kernel(const InstCtl* ctl, const double* stage_words, int nstage, int nb, double* info,
                            double* quality) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const InstCtl c = ctl[b];
  if (quality) quality[b] = c.quality;
  if (!info) return;
  double* o = info + (size_t)b * TWOACE_INFO_WORDS;
  const double* sw = stage_words + (size_t)b * nstage * STAGE_SCAL;
  const double* rf = sw + (size_t)(nstage - 1) * STAGE_SCAL;
  double tot = 0.0;
  for (int s = 0; s < nstage; ++s) tot += sw[(size_t)s * STAGE_SCAL + SC_ITERS];
  o[0] = c.quality; o[1] = c.similarity; o[2] = c.use_rank_one; o[3] = c.rolled_back;
  o[4] = c.best_trial; o[5] = c.y_rows; o[6] = c.trial_r1_mask;
  o[7] = c.trial_quality[0]; o[8] = c.trial_quality[1]; o[9] = c.trial_quality[2];
  o[10] = c.max_quality; o[11] = rf[SC_ITERS]; o[12] = rf[SC_OPT_ITER]; o[13] = rf[SC_MU];
  o[14] = rf[SC_BUMPS]; o[15] = tot;
}

static DevParams make_dev_params(const twoace_params& p) {
  DevParams d;
  d.mu0 = p.mu0; d.rho = p.rho; d.tol_rel = p.tol_rel; d.tol_abs = p.tol_abs; d.maxiter = p.maxiter;
  d.need_dual = (p.tol_rel != 0.0 || p.tol_abs != 0.0) ? 1 : 0;
  return d;
}

static int check_params(twoace_ctx* ctx, const twoace_params& p, int tx, int rx) {
  if (p.lambda != 0.0) FAIL(TWOACE_E_UNSUPPORTED, "lambda != 0 is not supported (dead path in the reference)");
  if (p.r < 1 || p.r > SMALL_DMAX) FAIL(TWOACE_E_INVALID, "r must be in [1,%d]", SMALL_DMAX);
  if (!(p.cc_frac > 0.0 && p.cc_frac < 1.0)) FAIL(TWOACE_E_INVALID, "cc_frac must be in (0,1)");
  if (p.maxiter < 1) FAIL(TWOACE_E_INVALID, "maxiter must be >= 1");
  if (!(p.mu0 > 0.0) || !(p.rho > 0.0)) FAIL(TWOACE_E_INVALID, "mu0 and rho must be positive");
  if (tx < 4 || tx > SMALL_DMAX || (tx % 4) != 0) FAIL(TWOACE_E_UNSUPPORTED, "tx must be a multiple of 4 in [4,%d]", SMALL_DMAX);
  if (rx < 1) FAIL(TWOACE_E_INVALID, "rx must be >= 1");
  return 0;
}

// ------------------------------------------------------------------------------------------
struct ChunkIn {
  int variant, nb, tx, rx;
  const int32_t* m;          // host
  const cd* dA;              // device, dense mode (concat col-major) or nullptr
  const int32_t* cb_rows;    // host, codebook mode or nullptr
  double row_scale;
  const double* dB;          // device
  const int32_t* train_idx;  // host
  twoace_params p;
  cd* dX; cd* dY; double* dQ; double* dInfo; double* dStage;  // device outputs (dInfo/dStage may be null)
  double* dTrace;            // device, [nb][nstage][maxiter] or nullptr
};

static int solve_chunk(twoace_ctx* ctx, const ChunkIn& in) {
  const int nb = in.nb, n = in.tx * in.rx;
  const bool multi = in.variant == TWOACE_V4_MULTI;
  const int nuclear = in.variant == TWOACE_NUCLEAR ? 1 : 0;
  const int T = multi ? 3 : 1;
  const int nstage = 4 * T + 1;
  // older solver versions (ADMM_v2.m:26-31): no rank-one rerun; V1 / V2 refine only when quality > 0.6; their
  // rank profile travels in StageTask.rank_one (2: inferLowRank.m, 3: inferLowRankV2.m -- identical to the V4
  // profile once ceil(0.7 sqrt(min(tx,rx))) > 2, i.e. from 9 antennas on)
  // inferMinL2.m (ADMM_v2.m:23, version 0): the same shell (train solve, held-out quality, refine only if quality > 0.6,
  // roll-back) around a solver without the low-rank variable; train rows = ceil(0.95 m) (:34), rank from the 90 % rule
  const bool minl2 = in.variant == TWOACE_MINL2;
  const bool older = in.variant == TWOACE_V3 || in.variant == TWOACE_V2 || in.variant == TWOACE_V1 || minl2;
  const int refine_if_good = (in.variant == TWOACE_V2 || in.variant == TWOACE_V1 || minl2) ? 1 : 0;
  const int prof = in.variant == TWOACE_V1 ? 2 : (in.variant == TWOACE_V2 && std::min(in.tx, in.rx) < 9) ? 3 : 0;
  const bool dense = in.dA != nullptr;

  // ---- host bookkeeping (integer index work is bit-exact host code)
  std::vector<size_t> a_off(nb + 1, 0), b_off(nb + 1, 0), tr_off(nb + 1, 0), te_off(nb + 1, 0);
  std::vector<int> mtr(nb), mte(nb), rb(nb);
  int maxr = 1;
  for (int b = 0; b < nb; ++b) {
    const int m = in.m[b];
    if (m < 2) FAIL(TWOACE_E_INVALID, "instance %d: m = %d (need m >= 2)", b, m);
    mtr[b] = minl2 ? (int)std::ceil((double)m * 0.95) : (int)std::floor((double)m * in.p.cc_frac);
    mte[b] = m - mtr[b];
    // (inferMinL2 with m <= 20 has an empty test set: quality = 1 - 0/0 = NaN, no refinement, as in MATLAB)
    if (mtr[b] < 1 || (mte[b] < 1 && !minl2)) FAIL(TWOACE_E_INVALID, "instance %d: empty train or test split", b);
    // r = min([r m n]) (:12); versions 1-3 re-evaluate it inside inferLowRankImpl with m = m_train, BEFORE their
    // spectral init (inferLowRankV3.m:198-224), V4 / _multi / _Nuclear initialise outside with the full-m clamp
    rb[b] = std::min(std::min((int)in.p.r, (older && !minl2) ? mtr[b] : m), n);
    maxr = std::max(maxr, rb[b]);
    a_off[b + 1] = a_off[b] + (size_t)m * n;
    b_off[b + 1] = b_off[b] + m;
    tr_off[b + 1] = tr_off[b] + (size_t)T * mtr[b];
    te_off[b + 1] = te_off[b] + (size_t)T * mte[b];
  }
  // index lists: trainB/testB index the instance's B (and dense A rows); trainA/testA index the
  // matrix the AView points at (== trainB/testB in dense mode, codebook rows otherwise)
  std::vector<int32_t> trainB(tr_off[nb]), testB(te_off[nb]), trainA, testA, fullA;
  if (!dense) { trainA.resize(tr_off[nb]); testA.resize(te_off[nb]); fullA.resize(b_off[nb]); }
  std::vector<char> mark;
  for (int b = 0; b < nb; ++b) {
    const int m = in.m[b];
    for (int t = 0; t < T; ++t) {
      const int32_t* tr = in.train_idx + tr_off[b] + (size_t)t * mtr[b];
      mark.assign(m, 0);
      for (int i = 0; i < mtr[b]; ++i) {
        const int v = tr[i];
        if (v < 0 || v >= m || mark[v]) FAIL(TWOACE_E_INVALID, "instance %d trial %d: train_idx must be unique and in [0,m)", b, t);
        mark[v] = 1;
        trainB[tr_off[b] + (size_t)t * mtr[b] + i] = v;
      }
      int k = 0;
      for (int v = 0; v < m; ++v)
        if (!mark[v]) testB[te_off[b] + (size_t)t * mte[b] + k++] = v;   // setdiff(1:m, train): ascending
    }
    if (!dense) {
      const int32_t* rows = in.cb_rows + b_off[b];
      for (int i = 0; i < m; ++i) {
        if (rows[i] < 0 || rows[i] >= ctx->cb_rows) FAIL(TWOACE_E_INVALID, "instance %d: codebook row %d out of range", b, rows[i]);
        fullA[b_off[b] + i] = rows[i];
      }
      for (size_t i = 0; i < (size_t)T * mtr[b]; ++i) trainA[tr_off[b] + i] = rows[trainB[tr_off[b] + i]];
      for (size_t i = 0; i < (size_t)T * mte[b]; ++i) testA[te_off[b] + i] = rows[testB[te_off[b] + i]];
    }
  }

  // ---- device arena layout
  Bump bp;
  const size_t o_trainB = bp.take(trainB.size() * 4), o_testB = bp.take(testB.size() * 4);
  const size_t o_trainA = dense ? o_trainB : bp.take(trainA.size() * 4);
  const size_t o_testA = dense ? o_testB : bp.take(testA.size() * 4);
  const size_t o_fullA = dense ? 0 : bp.take(fullA.size() * 4);
  const size_t o_ctl = bp.take((size_t)nb * sizeof(InstCtl));
  const size_t o_Arm = dense ? bp.take(a_off[nb] * sizeof(cd)) : 0;
  const size_t xstride = (size_t)n * maxr;
  const size_t o_Xs = bp.take((size_t)nb * xstride * sizeof(cd));
  const size_t o_Xa = bp.take((size_t)nb * xstride * sizeof(cd));
  const size_t o_xb = bp.take((size_t)nb * n * sizeof(cd));
  const size_t o_xmax = bp.take((size_t)nb * n * sizeof(cd));
  const size_t o_xr = bp.take((size_t)nb * n * sizeof(cd));
  const size_t o_yb = bp.take(b_off[nb] * sizeof(cd));
  const size_t o_ymax = bp.take(b_off[nb] * sizeof(cd));
  const size_t o_yr = bp.take(b_off[nb] * sizeof(cd));
  const size_t o_sw = bp.take((size_t)nb * nstage * STAGE_SCAL * sizeof(double));
  // 2-bit codes of A: the cluster kernels need them (16 x 16 arrays), the general kernel keeps them in shared memory
  const bool try_fast = ctx->opt_fast && n == FN && in.tx == FTX && in.rx == FTX;
  const bool try_codes = ctx->opt_fast && n % 16 == 0;
  const int wpr = n / 16;
  const size_t o_codes = (dense && try_codes) ? bp.take(b_off[nb] * wpr * sizeof(uint32_t)) : 0;
  const size_t o_qflag = (dense && try_codes) ? bp.take((size_t)nb * sizeof(int)) : 0;
  // per-instance store of (I + A_train A_train')^-1: the stages of a trial (over-parameterised, refinement and their
  // rank-one reruns) share the training rows, so only the first of them inverts (instances on the cluster kernels;
  // m x m Woodbury core for m <= 256, the n x n matrix (A'A + I)^-1 of the chunked kernel above)
  std::vector<size_t> sv_off(nb + 1, 0);
  for (int b = 0; b < nb; ++b) sv_off[b + 1] = sv_off[b] + ((try_fast && ctx->opt_cache_sinv && mtr[b] <= BIG_MMAX) ? (mtr[b] <= 256 ? (size_t)mtr[b] * mtr[b] : (size_t)FN * FN) : 0);
  const size_t o_sinv = sv_off[nb] ? bp.take(sv_off[nb] * sizeof(cd)) : 0;
  int rc = ensure(ctx, ctx->arena, bp.off + 256);
  if (rc) return rc;
  char* base = (char*)ctx->arena.p;
  int* d_trainB = (int*)(base + o_trainB); int* d_testB = (int*)(base + o_testB);
  int* d_trainA = (int*)(base + o_trainA); int* d_testA = (int*)(base + o_testA);
  int* d_fullA = dense ? nullptr : (int*)(base + o_fullA);
  InstCtl* d_ctl = (InstCtl*)(base + o_ctl);
  cd* d_Arm = dense ? (cd*)(base + o_Arm) : nullptr;
  cd* d_Xs = (cd*)(base + o_Xs); cd* d_Xa = (cd*)(base + o_Xa);
  cd* d_xb = (cd*)(base + o_xb); cd* d_xmax = (cd*)(base + o_xmax); cd* d_xr = (cd*)(base + o_xr);
  cd* d_yb = (cd*)(base + o_yb); cd* d_ymax = (cd*)(base + o_ymax); cd* d_yr = (cd*)(base + o_yr);
  double* d_sw = (double*)(base + o_sw);

  CK(cudaMemcpyAsync(d_trainB, trainB.data(), trainB.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_testB, testB.data(), testB.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (!dense) {
    CK(cudaMemcpyAsync(d_trainA, trainA.data(), trainA.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_testA, testA.data(), testA.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_fullA, fullA.data(), fullA.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  }
  CK(cudaMemsetAsync(d_sw, 0, (size_t)nb * nstage * STAGE_SCAL * sizeof(double), ctx->stream));

  // task buffer: generous upper bound for all task arrays of this chunk
  const size_t task_bytes = (size_t)nb * (sizeof(PrepTask) + sizeof(QuantTask) + T * (sizeof(SpecTask) + 8 * sizeof(StageTask) +
                            2 * sizeof(OrthoTask) + 2 * sizeof(QualTask)) + 2 * sizeof(StageTask) + sizeof(FinalTask)) +
                            256 * (16 * T + 16);
  rc = ensure(ctx, ctx->taskbuf, task_bytes);
  if (rc) return rc;
  size_t cursor = 0;
  const DevParams prm = make_dev_params(in.p);
  const cd* Abase_all = dense ? d_Arm : ctx->cb_rm;
  std::vector<char> inst_codes(nb, 0);     // per instance: the 2-bit representation of its A exists

  // ---- pre-processing
  {
    std::vector<PrepTask> pt(nb);
    for (int b = 0; b < nb; ++b) {
      PrepTask& t = pt[b];
      t.A_cm = dense ? in.dA + a_off[b] : nullptr;
      t.A_rm = dense ? d_Arm + a_off[b] : nullptr;
      t.cb = ctx->cb_rm; t.cbrows = dense ? nullptr : d_fullA + b_off[b];
      t.row_scale = in.row_scale; t.B = in.dB + b_off[b]; t.m = in.m[b]; t.ctl = d_ctl + b;
      t.code_mag = dense ? 0.0 : ctx->cb_mag;
    }
    const PrepTask* dt = nullptr;
    rc = upload_tasks(ctx, pt, cursor, &dt);
    if (rc) return rc;
    prep_kernel<<<std::min(nb, 8 * ctx->num_sms), NT, 0, ctx->stream>>>(dt, nb, n, in.p.tol_abs);
    CK(cudaGetLastError());
    ctx->launches++;
    if (dense && try_codes) {   // 2-bit phase codes of every instance's A (one batched flag read per chunk)
      std::vector<QuantTask> qt(nb);
      for (int b = 0; b < nb; ++b) {
        QuantTask& q = qt[b];
        q.A_rm = d_Arm + a_off[b]; q.rows = in.m[b]; q.wpr = wpr; q.codes = (uint32_t*)(base + o_codes) + b_off[b] * wpr;
        q.mag_out = nullptr; q.flag_out = (int*)(base + o_qflag) + b; q.ctl = d_ctl + b;
      }
      const QuantTask* dq = nullptr;
      rc = upload_tasks(ctx, qt, cursor, &dq);
      if (rc) return rc;
      quant_kernel<<<std::min(nb, 8 * ctx->num_sms), NT, 0, ctx->stream>>>(dq, nb);
      CK(cudaGetLastError());
      ctx->launches++;
      std::vector<int> flags(nb);
      CK(cudaMemcpyAsync(flags.data(), base + o_qflag, (size_t)nb * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      // decided per instance: the kernel an instance runs on never depends on its batch mates
      for (int b = 0; b < nb; ++b) inst_codes[b] = flags[b] == 1;
    } else if (!dense && try_codes) {
      std::fill(inst_codes.begin(), inst_codes.end(), (char)(ctx->cb_codes != nullptr));
    }
    fill_nan_kernel<<<std::min(4 * ctx->num_sms, (int)(((size_t)nb * n + 255) / 256)), 256, 0, ctx->stream>>>(d_xmax, (size_t)nb * n);
    CK(cudaGetLastError());
    ctx->launches++;
  }

  auto aview = [&](int b, const int* rows) {
    AView v;
    v.base = dense ? d_Arm + a_off[b] : Abase_all;
    v.rows = rows;
    v.scale = &d_ctl[b].a_scale;
    return v;
  };

  for (int t = 0; t < T; ++t) {
    // ---- spectral initialisation on the training rows
    {
      std::vector<SpecTask> st(nb);
      for (int b = 0; b < nb; ++b) {
        SpecTask& s = st[b];
        s.A = aview(b, d_trainA + tr_off[b] + (size_t)t * mtr[b]);
        s.B = in.dB + b_off[b]; s.brows = d_trainB + tr_off[b] + (size_t)t * mtr[b];
        s.bscale = &d_ctl[b].b_scale; s.m = mtr[b]; s.r = rb[b];
        s.Xs = d_Xs + (size_t)b * xstride; s.sweeps = nullptr;
        s.r_out = minl2 ? &d_ctl[b].r_eff : nullptr;
      }
      rc = launch_spectral(ctx, st, n, cursor);
      if (rc) return rc;
    }
    for (int pass = 0; pass < 2; ++pass) {   // pass 1 = rank-one rerun, masked by ctl.need_r1
      std::vector<StageTask> sa(nb), sb(nb);
      std::vector<OrthoTask> ot(nb);
      std::vector<QualTask> qt(nb);
      for (int b = 0; b < nb; ++b) {
        StageTask& a = sa[b];
        a.A = aview(b, d_trainA + tr_off[b] + (size_t)t * mtr[b]);
        a.B = in.dB + b_off[b]; a.brows = d_trainB + tr_off[b] + (size_t)t * mtr[b];
        a.bscale = &d_ctl[b].b_scale; a.m = mtr[b]; a.r = rb[b];
        a.X0 = d_Xs + (size_t)b * xstride; a.Xout = d_Xa + (size_t)b * xstride; a.Yout = nullptr;
        a.sbr = 1; a.rank_one = pass ? 1 : prof; a.nuclear = minl2 ? 2 : nuclear; a.rank_one_ptr = nullptr;
        a.r_ptr = minl2 ? &d_ctl[b].r_eff : nullptr;
        a.active = pass ? &d_ctl[b].need_r1 : nullptr; a.active_expect = 1;
        a.scal = d_sw + ((size_t)b * nstage + 4 * t + 2 * pass) * STAGE_SCAL; a.state = nullptr;
        a.trace = in.dTrace ? in.dTrace + ((size_t)b * nstage + 4 * t + 2 * pass) * in.p.maxiter : nullptr;
        const bool use_codes = inst_codes[b] != 0;
        a.codes = !use_codes ? nullptr : (dense ? (const uint32_t*)(base + o_codes) + b_off[b] * wpr : ctx->cb_codes);
        a.cscale = use_codes ? &d_ctl[b].c_scale : nullptr;
        if (use_codes && sv_off[b + 1] > sv_off[b]) { a.sinv = (cd*)(base + o_sinv) + sv_off[b]; a.sinv_state = pass ? 1 : 0; }
        StageTask& s2 = sb[b];
        s2 = a;
        if (s2.sinv) s2.sinv_state = 1;
        s2.X0 = d_Xa + (size_t)b * xstride; s2.Xout = d_xb + (size_t)b * n; s2.Yout = d_yb + b_off[b];
        s2.sbr = 0;
        s2.scal = d_sw + ((size_t)b * nstage + 4 * t + 2 * pass + 1) * STAGE_SCAL;
        s2.trace = in.dTrace ? in.dTrace + ((size_t)b * nstage + 4 * t + 2 * pass + 1) * in.p.maxiter : nullptr;
        OrthoTask& o = ot[b];
        o.X = d_Xa + (size_t)b * xstride; o.r = rb[b]; o.r_ptr = a.r_ptr; o.active = a.active; o.active_expect = 1;
        QualTask& q = qt[b];
        q.A = aview(b, d_testA + te_off[b] + (size_t)t * mte[b]);
        q.B = in.dB + b_off[b]; q.brows = d_testB + te_off[b] + (size_t)t * mte[b];
        q.bscale = &d_ctl[b].b_scale; q.mte = mte[b]; q.x = d_xb + (size_t)b * n; q.y = d_yb + b_off[b];
        q.mtr = mtr[b]; q.xmax = d_xmax + (size_t)b * n; q.ymax = d_ymax + b_off[b];
        q.ctl = d_ctl + b; q.trial = t; q.pass = pass; q.multi = multi ? 1 : 0;
        q.allow_r1 = older ? 0 : 1; q.refine_if_good = refine_if_good;
      }
      // Older versions have no rank-one rerun; pass 1 only does the bookkeeping.  inferLowRank_Nuclear.m:69-70
      // reruns inferLowRankImpl with use_rank_one = true, but its ArgMinZ (:411-419) ignores that flag, so the
      // rerun recomputes the first run bit for bit; with the opt-in "dedup_nuclear_rerun" it is not re-executed
      // (same X, Y, quality and flags; the stage words of the elided stages stay zero).
      if (!(pass == 1 && (older || (nuclear && ctx->opt_dedup_nuclear)))) {
        rc = launch_stage(ctx, sa, prm, n, in.tx, in.rx, cursor);
        if (rc) return rc;
        rc = launch_ortho(ctx, ot, n, cursor);
        if (rc) return rc;
        rc = launch_stage(ctx, sb, prm, n, in.tx, in.rx, cursor);
        if (rc) return rc;
      }
      const QualTask* dq = nullptr;
      rc = upload_tasks(ctx, qt, cursor, &dq);
      if (rc) return rc;
      quality_kernel<<<std::min(nb, 8 * ctx->num_sms), NT, 0, ctx->stream>>>(dq, nb, n);
      CK(cudaGetLastError());
      ctx->launches++;
    }
  }
  // ---- refine on all rows (:68-80), then roll-back / rescale (:72-86)
  {
    std::vector<StageTask> sr(nb);
    std::vector<FinalTask> ft(nb);
    for (int b = 0; b < nb; ++b) {
      StageTask& a = sr[b];
      a.A = aview(b, dense ? nullptr : d_fullA + b_off[b]);
      a.B = in.dB + b_off[b]; a.brows = nullptr; a.bscale = &d_ctl[b].b_scale;
      a.m = in.m[b]; a.r = 1; a.X0 = d_xmax + (size_t)b * n; a.Xout = d_xr + (size_t)b * n; a.Yout = d_yr + b_off[b];
      a.sbr = 1; a.rank_one = prof; a.nuclear = minl2 ? 2 : nuclear; a.rank_one_ptr = older ? nullptr : &d_ctl[b].use_rank_one;
      a.active = refine_if_good ? &d_ctl[b].refine_on : nullptr; a.active_expect = 1;
      a.scal = d_sw + ((size_t)b * nstage + (nstage - 1)) * STAGE_SCAL; a.state = nullptr;
      a.trace = in.dTrace ? in.dTrace + ((size_t)b * nstage + (nstage - 1)) * in.p.maxiter : nullptr;
      const bool use_codes = inst_codes[b] != 0;
      a.codes = !use_codes ? nullptr : (dense ? (const uint32_t*)(base + o_codes) + b_off[b] * wpr : ctx->cb_codes);
      a.cscale = use_codes ? &d_ctl[b].c_scale : nullptr;
      FinalTask& f = ft[b];
      f.x0 = d_xmax + (size_t)b * n; f.y0 = d_ymax + b_off[b]; f.xr = d_xr + (size_t)b * n; f.yr = d_yr + b_off[b];
      f.m = in.m[b]; f.mtr = mtr[b]; f.Xout = in.dX + (size_t)b * n; f.Yout = in.dY + b_off[b];
      f.quality_out = in.dQ + b; f.ctl = d_ctl + b; f.refine_if_good = refine_if_good;
    }
    rc = launch_stage(ctx, sr, prm, n, in.tx, in.rx, cursor);
    if (rc) return rc;
    const FinalTask* df = nullptr;
    rc = upload_tasks(ctx, ft, cursor, &df);
    if (rc) return rc;
    final_kernel<<<std::min(nb, 8 * ctx->num_sms), NT, 0, ctx->stream>>>(df, nb, n);
    CK(cudaGetLastError());
    ctx->launches++;
  }
  if (in.dInfo) {
    info_kernel<<<(nb + 127) / 128, 128, 0, ctx->stream>>>(d_ctl, d_sw, nstage, nb, in.dInfo, nullptr);
    CK(cudaGetLastError());
    ctx->launches++;
  }
  if (in.dStage)
    CK(cudaMemcpyAsync(in.dStage, d_sw, (size_t)nb * nstage * STAGE_SCAL * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  return 0;
}

// Host-pointer staging helper: device mirrors of inputs/outputs for one call.  The buffers come from a small per-context
// cache and go back to it when the call ends (cudaMalloc / cudaFree per call cost up to several hundred ms of host time
// now and then on the e2e path; every user of a staged buffer is ordered on the context's stream, so reuse is safe).
struct Staging {
  twoace_ctx* ctx = nullptr;
  std::vector<std::pair<void*, size_t>> owned;
  ~Staging() {
    for (auto& pr : owned) {
      if (ctx) ctx->stage_cache.push_back(pr);
      else cudaFree(pr.first);
    }
  }
};

static void* stage_alloc(twoace_ctx* ctx, Staging& st, size_t bytes) {
  st.ctx = ctx;
  const size_t want = (bytes + 262143) / 262144 * 262144;
  int best = -1;
  for (int i = 0; i < (int)ctx->stage_cache.size(); ++i) {
    const size_t cap = ctx->stage_cache[i].second;
    if (cap >= want && cap <= 2 * want + (1u << 20) && (best < 0 || cap < ctx->stage_cache[best].second)) best = i;
  }
  if (best >= 0) {
    auto pr = ctx->stage_cache[best];
    ctx->stage_cache.erase(ctx->stage_cache.begin() + best);
    st.owned.push_back(pr);
    return pr.first;
  }
  void* p = nullptr;
  if (cudaMalloc(&p, want) != cudaSuccess) {
    (void)cudaGetLastError();
    // release the cache and retry once
    for (auto& pr : ctx->stage_cache) cudaFree(pr.first);
    ctx->stage_cache.clear();
    if (cudaMalloc(&p, want) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
  }
  st.owned.emplace_back(p, want);
  return p;
}

static int dev_in(twoace_ctx* ctx, Staging& st, int mem, const void* src, size_t bytes, const void** dst) {
  if (mem == TWOACE_MEM_DEVICE || bytes == 0) { *dst = src; return 0; }
  void* p = stage_alloc(ctx, st, bytes);
  if (!p) FAIL(TWOACE_E_NOMEM, "cudaMalloc(%zu) for input staging failed", bytes);
  CK(cudaMemcpyAsync(p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  *dst = p;
  return 0;
}
static int dev_out(twoace_ctx* ctx, Staging& st, int mem, void* host, size_t bytes, void** dst) {
  if (host == nullptr) { *dst = nullptr; return 0; }
  if (mem == TWOACE_MEM_DEVICE) { *dst = host; return 0; }
  void* p = stage_alloc(ctx, st, bytes);
  if (!p) FAIL(TWOACE_E_NOMEM, "cudaMalloc(%zu) for output staging failed", bytes);
  *dst = p;
  return 0;
}
static int host_back(twoace_ctx* ctx, int mem, void* host, const void* dev, size_t bytes) {
  if (mem == TWOACE_MEM_DEVICE || host == nullptr || bytes == 0) return 0;
  CK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return 0;
}

// train rows per draw: floor(m cc_frac) (inferLowRankV4.m:36), ceil(0.95 m) for inferMinL2 (inferMinL2.m:34)
static size_t train_rows(int variant, int m, double cc_frac) {
  return variant == TWOACE_MINL2 ? (size_t)std::ceil((double)m * 0.95) : (size_t)std::floor((double)m * cc_frac);
}

static int solve_common(twoace_ctx* ctx, int variant, int mem, int nb, int tx, int rx, const int32_t* m,
                        const double* A, const int32_t* cb_rows, double row_scale, const double* B,
                        const int32_t* train_idx, const twoace_params* params, double* X, double* Y,
                        double* quality, double* info, double* stage_words) {
  if (!ctx) return TWOACE_E_INVALID;
  ctx->err.clear();
  if (variant < TWOACE_V4 || variant > TWOACE_MINL2) FAIL(TWOACE_E_INVALID, "unknown variant %d", variant);
  if (nb < 0 || !m || !B || !train_idx || !X || !Y || !quality) FAIL(TWOACE_E_INVALID, "null argument");
  if (!A && !cb_rows) FAIL(TWOACE_E_INVALID, "neither dense A nor codebook rows given");
  if (mem != TWOACE_MEM_HOST && mem != TWOACE_MEM_DEVICE) FAIL(TWOACE_E_INVALID, "bad mem flag");
  twoace_params p;
  if (params) p = *params; else twoace_default_params(&p);
  int rc = check_params(ctx, p, tx, rx);
  if (rc) return rc;
  if (nb == 0) return TWOACE_OK;
  CK(cudaSetDevice(ctx->device));
  const int n = tx * rx;
  if (!A) {
    if (!ctx->cb_rm) FAIL(TWOACE_E_INVALID, "no codebook registered (twoace_set_codebook)");
    if (ctx->cb_n != n) FAIL(TWOACE_E_INVALID, "codebook has n = %d, call has tx*rx = %d", ctx->cb_n, n);
  }
  const int T = variant == TWOACE_V4_MULTI ? 3 : 1;
  const int nstage = 4 * T + 1;
  size_t sum_m = 0, sum_tr = 0;
  for (int b = 0; b < nb; ++b) {
    if (m[b] < 2) FAIL(TWOACE_E_INVALID, "instance %d: m = %d (need m >= 2)", b, m[b]);
    sum_m += m[b];
    sum_tr += (size_t)T * train_rows(variant, m[b], p.cc_frac);
  }
  Staging st;
  const void *dA = nullptr, *dB = nullptr;
  void *dX = nullptr, *dY = nullptr, *dQ = nullptr, *dI = nullptr, *dS = nullptr;
  if (A) { rc = dev_in(ctx, st, mem, A, sum_m * n * sizeof(cd), &dA); if (rc) return rc; }
  rc = dev_in(ctx, st, mem, B, sum_m * sizeof(double), &dB); if (rc) return rc;
  rc = dev_out(ctx, st, mem, X, (size_t)nb * n * sizeof(cd), &dX); if (rc) return rc;
  rc = dev_out(ctx, st, mem, Y, sum_m * sizeof(cd), &dY); if (rc) return rc;
  rc = dev_out(ctx, st, mem, quality, (size_t)nb * sizeof(double), &dQ); if (rc) return rc;
  rc = dev_out(ctx, st, mem, info, (size_t)nb * TWOACE_INFO_WORDS * sizeof(double), &dI); if (rc) return rc;
  rc = dev_out(ctx, st, mem, stage_words, (size_t)nb * nstage * STAGE_SCAL * sizeof(double), &dS); if (rc) return rc;
  void* dT = nullptr;
  const size_t trace_n = (size_t)nb * nstage * p.maxiter;
  if (ctx->trace_user) {
    if (trace_n > ctx->trace_cap) FAIL(TWOACE_E_INVALID, "trace buffer holds %zu doubles, this call needs %zu", ctx->trace_cap, trace_n);
    rc = dev_out(ctx, st, ctx->trace_mem, ctx->trace_user, (trace_n + 1) * sizeof(double), &dT); if (rc) return rc;
    if (ctx->trace_mem == TWOACE_MEM_DEVICE && (trace_n & 1)) FAIL(TWOACE_E_INVALID, "device trace buffers need an even element count (nb * stages * maxiter = %zu)", trace_n);
    fill_nan_kernel<<<4 * ctx->num_sms, 256, 0, ctx->stream>>>((cd*)dT, (trace_n + 1) / 2);
    CK(cudaGetLastError());
    ctx->launches++;
  }

  size_t ao = 0, bo = 0, to = 0;
  for (int b0 = 0; b0 < nb; b0 += ctx->chunk) {
    const int cnt = std::min(ctx->chunk, nb - b0);
    ChunkIn in;
    in.variant = variant; in.nb = cnt; in.tx = tx; in.rx = rx; in.m = m + b0;
    in.dA = A ? (const cd*)dA + ao : nullptr;
    in.cb_rows = cb_rows ? cb_rows + bo : nullptr;
    in.row_scale = row_scale;
    in.dB = (const double*)dB + bo; in.train_idx = train_idx + to; in.p = p;
    in.dX = (cd*)dX + (size_t)b0 * n; in.dY = (cd*)dY + bo; in.dQ = (double*)dQ + b0;
    in.dInfo = dI ? (double*)dI + (size_t)b0 * TWOACE_INFO_WORDS : nullptr;
    in.dStage = dS ? (double*)dS + (size_t)b0 * nstage * STAGE_SCAL : nullptr;
    in.dTrace = dT ? (double*)dT + (size_t)b0 * nstage * p.maxiter : nullptr;
    rc = solve_chunk(ctx, in);
    if (rc) return rc;
    for (int b = b0; b < b0 + cnt; ++b) {
      ao += (size_t)m[b] * n; bo += m[b];
      to += (size_t)T * train_rows(variant, m[b], p.cc_frac);
    }
  }
  rc = host_back(ctx, mem, X, dX, (size_t)nb * n * sizeof(cd)); if (rc) return rc;
  rc = host_back(ctx, mem, Y, dY, sum_m * sizeof(cd)); if (rc) return rc;
  rc = host_back(ctx, mem, quality, dQ, (size_t)nb * sizeof(double)); if (rc) return rc;
  rc = host_back(ctx, mem, info, dI, (size_t)nb * TWOACE_INFO_WORDS * sizeof(double)); if (rc) return rc;
  rc = host_back(ctx, mem, stage_words, dS, (size_t)nb * nstage * STAGE_SCAL * sizeof(double)); if (rc) return rc;
  if (dT) { rc = host_back(ctx, ctx->trace_mem, ctx->trace_user, dT, trace_n * sizeof(double)); if (rc) return rc; }
  if (mem == TWOACE_MEM_HOST || (dT && ctx->trace_mem == TWOACE_MEM_HOST)) CK(cudaStreamSynchronize(ctx->stream));
  return TWOACE_OK;
}

// One GPU: solve_common.  Several (twoace_create_multi): every GPU solves a contiguous slice of the batch on its own
// host thread; the outputs land in the caller's (host) buffers at the slice offsets.
static int solve_multi(twoace_ctx* ctx, int variant, int mem, int nb, int tx, int rx, const int32_t* m,
                       const double* A, const int32_t* cb_rows, double row_scale, const double* B,
                       const int32_t* train_idx, const twoace_params* params, double* X, double* Y,
                       double* quality, double* info, double* stage_words) {
  if (!ctx || ctx->peers.empty() || nb < 2 || !m)
    return solve_common(ctx, variant, mem, nb, tx, rx, m, A, cb_rows, row_scale, B, train_idx, params, X, Y, quality, info, stage_words);
  ctx->err.clear();
  int rc = multi_guard(ctx, mem);
  if (rc) return rc;
  if (!B || !train_idx || !X || !Y || !quality) FAIL(TWOACE_E_INVALID, "null argument");
  twoace_params p;
  if (params) p = *params; else twoace_default_params(&p);
  const int T = variant == TWOACE_V4_MULTI ? 3 : 1, nstage = 4 * T + 1;
  const size_t n = (size_t)tx * rx;
  for (int b = 0; b < nb; ++b) if (m[b] < 2) FAIL(TWOACE_E_INVALID, "instance %d: m = %d (need m >= 2)", b, m[b]);
  std::vector<size_t> tr_off(nb + 1, 0);
  for (int b = 0; b < nb; ++b) tr_off[b + 1] = tr_off[b] + (size_t)T * train_rows(variant, m[b], p.cc_frac);
  const auto sl = split_batch(m, nb, 1 + (int)ctx->peers.size());
  return run_on_all(ctx, sl, [&](twoace_ctx* c, const BatchSlice& s) {
    return solve_common(c, variant, mem, s.b1 - s.b0, tx, rx, m + s.b0, A ? A + 2 * s.rows0 * n : nullptr,
                        cb_rows ? cb_rows + s.rows0 : nullptr, row_scale, B + s.rows0, train_idx + tr_off[s.b0], &p,
                        X + 2 * (size_t)s.b0 * n, Y + 2 * s.rows0, quality + s.b0,
                        info ? info + (size_t)s.b0 * TWOACE_INFO_WORDS : nullptr,
                        stage_words ? stage_words + (size_t)s.b0 * nstage * STAGE_SCAL : nullptr);
  });
}

extern "C" int twoace_set_trace(twoace_ctx* ctx, int mem, double* trace, int64_t capacity) {
  if (!ctx) return TWOACE_E_INVALID;
  ctx->err.clear();
  if (mem != TWOACE_MEM_HOST && mem != TWOACE_MEM_DEVICE) FAIL(TWOACE_E_INVALID, "bad mem flag");
  if (trace && capacity < 1) FAIL(TWOACE_E_INVALID, "trace capacity must be positive");
  ctx->trace_user = trace; ctx->trace_mem = mem; ctx->trace_cap = trace ? (size_t)capacity : 0;
  return TWOACE_OK;
}

extern "C" int twoace_solve_batch(twoace_ctx* ctx, int variant, int mem, int nb, int tx, int rx,
                                  const int32_t* m, const double* A, const double* B, const int32_t* train_idx,
                                  const twoace_params* params, double* X, double* Y, double* quality,
                                  double* info, double* stage_words) {
  if (ctx && !A) { ctx->err = "A is null"; return TWOACE_E_INVALID; }
  return solve_multi(ctx, variant, mem, nb, tx, rx, m, A, nullptr, 1.0, B, train_idx, params, X, Y, quality, info, stage_words);
}

extern "C" int twoace_solve_batch_codebook(twoace_ctx* ctx, int variant, int mem, int nb, int tx, int rx,
                                           const int32_t* m, const int32_t* cb_rows, double row_scale,
                                           const double* B, const int32_t* train_idx, const twoace_params* params,
                                           double* X, double* Y, double* quality, double* info, double* stage_words) {
  if (ctx && !cb_rows) { ctx->err = "cb_rows is null"; return TWOACE_E_INVALID; }
  return solve_multi(ctx, variant, mem, nb, tx, rx, m, nullptr, cb_rows, row_scale, B, train_idx, params, X, Y, quality, info, stage_words);
}

__global__ void transpose_cm_to_rm(const cd* __restrict__ src, cd* __restrict__ dst, int rows, int n) {
  // src: rows x n column-major; dst: rows x n row-major
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < (size_t)rows * n; idx += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % n);
    const size_t i = idx / n;
    dst[idx] = src[i + (size_t)rows * k];
  }
}

extern "C" int twoace_set_codebook(twoace_ctx* ctx, int mem, int rows, int n, const double* cb) {
  if (!ctx) return TWOACE_E_INVALID;
  ctx->err.clear();
  if (rows < 1 || n < 1 || !cb) FAIL(TWOACE_E_INVALID, "bad codebook arguments");
  if (!ctx->peers.empty()) {   // every GPU of a multi-GPU context holds its own copy
    if (mem != TWOACE_MEM_HOST) FAIL(TWOACE_E_INVALID, "a multi-GPU context takes the codebook from host memory");
    for (twoace_ctx* p : ctx->peers) {
      const int rc = twoace_set_codebook(p, mem, rows, n, cb);
      if (rc) { ctx->err = p->err; return rc; }
    }
  }
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->cb_rm) { CK(cudaFree(ctx->cb_rm)); ctx->cb_rm = nullptr; }
  const size_t bytes = (size_t)rows * n * sizeof(cd);
  CK(cudaMalloc((void**)&ctx->cb_rm, bytes));
  Staging st;
  const void* dsrc = nullptr;
  int rc = dev_in(ctx, st, mem, cb, bytes, &dsrc);
  if (rc) return rc;
  transpose_cm_to_rm<<<4 * ctx->num_sms, 256, 0, ctx->stream>>>((const cd*)dsrc, ctx->cb_rm, rows, n);
  CK(cudaGetLastError());
  ctx->launches++;
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->cb_rows = rows; ctx->cb_n = n;
  if (ctx->cb_codes) { CK(cudaFree(ctx->cb_codes)); ctx->cb_codes = nullptr; }
  ctx->cb_mag = 0.0;
  if (n % 16 == 0) {   // try the 2-bit representation (every shipped codebook qualifies)
    uint32_t* codes = nullptr;
    CK(cudaMalloc((void**)&codes, (size_t)rows * (n / 16) * sizeof(uint32_t) + 64));
    char* aux = nullptr;
    CK(cudaMalloc((void**)&aux, 64 + sizeof(QuantTask)));
    QuantTask q;
    q.A_rm = ctx->cb_rm; q.rows = rows; q.wpr = n / 16; q.codes = codes; q.mag_out = (double*)aux; q.flag_out = (int*)(aux + 8); q.ctl = nullptr;
    CK(cudaMemcpyAsync(aux + 64, &q, sizeof q, cudaMemcpyHostToDevice, ctx->stream));
    quant_kernel<<<1, NT, 0, ctx->stream>>>((const QuantTask*)(aux + 64), 1);
    CK(cudaGetLastError());
    ctx->launches++;
    struct { double mag; int flag; int pad; } res;
    CK(cudaMemcpyAsync(&res, aux, 16, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaFree(aux));
    if (res.flag == 1) { ctx->cb_codes = codes; ctx->cb_mag = res.mag; }
    else CK(cudaFree(codes));
  }
  return TWOACE_OK;
}

// Fill the constant scalars the stand-alone stage / spectral entry points need (scale = 1).
__global__ void set_ones_kernel(double* p, int cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cnt) p[i] = 1.0;
}

extern "C" int twoace_infer_admm_batch(twoace_ctx* ctx, int mem, int nb, int tx, int rx, const int32_t* m,
                                       const double* A, const double* B, int r, const double* X0,
                                       int scale_by_row, int use_rank_one, int nuclear,
                                       const twoace_params* params, double* X, double* Y, double* state,
                                       double* words) {
  if (!ctx) return TWOACE_E_INVALID;
  ctx->err.clear();
  if (nb < 0 || !m || !A || !B || !X0 || !X || !Y) FAIL(TWOACE_E_INVALID, "null argument");
  twoace_params p;
  if (params) p = *params; else twoace_default_params(&p);
  p.r = std::min(std::max(r, 1), SMALL_DMAX);
  int rc = check_params(ctx, p, tx, rx);
  if (rc) return rc;
  if (r < 1 || r > SMALL_DMAX) FAIL(TWOACE_E_INVALID, "r must be in [1,%d]", SMALL_DMAX);
  if (nb == 0) return TWOACE_OK;
  CK(cudaSetDevice(ctx->device));
  const int n = tx * rx;
  const int rout = scale_by_row ? r : 1;
  std::vector<size_t> a_off(nb + 1, 0), b_off(nb + 1, 0);
  for (int b = 0; b < nb; ++b) {
    if (m[b] < 1) FAIL(TWOACE_E_INVALID, "instance %d: m < 1", b);
    a_off[b + 1] = a_off[b] + (size_t)m[b] * n;
    b_off[b + 1] = b_off[b] + m[b];
  }
  const size_t sum_m = b_off[nb];
  std::vector<size_t> st_off(nb + 1, 0);
  for (int b = 0; b < nb; ++b) st_off[b + 1] = st_off[b] + 3 * (size_t)n * r + 2 * (size_t)m[b] * r;
  Staging st;
  const void *dA, *dB, *dX0;
  void *dX, *dY, *dSt, *dW;
  rc = dev_in(ctx, st, mem, A, a_off[nb] * sizeof(cd), &dA); if (rc) return rc;
  rc = dev_in(ctx, st, mem, B, sum_m * sizeof(double), &dB); if (rc) return rc;
  rc = dev_in(ctx, st, mem, X0, (size_t)nb * n * r * sizeof(cd), &dX0); if (rc) return rc;
  rc = dev_out(ctx, st, mem, X, (size_t)nb * n * rout * sizeof(cd), &dX); if (rc) return rc;
  rc = dev_out(ctx, st, mem, Y, sum_m * rout * sizeof(cd), &dY); if (rc) return rc;
  rc = dev_out(ctx, st, mem, state, st_off[nb] * sizeof(cd), &dSt); if (rc) return rc;
  rc = dev_out(ctx, st, mem, words, (size_t)nb * STAGE_SCAL * sizeof(double), &dW); if (rc) return rc;
  // arena: row-major A + the unit scalar
  Bump bp;
  const size_t o_Arm = bp.take(a_off[nb] * sizeof(cd)), o_one = bp.take(sizeof(double));
  const size_t o_ctl = bp.take((size_t)nb * sizeof(InstCtl));
  const bool try_codes = ctx->opt_fast && n % 16 == 0;
  const int wpr = n / 16;
  const size_t o_codes = try_codes ? bp.take(b_off[nb] * wpr * sizeof(uint32_t)) : 0;
  const size_t o_qflag = try_codes ? bp.take((size_t)nb * sizeof(int)) : 0;
  const size_t o_mag = try_codes ? bp.take((size_t)nb * sizeof(double)) : 0;
  rc = ensure(ctx, ctx->arena, bp.off + 256); if (rc) return rc;
  char* base = (char*)ctx->arena.p;
  cd* d_Arm = (cd*)(base + o_Arm);
  double* d_one = (double*)(base + o_one);
  InstCtl* d_ctl = (InstCtl*)(base + o_ctl);
  rc = ensure(ctx, ctx->taskbuf, (size_t)nb * (sizeof(PrepTask) + sizeof(QuantTask) + 2 * sizeof(StageTask)) + 8192); if (rc) return rc;
  size_t cursor = 0;
  set_ones_kernel<<<1, 32, 0, ctx->stream>>>(d_one, 1);
  CK(cudaGetLastError());
  ctx->launches++;
  {
    std::vector<PrepTask> pt(nb);
    for (int b = 0; b < nb; ++b) {
      PrepTask& t = pt[b];
      t = PrepTask{};
      t.A_cm = (const cd*)dA + a_off[b]; t.A_rm = d_Arm + a_off[b]; t.cb = nullptr; t.cbrows = nullptr;
      t.row_scale = 1.0; t.B = (const double*)dB + b_off[b]; t.m = m[b]; t.ctl = d_ctl + b;
    }
    const PrepTask* dt = nullptr;
    rc = upload_tasks(ctx, pt, cursor, &dt); if (rc) return rc;
    prep_kernel<<<std::min(nb, 8 * ctx->num_sms), NT, 0, ctx->stream>>>(dt, nb, n, 0.0);   // only the transpose is used
    CK(cudaGetLastError());
    ctx->launches++;
  }
  std::vector<char> inst_codes(nb, 0);
  if (try_codes) {
    std::vector<QuantTask> qt(nb);
    for (int b = 0; b < nb; ++b) {
      QuantTask& q = qt[b];
      q.A_rm = d_Arm + a_off[b]; q.rows = m[b]; q.wpr = wpr; q.codes = (uint32_t*)(base + o_codes) + b_off[b] * wpr;
      q.mag_out = (double*)(base + o_mag) + b; q.flag_out = (int*)(base + o_qflag) + b; q.ctl = nullptr;
    }
    const QuantTask* dq = nullptr;
    rc = upload_tasks(ctx, qt, cursor, &dq); if (rc) return rc;
    quant_kernel<<<std::min(nb, 8 * ctx->num_sms), NT, 0, ctx->stream>>>(dq, nb);
    CK(cudaGetLastError());
    ctx->launches++;
    std::vector<int> flags(nb);
    CK(cudaMemcpyAsync(flags.data(), base + o_qflag, (size_t)nb * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int b = 0; b < nb; ++b) inst_codes[b] = flags[b] == 1;      // per instance
  }
  std::vector<StageTask> tasks(nb);
  for (int b = 0; b < nb; ++b) {
    StageTask& a = tasks[b];
    const bool use_codes = inst_codes[b] != 0;
    a.codes = use_codes ? (const uint32_t*)(base + o_codes) + b_off[b] * wpr : nullptr;
    a.cscale = use_codes ? (const double*)(base + o_mag) + b : nullptr;
    a.A.base = d_Arm + a_off[b]; a.A.rows = nullptr; a.A.scale = d_one;
    a.B = (const double*)dB + b_off[b]; a.brows = nullptr; a.bscale = d_one;
    a.m = m[b]; a.r = r; a.X0 = (const cd*)dX0 + (size_t)b * n * r;
    a.Xout = (cd*)dX + (size_t)b * n * rout; a.Yout = (cd*)dY + b_off[b] * rout;
    a.sbr = scale_by_row ? 1 : 0; a.rank_one = use_rank_one ? 1 : 0; a.nuclear = nuclear == 2 ? 2 : (nuclear ? 1 : 0);   // 2: the inferMinL2.m iteration (no low-rank variable)
    a.rank_one_ptr = nullptr; a.active = nullptr; a.active_expect = 1;
    a.scal = dW ? (double*)dW + (size_t)b * STAGE_SCAL : nullptr;
    a.state = dSt ? (cd*)dSt + st_off[b] : nullptr;
  }
  rc = launch_stage(ctx, tasks, make_dev_params(p), n, tx, rx, cursor); if (rc) return rc;
  rc = host_back(ctx, mem, X, dX, (size_t)nb * n * rout * sizeof(cd)); if (rc) return rc;
  rc = host_back(ctx, mem, Y, dY, sum_m * rout * sizeof(cd)); if (rc) return rc;
  rc = host_back(ctx, mem, state, dSt, st_off[nb] * sizeof(cd)); if (rc) return rc;
  rc = host_back(ctx, mem, words, dW, (size_t)nb * STAGE_SCAL * sizeof(double)); if (rc) return rc;
  if (mem == TWOACE_MEM_HOST) CK(cudaStreamSynchronize(ctx->stream));
  return TWOACE_OK;
}

extern "C" int twoace_spectral_init_batch(twoace_ctx* ctx, int mem, int nb, int n, const int32_t* m,
                                          const double* A, const double* B, int r, double* Xs) {
  if (!ctx) return TWOACE_E_INVALID;
  ctx->err.clear();
  if (nb < 0 || !m || !A || !B || !Xs || n < 1 || r < 1) FAIL(TWOACE_E_INVALID, "bad argument");
  if (nb == 0) return TWOACE_OK;
  CK(cudaSetDevice(ctx->device));
  std::vector<size_t> a_off(nb + 1, 0), b_off(nb + 1, 0);
  for (int b = 0; b < nb; ++b) {
    if (m[b] < 1) FAIL(TWOACE_E_INVALID, "instance %d: m < 1", b);
    a_off[b + 1] = a_off[b] + (size_t)m[b] * n;
    b_off[b + 1] = b_off[b] + m[b];
  }
  Staging st;
  const void *dA, *dB;
  void* dXs;
  int rc = dev_in(ctx, st, mem, A, a_off[nb] * sizeof(cd), &dA); if (rc) return rc;
  rc = dev_in(ctx, st, mem, B, b_off[nb] * sizeof(double), &dB); if (rc) return rc;
  rc = dev_out(ctx, st, mem, Xs, (size_t)nb * n * r * sizeof(cd), &dXs); if (rc) return rc;
  Bump bp;
  const size_t o_Arm = bp.take(a_off[nb] * sizeof(cd)), o_one = bp.take(sizeof(double));
  const size_t o_ctl = bp.take((size_t)nb * sizeof(InstCtl));
  rc = ensure(ctx, ctx->arena, bp.off + 256); if (rc) return rc;
  char* base = (char*)ctx->arena.p;
  cd* d_Arm = (cd*)(base + o_Arm);
  double* d_one = (double*)(base + o_one);
  InstCtl* d_ctl = (InstCtl*)(base + o_ctl);
  rc = ensure(ctx, ctx->taskbuf, (size_t)nb * (sizeof(PrepTask) + sizeof(SpecTask)) + 4096); if (rc) return rc;
  size_t cursor = 0;
  set_ones_kernel<<<1, 32, 0, ctx->stream>>>(d_one, 1);
  CK(cudaGetLastError());
  ctx->launches++;
  {
    std::vector<PrepTask> pt(nb);
    for (int b = 0; b < nb; ++b) {
      PrepTask& t = pt[b];
      t = PrepTask{};
      t.A_cm = (const cd*)dA + a_off[b]; t.A_rm = d_Arm + a_off[b]; t.cb = nullptr; t.cbrows = nullptr;
      t.row_scale = 1.0; t.B = (const double*)dB + b_off[b]; t.m = m[b]; t.ctl = d_ctl + b;
    }
    const PrepTask* dt = nullptr;
    rc = upload_tasks(ctx, pt, cursor, &dt); if (rc) return rc;
    prep_kernel<<<std::min(nb, 8 * ctx->num_sms), NT, 0, ctx->stream>>>(dt, nb, n, 0.0);
    CK(cudaGetLastError());
    ctx->launches++;
  }
  std::vector<SpecTask> tasks(nb);
  for (int b = 0; b < nb; ++b) {
    SpecTask& s = tasks[b];
    s.A.base = d_Arm + a_off[b]; s.A.rows = nullptr; s.A.scale = d_one;
    s.B = (const double*)dB + b_off[b]; s.brows = nullptr; s.bscale = d_one;
    s.m = m[b]; s.r = std::min(r, std::min(m[b], n)); s.Xs = (cd*)dXs + (size_t)b * n * r; s.sweeps = nullptr;
  }
  if (true) {   // columns beyond min(r, m, n) stay zero
    CK(cudaMemsetAsync(dXs, 0, (size_t)nb * n * r * sizeof(cd), ctx->stream));
  }
  rc = launch_spectral(ctx, tasks, n, cursor); if (rc) return rc;
  rc = host_back(ctx, mem, Xs, dXs, (size_t)nb * n * r * sizeof(cd)); if (rc) return rc;
  if (mem == TWOACE_MEM_HOST) CK(cudaStreamSynchronize(ctx->stream));
  return TWOACE_OK;
}


// ------------------------------------------------------------------------------------------
// PhaseLift (MyPhaseLift.m:69-107)
extern "C" void twoace_pl_default_opts(twoace_pl_opts* o) {
  o->maxIts = 4000; o->tol = 1e-10; o->restart = 200; o->lambda = 5e-2; o->alpha = 0.9; o->beta = 0.5;
  o->L0 = 1.0; o->cntr_reset = 50; o->backtrack_tol = 1e-10; o->reduce = 1;
}

static int phaselift_single(twoace_ctx* ctx, int mem, int nb, int n, const int32_t* m, const double* A,
                                      const int32_t* cb_rows, double row_scale, const double* y,
                                      const twoace_pl_opts* opts, double* sig, double* info) {
  if (!ctx) return TWOACE_E_INVALID;
  ctx->err.clear();
  if (nb < 0 || !m || !y || !sig) FAIL(TWOACE_E_INVALID, "null argument");
  if (!A && !cb_rows) FAIL(TWOACE_E_INVALID, "neither dense A nor codebook rows given");
  if (mem != TWOACE_MEM_HOST && mem != TWOACE_MEM_DEVICE) FAIL(TWOACE_E_INVALID, "bad mem flag");
  if (n < 1) FAIL(TWOACE_E_INVALID, "n = %d", n);
  if (n > PL_NMAX) FAIL(TWOACE_E_UNSUPPORTED, "PhaseLift is built for n <= %d, the reference's non-largescale range (n = %d)", PL_NMAX, n);
  twoace_pl_opts o;
  if (opts) o = *opts; else twoace_pl_default_opts(&o);
  if (o.maxIts < 1 || !(o.lambda > 0.0) || !(o.L0 > 0.0) || !(o.alpha > 0.0) || !(o.beta > 0.0) || o.restart < 1 ||
      o.cntr_reset < 1 || !(o.tol >= 0.0))
    FAIL(TWOACE_E_INVALID, "bad PhaseLift option (maxIts %d, lambda %g, L0 %g, alpha %g, beta %g, restart %d)",
         o.maxIts, o.lambda, o.L0, o.alpha, o.beta, o.restart);
  if (nb == 0) return TWOACE_OK;
  CK(cudaSetDevice(ctx->device));
  if (!A) {
    if (!ctx->cb_rm) FAIL(TWOACE_E_INVALID, "no codebook registered (twoace_set_codebook)");
    if (ctx->cb_n != n) FAIL(TWOACE_E_INVALID, "codebook has n = %d, call has n = %d", ctx->cb_n, n);
  }
  std::vector<size_t> a_off(nb + 1, 0), b_off(nb + 1, 0);
  int maxm = 0;
  for (int b = 0; b < nb; ++b) {
    if (m[b] < 1) FAIL(TWOACE_E_INVALID, "instance %d: m = %d", b, m[b]);
    if (m[b] > 4096) FAIL(TWOACE_E_UNSUPPORTED, "instance %d: m = %d > 4096", b, m[b]);
    a_off[b + 1] = a_off[b] + (size_t)m[b] * n;
    b_off[b + 1] = b_off[b] + m[b];
    maxm = std::max(maxm, m[b]);
  }
  if (!A) {
    for (size_t i = 0; i < b_off[nb]; ++i)
      if (cb_rows[i] < 0 || cb_rows[i] >= ctx->cb_rows) FAIL(TWOACE_E_INVALID, "codebook row id %d out of range", cb_rows[i]);
  }
  // n > 256: the iteration runs in the row space of A only (d = m), so every instance needs m <= 256 and reduce = 1
  if (n > PL_DMAX && (!o.reduce || maxm > PL_DMAX))
    FAIL(TWOACE_E_UNSUPPORTED, "PhaseLift with n = %d > %d runs in the row space of A: needs reduce = 1 and m <= %d (m = %d)",
         n, PL_DMAX, PL_DMAX, maxm);
  const int dcap = pl_dcap(n, maxm);
  Staging st;
  const void *dA = nullptr, *dY = nullptr;
  void *dSig = nullptr, *dInfo = nullptr;
  int rc;
  if (A) { rc = dev_in(ctx, st, mem, A, a_off[nb] * sizeof(cd), &dA); if (rc) return rc; }
  rc = dev_in(ctx, st, mem, y, b_off[nb] * sizeof(double), &dY); if (rc) return rc;
  rc = dev_out(ctx, st, mem, sig, (size_t)nb * n * sizeof(cd), &dSig); if (rc) return rc;
  rc = dev_out(ctx, st, mem, info, (size_t)nb * PL_INFO * sizeof(double), &dInfo); if (rc) return rc;

  const size_t smem = pl_smem_bytes(n, maxm);
  if (smem > 227 * 1024) FAIL(TWOACE_E_UNSUPPORTED, "PhaseLift: m = %d needs %zu bytes of shared memory", maxm, smem);
  CK(pl_kernel_set_smem(smem));
  int per_sm = 0;
  CK(pl_kernel_occupancy(&per_sm, smem));
  if (per_sm < 1) FAIL(TWOACE_E_CUDA, "PhaseLift kernel does not fit on an SM");
  // longest solves first is not knowable up front; instances are handed out dynamically through a counter
  const int grid = std::min(nb, per_sm * ctx->num_sms);
  const size_t ws_stride = (pl_ws_elems(dcap, maxm) + 15) / 16 * 16;
  Bump bp;
  const size_t o_ws = bp.take((size_t)grid * ws_stride * sizeof(cd));
  const size_t o_rows = bp.take(b_off[nb] * sizeof(int32_t));
  const size_t o_cnt = bp.take(sizeof(int));
  rc = ensure(ctx, ctx->arena, bp.off + 256); if (rc) return rc;
  char* base = (char*)ctx->arena.p;
  int32_t* d_rows = (int32_t*)(base + o_rows);
  int* d_cnt = (int*)(base + o_cnt);
  if (!A) CK(cudaMemcpyAsync(d_rows, cb_rows, b_off[nb] * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync(d_cnt, 0, sizeof(int), ctx->stream));
  rc = ensure(ctx, ctx->taskbuf, (size_t)nb * sizeof(PlTask) + 4096); if (rc) return rc;
  std::vector<PlTask> tasks(nb);
  for (int b = 0; b < nb; ++b) {
    PlTask& t = tasks[b];
    t.A_cm = A ? (const cd*)dA + a_off[b] : nullptr;
    t.cb = A ? nullptr : ctx->cb_rm;
    t.rows = A ? nullptr : d_rows + b_off[b];
    t.scale = row_scale;
    t.y = (const double*)dY + b_off[b];
    t.m = m[b];
    t.sig = (cd*)dSig + (size_t)b * n;
    t.info = dInfo ? (double*)dInfo + (size_t)b * PL_INFO : nullptr;
  }
  size_t cursor = 0;
  const PlTask* dt = nullptr;
  rc = upload_tasks(ctx, tasks, cursor, &dt); if (rc) return rc;
  PlOpts po;
  po.maxIts = o.maxIts; po.tol = o.tol; po.restart = o.restart; po.lam = o.lambda; po.alpha = o.alpha;
  po.beta = o.beta; po.L0 = o.L0; po.cntr_reset = o.cntr_reset; po.backtrack_tol = o.backtrack_tol;
  po.reduce = o.reduce;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (ctx->timing) {
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, ctx->stream));
  }
  CK(pl_kernel_launch(grid, smem, ctx->stream, dt, nb, n, maxm, po, (cd*)(base + o_ws), ws_stride, d_cnt));
  ctx->launches++;
  if (ctx->timing) {
    CK(cudaEventRecord(e1, ctx->stream));
    push_stage_event(ctx, e0, e1, "phaselift_kernel");
  }
  rc = host_back(ctx, mem, sig, dSig, (size_t)nb * n * sizeof(cd)); if (rc) return rc;
  rc = host_back(ctx, mem, info, dInfo, (size_t)nb * PL_INFO * sizeof(double)); if (rc) return rc;
  if (mem == TWOACE_MEM_HOST) {
    CK(cudaStreamSynchronize(ctx->stream));
    if (n > dcap && info)
      for (int b = 0; b < nb; ++b)
        if (info[(size_t)b * PL_INFO + 3] == -2.0)
          FAIL(TWOACE_E_UNSUPPORTED, "instance %d: linearly dependent measurement rows at n = %d > %d (no row-space factor)",
               b, n, PL_DMAX);
  }
  return TWOACE_OK;
}

// ------------------------------------------------------------------------------------------
// Evaluation metrics (Evaluation_H.m:81-115)
static int metrics_single(twoace_ctx* ctx, int mem, int nb, int tx, int rx, const double* X_est,
                          const double* X_true, int phase_bit, double* out) {
  if (!ctx) return TWOACE_E_INVALID;
  ctx->err.clear();
  if (nb < 0 || !X_est || !X_true || !out) FAIL(TWOACE_E_INVALID, "null argument");
  if (mem != TWOACE_MEM_HOST && mem != TWOACE_MEM_DEVICE) FAIL(TWOACE_E_INVALID, "bad mem flag");
  if (tx < 1 || rx < 1 || tx > MET_DMAX || rx > MET_DMAX) FAIL(TWOACE_E_UNSUPPORTED, "metrics: tx, rx must be in 1..%d", MET_DMAX);
  if (phase_bit < 1 || phase_bit > 8) FAIL(TWOACE_E_INVALID, "phase_bit = %d", phase_bit);
  if (nb == 0) return TWOACE_OK;
  CK(cudaSetDevice(ctx->device));
  const size_t n = (size_t)tx * rx;
  Staging st;
  const void *dE = nullptr, *dT = nullptr;
  void* dO = nullptr;
  int rc = dev_in(ctx, st, mem, X_est, (size_t)nb * n * sizeof(cd), &dE); if (rc) return rc;
  rc = dev_in(ctx, st, mem, X_true, (size_t)nb * n * sizeof(cd), &dT); if (rc) return rc;
  rc = dev_out(ctx, st, mem, out, (size_t)nb * MET_WORDS * sizeof(double), &dO); if (rc) return rc;
  const size_t smem = met_smem_bytes();
  CK(cudaFuncSetAttribute(metrics_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  metrics_kernel<<<std::min(nb, 8 * ctx->num_sms), NT, smem, ctx->stream>>>((const cd*)dE, (const cd*)dT, nb, tx, rx,
                                                                         phase_bit, (double*)dO);
  CK(cudaGetLastError());
  ctx->launches++;
  rc = host_back(ctx, mem, out, dO, (size_t)nb * MET_WORDS * sizeof(double)); if (rc) return rc;
  if (mem == TWOACE_MEM_HOST) CK(cudaStreamSynchronize(ctx->stream));
  return TWOACE_OK;
}

extern "C" void twoace_synth_default_params(twoace_synth_params* p, int nt, int nr) {
  p->nt = nt; p->nr = nr; p->L = 3; p->searching_area = 95.0; p->wavelength = 3e8 / 60.48e9; p->spacing = 3.055e-3;
  p->row_scale = 1.0 / std::sqrt((double)nt * nr); p->cc_frac = 0.95; p->ntrain = 1; p->seed = 58659179ull;
}

static int synth_single(twoace_ctx* ctx, int mem, int nb, const twoace_synth_params* sp, const int32_t* m,
                        const double* snr_db, const int32_t* row_lo, const int32_t* row_hi,
                        const int64_t* trial_id, int32_t* cb_rows, int32_t* train_idx, double* B,
                        double* vecH, double* angles) {
  if (!ctx) return TWOACE_E_INVALID;
  ctx->err.clear();
  if (nb < 0 || !sp || !m || !snr_db || !row_lo || !row_hi || !trial_id || !cb_rows || !train_idx || !B || !vecH)
    FAIL(TWOACE_E_INVALID, "null argument");
  if (mem != TWOACE_MEM_HOST && mem != TWOACE_MEM_DEVICE) FAIL(TWOACE_E_INVALID, "bad mem flag");
  if (!ctx->cb_rm) FAIL(TWOACE_E_INVALID, "no codebook registered (twoace_set_codebook)");
  const int n = sp->nt * sp->nr;
  if (sp->nt < 1 || sp->nr < 1 || n != ctx->cb_n) FAIL(TWOACE_E_INVALID, "codebook has n = %d, synthesis asks for nt*nr = %d", ctx->cb_n, n);
  if (sp->L < 1 || sp->L > 32) FAIL(TWOACE_E_INVALID, "L must be in [1,32]");
  if (sp->ntrain < 1 || sp->ntrain > 8) FAIL(TWOACE_E_INVALID, "ntrain must be in [1,8]");
  if (!(sp->cc_frac > 0.0 && sp->cc_frac < 1.0)) FAIL(TWOACE_E_INVALID, "cc_frac must be in (0,1)");
  if (nb == 0) return TWOACE_OK;
  CK(cudaSetDevice(ctx->device));
  std::vector<size_t> b_off(nb + 1, 0), t_off(nb + 1, 0);
  int pmax = 2;
  for (int b = 0; b < nb; ++b) {
    const int R = row_hi[b] - row_lo[b];
    if (row_lo[b] < 0 || row_hi[b] > ctx->cb_rows || R < 1) FAIL(TWOACE_E_INVALID, "instance %d: row range [%d,%d) outside the codebook", b, row_lo[b], row_hi[b]);
    if (m[b] < 2 || m[b] > R) FAIL(TWOACE_E_INVALID, "instance %d: m = %d probes from %d candidate rows", b, m[b], R);
    if (R > 8192) FAIL(TWOACE_E_UNSUPPORTED, "instance %d: more than 8192 candidate rows", b);
    while (pmax < R) pmax <<= 1;
    const int mtr = (int)std::floor((double)m[b] * sp->cc_frac);
    b_off[b + 1] = b_off[b] + m[b];
    t_off[b + 1] = t_off[b] + (size_t)sp->ntrain * mtr;
  }
  Bump bp;
  const size_t o_rows = bp.take(b_off[nb] * 4), o_train = bp.take(t_off[nb] * 4), o_tasks = bp.take((size_t)nb * sizeof(SynthTask));
  int rc = ensure(ctx, ctx->arena, bp.off + 256);
  if (rc) return rc;
  char* base = (char*)ctx->arena.p;
  Staging st;
  void *dB = nullptr, *dH = nullptr, *dAng = nullptr;
  rc = dev_out(ctx, st, mem, B, b_off[nb] * sizeof(double), &dB); if (rc) return rc;
  rc = dev_out(ctx, st, mem, vecH, (size_t)nb * n * sizeof(cd), &dH); if (rc) return rc;
  rc = dev_out(ctx, st, mem, angles, (size_t)nb * 2 * sp->L * sizeof(double), &dAng); if (rc) return rc;
  std::vector<SynthTask> tasks(nb);
  for (int b = 0; b < nb; ++b) {
    SynthTask& t = tasks[b];
    t.m = m[b]; t.row_lo = row_lo[b]; t.row_hi = row_hi[b]; t.snr_db = snr_db[b];
    t.trial = (unsigned long long)trial_id[b];
    t.rows_out = (int*)(base + o_rows) + b_off[b];
    t.train_out = (int*)(base + o_train) + t_off[b];
    t.B_out = (double*)dB + b_off[b];
    t.vecH_out = (cd*)dH + (size_t)b * n;
    t.angles_out = dAng ? (double*)dAng + (size_t)b * 2 * sp->L : nullptr;
  }
  CK(cudaMemcpyAsync(base + o_tasks, tasks.data(), (size_t)nb * sizeof(SynthTask), cudaMemcpyHostToDevice, ctx->stream));
  SynthDims dm = {};
  dm.nt = sp->nt; dm.nr = sp->nr; dm.L = sp->L; dm.ntrain = sp->ntrain;
  dm.area = sp->searching_area; dm.k_phase = 2.0 * 3.14159265358979323846 / sp->wavelength * sp->spacing;
  dm.cc_frac = sp->cc_frac; dm.row_scale = sp->row_scale;
  dm.key0 = (unsigned int)(sp->seed & 0xffffffffull); dm.key1 = (unsigned int)(sp->seed >> 32);
  dm.pmax = pmax;
  const size_t smem = synth_smem_bytes(dm);
  CK(cudaFuncSetAttribute(synth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  synth_kernel<<<std::min(nb, 8 * ctx->num_sms), NT, smem, ctx->stream>>>((const SynthTask*)(base + o_tasks), nb, dm, ctx->cb_rm, n);
  CK(cudaGetLastError());
  ctx->launches++;
  CK(cudaMemcpyAsync(cb_rows, base + o_rows, b_off[nb] * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(train_idx, base + o_train, t_off[nb] * 4, cudaMemcpyDeviceToHost, ctx->stream));
  rc = host_back(ctx, mem, B, dB, b_off[nb] * sizeof(double)); if (rc) return rc;
  rc = host_back(ctx, mem, vecH, dH, (size_t)nb * n * sizeof(cd)); if (rc) return rc;
  rc = host_back(ctx, mem, angles, dAng, (size_t)nb * 2 * sp->L * sizeof(double)); if (rc) return rc;
  CK(cudaStreamSynchronize(ctx->stream));   // the index lists are host outputs
  return TWOACE_OK;
}

extern "C" int twoace_phaselift_batch(twoace_ctx* ctx, int mem, int nb, int n, const int32_t* m, const double* A,
                                      const int32_t* cb_rows, double row_scale, const double* y,
                                      const twoace_pl_opts* opts, double* sig, double* info) {
  if (!ctx || ctx->peers.empty() || nb < 2 || !m)
    return phaselift_single(ctx, mem, nb, n, m, A, cb_rows, row_scale, y, opts, sig, info);
  ctx->err.clear();
  int rc = multi_guard(ctx, mem);
  if (rc) return rc;
  if (!y || !sig) FAIL(TWOACE_E_INVALID, "null argument");
  const auto sl = split_batch(m, nb, 1 + (int)ctx->peers.size());
  return run_on_all(ctx, sl, [&](twoace_ctx* c, const BatchSlice& s) {
    return phaselift_single(c, mem, s.b1 - s.b0, n, m + s.b0, A ? A + 2 * s.rows0 * (size_t)n : nullptr,
                            cb_rows ? cb_rows + s.rows0 : nullptr, row_scale, y + s.rows0, opts,
                            sig + 2 * (size_t)s.b0 * n, info ? info + (size_t)s.b0 * TWOACE_PL_INFO_WORDS : nullptr);
  });
}

extern "C" int twoace_metrics_batch(twoace_ctx* ctx, int mem, int nb, int tx, int rx, const double* X_est,
                                    const double* X_true, int phase_bit, double* out) {
  if (!ctx || ctx->peers.empty() || nb < 2) return metrics_single(ctx, mem, nb, tx, rx, X_est, X_true, phase_bit, out);
  ctx->err.clear();
  int rc = multi_guard(ctx, mem);
  if (rc) return rc;
  if (!X_est || !X_true || !out) FAIL(TWOACE_E_INVALID, "null argument");
  const size_t n = (size_t)tx * rx;
  const auto sl = split_batch(nullptr, nb, 1 + (int)ctx->peers.size());
  return run_on_all(ctx, sl, [&](twoace_ctx* c, const BatchSlice& s) {
    return metrics_single(c, mem, s.b1 - s.b0, tx, rx, X_est + 2 * (size_t)s.b0 * n, X_true + 2 * (size_t)s.b0 * n, phase_bit,
                          out + (size_t)s.b0 * MET_WORDS);
  });
}

extern "C" int twoace_synth_batch(twoace_ctx* ctx, int mem, int nb, const twoace_synth_params* sp, const int32_t* m,
                                  const double* snr_db, const int32_t* row_lo, const int32_t* row_hi,
                                  const int64_t* trial_id, int32_t* cb_rows, int32_t* train_idx, double* B,
                                  double* vecH, double* angles) {
  if (!ctx || ctx->peers.empty() || nb < 2 || !m || !sp)
    return synth_single(ctx, mem, nb, sp, m, snr_db, row_lo, row_hi, trial_id, cb_rows, train_idx, B, vecH, angles);
  ctx->err.clear();
  int rc = multi_guard(ctx, mem);
  if (rc) return rc;
  if (!snr_db || !row_lo || !row_hi || !trial_id || !cb_rows || !train_idx || !B || !vecH) FAIL(TWOACE_E_INVALID, "null argument");
  if (sp->ntrain < 1 || !(sp->cc_frac > 0.0 && sp->cc_frac < 1.0)) FAIL(TWOACE_E_INVALID, "bad synthesis parameters");
  std::vector<size_t> tr_off(nb + 1, 0);
  for (int b = 0; b < nb; ++b) tr_off[b + 1] = tr_off[b] + (size_t)sp->ntrain * (size_t)std::floor((double)std::max(0, m[b]) * sp->cc_frac);
  const size_t n = (size_t)sp->nt * sp->nr;
  const auto sl = split_batch(m, nb, 1 + (int)ctx->peers.size());
  return run_on_all(ctx, sl, [&](twoace_ctx* c, const BatchSlice& s) {
    return synth_single(c, mem, s.b1 - s.b0, sp, m + s.b0, snr_db + s.b0, row_lo + s.b0, row_hi + s.b0, trial_id + s.b0,
                        cb_rows + s.rows0, train_idx + tr_off[s.b0], B + s.rows0, vecH + 2 * (size_t)s.b0 * n,
                        angles ? angles + (size_t)s.b0 * 2 * sp->L : nullptr);
  });
}

static int angle_single(twoace_ctx* ctx, int mem, int nb, int nt, int nr, int L, int nqt, int nqr,
                        double searching_area, double wavelength, double spacing, const double* X_est,
                        const double* angles_true, double* out) {
  if (!ctx) return TWOACE_E_INVALID;
  ctx->err.clear();
  if (nb < 0 || !X_est || !angles_true || !out) FAIL(TWOACE_E_INVALID, "null argument");
  if (mem != TWOACE_MEM_HOST && mem != TWOACE_MEM_DEVICE) FAIL(TWOACE_E_INVALID, "bad mem flag");
  if (nt < 1 || nr < 1 || nt > MET_DMAX || nr > MET_DMAX) FAIL(TWOACE_E_UNSUPPORTED, "angle metrics: nt, nr must be in 1..%d", MET_DMAX);
  if (L < 1 || L > ANG_LMAX) FAIL(TWOACE_E_INVALID, "L must be in [1,%d]", ANG_LMAX);
  if (nqt < 1 || nqr < 1 || nqt > 4096 || nqr > 4096) FAIL(TWOACE_E_INVALID, "grid sizes must be in [1,4096]");
  if (!(wavelength > 0.0) || !(spacing > 0.0) || !(searching_area > 0.0 && searching_area <= 180.0)) FAIL(TWOACE_E_INVALID, "bad array geometry");
  if (nb == 0) return TWOACE_OK;
  CK(cudaSetDevice(ctx->device));
  AngDims dm = {};
  dm.nt = nt; dm.nr = nr; dm.L = L; dm.nqt = nqt; dm.nqr = nqr;
  dm.kph = 2.0 * 3.14159265358979323846 * spacing / wavelength;
  // Sparse_Channel_Formulation.m:120-135: nearest grid point (first minimum) of the two ends of the searching area
  auto nearest = [&](int nq, double x) {
    int pos = 0;
    double best = INFINITY;
    for (int q = 0; q < nq; ++q) { const double e = std::fabs(dm.kph * (-1.0 + 2.0 * q / (double)nq) - x); if (e < best) { best = e; pos = q; } }
    return pos;
  };
  const double lo = dm.kph * std::sin(-searching_area / 2.0 * 3.14159265358979323846 / 180.0);
  const double hi = dm.kph * std::sin(searching_area / 2.0 * 3.14159265358979323846 / 180.0);
  dm.u0 = nearest(nqt, lo); dm.u1 = nearest(nqt, hi); dm.v0 = nearest(nqr, lo); dm.v1 = nearest(nqr, hi);
  const int nu = dm.u1 - dm.u0 + 1, nv = dm.v1 - dm.v0 + 1;
  if ((size_t)nu * nv < (size_t)L) FAIL(TWOACE_E_INVALID, "the searching area holds fewer than L grid points");
  dm.ws_stride = ((size_t)2 * nv * nt + (size_t)nv * nu + 15) / 16 * 16;
  const int grid = std::min(nb, 4 * ctx->num_sms);
  int rc = ensure(ctx, ctx->ws, (size_t)grid * dm.ws_stride * sizeof(double));
  if (rc) return rc;
  const size_t n = (size_t)nt * nr;
  Staging st;
  const void *dE = nullptr, *dA = nullptr;
  void* dO = nullptr;
  rc = dev_in(ctx, st, mem, X_est, (size_t)nb * n * sizeof(cd), &dE); if (rc) return rc;
  rc = dev_in(ctx, st, mem, angles_true, (size_t)nb * 2 * L * sizeof(double), &dA); if (rc) return rc;
  rc = dev_out(ctx, st, mem, out, (size_t)nb * ANG_WORDS * sizeof(double), &dO); if (rc) return rc;
  angle_metrics_kernel<<<grid, NT, 0, ctx->stream>>>((const cd*)dE, (const double*)dA, nb, dm, (double*)dO, (double*)ctx->ws.p);
  CK(cudaGetLastError());
  ctx->launches++;
  rc = host_back(ctx, mem, out, dO, (size_t)nb * ANG_WORDS * sizeof(double)); if (rc) return rc;
  if (mem == TWOACE_MEM_HOST) CK(cudaStreamSynchronize(ctx->stream));
  return TWOACE_OK;
}

extern "C" int twoace_angle_metrics_batch(twoace_ctx* ctx, int mem, int nb, int nt, int nr, int L, int nqt, int nqr,
                                          double searching_area, double wavelength, double spacing, const double* X_est,
                                          const double* angles_true, double* out) {
  if (!ctx || ctx->peers.empty() || nb < 2)
    return angle_single(ctx, mem, nb, nt, nr, L, nqt, nqr, searching_area, wavelength, spacing, X_est, angles_true, out);
  ctx->err.clear();
  int rc = multi_guard(ctx, mem);
  if (rc) return rc;
  if (!X_est || !angles_true || !out || L < 1) FAIL(TWOACE_E_INVALID, "null argument");
  const size_t n = (size_t)nt * nr;
  const auto sl = split_batch(nullptr, nb, 1 + (int)ctx->peers.size());
  return run_on_all(ctx, sl, [&](twoace_ctx* c, const BatchSlice& s) {
    return angle_single(c, mem, s.b1 - s.b0, nt, nr, L, nqt, nqr, searching_area, wavelength, spacing,
                        X_est + 2 * (size_t)s.b0 * n, angles_true + (size_t)s.b0 * 2 * L, out + (size_t)s.b0 * ANG_WORDS);
  });
}

extern "C" int twoace_set_timing(twoace_ctx* ctx, int on) {
  if (!ctx) return TWOACE_E_INVALID;
  ctx->timing = on != 0;
  return TWOACE_OK;
}

extern "C" int twoace_timing_collect(twoace_ctx* ctx, double* stage_ms, int64_t* stage_launches) {
  if (!ctx) return TWOACE_E_INVALID;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  double tot = 0.0;
  const bool trace = getenv("TWOACE_TRACE_LAUNCHES") != nullptr;
  size_t li = 0;
  for (auto& pr : ctx->stage_events) {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, pr.first, pr.second));
    if (trace) fprintf(stderr, "[twoace] %9.3f ms  %s\n", ms, li < ctx->stage_labels.size() ? ctx->stage_labels[li].c_str() : "stage kernel");
    const bool side = li < ctx->stage_side.size() && ctx->stage_side[li];
    ++li;
    if (!side) tot += ms;          // side-stream launches overlap main-stream ones: not part of the serial sum
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  if (stage_ms) *stage_ms = tot;
  if (stage_launches) *stage_launches = (int64_t)ctx->stage_events.size();
  ctx->stage_events.clear();
  ctx->stage_labels.clear();
  ctx->stage_side.clear();
  return TWOACE_OK;
}

// 8 independent DFMA chains per thread, 2048 threads per SM: saturates the FP64 pipe.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

extern "C" int twoace_fp64_peak(twoace_ctx* ctx, double* tflops) {
  if (!ctx || !tflops) return TWOACE_E_INVALID;
  CK(cudaSetDevice(ctx->device));
  const int blocks = ctx->num_sms * 8, iters = 4096;
  double* d = nullptr;
  CK(cudaMalloc((void**)&d, (size_t)blocks * 256 * sizeof(double)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(e0, ctx->stream));
    dfma_peak_kernel<<<blocks, 256, 0, ctx->stream>>>(d, iters, 0.999999, 1e-9);
    CK(cudaGetLastError());
    CK(cudaEventRecord(e1, ctx->stream));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double fl = 2.0 * 64.0 * iters * (double)blocks * 256.0;
    if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *tflops = best;
  return TWOACE_OK;
}

extern "C" int twoace_set_option(twoace_ctx* ctx, const char* key, int value) {
  if (!ctx || !key) return TWOACE_E_INVALID;
  ctx->err.clear();
  for (twoace_ctx* p : ctx->peers) { const int rc = twoace_set_option(p, key, value); if (rc) { ctx->err = p->err; return rc; } }
  const std::string k(key);
  if (k == "fast") ctx->opt_fast = value ? 1 : 0;
  else if (k == "fast_cs") { if (value != 2 && value != 4) FAIL(TWOACE_E_INVALID, "fast_cs must be 2 or 4"); ctx->opt_fast_cs = value; }
  else if (k == "chunk") { if (value < 1) FAIL(TWOACE_E_INVALID, "chunk must be >= 1"); ctx->chunk = value; }
  else if (k == "dedup_nuclear_rerun") ctx->opt_dedup_nuclear = value ? 1 : 0;
  else if (k == "tensor") ctx->opt_tensor = value ? 1 : 0;
  else if (k == "cache_sinv") ctx->opt_cache_sinv = value ? 1 : 0;
  else if (k == "spectral_jacobi") ctx->opt_spectral_jacobi = value ? 1 : 0;
  else if (k == "overlap") ctx->opt_overlap = value ? 1 : 0;
  else FAIL(TWOACE_E_INVALID, "unknown option %s", key);
  return TWOACE_OK;
}

extern "C" int64_t twoace_fast_launch_count(const twoace_ctx* ctx) {
  if (!ctx) return 0;
  int64_t n = ctx->fast_launches;
  for (const twoace_ctx* p : ctx->peers) n += p->fast_launches;
  return n;
}
extern "C" int64_t twoace_tensor_launch_count(const twoace_ctx* ctx) {
  if (!ctx) return 0;
  int64_t n = ctx->tc_launches;
  for (const twoace_ctx* p : ctx->peers) n += p->tc_launches;
  return n;
}
