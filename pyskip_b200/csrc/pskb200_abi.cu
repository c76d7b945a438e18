// pskb200_abi.cu -- implementation of include/pskb200.h: store management, plan upload and
// kernel launches for the run-length-encoded array evaluation path on one B200.
//
// Host-side counterpart of eval::eval_generic / eval_simple (include/pyskip/detail/eval.hpp:
// 609-659): where the reference allocates a worst-case Store, partitions the pool over CPU
// threads and fuses the partial stores, this file allocates the worst-case device arrays,
// uploads a DPlan and launches k_prepare / k_sample_rank / k_tile_bounds / k_merge_eval on
// the library stream.  There is no CPU compute path in this file.
#include <atomic>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "psk_conv.cuh"
#include "psk_convert.cuh"
#include "psk_eval.cuh"
#ifndef PSK_CPU_SIM
#include "psk_jit.h"
#endif

using namespace psk;

struct psk_store_s {
  std::atomic<int> ref{1};
  int32_t dtype = 0;
  int64_t nruns = 0;
  int64_t capacity = 0;
  int32_t span = 0;
  int2* d_runs = nullptr;       // (end, payload) pairs: capacity + kStorePad pairs; [nruns], [nruns+1] sentinels
  int32_t* d_count = nullptr;   // the run count on the device (inside the padding of d_runs)
  int32_t* d_index = nullptr;   // ends of every kIndexStride-th run (behind the padding, same allocation)
  bool indexed = false;         // ... filled (on first use as a source of an evaluation)
  cudaEvent_t ready = nullptr;  // produced on another stream (async upload): consumers wait for it
  cudaEvent_t busy = nullptr;   // async download finished (waited on before the memory is freed)
  int32_t pending_slot = -1;    // result of psk_eval_async whose run count has not been read yet
  int32_t failed = 0;           // status of a psk_eval_async that turned out to have failed
};

namespace {

thread_local std::string g_err;

struct Ctx {
  bool inited = false;
  int device = -1;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t h2d = nullptr, d2h = nullptr;  // copy streams of the *_async entry points
  int32_t* d_async_err = nullptr;             // device flag: an async upload failed validation
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t pev0 = nullptr, pev1 = nullptr;  // bracket the merge/eval kernel when profiling
  bool profile = false;
  double prof_ms = 0.0;
  bool profile_partition = false;
  int sample_div = 32;  // sample stride S = tile capacity / sample_div (tuning aid: PSKB200_SAMPLE_DIV)
  int64_t prof_n = 0;
  int64_t prof_runs_in = 0, prof_runs_out = 0;
  unsigned long long phase_cycles[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int64_t jit_min_runs = 1 << 14;  // evaluations at least this large get a specialised kernel
  int64_t jit_launches = 0;
  bool convert_vec = true;         // dense <-> RLE with 16-byte accesses (psk_set_vector_conversion)
  int64_t small_partition = 2048;  // samples up to which one block does the whole partition
  int stagger_ns = 0;
  int jit_T = 256, jit_CAP = 4096, jit_regions = kParkDepth;  // tile geometry / park depth of the specialised kernels (PSKB200_TILE_*)
  cudaEvent_t xev = nullptr;                     // orders the download stream behind the library stream
  unsigned long long* trace_buf = nullptr;  // PSK_PHASE_TIMING builds: device trace buffer
  int64_t trace_cap = 0, trace_tiles = 0;
  // psk_eval_async: a ring of in-flight evaluations, each with its own pinned plan staging,
  // pinned read-back words and completion event
  static constexpr int kAsyncSlots = 32;
  struct AsyncSlot {
    cudaEvent_t done = nullptr;
    cudaEvent_t t0 = nullptr, t1 = nullptr;  // bracket the merge kernel when profiling
    bool timed = false;
    int64_t runs_in = 0;
    psk_store_s* owner = nullptr;
  };
  AsyncSlot slots[kAsyncSlots];
  DPlan* h_plans = nullptr;     // [kAsyncSlots] pinned
  int32_t* h_async = nullptr;   // [kAsyncSlots * 8] pinned
  int next_slot = 0;
  DPlan* h_plan = nullptr;      // pinned staging copy of the plan
  int32_t* h_ctrl = nullptr;    // pinned read-back of CTRL words / small results
  std::atomic<int64_t> launches{0};
  std::atomic<int64_t> live_bytes{0};
  std::mutex mu;
  std::recursive_mutex api_mu;  // entry points serialise on the single library stream
};
Ctx g;

psk_status fail(psk_status code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define PSK_CHECK_ARG(cond)                                                              \
  do {                                                                                   \
    if (!(cond)) return fail(PSK_EINVAL, std::string("invalid argument: ") + #cond);     \
  } while (0)
#define PSK_CHECK_STATE(cond)                                                            \
  do {                                                                                   \
    if (!(cond)) return fail(PSK_ESTATE, std::string("invalid state: ") + #cond);        \
  } while (0)
#define CUDA_TRY(expr)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (expr);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return fail(PSK_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));        \
  } while (0)
#define NEED_INIT()                                                                      \
  std::lock_guard<std::recursive_mutex> api_lock_(g.api_mu);                             \
  do {                                                                                   \
    psk_status s_ = ensure_init();                                                       \
    if (s_ != PSK_OK) return s_;                                                         \
  } while (0)
#define LAUNCH(kernel, grid, block, smem, ...)                                           \
  do {                                                                                   \
    PSK_LAUNCH(kernel, grid, block, smem, g.stream, __VA_ARGS__);                        \
    g.launches.fetch_add(1, std::memory_order_relaxed);                                  \
  } while (0)

psk_status do_init(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0)
    return fail(PSK_ECUDA, std::string("no CUDA device available (") +
                               (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                               "); pskb200 has no CPU path");
  if (device < 0 || device >= count) return fail(PSK_EINVAL, "device index out of range");
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&g.h2d, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&g.d2h, cudaStreamNonBlocking));
  CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&g.d_async_err), 16));
  CUDA_TRY(cudaMemset(g.d_async_err, 0, 16));
  CUDA_TRY(cudaEventCreate(&g.ev0));
  CUDA_TRY(cudaEventCreate(&g.ev1));
  CUDA_TRY(cudaEventCreate(&g.pev0));
  CUDA_TRY(cudaEventCreate(&g.pev1));
  CUDA_TRY(cudaMallocHost(reinterpret_cast<void**>(&g.h_plan), sizeof(DPlan)));
  CUDA_TRY(cudaMallocHost(reinterpret_cast<void**>(&g.h_plans), sizeof(DPlan) * Ctx::kAsyncSlots));
  CUDA_TRY(cudaMallocHost(reinterpret_cast<void**>(&g.h_async), sizeof(int32_t) * 8 * Ctx::kAsyncSlots));
  for (int i = 0; i < Ctx::kAsyncSlots; ++i)
  {
    CUDA_TRY(cudaEventCreateWithFlags(&g.slots[i].done, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreate(&g.slots[i].t0));
    CUDA_TRY(cudaEventCreate(&g.slots[i].t1));
  }
  CUDA_TRY(cudaMallocHost(reinterpret_cast<void**>(&g.h_ctrl), 64 * sizeof(int32_t)));
#ifndef PSK_CPU_SIM
  cudaDeviceGetAttribute(&g.sm_count, cudaDevAttrMultiProcessorCount, device);
  {
    // keep freed blocks cached in the stream-ordered pool (no cudaFree round trips)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long thr = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
  }
#endif
  const int smem = (int)kMergeEvalSmemBytes;
  CUDA_TRY(cudaFuncSetAttribute(k_merge_eval_interp<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CUDA_TRY(cudaFuncSetAttribute(k_merge_eval_interp<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CUDA_TRY(cudaFuncSetAttribute(k_merge_eval_interp<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CUDA_TRY(cudaFuncSetAttribute(k_merge_eval_interp<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CUDA_TRY(cudaFuncSetAttribute(k_merge_eval_interp<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  if (const char* sv = getenv("PSKB200_TILE_THREADS")) {  // tuning aid: geometry of the NVRTC kernels
    const int t = atoi(sv);
    if (t >= 32 && t <= 1024 && t % 32 == 0) g.jit_T = t;
  }
  if (const char* sv = getenv("PSKB200_PARK_DEPTH")) {
    const int r = atoi(sv);
    if (r >= 1 && r <= 16) g.jit_regions = r;
  }
  if (const char* sv = getenv("PSKB200_TILE_CAP")) {
    const int c = atoi(sv);
    if (c >= 512 && c <= 13000 && c % 64 == 0) g.jit_CAP = c;
  }
  if (const char* sv = getenv("PSKB200_PROFILE_PARTITION")) g.profile_partition = atoi(sv) != 0;
  if (const char* sv = getenv("PSKB200_SAMPLE_DIV")) {
    const int d = atoi(sv);
    if (d >= 8 && d <= 128) g.sample_div = d;
  }
  g.device = device;
  g.inited = true;
  return PSK_OK;
}

psk_status ensure_init() {
  if (g.inited) return PSK_OK;
  std::lock_guard<std::mutex> l(g.mu);
  if (g.inited) return PSK_OK;
#ifdef PSK_CPU_SIM
  return do_init(0);
#else
  const char* env = getenv("LOCAL_RANK");  // one process per GPU (torchrun)
  return do_init(env ? atoi(env) : 0);
#endif
}

psk_status dmalloc(void** p, size_t bytes) {
  cudaError_t e = cudaMallocAsync(p, bytes ? bytes : 16, g.stream);
  if (e != cudaSuccess) return fail(PSK_ENOMEM, std::string("cudaMallocAsync: ") + cudaGetErrorString(e));
  return PSK_OK;
}

int elem_size(int dtype) { return (dtype == PSK_I32 || dtype == PSK_F32) ? 4 : 1; }

constexpr int64_t kStorePad = 4;  // pairs behind the capacity: 2 sentinels, tile-load slack, the count word

// one allocation per store: capacity + kStorePad pairs, then the compact index
size_t store_bytes(int64_t capacity) {
  return (size_t)(capacity + kStorePad) * 8 + (size_t)(capacity / kIndexStride + 1) * 4;
}

psk_status new_store(int dtype, int64_t capacity, psk_store_t* out) {
  psk_store_s* s = new (std::nothrow) psk_store_s();
  if (!s) return fail(PSK_ENOMEM, "out of host memory");
  s->dtype = dtype;
  s->capacity = capacity;
  psk_status st = dmalloc(reinterpret_cast<void**>(&s->d_runs), store_bytes(capacity));
  if (st != PSK_OK) {
    delete s;
    return st;
  }
  // the last padding pair holds the run count for device-side consumers (queued results)
  s->d_count = reinterpret_cast<int32_t*>(s->d_runs + capacity + kStorePad - 1);
  s->d_index = reinterpret_cast<int32_t*>(s->d_runs + capacity + kStorePad);
  g.live_bytes += capacity * 8;
  *out = s;
  return PSK_OK;
}

// a store produced on a copy stream: make the library stream wait for it (idempotent; the
// event lives until the store is released)
void wait_ready(psk_store_s* s) {
  if (s->ready) cudaStreamWaitEvent(g.stream, s->ready, 0);
}

// Shrink a result whose worst-case allocation is much larger than what it holds (the count is
// known on the host here): copy into an exact-size allocation, free the big one.  Both are
// stream-ordered, so kernels already queued on the old pointer are unaffected.
void right_size(psk_store_s* s) {
  if (s->nruns <= 0 || s->capacity < 4096 || s->capacity <= 2 * s->nruns) return;
  int2* small = nullptr;
  if (cudaMallocAsync(reinterpret_cast<void**>(&small), store_bytes(s->nruns), g.stream) != cudaSuccess) {
    cudaGetLastError();
    return;
  }
  // runs + the two sentinel pairs; the count word moves to the new padding
  cudaMemcpyAsync(small, s->d_runs, (size_t)(s->nruns + 2) * 8, cudaMemcpyDeviceToDevice, g.stream);
  int32_t* new_count = reinterpret_cast<int32_t*>(small + s->nruns + kStorePad - 1);
  cudaMemcpyAsync(new_count, s->d_count, 4, cudaMemcpyDeviceToDevice, g.stream);
  cudaFreeAsync(s->d_runs, g.stream);
  g.live_bytes -= (s->capacity - s->nruns) * 8;
  s->d_runs = small;
  s->d_count = new_count;
  s->d_index = reinterpret_cast<int32_t*>(small + s->nruns + kStorePad);
  s->indexed = false;
  s->capacity = s->nruns;
}

// ---- view descriptors ----------------------------------------------------------------
psk_status make_dview(const psk_view& v, DView* d) {
  memset(d, 0, sizeof(*d));
  d->kind = v.kind;
  if (v.kind == PSK_VIEW_AFFINE) {
    PSK_CHECK_ARG(v.shape[0] > 0 && v.start[0] > 0 && v.stop[0] >= 0);
    d->ndim = 1;
    d->n = (uint32_t)v.shape[0];
    d->c0[0] = (uint32_t)v.start[0];
    d->c1[0] = (uint32_t)v.stop[0];
    long long top = (long long)v.stop[0] + (long long)v.start[0] * v.shape[0];
    PSK_CHECK_ARG(top <= 0x7fffffffLL);
    d->m = (uint32_t)top;
    return PSK_OK;
  }
  PSK_CHECK_ARG(v.kind == PSK_VIEW_GATHER || v.kind == PSK_VIEW_SCATTER);
  PSK_CHECK_ARG(v.ndim >= 1 && v.ndim <= PSK_MAX_DIM);
  d->ndim = v.ndim;
  long long n = 1, m = 1;
  for (int i = 0; i < v.ndim; ++i) {
    // TensorSlice ctor checks, tensor.hpp:62-69, and valid(), :76-84
    PSK_CHECK_ARG(v.shape[i] > 0);
    PSK_CHECK_ARG(0 <= v.start[i] && v.start[i] <= v.stop[i] && v.stride[i] > 0);
    PSK_CHECK_ARG(v.stop[i] <= v.shape[i]);
    d->shape[i] = v.shape[i];
    d->c0[i] = v.start[i];
    d->c1[i] = v.stop[i];
    d->cs[i] = v.stride[i];
    d->cnt[i] = (uint32_t)((v.stop[i] - v.start[i] + v.stride[i] - 1) / v.stride[i]);
    d->scale[i] = (uint32_t)n;
    d->inner[i] = (uint32_t)m;
    n *= v.shape[i];
    m *= d->cnt[i];
    PSK_CHECK_ARG(n <= 0x7fffffffLL);
  }
  PSK_CHECK_ARG(m > 0);  // empty selections never reach eval (array.hpp:118-120)
  d->n = (uint32_t)n;
  d->m = (uint32_t)m;
  return PSK_OK;
}

// span of the view's input and output
void dview_spans(const DView& d, int64_t* in, int64_t* out) {
  if (d.kind == PSK_VIEW_GATHER) { *in = d.n; *out = d.m; }
  else if (d.kind == PSK_VIEW_SCATTER) { *in = d.m; *out = d.n; }
  else { *in = d.n; *out = d.m; }
}

int pick_K(int k) { return k <= 2 ? 2 : k <= 4 ? 4 : k <= 8 ? 8 : k <= 16 ? 16 : 32; }

void launch_merge_eval(int K, int grid, DPlan* d_plan) {
  const size_t smem = kMergeEvalSmemBytes;
  switch (K) {
    case 2: LAUNCH(k_merge_eval_interp<2>, grid, kTileThreads, smem, d_plan); break;
    case 4: LAUNCH(k_merge_eval_interp<4>, grid, kTileThreads, smem, d_plan); break;
    case 8: LAUNCH(k_merge_eval_interp<8>, grid, kTileThreads, smem, d_plan); break;
    case 16: LAUNCH(k_merge_eval_interp<16>, grid, kTileThreads, smem, d_plan); break;
    default: LAUNCH(k_merge_eval_interp<32>, grid, kTileThreads, smem, d_plan); break;
  }
}

psk_status validate_program(int k, int nprog, const psk_node* prog) {
  PSK_CHECK_ARG(nprog >= 1 && nprog <= PSK_MAX_PROGRAM);
  int depth = 0;
  for (int i = 0; i < nprog; ++i) {
    const int op = prog[i].op;
    PSK_CHECK_ARG(op_known(op));
    const int ar = op_arity(op);
    if (op == PSK_OP_SRC) PSK_CHECK_ARG(prog[i].arg < (uint32_t)k);
    PSK_CHECK_ARG(depth >= ar);
    depth += 1 - ar;
    PSK_CHECK_ARG(depth <= kMaxStack);  // lang.hpp:753 stack capacity check
  }
  PSK_CHECK_ARG(depth == 1);
  return PSK_OK;
}

}  // namespace

namespace {
// host / device SoA run arrays -> the store's pair layout; checks 0 < ends[0] < ends[1] < ...
// (flag[0] |= 1 when violated), widens 1-byte payloads, appends the sentinel pairs and the count
template <int ELEM>  // payload bytes: 4, 1 (char, sign-extended), -1 (bool)
__global__ void k_interleave(const int32_t* ends, const void* vals, int32_t n, int2* runs, int32_t* d_count,
                             int32_t* flag) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  // (the arrays may be pinned HOST memory read over PCIe: every element is loaded once, the
  // predecessor's end comes from the neighbouring lane)
  const int32_t e = i < n ? ends[i] : 0;
  int32_t prev = __shfl_up_sync(0xffffffffu, e, 1);
  if ((threadIdx.x & 31) == 0) prev = (i > 0 && i < n) ? ends[i - 1] : 0;
  if (i < n) {
    if (e <= prev) atomicOr(flag, 1);
    uint32_t v;
    if (ELEM == 4) v = reinterpret_cast<const uint32_t*>(vals)[i];
    else if (ELEM == 1) v = (uint32_t)(int32_t) reinterpret_cast<const int8_t*>(vals)[i];
    else v = reinterpret_cast<const uint8_t*>(vals)[i] ? 1u : 0u;
    runs[i] = make_int2(e, (int32_t)v);
  }
  if (i == 0) {
    runs[n] = make_int2(kSentinel, 0);
    runs[n + 1] = make_int2(kSentinel, 0);
    *d_count = n;
  }
}

// the store's pairs -> SoA arrays (Array.runs(), type_binds.hpp:145-151)
template <int ELEM>
__global__ void k_deinterleave(const int2* runs, int32_t n, int32_t* ends, void* vals) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int2 r = runs[i];
    ends[i] = r.x;
    if (ELEM == 4) reinterpret_cast<uint32_t*>(vals)[i] = (uint32_t)r.y;
    else reinterpret_cast<uint8_t*>(vals)[i] = (uint8_t)r.y;
  }
}

// pairs already in the store layout (another GPU's runs, received into a device buffer): copy +
// validate; the ends are shifted by `offset`
__global__ void k_copy_pairs(const int2* in, int32_t n, int32_t offset, int2* out, int32_t* flag) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int2 r = in[i];
    const int32_t prev = i ? in[i - 1].x : 0;
    if (flag && r.x <= prev) atomicOr(flag, 1);
    out[i] = make_int2(r.x + offset, r.y);
  }
}

// sentinel pairs + device-side count behind runs that a kernel wrote (encode, concat, ...)
__global__ void k_seal(int2* runs, int32_t n, int32_t* d_count) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    runs[n] = make_int2(kSentinel, 0);
    runs[n + 1] = make_int2(kSentinel, 0);
    *d_count = n;
  }
}

// (first payload, last payload, run count) of a store, as int64, for the stitch all-gather
__global__ void k_boundary(const int2* runs, int64_t nruns, const int32_t* d_count, long long* out3) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (d_count) nruns = *d_count;  // result of psk_eval_async: the count is still on the device
    out3[0] = (long long)(uint32_t)runs[0].y;
    out3[1] = (long long)(uint32_t)runs[nruns - 1].y;
    out3[2] = (long long)nruns;
  }
}

// psk_store_slice: the runs that cover [start, stop) -- out2[0] = first run whose end is > start,
// out2[1] = first run whose end is >= stop (one warp, 32 probes per round) ...
__global__ void k_slice_find(const int2* runs, int32_t nruns, int32_t start, int32_t stop, int32_t* out2) {
  const int lane = threadIdx.x & 31;
  for (int which = 0; which < 2; ++which) {
    const int32_t key = which == 0 ? start : stop - 1;  // first run whose end is > key
    int32_t lo = 0, hi = nruns;
    while (hi - lo > 32) {
      const int32_t idx = lo + (int32_t)(((long long)(hi - lo) * (lane + 1)) / 33);
      const bool before = ld_nc(runs + idx).x <= key;
      const int nb = __popc(__ballot_sync(0xffffffffu, before));
      const int32_t last_before = __shfl_sync(0xffffffffu, idx, nb > 0 ? nb - 1 : 0);
      const int32_t first_after = __shfl_sync(0xffffffffu, idx, nb < 32 ? nb : 31);
      if (nb < 32) hi = first_after;
      if (nb > 0) lo = last_before + 1;
    }
    const bool before = lo + lane < hi && ld_nc(runs + lo + lane).x <= key;
    const int32_t r = lo + __popc(__ballot_sync(0xffffffffu, before));
    if (lane == 0) out2[which] = r;
  }
}

// ... and their copy, ends relative to start and clipped to the slice
__global__ void k_slice_copy(const int2* runs, int32_t first, int32_t n, int32_t start, int32_t stop, int2* out) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int2 r = ld_nc(runs + first + i);
    out[i] = make_int2((r.x < stop ? r.x : stop) - start, r.y);
  }
}

// index[i] = end of run kIndexStride (i + 1) - 1  (see DSource::index)
__global__ void k_build_index(const int2* runs, int32_t nruns, const int32_t* d_count, int32_t* index) {
  const int32_t n = d_count ? ld_nc(d_count) : nruns;
  const int32_t i = (int32_t)(blockIdx.x * blockDim.x + threadIdx.x);
  if (i < n / kIndexStride) index[i] = ld_nc(runs + (kIndexStride * (i + 1) - 1)).x;
}

__global__ void k_fill_one(int2* runs, int32_t* d_count, int32_t span, uint32_t bits) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    runs[0] = make_int2(span, (int32_t)bits);
    runs[1] = make_int2(kSentinel, 0);
    runs[2] = make_int2(kSentinel, 0);
    *d_count = 1;
  }
}

// d_ends / d_vals (device, SoA) -> pairs of store s on `stream`; *flag accumulates violations
void launch_interleave(psk_store_s* s, const int32_t* d_ends, const void* d_vals, int64_t nruns, int32_t* flag,
                       cudaStream_t stream) {
  const unsigned grid = (unsigned)((nruns + 255) / 256);
  if (elem_size(s->dtype) == 4)
    PSK_LAUNCH(k_interleave<4>, grid, 256, 0, stream, d_ends, d_vals, (int32_t)nruns, s->d_runs, s->d_count, flag);
  else if (s->dtype == PSK_BOOL)
    PSK_LAUNCH(k_interleave<-1>, grid, 256, 0, stream, d_ends, d_vals, (int32_t)nruns, s->d_runs, s->d_count, flag);
  else
    PSK_LAUNCH(k_interleave<1>, grid, 256, 0, stream, d_ends, d_vals, (int32_t)nruns, s->d_runs, s->d_count, flag);
  g.launches.fetch_add(1, std::memory_order_relaxed);
}

// synchronous tail of the validated store constructors: read the flag and the span
psk_status finish_runs_store(psk_store_s* s, int64_t nruns, int32_t* d_flag) {
  CUDA_TRY(cudaMemcpyAsync(g.h_ctrl, d_flag, 4, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaMemcpyAsync(g.h_ctrl + 1, &s->d_runs[nruns - 1].x, 4, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  CUDA_TRY(cudaFreeAsync(d_flag, g.stream));
  if (g.h_ctrl[0] != 0) return fail(PSK_EINVAL, "run ends must be positive and strictly increasing");
  s->nruns = nruns;
  s->span = g.h_ctrl[1];
  return PSK_OK;
}
}  // namespace

namespace {
template <int DT>
psk_status encode_device(const void* d_dense, int64_t n, psk_store_t* out) {
  // 16-byte accesses for 4-byte elements in an aligned buffer
  const bool vec = g.convert_vec && (DT == PSK_I32 || DT == PSK_F32) &&
                   (reinterpret_cast<uintptr_t>(d_dense) & 15) == 0;
  const int32_t nb = (int32_t)((n + (vec ? kEncChunk : kConvChunk) - 1) / (vec ? kEncChunk : kConvChunk));
  if (vec) {
    // one pass over the buffer; the output is sized for one run per two elements and the pass is
    // repeated with the exact size in the (dense) case that this is not enough
    int64_t cap = n / 2 + 4096;
    if (cap > n) cap = n;
    unsigned char* d_scr = nullptr;
    const size_t scr_bytes = (size_t)nb * 8 * kEncStride + 16;
    psk_status st = dmalloc(reinterpret_cast<void**>(&d_scr), scr_bytes);
    if (st != PSK_OK) return st;
    psk_store_t s = nullptr;
    int64_t nruns = 0;
    for (int pass = 0; pass < 2; ++pass) {
      st = new_store(DT, cap, &s);
      if (st != PSK_OK) break;
      cudaMemsetAsync(d_scr, 0, scr_bytes, g.stream);
      int32_t* ctl = reinterpret_cast<int32_t*>(d_scr + (size_t)nb * 8 * kEncStride);
      LAUNCH(k_encode_onepass<DT>, nb, kEncThreads, 0, d_dense, (long long)n, nb,
             reinterpret_cast<unsigned long long*>(d_scr), ctl, s->d_runs, (int32_t)cap);
      cudaError_t e = cudaMemcpyAsync(g.h_ctrl, ctl, 4, cudaMemcpyDeviceToHost, g.stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
      if (e != cudaSuccess) {
        st = fail(PSK_ECUDA, std::string("psk_store_from_dense: ") + cudaGetErrorString(e));
        break;
      }
      nruns = g.h_ctrl[0];
      if (nruns <= cap) break;
      psk_store_release(s);
      s = nullptr;
      cap = nruns;
    }
    cudaFreeAsync(d_scr, g.stream);
    if (st != PSK_OK) {
      if (s) psk_store_release(s);
      return st;
    }
    if (nruns <= 0) {
      psk_store_release(s);
      return fail(PSK_ESTATE, "psk_store_from_dense: produced no runs");
    }
    LAUNCH(k_seal, 1, 32, 0, s->d_runs, (int32_t)nruns, s->d_count);
    s->nruns = nruns;
    s->span = (int32_t)n;
    right_size(s);
    *out = s;
    return PSK_OK;
  }
  int32_t* d_counts = nullptr;
  psk_status st = dmalloc(reinterpret_cast<void**>(&d_counts), (size_t)(nb + 1) * 4);
  if (st != PSK_OK) return st;
  LAUNCH(k_encode_count<DT>, nb, kConvThreads, 0, d_dense, (long long)n, d_counts);
  LAUNCH(k_scan_counts, 1, 1024, 0, d_counts, nb, d_counts + nb);
  CUDA_TRY(cudaMemcpyAsync(g.h_ctrl, d_counts + nb, 4, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  const int64_t nruns = g.h_ctrl[0];
  PSK_CHECK_STATE(nruns > 0);
  psk_store_t s;
  st = new_store(DT, nruns, &s);
  if (st != PSK_OK) {
    cudaFreeAsync(d_counts, g.stream);
    return st;
  }
  LAUNCH(k_encode_write<DT>, nb, kConvThreads, 0, d_dense, (long long)n, d_counts, s->d_runs);
  LAUNCH(k_seal, 1, 32, 0, s->d_runs, (int32_t)nruns, s->d_count);
  CUDA_TRY(cudaFreeAsync(d_counts, g.stream));
  s->nruns = nruns;
  s->span = (int32_t)n;
  *out = s;
  return PSK_OK;
}
}  // namespace

namespace {
template <int KW, int KH, int KD>
void launch_conv(bool is_f, unsigned grid, const ConvParams& cp) {
  auto kf = k_conv_rows<KW, KH, KD, true>;
  auto ki = k_conv_rows<KW, KH, KD, false>;
  if (is_f) LAUNCH(kf, grid, kConvRowThreads, 0, cp);
  else LAUNCH(ki, grid, kConvRowThreads, 0, cp);
}
}  // namespace

// ========================================================================================
extern "C" {

// Completes a store returned by psk_eval_async: waits for its kernels, reads the run count.
psk_status resolve(psk_store_s* s) {
  if (s->pending_slot >= 0) {
    Ctx::AsyncSlot& slot = g.slots[s->pending_slot];
    const int32_t* w = g.h_async + 8 * s->pending_slot;
    cudaError_t e = cudaEventSynchronize(slot.done);
    slot.owner = nullptr;
    s->pending_slot = -1;
    if (e != cudaSuccess) {
      s->failed = PSK_ECUDA;
    } else if (w[4] != 0) {
      cudaMemsetAsync(g.d_async_err, 0, 4, g.stream);
      s->failed = PSK_EINVAL;
    } else if (w[CTRL_ERROR] != 0) {
      s->failed = PSK_ESTATE;
    } else {
      s->nruns = w[CTRL_COUNT];
      float ms = 0.f;
      if (slot.timed && cudaEventElapsedTime(&ms, slot.t0, slot.t1) == cudaSuccess) {
        g.prof_ms += ms;
        g.prof_n += 1;
        g.prof_runs_in += slot.runs_in;
        g.prof_runs_out += s->nruns;
      }
    }
    slot.timed = false;
  }
  if (s->failed == PSK_EINVAL)
    return fail(PSK_EINVAL, "run ends of an asynchronously uploaded store are not positive and strictly increasing");
  if (s->failed == PSK_ESTATE)
    return fail(PSK_ESTATE, "psk_eval: tile capacity exceeded (internal partition bound violated)");
  if (s->failed) return fail(PSK_ECUDA, "psk_eval_async: the evaluation failed on the device");
  return PSK_OK;
}
#define READY(s)                             \
  do {                                       \
    psk_status r_ = resolve(s);              \
    if (r_ != PSK_OK) return r_;             \
    wait_ready(s);                           \
  } while (0)

const char* psk_last_error(void) { return g_err.c_str(); }
const char* psk_version(void) { return "pskb200 0.2 (sm_100a)"; }
int32_t psk_is_simulator(void) {
#ifdef PSK_CPU_SIM
  return 1;
#else
  return 0;
#endif
}

psk_status psk_init(int32_t device) {
  std::lock_guard<std::mutex> l(g.mu);
  if (g.inited) {
    if (device == g.device) return PSK_OK;
    return fail(PSK_ESTATE, "pskb200 already bound to another device in this process");
  }
  return do_init(device);
}

psk_status psk_device_count(int32_t* count) {
  PSK_CHECK_ARG(count != nullptr);
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  *count = (e == cudaSuccess) ? c : 0;
  return PSK_OK;
}

psk_status psk_synchronize(void) {
  NEED_INIT();
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  return PSK_OK;
}

psk_status psk_get_stream(void** stream) {
  NEED_INIT();
  PSK_CHECK_ARG(stream != nullptr);
  *stream = (void*)g.stream;
  return PSK_OK;
}

psk_status psk_set_stream(void* stream) {
  NEED_INIT();
  // adopt a caller-owned stream (e.g. torch's current stream) so that library work and the
  // caller's collectives are ordered on -- and timed by events of -- one stream
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  g.stream = (cudaStream_t)stream;
  return PSK_OK;
}

psk_status psk_timer_start(void) {
  NEED_INIT();
  CUDA_TRY(cudaEventRecord(g.ev0, g.stream));
  return PSK_OK;
}

psk_status psk_timer_stop(float* elapsed_ms) {
  NEED_INIT();
  PSK_CHECK_ARG(elapsed_ms != nullptr);
  CUDA_TRY(cudaEventRecord(g.ev1, g.stream));
  CUDA_TRY(cudaEventSynchronize(g.ev1));
  CUDA_TRY(cudaEventElapsedTime(elapsed_ms, g.ev0, g.ev1));
  return PSK_OK;
}

int64_t psk_launch_count(int32_t reset) {
  int64_t v = g.launches.load();
  if (reset) g.launches.store(0);
  return v;
}

int64_t psk_live_bytes(void) { return g.live_bytes.load(); }

psk_status psk_set_vector_conversion(int32_t enable) {
  std::lock_guard<std::recursive_mutex> api_lock_(g.api_mu);
  g.convert_vec = enable != 0;
  return PSK_OK;
}

psk_status psk_set_small_partition(int64_t max_samples) {
  std::lock_guard<std::recursive_mutex> api_lock_(g.api_mu);
  g.small_partition = max_samples < 0 ? 0 : max_samples;
  return PSK_OK;
}

psk_status psk_set_jit_threshold(int64_t min_total_runs) {
  g.jit_min_runs = min_total_runs;
  return PSK_OK;
}

int64_t psk_jit_launch_count(void) { return g.jit_launches; }

psk_status psk_set_tile_geometry(int32_t threads, int32_t cap, int32_t regions) {
  std::lock_guard<std::recursive_mutex> api_lock_(g.api_mu);
  PSK_CHECK_ARG(threads >= 32 && threads <= 1024 && threads % 32 == 0);
  PSK_CHECK_ARG(cap >= 512 && cap <= 13000 && cap % 64 == 0);
  PSK_CHECK_ARG(regions >= 1 && regions <= 16);  // park depth
  PSK_CHECK_ARG((size_t)kTileRegions * (size_t)(cap + kTileSlack) * 8u + 6144 <= 227 * 1024);
  g.jit_T = threads;
  g.jit_CAP = cap;
  g.jit_regions = regions;
  return PSK_OK;
}

int32_t psk_jit_compile_check(int32_t k, int32_t nprog, const psk_node* prog, char* log, int32_t loglen) {
#ifdef PSK_CPU_SIM
  (void)k; (void)nprog; (void)prog; (void)log; (void)loglen;
  return -1;
#else
  if (validate_program(k, nprog, prog) != PSK_OK) return 2;
  return jit_compile_check(k, false, g.jit_T, g.jit_CAP, g.jit_regions, nprog, prog, log, loglen);
#endif
}

psk_status psk_profile(int32_t enable) {
  g.profile = enable != 0;
  g.prof_ms = 0.0;
  g.prof_n = 0;
  g.prof_runs_in = g.prof_runs_out = 0;
  return PSK_OK;
}

#ifdef PSK_PHASE_TIMING
// development builds only (tools/phase_profile.py); not part of include/pskb200.h
PSK_API void psk_phase_read(unsigned long long* out8, int32_t reset) {
  for (int q = 0; q < 8; ++q) {
    out8[q] = g.phase_cycles[q];
    if (reset) g.phase_cycles[q] = 0;
  }
}
#endif

psk_status psk_profile_read(double* merge_eval_ms, int64_t* evals, int64_t* runs_in, int64_t* runs_out) {
  PSK_CHECK_ARG(merge_eval_ms && evals && runs_in && runs_out);
  *merge_eval_ms = g.prof_ms;
  *evals = g.prof_n;
  *runs_in = g.prof_runs_in;
  *runs_out = g.prof_runs_out;
  return PSK_OK;
}

// ---- stores ----------------------------------------------------------------------------

psk_status psk_store_from_runs(psk_dtype dtype, int64_t nruns, const int32_t* ends, const void* vals,
                               psk_store_t* out) {
  NEED_INIT();
  PSK_CHECK_ARG(out != nullptr && ends != nullptr && vals != nullptr);
  PSK_CHECK_ARG(dtype >= PSK_I32 && dtype <= PSK_CHAR);
  PSK_CHECK_ARG(nruns > 0 && nruns < 0x7fffffffLL);  // core.hpp:33 (n > 0)
  psk_store_t s;
  psk_status st = new_store(dtype, nruns, &s);
  if (st != PSK_OK) return st;
  // staging: [flag | ends | vals] on the device, interleaved into the store by one kernel
  const size_t vbytes = (size_t)nruns * elem_size(dtype);
  const size_t o_ends = 16, o_vals = o_ends + (((size_t)nruns * 4 + 15) & ~(size_t)15);
  unsigned char* tmp = nullptr;
  st = dmalloc(reinterpret_cast<void**>(&tmp), o_vals + vbytes);
  if (st != PSK_OK) {
    psk_store_release(s);
    return st;
  }
  cudaError_t e = cudaMemsetAsync(tmp, 0, 16, g.stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(tmp + o_ends, ends, (size_t)nruns * 4, cudaMemcpyHostToDevice, g.stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(tmp + o_vals, vals, vbytes, cudaMemcpyHostToDevice, g.stream);
  if (e != cudaSuccess) {
    cudaFreeAsync(tmp, g.stream);
    psk_store_release(s);
    return fail(PSK_ECUDA, std::string("H2D copy: ") + cudaGetErrorString(e));
  }
  launch_interleave(s, reinterpret_cast<int32_t*>(tmp + o_ends), tmp + o_vals, nruns, reinterpret_cast<int32_t*>(tmp),
                    g.stream);
  st = finish_runs_store(s, nruns, reinterpret_cast<int32_t*>(tmp));
  if (st != PSK_OK) {
    psk_store_release(s);
    return st;
  }
  *out = s;
  return PSK_OK;
}

psk_status psk_store_from_runs_async(psk_dtype dtype, int64_t nruns, const int32_t* ends, const void* vals,
                                     psk_store_t* out) {
  NEED_INIT();
  PSK_CHECK_ARG(out != nullptr && ends != nullptr && vals != nullptr);
  PSK_CHECK_ARG(dtype == PSK_I32 || dtype == PSK_F32);  // 4-byte payloads only
  PSK_CHECK_ARG(nruns > 0 && nruns < 0x7fffffffLL);
  psk_store_s* s = new (std::nothrow) psk_store_s();
  if (!s) return fail(PSK_ENOMEM, "out of host memory");
  s->dtype = dtype;
  s->capacity = nruns;
  s->nruns = nruns;
  s->span = ends[nruns - 1];  // host memory: readable here
  // everything on the upload stream: the copy engine moves the two host arrays into a staging
  // buffer (SMs stay free for the evaluation kernels of the step in front), one kernel
  // interleaves + validates.  (Measured: letting that kernel read pinned host memory in place
  // over PCIe is slower -- 13.2 vs 9.5 ms per bench step -- because its blocks sit on the SMs
  // while they wait for PCIe and delay the persistent merge kernel.)
  unsigned char* tmp = nullptr;
  const size_t half = ((size_t)nruns * 4 + 15) & ~(size_t)15;
  cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&s->d_runs), store_bytes(nruns), g.h2d);
  if (e == cudaSuccess) e = cudaMallocAsync(reinterpret_cast<void**>(&tmp), 2 * half, g.h2d);
  if (e == cudaSuccess) e = cudaMemcpyAsync(tmp, ends, (size_t)nruns * 4, cudaMemcpyHostToDevice, g.h2d);
  if (e == cudaSuccess) e = cudaMemcpyAsync(tmp + half, vals, (size_t)nruns * 4, cudaMemcpyHostToDevice, g.h2d);
  if (e == cudaSuccess) {
    s->d_count = reinterpret_cast<int32_t*>(s->d_runs + nruns + kStorePad - 1);
    s->d_index = reinterpret_cast<int32_t*>(s->d_runs + nruns + kStorePad);
    launch_interleave(s, reinterpret_cast<int32_t*>(tmp), tmp + half, nruns, g.d_async_err, g.h2d);
    e = cudaFreeAsync(tmp, g.h2d);
  }
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ready, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventRecord(s->ready, g.h2d);
  if (e != cudaSuccess) {
    delete s;
    return fail(PSK_ECUDA, std::string("psk_store_from_runs_async: ") + cudaGetErrorString(e));
  }
  g.live_bytes += nruns * 8;
  *out = s;
  return PSK_OK;
}

psk_status psk_store_runs_async(psk_store_t s, int32_t* ends, void* vals) {
  NEED_INIT();
  PSK_CHECK_ARG(s != nullptr && ends != nullptr && vals != nullptr);
  READY(s);
  PSK_CHECK_ARG(s->dtype == PSK_I32 || s->dtype == PSK_F32);
  const size_t n = (size_t)s->nruns;
  // the download stream waits for everything queued on the library stream so far (the kernels
  // that produce the store) and for the upload, if the store came from one
  if (!g.xev) CUDA_TRY(cudaEventCreateWithFlags(&g.xev, cudaEventDisableTiming));
  CUDA_TRY(cudaEventRecord(g.xev, g.stream));
  CUDA_TRY(cudaStreamWaitEvent(g.d2h, g.xev, 0));
  if (s->ready) CUDA_TRY(cudaStreamWaitEvent(g.d2h, s->ready, 0));
  unsigned char* tmp = nullptr;
  const size_t half = (n * 4 + 15) & ~(size_t)15;
  CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&tmp), 2 * half, g.d2h));
  PSK_LAUNCH(k_deinterleave<4>, (unsigned)((n + 255) / 256), 256, 0, g.d2h, s->d_runs, (int32_t)n,
             reinterpret_cast<int32_t*>(tmp), tmp + half);
  g.launches.fetch_add(1, std::memory_order_relaxed);
  CUDA_TRY(cudaMemcpyAsync(ends, tmp, n * 4, cudaMemcpyDeviceToHost, g.d2h));
  CUDA_TRY(cudaMemcpyAsync(vals, tmp + half, n * 4, cudaMemcpyDeviceToHost, g.d2h));
  CUDA_TRY(cudaFreeAsync(tmp, g.d2h));
  if (!s->busy) CUDA_TRY(cudaEventCreateWithFlags(&s->busy, cudaEventDisableTiming));
  CUDA_TRY(cudaEventRecord(s->busy, g.d2h));
  return PSK_OK;
}

psk_status psk_copy_synchronize(void) {
  NEED_INIT();
  CUDA_TRY(cudaStreamSynchronize(g.h2d));
  CUDA_TRY(cudaStreamSynchronize(g.d2h));
  return PSK_OK;
}

psk_status psk_store_from_device_runs(psk_dtype dtype, int64_t nruns, const int32_t* d_ends,
                                      const uint32_t* d_vals, psk_store_t* out) {
  NEED_INIT();
  PSK_CHECK_ARG(out != nullptr && d_ends != nullptr && d_vals != nullptr);
  PSK_CHECK_ARG(dtype >= PSK_I32 && dtype <= PSK_CHAR);
  PSK_CHECK_ARG(nruns > 0 && nruns < 0x7fffffffLL);
  psk_store_t s;
  psk_status st = new_store(dtype, nruns, &s);
  if (st != PSK_OK) return st;
  int32_t* d_flag = nullptr;
  st = dmalloc(reinterpret_cast<void**>(&d_flag), 16);
  if (st != PSK_OK) {
    psk_store_release(s);
    return st;
  }
  cudaMemsetAsync(d_flag, 0, 16, g.stream);
  // device payloads are always 32-bit words
  PSK_LAUNCH(k_interleave<4>, (unsigned)((nruns + 255) / 256), 256, 0, g.stream, d_ends, (const void*)d_vals,
             (int32_t)nruns, s->d_runs, s->d_count, d_flag);
  g.launches.fetch_add(1, std::memory_order_relaxed);
  st = finish_runs_store(s, nruns, d_flag);
  if (st != PSK_OK) {
    psk_store_release(s);
    return st;
  }
  *out = s;
  return PSK_OK;
}

psk_status psk_store_export_pairs(psk_store_t s, void* d_pairs) {
  NEED_INIT();
  PSK_CHECK_ARG(s != nullptr && d_pairs != nullptr);
  READY(s);
  CUDA_TRY(cudaMemcpyAsync(d_pairs, s->d_runs, (size_t)s->nruns * 8, cudaMemcpyDeviceToDevice, g.stream));
  return PSK_OK;
}

psk_status psk_store_from_device_pairs(psk_dtype dtype, int64_t nruns, const void* d_pairs, psk_store_t* out) {
  NEED_INIT();
  PSK_CHECK_ARG(out != nullptr && d_pairs != nullptr);
  PSK_CHECK_ARG(dtype >= PSK_I32 && dtype <= PSK_CHAR);
  PSK_CHECK_ARG(nruns > 0 && nruns < 0x7fffffffLL);
  psk_store_t s;
  psk_status st = new_store(dtype, nruns, &s);
  if (st != PSK_OK) return st;
  int32_t* d_flag = nullptr;
  st = dmalloc(reinterpret_cast<void**>(&d_flag), 16);
  if (st != PSK_OK) {
    psk_store_release(s);
    return st;
  }
  cudaMemsetAsync(d_flag, 0, 16, g.stream);
  LAUNCH(k_copy_pairs, (unsigned)((nruns + 255) / 256), 256, 0, reinterpret_cast<const int2*>(d_pairs), (int32_t)nruns, 0,
         s->d_runs, d_flag);
  LAUNCH(k_seal, 1, 32, 0, s->d_runs, (int32_t)nruns, s->d_count);
  st = finish_runs_store(s, nruns, d_flag);
  if (st != PSK_OK) {
    psk_store_release(s);
    return st;
  }
  *out = s;
  return PSK_OK;
}

psk_status psk_store_concat(int32_t n, const psk_store_t* parts, psk_store_t* out) {
  NEED_INIT();
  PSK_CHECK_ARG(n >= 1 && parts != nullptr && out != nullptr);
  int64_t total = 0, span = 0;
  for (int i = 0; i < n; ++i) {
    PSK_CHECK_ARG(parts[i] != nullptr);
    READY(parts[i]);
    PSK_CHECK_ARG(parts[i]->dtype == parts[0]->dtype);
    total += parts[i]->nruns;
    span += parts[i]->span;
  }
  PSK_CHECK_ARG(span <= 0x7fffffffLL && total < 0x7fffffffLL);
  psk_store_t s;
  psk_status st = new_store(parts[0]->dtype, total, &s);
  if (st != PSK_OK) return st;
  int64_t r = 0, off = 0;
  for (int i = 0; i < n; ++i) {
    LAUNCH(k_copy_pairs, (unsigned)((parts[i]->nruns + 255) / 256), 256, 0, parts[i]->d_runs, (int32_t)parts[i]->nruns,
           (int32_t)off, s->d_runs + r, (int32_t*)nullptr);
    r += parts[i]->nruns;
    off += parts[i]->span;
  }
  LAUNCH(k_seal, 1, 32, 0, s->d_runs, (int32_t)total, s->d_count);
  s->nruns = total;
  s->span = (int32_t)span;
  *out = s;
  return PSK_OK;
}

psk_status psk_store_fill(psk_dtype dtype, int32_t span, uint32_t fill_bits, psk_store_t* out) {
  NEED_INIT();
  PSK_CHECK_ARG(out != nullptr);
  PSK_CHECK_ARG(dtype >= PSK_I32 && dtype <= PSK_CHAR);
  PSK_CHECK_ARG(span > 0);  // core.hpp:106
  psk_store_t s;
  psk_status st = new_store(dtype, 1, &s);
  if (st != PSK_OK) return st;
  LAUNCH(k_fill_one, 1, 32, 0, s->d_runs, s->d_count, span, fill_bits);
  s->nruns = 1;
  s->span = span;
  *out = s;
  return PSK_OK;
}

psk_status psk_store_from_dense_device(psk_dtype dtype, int64_t n, const void* d_dense, psk_store_t* out) {
  NEED_INIT();
  PSK_CHECK_ARG(out != nullptr && d_dense != nullptr);
  PSK_CHECK_ARG(n > 0 && n <= 0x7fffffffLL);  // conv.hpp:17 (size > 0)
  switch (dtype) {
    case PSK_I32: return encode_device<PSK_I32>(d_dense, n, out);
    case PSK_F32: return encode_device<PSK_F32>(d_dense, n, out);
    case PSK_BOOL: return encode_device<PSK_BOOL>(d_dense, n, out);
    case PSK_CHAR: return encode_device<PSK_CHAR>(d_dense, n, out);
    default: return fail(PSK_EINVAL, "bad dtype");
  }
}

psk_status psk_store_from_dense(psk_dtype dtype, int64_t n, const void* dense, psk_store_t* out) {
  NEED_INIT();
  PSK_CHECK_ARG(out != nullptr && dense != nullptr);
  PSK_CHECK_ARG(dtype >= PSK_I32 && dtype <= PSK_CHAR);
  PSK_CHECK_ARG(n > 0 && n <= 0x7fffffffLL);
  void* d = nullptr;
  const size_t bytes = (size_t)n * elem_size(dtype);
  psk_status st = dmalloc(&d, bytes);
  if (st != PSK_OK) return st;
  cudaError_t e = cudaMemcpyAsync(d, dense, bytes, cudaMemcpyHostToDevice, g.stream);
  if (e != cudaSuccess) {
    cudaFreeAsync(d, g.stream);
    return fail(PSK_ECUDA, std::string("H2D copy: ") + cudaGetErrorString(e));
  }
  st = psk_store_from_dense_device(dtype, n, d, out);
  cudaFreeAsync(d, g.stream);
  return st;
}

psk_status psk_store_to_dense_device(psk_store_t s, void* d_out) {
  NEED_INIT();
  PSK_CHECK_ARG(s != nullptr && d_out != nullptr);
  READY(s);
  const long long n = s->span;
  const unsigned nb = (unsigned)((n + kConvChunk - 1) / kConvChunk);
  if (g.convert_vec && (s->dtype == PSK_I32 || s->dtype == PSK_F32) &&
      (reinterpret_cast<uintptr_t>(d_out) & 15) == 0) {
    // persistent grid, 16-byte stores (the two 4-byte types share the payload path)
    const int32_t nchunks = (int32_t)((n + kExpChunk - 1) / kExpChunk);
    int32_t* d_table = nullptr;
    psk_status st = dmalloc(reinterpret_cast<void**>(&d_table), (size_t)(nchunks + 1) * 4);
    if (st != PSK_OK) return st;
    LAUNCH(k_expand_table, (unsigned)((nchunks + 1 + 255) / 256), 256, 0, (const int2*)s->d_runs, (int32_t)s->nruns, nchunks,
           d_table);
    long long grid = (long long)g.sm_count * 8;
#ifdef PSK_CPU_SIM
    grid = 3;  // (so that the simulator's small cases walk several chunks per block)
#endif
    if (grid > nchunks) grid = nchunks;
    LAUNCH(k_expand_chunks<PSK_I32>, (unsigned)grid, kConvThreads, 0, (const int2*)s->d_runs, (int32_t)s->nruns, n,
           (const int32_t*)d_table, nchunks, d_out);
    cudaFreeAsync(d_table, g.stream);
    CUDA_TRY(cudaGetLastError());
    return PSK_OK;
  }
  switch (s->dtype) {
    case PSK_I32: LAUNCH(k_expand<PSK_I32>, nb, kConvThreads, 0, s->d_runs, (int32_t)s->nruns, n, d_out); break;
    case PSK_F32: LAUNCH(k_expand<PSK_F32>, nb, kConvThreads, 0, s->d_runs, (int32_t)s->nruns, n, d_out); break;
    case PSK_BOOL: LAUNCH(k_expand<PSK_BOOL>, nb, kConvThreads, 0, s->d_runs, (int32_t)s->nruns, n, d_out); break;
    default: LAUNCH(k_expand<PSK_CHAR>, nb, kConvThreads, 0, s->d_runs, (int32_t)s->nruns, n, d_out); break;
  }
  CUDA_TRY(cudaGetLastError());
  return PSK_OK;
}

psk_status psk_store_to_dense(psk_store_t s, void* out_host) {
  NEED_INIT();
  PSK_CHECK_ARG(s != nullptr && out_host != nullptr);
  READY(s);
  void* d = nullptr;
  const size_t bytes = (size_t)s->span * elem_size(s->dtype);
  psk_status st = dmalloc(&d, bytes);
  if (st != PSK_OK) return st;
  st = psk_store_to_dense_device(s, d);
  if (st == PSK_OK) {
    cudaError_t e = cudaMemcpyAsync(out_host, d, bytes, cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    if (e != cudaSuccess) st = fail(PSK_ECUDA, std::string("D2H copy: ") + cudaGetErrorString(e));
  }
  cudaFreeAsync(d, g.stream);
  return st;
}

psk_status psk_store_runs(psk_store_t s, int32_t* ends, void* vals) {
  NEED_INIT();
  PSK_CHECK_ARG(s != nullptr && ends != nullptr && vals != nullptr);
  READY(s);
  const size_t n = (size_t)s->nruns;
  const size_t esz = (size_t)elem_size(s->dtype);
  unsigned char* tmp = nullptr;
  const size_t half = (n * 4 + 15) & ~(size_t)15;
  psk_status st = dmalloc(reinterpret_cast<void**>(&tmp), half + n * esz);
  if (st != PSK_OK) return st;
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (esz == 4) LAUNCH(k_deinterleave<4>, grid, 256, 0, s->d_runs, (int32_t)n, reinterpret_cast<int32_t*>(tmp), tmp + half);
  else LAUNCH(k_deinterleave<1>, grid, 256, 0, s->d_runs, (int32_t)n, reinterpret_cast<int32_t*>(tmp), tmp + half);
  CUDA_TRY(cudaMemcpyAsync(ends, tmp, n * 4, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaMemcpyAsync(vals, tmp + half, n * esz, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaFreeAsync(tmp, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  return PSK_OK;
}

int64_t psk_store_nruns(psk_store_t s) {
  if (!s) return 0;
  if (s->pending_slot >= 0 || s->failed) {
    std::lock_guard<std::recursive_mutex> api_lock_(g.api_mu);
    if (resolve(s) != PSK_OK) return -1;
  }
  return s->nruns;
}
int32_t psk_store_span(psk_store_t s) { return s ? s->span : 0; }
int32_t psk_store_pending(psk_store_t s) { return (s && s->pending_slot >= 0) ? 1 : 0; }
int32_t psk_store_dtype(psk_store_t s) { return s ? s->dtype : -1; }

psk_status psk_store_device_pairs(psk_store_t s, const void** d_pairs) {
  PSK_CHECK_ARG(s != nullptr && d_pairs != nullptr);
  {
    std::lock_guard<std::recursive_mutex> api_lock_(g.api_mu);
    READY(s);
  }
  *d_pairs = s->d_runs;
  return PSK_OK;
}

psk_status psk_store_first_value(psk_store_t s, uint32_t* bits) {
  NEED_INIT();
  PSK_CHECK_ARG(s != nullptr && bits != nullptr);
  READY(s);
  CUDA_TRY(cudaMemcpyAsync(g.h_ctrl, &s->d_runs[0].y, 4, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  *bits = (uint32_t)g.h_ctrl[0];
  return PSK_OK;
}

psk_status psk_store_boundary(psk_store_t s, uint32_t* first_bits, uint32_t* last_bits, int64_t* nruns) {
  NEED_INIT();
  PSK_CHECK_ARG(s != nullptr && first_bits != nullptr && last_bits != nullptr && nruns != nullptr);
  READY(s);
  CUDA_TRY(cudaMemcpyAsync(g.h_ctrl, &s->d_runs[0].y, 4, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaMemcpyAsync(g.h_ctrl + 1, &s->d_runs[s->nruns - 1].y, 4, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  *first_bits = (uint32_t)g.h_ctrl[0];
  *last_bits = (uint32_t)g.h_ctrl[1];
  *nruns = s->nruns;
  return PSK_OK;
}

psk_status psk_store_boundary_device(psk_store_t s, int64_t* d_out3) {
  NEED_INIT();
  PSK_CHECK_ARG(s != nullptr && d_out3 != nullptr);
  // a deferred result is NOT completed here: the kernel reads its run count from the device
  // word of the store, so the caller's pipeline stays asynchronous
  const int32_t* d_count = s->pending_slot >= 0 ? s->d_count : nullptr;
  if (!d_count) READY(s);
  LAUNCH(k_boundary, 1, 32, 0, s->d_runs, (int64_t)s->nruns, d_count, reinterpret_cast<long long*>(d_out3));
  return PSK_OK;
}

void psk_store_retain(psk_store_t s) {
  if (s) s->ref.fetch_add(1);
}

void psk_store_release(psk_store_t s) {
  if (!s) return;
  if (s->ref.fetch_sub(1) == 1) {
    std::lock_guard<std::recursive_mutex> api_lock_(g.api_mu);
    const int32_t slot = s->pending_slot;  // result never looked at
    if (slot >= 0 && slot < Ctx::kAsyncSlots && g.slots[slot].owner == s) g.slots[slot].owner = nullptr;
    // stream-ordered free on the library stream, after any copy still reading / writing it
    if (s->ready) { cudaStreamWaitEvent(g.stream, s->ready, 0); cudaEventDestroy(s->ready); }
    if (s->busy) { cudaStreamWaitEvent(g.stream, s->busy, 0); cudaEventDestroy(s->busy); }
    if (s->d_runs) cudaFreeAsync(s->d_runs, g.stream);
    g.live_bytes -= s->capacity * 8;
    delete s;
  }
}

// ---- views -----------------------------------------------------------------------------
psk_status psk_view_span(int32_t span_in, int32_t nviews, const psk_view* views, int32_t* span_out) {
  PSK_CHECK_ARG(span_out != nullptr && nviews >= 0 && nviews <= PSK_MAX_VIEWS);
  int64_t cur = span_in;
  for (int i = 0; i < nviews; ++i) {
    DView d;
    psk_status st = make_dview(views[i], &d);
    if (st != PSK_OK) return st;
    int64_t in, out;
    dview_spans(d, &in, &out);
    // Array::get / Tensor::get argument checks (array.hpp:117, tensor.hpp:272, :293)
    if (in != cur) return fail(PSK_EINVAL, "view does not match the span of its input");
    cur = out;
  }
  *span_out = (int32_t)cur;
  return PSK_OK;
}

// ---- the hot path ----------------------------------------------------------------------
namespace {

psk_status eval_impl(int32_t k, const psk_source* sources, int32_t nprog, const psk_node* prog,
                     psk_dtype out_dtype, psk_eq eq, bool deferred, psk_store_t* out) {
  PSK_CHECK_ARG(out != nullptr && sources != nullptr && prog != nullptr);
  PSK_CHECK_ARG(k >= 1 && k <= PSK_MAX_SOURCES);  // lang.hpp:789-857
  PSK_CHECK_ARG(out_dtype >= PSK_I32 && out_dtype <= PSK_CHAR);
  psk_status st = validate_program(k, nprog, prog);
  if (st != PSK_OK) return st;
  for (int i = 0; i < k; ++i) {
    PSK_CHECK_ARG(sources[i].store != nullptr);
    // a source that is itself a queued result is NOT waited for: its run count is read on the
    // device (k_prepare) and everything here is sized by its capacity.  Only a result that
    // already failed is reported.
    if (sources[i].store->failed) READY(sources[i].store);
  }
#ifdef PSK_PHASE_TIMING
  deferred = false;
#endif
  // a deferred evaluation stages its plan and reads back its counters in its own ring slot
  int slot = -1;
  if (deferred) {
    slot = g.next_slot;
    g.next_slot = (g.next_slot + 1) % Ctx::kAsyncSlots;
    if (g.slots[slot].owner) resolve(g.slots[slot].owner);  // ring full: finish the oldest
    else cudaEventSynchronize(g.slots[slot].done);          // its owner was released unread
  }

  DPlan& hp = deferred ? g.h_plans[slot] : *g.h_plan;
  hp.k = k;
  hp.eq = (eq == PSK_EQ_TYPED && out_dtype == PSK_F32) ? 1 : 0;
  hp.nprog = nprog;
  hp.epoch = 0;
  memcpy(hp.prog, prog, sizeof(psk_node) * (size_t)nprog);

  int64_t span = -1;
  int64_t total_runs = 0;  // upper bound when a source is a queued result
  bool any_views = false, any_queued = false;
  bool used_async = false;
  for (int i = 0; i < k; ++i) {
    const psk_source& s = sources[i];
    PSK_CHECK_ARG(s.nviews >= 0 && s.nviews <= PSK_MAX_VIEWS);
    DSource& d = hp.src[i];
    d.runs = s.store->d_runs;
    const bool queued = s.store->pending_slot >= 0;
    d.d_count = queued ? s.store->d_count : nullptr;
    d.nruns = (int32_t)(queued ? s.store->capacity : s.store->nruns);
    d.nviews = s.nviews;
    int64_t cur = s.store->span;
    for (int v = 0; v < s.nviews; ++v) {
      st = make_dview(s.views[v], &d.v[v]);
      if (st != PSK_OK) return st;
      int64_t in, o;
      dview_spans(d.v[v], &in, &o);
      if (in != cur) return fail(PSK_EINVAL, "view does not match the span of its input");
      cur = o;
    }
    if (s.nviews) any_views = true;
    if (queued) any_queued = true;
    if (s.store->ready) {  // uploaded by psk_store_from_runs_async: order the kernels after the copy
      wait_ready(s.store);
      used_async = true;
    }
    if (!s.store->indexed && s.store->d_index != nullptr && d.nruns >= kIndexStride) {  // first use as a source: build the compact index
      const int32_t entries = d.nruns / kIndexStride;
      LAUNCH(k_build_index, (unsigned)((entries + 255) / 256), 256, 0, (const int2*)s.store->d_runs, d.nruns,
             (const int32_t*)d.d_count, s.store->d_index);
      s.store->indexed = true;
    }
    d.index = s.store->d_index;
    // every source of a pool must share the output span (eval.hpp:400-403, lang.hpp:198)
    if (span < 0) span = cur;
    else if (span != cur) return fail(PSK_EINVAL, "sources have different spans");
    total_runs += d.nruns;
  }
  PSK_CHECK_ARG(span > 0);  // eval.hpp:611
  PSK_CHECK_ARG(total_runs < 0x7ffffff0LL);
  hp.span = (int32_t)span;
  hp.any_views = any_views ? 1 : 0;
  const bool on_device = any_views || any_queued;
  hp.on_device = on_device ? 1 : 0;

  // ---- which kernel: the specialised one (NVRTC) for large evaluations ---------------------
  int T = kTileThreads, CAP = kTileCap, park_depth = kParkDepth;
#ifndef PSK_CPU_SIM
  const JitKernel* jk = nullptr;
  if (total_runs >= g.jit_min_runs) {
    // K = k exactly: no padding sources in the specialised kernel
    jk = jit_get_merge_eval(k, hp.eq != 0, any_views, g.jit_T, g.jit_CAP, g.jit_regions, nprog, prog, nullptr);
    if (jk) {
      T = g.jit_T;
      CAP = g.jit_CAP;
      park_depth = g.jit_regions;
    }
  }
#endif
  const size_t smem = (size_t)kTileRegions * (size_t)(CAP + kTileSlack) * 8u;
  int ctas = (int)((227 * 1024) / (smem + 6144));
  if (ctas > 2048 / T) ctas = 2048 / T;
  if (ctas < 1) ctas = 1;

  // ---- partition parameters: S * (m + k) <= CAP bounds the runs of one tile ---------------
  // (a tile holds m samples; source q with c_q of them contributes at most (c_q + 1) S - 1 runs)
  int S = CAP / g.sample_div;
  if (k > 8) S = CAP / (2 * g.sample_div);
  while (S > 16 && total_runs / S < 1024) S >>= 1;
  while (S > 4 && CAP / S - k < 4) S >>= 1;
  if (S >= kIndexStride) S -= S % kIndexStride;  // samples on the grid of the stores' compact index
  int64_t max_samples = 0;
  for (int i = 0; i < k; ++i) max_samples += hp.src[i].nruns / S + 2;  // (a source trimmed on the device: its first interval is short)
  const int m_max = CAP / S - k;
  int m = (int)(max_samples / (2 * (int64_t)ctas * g.sm_count));  // aim at >= 2 tiles per resident CTA ...
  const int m_small = (CAP / 4) / S;                              // ... but not at tiles below CAP/4 runs
  if (m < m_small) m = m_small;
  if (m > m_max) m = m_max;
  if (m < 1) m = 1;
  const int64_t max_tiles = (max_samples + m - 1) / m + 1;
  hp.S = S;
  hp.m = m;
  hp.max_tiles = (int32_t)max_tiles;
  {
    int32_t off = 0, blk = 0;
    for (int i = 0; i < k; ++i) {
      DSource& d = hp.src[i];
      d.a = 0;
      d.e = d.nruns;
      d.g0 = first_sample(0, S);
      d.ns = sample_count(0, d.nruns, S);
      d.blk_off = blk;
      blk += (d.ns + kPartBlock - 1) / kPartBlock;
      off += d.ns;
    }
    hp.total_samples = off;
    hp.n_tiles = (off + m - 1) / m;
    if (hp.n_tiles < 1) hp.n_tiles = 1;
  }

  // ---- scratch: [tile_state | phase | ctrl | bounds | tile_hi | lookahead | plan] ------------
  size_t o_state = 0;
  size_t o_phase = o_state + (size_t)max_tiles * 8 * kStateStride;
  size_t o_ctrl = o_phase + 8 * 8;
  size_t o_bounds = o_ctrl + CTRL_WORDS * 4;
  size_t o_hi = o_bounds + (size_t)(max_tiles + 1) * k * 4;
  size_t o_la = (o_hi + (size_t)max_tiles * 4 + 15) & ~(size_t)15;
  size_t o_plan = (o_la + (any_views ? (size_t)max_tiles * k * 16 : 0) + 15) & ~(size_t)15;
  const size_t plan_bytes = offsetof(DPlan, src) + sizeof(DSource) * (size_t)k;
  // persistent grid: as many CTAs as stay resident (shared memory / thread limits)
  int64_t grid = (int64_t)ctas * g.sm_count;
  if (grid > max_tiles) grid = max_tiles;
  // ring of parked tiles per CTA (no views): park_depth slots of CAP pairs
  const size_t o_park = (o_plan + sizeof(DPlan) + 255) & ~(size_t)255;
  const size_t park_bytes = any_views ? 0 : (size_t)grid * (size_t)park_depth * (size_t)CAP * 8u;
  size_t scratch_bytes = o_park + park_bytes;
  unsigned char* scratch = nullptr;
  st = dmalloc(reinterpret_cast<void**>(&scratch), scratch_bytes);
  if (st != PSK_OK) return st;
  psk_store_t res = nullptr;
  st = new_store(out_dtype, total_runs, &res);
  if (st != PSK_OK) {
    cudaFreeAsync(scratch, g.stream);
    return st;
  }
  hp.tile_state = reinterpret_cast<unsigned long long*>(scratch + o_state);
  hp.ctrl = reinterpret_cast<int32_t*>(scratch + o_ctrl);
  hp.phase = reinterpret_cast<unsigned long long*>(scratch + o_phase);
  hp.trace = nullptr;
  hp.bounds = reinterpret_cast<int32_t*>(scratch + o_bounds);
  hp.tile_hi = reinterpret_cast<int32_t*>(scratch + o_hi);
  hp.lookahead = reinterpret_cast<int2*>(scratch + o_la);
  hp.park = reinterpret_cast<int2*>(scratch + o_park);
  hp.out = res->d_runs;
  hp.out_count = res->d_count;
  hp.out_capacity = (int32_t)total_runs;
  DPlan* d_plan = reinterpret_cast<DPlan*>(scratch + o_plan);

  cudaError_t e = cudaMemcpyAsync(d_plan, &hp, plan_bytes, cudaMemcpyHostToDevice, g.stream);
#ifdef PSK_PHASE_TIMING
  if (e == cudaSuccess) e = cudaMemsetAsync(hp.phase, 0, 64, g.stream);
#endif
  if (e == cudaSuccess) {
    if (on_device) LAUNCH(k_prepare, 1, 32 * k, 0, d_plan);
    int64_t pgrid = 0;  // one source per block; a source trimmed on the device has at most nruns / S + 2 samples
    for (int i = 0; i < k; ++i) pgrid += (hp.src[i].nruns / S + 2 + kPartBlock - 1) / kPartBlock;
    // (development: PSKB200_PROFILE_PARTITION=1 makes the profile events bracket k_partition)
    if (g.profile && g.profile_partition) cudaEventRecord(deferred ? g.slots[slot].t0 : g.pev0, g.stream);
    LAUNCH(k_partition, (unsigned)pgrid, kPartBlock, 0, d_plan);
    if (g.profile)
      cudaEventRecord(deferred ? (g.profile_partition ? g.slots[slot].t1 : g.slots[slot].t0)
                               : (g.profile_partition ? g.pev1 : g.pev0), g.stream);
    bool launched = false;
#ifndef PSK_CPU_SIM
    if (jk) {
      if (jit_launch(jk, (unsigned)grid, (unsigned)T, smem, g.stream, d_plan) == 0) {
        launched = true;
        g.launches.fetch_add(1, std::memory_order_relaxed);
        g.jit_launches += 1;
      } else {
        e = cudaErrorUnknown;  // the partition was laid out for the specialised kernel's tiles
      }
    }
#endif
    if (!launched && e == cudaSuccess) launch_merge_eval(pick_K(k), (int)grid, d_plan);
    if (g.profile && !g.profile_partition) cudaEventRecord(deferred ? g.slots[slot].t1 : g.pev1, g.stream);
    if (deferred) {
      g.slots[slot].timed = g.profile;
      g.slots[slot].runs_in = total_runs;
    }
    if (e == cudaSuccess) e = cudaGetLastError();
  }
  if (deferred) {
    int32_t* w = g.h_async + 8 * slot;
    w[4] = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(w, hp.ctrl, CTRL_WORDS * 4, cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess && used_async)
      e = cudaMemcpyAsync(w + 4, g.d_async_err, 4, cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess) e = cudaEventRecord(g.slots[slot].done, g.stream);
    cudaFreeAsync(scratch, g.stream);
    if (e != cudaSuccess) {
      psk_store_release(res);
      return fail(PSK_ECUDA, std::string("psk_eval_async: ") + cudaGetErrorString(e));
    }
    res->nruns = 0;  // known once the store is first used (resolve)
    res->span = (int32_t)span;
    res->pending_slot = slot;
    g.slots[slot].owner = res;
    *out = res;
    return PSK_OK;
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(g.h_ctrl, hp.ctrl, CTRL_WORDS * 4, cudaMemcpyDeviceToHost, g.stream);
  g.h_ctrl[8] = 0;
  if (e == cudaSuccess && used_async)
    e = cudaMemcpyAsync(g.h_ctrl + 8, g.d_async_err, 4, cudaMemcpyDeviceToHost, g.stream);
#ifdef PSK_PHASE_TIMING
  if (e == cudaSuccess) e = cudaMemcpyAsync(g.h_ctrl + 16, hp.phase, 64, cudaMemcpyDeviceToHost, g.stream);
#endif
  if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
  cudaFreeAsync(scratch, g.stream);
  if (e != cudaSuccess) {
    psk_store_release(res);
    return fail(PSK_ECUDA, std::string("psk_eval: ") + cudaGetErrorString(e));
  }
  if (g.h_ctrl[8] != 0) {
    cudaMemsetAsync(g.d_async_err, 0, 4, g.stream);
    psk_store_release(res);
    return fail(PSK_EINVAL, "run ends of an asynchronously uploaded store are not positive and strictly increasing");
  }
  if (g.h_ctrl[CTRL_ERROR] != 0) {
    psk_store_release(res);
    return fail(PSK_ESTATE, "psk_eval: tile capacity exceeded (internal partition bound violated)");
  }
  res->nruns = g.h_ctrl[CTRL_COUNT];
  res->span = (int32_t)span;
#ifdef PSK_PHASE_TIMING
  for (int q = 0; q < 8; ++q) g.phase_cycles[q] += reinterpret_cast<unsigned long long*>(g.h_ctrl + 16)[q];
#endif
  if (g.profile) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g.pev0, g.pev1) == cudaSuccess) {
      g.prof_ms += ms;
      g.prof_n += 1;
      g.prof_runs_in += total_runs;
      g.prof_runs_out += res->nruns;
    }
  }
  if (res->nruns <= 0) {
    psk_store_release(res);
    return fail(PSK_ESTATE, "psk_eval: produced no runs");
  }
  right_size(res);
  *out = res;
  return PSK_OK;
}
}  // namespace

psk_status psk_eval(int32_t k, const psk_source* sources, int32_t nprog, const psk_node* prog,
                    psk_dtype out_dtype, psk_eq eq, psk_store_t* out) {
  NEED_INIT();
  return eval_impl(k, sources, nprog, prog, out_dtype, eq, false, out);
}

psk_status psk_eval_async(int32_t k, const psk_source* sources, int32_t nprog, const psk_node* prog,
                          psk_dtype out_dtype, psk_eq eq, psk_store_t* out) {
  NEED_INIT();
  return eval_impl(k, sources, nprog, prog, out_dtype, eq, true, out);
}


// ---- stencil over runs: conv_2d / conv_3d (src/pyskip/convolve.py:5-59) ------------------
namespace {


}  // namespace

int32_t psk_conv_supported(int32_t ndim, const int32_t* kshape) {
  if (!kshape) return 0;
  if (ndim == 3) return kshape[0] == 3 && kshape[1] == 3 && kshape[2] == 3;
  if (ndim == 2) return kshape[0] == 3 && kshape[1] == 3;
  return 0;
}

psk_status psk_conv_nd(psk_store_t s, int32_t ndim, const int32_t* shape, const int32_t* kshape,
                       const uint32_t* weights, const int32_t* pad, uint32_t fill_bits, psk_eq eq,
                       psk_store_t* out) {
  NEED_INIT();
  PSK_CHECK_ARG(s != nullptr && out != nullptr && shape && kshape && weights && pad);
  PSK_CHECK_ARG(ndim == 2 || ndim == 3);
  if (!psk_conv_supported(ndim, kshape))
    return fail(PSK_EINVAL, "psk_conv_nd: kernel shape not supported by the stencil kernel (3x3, 3x3x3)");
  PSK_CHECK_ARG(s->dtype == PSK_I32 || s->dtype == PSK_F32);  // arithmetic types (type_binds.hpp:174-236)
  if (s->failed) READY(s);
  const bool queued = s->pending_slot >= 0;
  if (!queued) wait_ready(s);
  ConvParams cp;
  memset(&cp, 0, sizeof(cp));
  int64_t n_in = 1, n_out = 1;
  int32_t ext[3] = {1, 1, 1}, pd[3] = {0, 0, 0}, ks[3] = {1, 1, 1}, oext[3] = {1, 1, 1};
  for (int i = 0; i < ndim; ++i) {
    PSK_CHECK_ARG(shape[i] > 0 && pad[i] >= 0);
    ext[i] = shape[i];
    pd[i] = pad[i];
    ks[i] = kshape[i];
    oext[i] = shape[i] + 2 * pad[i] - kshape[i] + 1;
    PSK_CHECK_ARG(oext[i] > 0);
    n_in *= ext[i];
    n_out *= oext[i];
  }
  PSK_CHECK_ARG(n_in == s->span);
  PSK_CHECK_ARG(n_in <= 0x7fffffffLL && n_out <= 0x7fffffffLL);
  // the padded tensor of the reference must be representable too (TensorShape::len, tensor.hpp:38-44)
  PSK_CHECK_ARG((int64_t)(ext[0] + 2 * pd[0]) * (ext[1] + 2 * pd[1]) * (ext[2] + 2 * pd[2]) <= 0x7fffffffLL);
  const int taps = ks[0] * ks[1] * ks[2];
  cp.runs = s->d_runs;
  cp.d_count = queued ? s->d_count : nullptr;
  cp.nruns = (int32_t)(queued ? s->capacity : s->nruns);
  cp.W = ext[0]; cp.H = ext[1]; cp.D = ext[2];
  cp.px = pd[0]; cp.py = pd[1]; cp.pz = pd[2];
  cp.OW = oext[0]; cp.OH = oext[1]; cp.OD = oext[2];
  cp.fill = fill_bits;
  cp.typed_float = (eq == PSK_EQ_TYPED && s->dtype == PSK_F32) ? 1 : 0;
  for (int t = 0; t < taps; ++t) cp.w[t] = weights[t];
  // every output row sees, per tap, the pieces of one input row: its runs + two padding pieces
  const int64_t in_rows = (int64_t)ext[1] * ext[2];
  const int64_t out_rows = (int64_t)oext[1] * oext[2];
  // worst case: every tap contributes all of its pieces to every output row
  int64_t cap_full = (int64_t)taps * ((int64_t)cp.nruns + 3 * in_rows) + 2 * out_rows;
  if (cap_full > n_out) cap_full = n_out;
  // Voxel data stays far below that (BASELINE configs 3 / 4: 4.5x / 8.2x the input runs), so the
  // output is first sized for 12 runs per input piece; the kernel counts every run even when it
  // cannot write them all, and a second pass with the exact size follows in that (dense) case.
  int64_t cap = 12 * ((int64_t)cp.nruns + in_rows) + 2 * out_rows;
  if (cap > cap_full) cap = cap_full;
  const int64_t n_tiles = (out_rows + kConvRowThreads - 1) / kConvRowThreads;
  // scratch: [tile_state | ctrl | row_start]
  const size_t o_ctrl = (size_t)n_tiles * 8 * kStateStride;
  const size_t o_rows = o_ctrl + 16 * 4;
  unsigned char* scratch = nullptr;
  psk_status st = dmalloc(reinterpret_cast<void**>(&scratch), o_rows + (size_t)in_rows * 4);
  if (st != PSK_OK) return st;
  cp.tile_state = reinterpret_cast<unsigned long long*>(scratch);
  cp.ctrl = reinterpret_cast<int32_t*>(scratch + o_ctrl);
  cp.row_start = reinterpret_cast<int32_t*>(scratch + o_rows);
  psk_store_t res = nullptr;
  for (int pass = 0; pass < 2; ++pass) {
    st = new_store(s->dtype, cap, &res);
    if (st != PSK_OK) {
      cudaFreeAsync(scratch, g.stream);
      return st;
    }
    cp.out = res->d_runs;
    cp.out_capacity = (int32_t)cap;
    cp.out_count = res->d_count;
    cudaError_t e = cudaMemsetAsync(scratch, 0, o_rows, g.stream);
    if (e == cudaSuccess) {
      if (pass == 0) LAUNCH(k_conv_row_starts, (unsigned)((in_rows + 255) / 256), 256, 0, cp);
      const bool is_f = s->dtype == PSK_F32;
      if (g.profile) cudaEventRecord(g.pev0, g.stream);
      if (ndim == 3) launch_conv<3, 3, 3>(is_f, (unsigned)n_tiles, cp);
      else launch_conv<3, 3, 1>(is_f, (unsigned)n_tiles, cp);
      if (g.profile) cudaEventRecord(g.pev1, g.stream);
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(g.h_ctrl, cp.ctrl, CTRL_WORDS * 4, cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    if (e != cudaSuccess) {
      cudaFreeAsync(scratch, g.stream);
      psk_store_release(res);
      return fail(PSK_ECUDA, std::string("psk_conv_nd: ") + cudaGetErrorString(e));
    }
    if (g.h_ctrl[CTRL_ERROR] == 0) break;
    // did not fit: the count is exact, repeat with it
    psk_store_release(res);
    res = nullptr;
    const int64_t need = g.h_ctrl[CTRL_COUNT];
    if (pass == 1 || need <= cap || need > cap_full) {
      cudaFreeAsync(scratch, g.stream);
      return fail(PSK_ESTATE, "psk_conv_nd: output capacity exceeded (internal bound violated)");
    }
    cap = need;
  }
  cudaFreeAsync(scratch, g.stream);
  res->nruns = g.h_ctrl[CTRL_COUNT];
  res->span = (int32_t)n_out;
  if (g.profile) {  // the stencil kernel is this entry point's dominant kernel
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g.pev0, g.pev1) == cudaSuccess) {
      g.prof_ms += ms;
      g.prof_n += 1;
      g.prof_runs_in += cp.nruns;
      g.prof_runs_out += res->nruns;
    }
  }
  right_size(res);
  *out = res;
  return PSK_OK;
}

// ---- masks -----------------------------------------------------------------------------
psk_status psk_mask_nd(int32_t ndim, const int32_t* shape, const int32_t* start, const int32_t* stop,
                       const int32_t* stride, psk_store_t* out) {
  NEED_INIT();
  PSK_CHECK_ARG(out != nullptr && shape && start && stop && stride);
  PSK_CHECK_ARG(ndim >= 1 && ndim <= PSK_MAX_DIM);
  psk_view v;
  memset(&v, 0, sizeof(v));
  v.kind = PSK_VIEW_GATHER;
  v.ndim = ndim;
  for (int i = 0; i < ndim; ++i) {
    v.shape[i] = shape[i];
    v.start[i] = start[i];
    v.stop[i] = stop[i];
    v.stride[i] = stride[i];
  }
  DView d;
  psk_status st = make_dview(v, &d);
  if (st != PSK_OK) return st;
  const uint32_t seg_len = d.cs[0] == 1 ? d.cnt[0] : 1;
  const int64_t nseg = d.m / seg_len;
  const int64_t nraw = 2 * nseg + 1;
  PSK_CHECK_ARG(nraw < 0x7fffffffLL);
  // raw (non-canonical) run list, then canonicalise with the eval kernel: an AFFINE
  // identity view routes it through the path that drops zero-length runs
  psk_store_t raw;
  st = new_store(PSK_BOOL, nraw, &raw);
  if (st != PSK_OK) return st;
  LAUNCH(k_mask_segments, (unsigned)((nseg + 255) / 256), 256, 0, d, seg_len, (int32_t)nseg, raw->d_runs);
  raw->nruns = nraw;
  raw->span = (int32_t)d.n;
  psk_source src;
  memset(&src, 0, sizeof(src));
  src.store = raw;
  src.nviews = 1;
  src.views[0].kind = PSK_VIEW_AFFINE;
  src.views[0].ndim = 1;
  src.views[0].shape[0] = (int32_t)d.n;
  src.views[0].start[0] = 1;
  src.views[0].stop[0] = 0;
  psk_node prog{PSK_OP_SRC, 0};
  st = psk_eval(1, &src, 1, &prog, PSK_BOOL, PSK_EQ_BITWISE, out);
  psk_store_release(raw);
  return st;
}

// ---- reductions ------------------------------------------------------------------------
psk_status psk_reduce(psk_store_t s, psk_reduce_op op, uint64_t* result_bits) {
  NEED_INIT();
  PSK_CHECK_ARG(s != nullptr && result_bits != nullptr);
  PSK_CHECK_ARG(op >= PSK_RED_SUM && op <= PSK_RED_PROD);
  // a queued result is reduced without completing it first: the kernel reads its run count
  // on the device and the grid is sized by the capacity
  const bool queued = s->pending_slot >= 0 && !s->failed;
  if (!queued) READY(s);
  const int64_t nmax = queued ? s->capacity : s->nruns;
  const bool is_f = s->dtype == PSK_F32;
  int nblocks = (int)((nmax + kRedThreads - 1) / kRedThreads);
  if (nblocks > kRedBlocks) nblocks = kRedBlocks;
  if (nblocks < 1) nblocks = 1;
  RedAcc* d_part = nullptr;
  psk_status st = dmalloc(reinterpret_cast<void**>(&d_part), sizeof(RedAcc) * (size_t)(nblocks + 1));
  if (st != PSK_OK) return st;
  const int32_t* d_count = queued ? s->d_count : nullptr;
  if (is_f) {
    LAUNCH(k_reduce_partial<true>, nblocks, kRedThreads, 0, s->d_runs, (int32_t)nmax, d_count, (int)op, d_part);
    LAUNCH(k_reduce_final<true>, 1, kRedThreads, 0, d_part, nblocks, (int)op, d_part + nblocks);
  } else {
    LAUNCH(k_reduce_partial<false>, nblocks, kRedThreads, 0, s->d_runs, (int32_t)nmax, d_count, (int)op, d_part);
    LAUNCH(k_reduce_final<false>, 1, kRedThreads, 0, d_part, nblocks, (int)op, d_part + nblocks);
  }
  RedAcc* h = reinterpret_cast<RedAcc*>(g.h_ctrl);
  CUDA_TRY(cudaMemcpyAsync(h, d_part + nblocks, sizeof(RedAcc), cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  CUDA_TRY(cudaFreeAsync(d_part, g.stream));
  if (op == PSK_RED_MAX) *result_bits = h->mx;
  else if (op == PSK_RED_MIN) *result_bits = h->mn;
  else if (is_f) memcpy(result_bits, &h->fsum, 8);
  else *result_bits = h->isum;
  return PSK_OK;
}

psk_status psk_reduce_segments(psk_store_t s, int32_t seg_len, psk_reduce_op op, psk_store_t* out) {
  NEED_INIT();
  PSK_CHECK_ARG(s != nullptr && out != nullptr);
  PSK_CHECK_ARG(op >= PSK_RED_SUM && op <= PSK_RED_PROD);
  PSK_CHECK_ARG(s->dtype == PSK_I32 || s->dtype == PSK_F32);
  PSK_CHECK_ARG(seg_len > 0 && s->span % seg_len == 0);
  const bool queued = s->pending_slot >= 0 && !s->failed;
  if (!queued) READY(s);
  const int32_t nseg = s->span / seg_len;
  uint32_t* d_dense = nullptr;
  psk_status st = dmalloc(reinterpret_cast<void**>(&d_dense), (size_t)nseg * 4);
  if (st != PSK_OK) return st;
  const unsigned grid = (unsigned)(((long long)nseg * 32 + kSegThreads - 1) / kSegThreads);
  const int32_t nmax = (int32_t)(queued ? s->capacity : s->nruns);
  if (s->dtype == PSK_F32)
    LAUNCH(k_reduce_segments<true>, grid, kSegThreads, 0, s->d_runs, nmax, queued ? s->d_count : nullptr, seg_len, nseg,
           (int)op, d_dense);
  else
    LAUNCH(k_reduce_segments<false>, grid, kSegThreads, 0, s->d_runs, nmax, queued ? s->d_count : nullptr, seg_len, nseg,
           (int)op, d_dense);
  // the per-segment values, run-length encoded (TYPED equality, as from_buffer: conv.hpp:21-25)
  st = psk_store_from_dense_device((psk_dtype)s->dtype, nseg, d_dense, out);
  cudaFreeAsync(d_dense, g.stream);
  return st;
}

psk_status psk_store_slice(psk_store_t s, int32_t start, int32_t stop, psk_store_t* out) {
  NEED_INIT();
  PSK_CHECK_ARG(s != nullptr && out != nullptr);
  PSK_CHECK_ARG(0 <= start && start < stop && stop <= s->span);
  // (the same store as an evaluation through a one-dimensional gather view, without the
  //  evaluation: two searches on the device, then a copy of exactly the runs that are needed)
  READY(s);
  int32_t* d_two = nullptr;
  psk_status st = dmalloc(reinterpret_cast<void**>(&d_two), 8);
  if (st != PSK_OK) return st;
  LAUNCH(k_slice_find, 1, 32, 0, (const int2*)s->d_runs, (int32_t)s->nruns, start, stop, d_two);
  cudaError_t e = cudaMemcpyAsync(g.h_ctrl, d_two, 8, cudaMemcpyDeviceToHost, g.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
  cudaFreeAsync(d_two, g.stream);
  if (e != cudaSuccess) return fail(PSK_ECUDA, std::string("psk_store_slice: ") + cudaGetErrorString(e));
  const int32_t first = g.h_ctrl[0], n = g.h_ctrl[1] - g.h_ctrl[0] + 1;
  PSK_CHECK_STATE(n > 0 && first + n <= s->nruns);
  psk_store_t res = nullptr;
  st = new_store(s->dtype, n, &res);
  if (st != PSK_OK) return st;
  LAUNCH(k_slice_copy, (unsigned)((n + 255) / 256), 256, 0, (const int2*)s->d_runs, first, n, start, stop, res->d_runs);
  LAUNCH(k_seal, 1, 32, 0, res->d_runs, n, res->d_count);
  res->nruns = n;
  res->span = stop - start;
  *out = res;
  return PSK_OK;
}

}  // extern "C"
