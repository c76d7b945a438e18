// cuda_sim.h -- TEST INFRASTRUCTURE ONLY.
//
// A minimal "one OS thread per CUDA thread" execution shim so that the kernel sources in
// pyskip_b200/csrc/*.cuh can be compiled with g++ and stepped through on a machine with no
// GPU.  It exists to debug kernel LOGIC (tile partition, merge loops, look-back chain,
// compaction offsets) before spending GPU minutes; it is not an implementation of the
// product, is not importable from the pyskip_b200 package, and reports
// psk_is_simulator() == 1 so that GPU tests and bench.py can refuse it.
//
// Model: blocks of a launch run one after the other (so a persistent kernel's first block
// drains every tile and the decoupled look-back never spins); the threads of a block are
// real threads joined by std::barrier for __syncthreads(); warp collectives are a 32-wide
// barrier plus an exchange array.  A thread that returns early drops out of the barriers,
// as on the GPU.
#pragma once

#include <atomic>
#include <barrier>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct int2 {
  int x, y;
};
inline int2 make_int2(int x, int y) { return int2{x, y}; }
struct alignas(16) uint4 {
  unsigned x, y, z, w;
};
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }

struct psk_sim_uint3 {
  unsigned x = 0, y = 0, z = 0;
};
extern thread_local psk_sim_uint3 threadIdx;
extern psk_sim_uint3 blockIdx, blockDim, gridDim;

namespace psk_sim {

struct BlockCtx {
  std::unique_ptr<std::barrier<>> block_bar;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
  std::vector<uint64_t> xchg;  // [nthreads]
};
extern BlockCtx* g_ctx;
extern unsigned char* g_dyn_smem;

void launch(unsigned grid, unsigned block, size_t smem, const std::function<void()>& fn);
inline unsigned char* dyn_smem() { return g_dyn_smem; }

inline void sync_block() { g_ctx->block_bar->arrive_and_wait(); }
inline void sync_warp() { g_ctx->warp_bar[threadIdx.x >> 5]->arrive_and_wait(); }

template <typename T>
inline T warp_exchange(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "");
  unsigned tid = threadIdx.x, w = tid >> 5;
  unsigned nthreads = blockDim.x;
  uint64_t bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  g_ctx->xchg[tid] = bits;
  sync_warp();
  unsigned src = (w << 5) + (unsigned)(src_lane & 31);
  T out = v;
  if (src < nthreads) {
    uint64_t b = g_ctx->xchg[src];
    std::memcpy(&out, &b, sizeof(T));
  }
  sync_warp();
  return out;
}

}  // namespace psk_sim

inline void __syncthreads() { psk_sim::sync_block(); }
inline void __syncwarp(unsigned = 0xffffffffu) { psk_sim::sync_warp(); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }

template <typename T>
inline T __shfl_sync(unsigned, T v, int src) {
  return psk_sim::warp_exchange(v, src);
}
template <typename T>
inline T __shfl_up_sync(unsigned, T v, unsigned d) {
  int lane = threadIdx.x & 31;
  T o = psk_sim::warp_exchange(v, lane - (int)d < 0 ? lane : lane - (int)d);
  return lane - (int)d < 0 ? v : o;
}
template <typename T>
inline T __shfl_down_sync(unsigned, T v, unsigned d) {
  int lane = threadIdx.x & 31;
  T o = psk_sim::warp_exchange(v, lane + (int)d > 31 ? lane : lane + (int)d);
  return lane + (int)d > 31 ? v : o;
}
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int m) {
  int lane = threadIdx.x & 31;
  return psk_sim::warp_exchange(v, lane ^ m);
}
inline unsigned __ballot_sync(unsigned, int pred) {
  unsigned tid = threadIdx.x, w = tid >> 5, nthreads = blockDim.x;
  psk_sim::g_ctx->xchg[tid] = pred ? 1 : 0;
  psk_sim::sync_warp();
  unsigned out = 0;
  for (unsigned l = 0; l < 32; l++) {
    unsigned t = (w << 5) + l;
    if (t < nthreads && psk_sim::g_ctx->xchg[t]) out |= 1u << l;
  }
  psk_sim::sync_warp();
  return out;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __all_sync(unsigned m, int pred) {
  unsigned w = threadIdx.x >> 5;
  unsigned n = blockDim.x - (w << 5);
  unsigned full = n >= 32 ? 0xffffffffu : ((1u << n) - 1);
  return (__ballot_sync(m, pred) & full) == full;
}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
inline int __ffs(int v) { return __builtin_ffs(v); }

inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
inline float __int_as_float(int u) { float f; std::memcpy(&f, &u, 4); return f; }
inline unsigned long long __double_as_longlong(double d) { unsigned long long u; std::memcpy(&u, &d, 8); return u; }
inline double __longlong_as_double(long long u) { double d; std::memcpy(&d, &u, 8); return d; }
inline float __int2float_rn(int v) { return (float)v; }

template <typename T>
inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <typename T>
inline T atomicExch(T* p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template <typename T>
inline T atomicMax(T* p, T v) {
  T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
template <typename T>
inline T atomicOr(T* p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }

// ---- the handful of runtime calls the ABI layer makes --------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorUnknown = 999 };
typedef void* cudaStream_t;
struct psk_sim_event { std::chrono::steady_clock::time_point t; };
typedef psk_sim_event* cudaEvent_t;
enum { cudaStreamNonBlocking = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };

inline const char* cudaGetErrorString(cudaError_t) { return "simulator"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? cudaSuccess : 2; }
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMallocAsync(void** p, size_t n, cudaStream_t) { return cudaMalloc(p, n); }
inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { return cudaFree(p); }
inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
inline cudaError_t cudaFreeHost(void* p) { return cudaFree(p); }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return cudaSuccess; }
enum { cudaEventDisableTiming = 2 };
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new psk_sim_event(); return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new psk_sim_event(); return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
  return cudaSuccess;
}
template <typename F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
