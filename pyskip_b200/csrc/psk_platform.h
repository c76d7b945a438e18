// psk_platform.h -- the few macros that let the SAME kernel sources compile either with
// nvcc for sm_100a (the product) or, for development without a GPU, with g++ against the
// thread-per-CUDA-thread simulator in tests/sim/cuda_sim.h (PSK_CPU_SIM; test-only, never
// loaded by the package -- see DESIGN.md "Kernel simulator").
#pragma once

#ifndef __CUDACC_RTC__
#include <stdint.h>
#endif

#ifdef PSK_CPU_SIM
#include "cuda_sim.h"
#define PSK_LAUNCH(kernel, grid, block, smem, stream, ...) \
  psk_sim::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); })
#define PSK_DYN_SMEM(name) unsigned char* name = psk_sim::dyn_smem()
#define PSK_DYN_SMEM_PAIRS(name) int2* name = reinterpret_cast<int2*>(psk_sim::dyn_smem())
#define PSK_HD inline
#define PSK_D inline
#else
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#endif
#define PSK_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define PSK_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define PSK_DYN_SMEM_PAIRS(name) extern __shared__ __align__(16) int2 name[]
#define PSK_HD __host__ __device__ __forceinline__
#define PSK_D __device__ __forceinline__
#endif

namespace psk {

PSK_HD int imin(int a, int b) { return a < b ? a : b; }
PSK_HD int imax(int a, int b) { return a > b ? a : b; }

// read-only global loads (ld.global.nc): never alias shared-memory stores, so the compiler
// may hoist them above the staging stores
PSK_D int32_t ld_nc(const int32_t* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}
PSK_D uint32_t ld_nc(const uint32_t* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}

PSK_D int2 ld_nc(const int2* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}

// load that bypasses L1 (data written by other threads of the SM through L2)
PSK_D int2 ld_cg(const int2* p) {
#if defined(__CUDA_ARCH__)
  return __ldcg(p);
#else
  return *p;
#endif
}

// predicated read-only load: *p if take, else dflt -- one @p LDG, never a branch
PSK_D int32_t ld_nc_if(const int32_t* p, bool take, int32_t dflt) {
#if defined(__CUDA_ARCH__)
  int32_t r;
  asm volatile("{\n .reg .pred q;\n setp.ne.s32 q, %2, 0;\n mov.b32 %0, %3;\n @q ld.global.nc.b32 %0, [%1];\n}"
               : "=&r"(r)
               : "l"(p), "r"((int)take), "r"(dflt));
  return r;
#else
  return take ? *p : dflt;
#endif
}

// The tile buffer in dynamic shared memory, addressed by 32-bit shared-window offsets with
// explicit ld.shared / st.shared: nvcc otherwise re-derives the generic address of dynamic
// shared memory (S2R SR_CgaCtaId + LEA) inside the merge loop.
struct SmemPairs {
#ifndef PSK_CPU_SIM
  uint32_t base;
  PSK_D explicit SmemPairs(void* p) {
    // volatile: computed once and kept in a register (never rematerialised in the loops)
    asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(base) : "l"(p));
  }
  PSK_D int2 ld(int i) const {
    int2 r;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(base + (uint32_t)i * 8u));
    return r;
  }
  PSK_D int ldx(int i) const {
    int r;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(r) : "r"(base + (uint32_t)i * 8u));
    return r;
  }
  PSK_D void st(int i, int x, int y) const {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(base + (uint32_t)i * 8u), "r"(x), "r"(y) : "memory");
  }
  // A cursor is the shared-window byte address of a pair: loops that walk the buffer keep
  // cursors in registers and never recompute base + 8 * index.
  PSK_D uint32_t at(int i) const { return base + (uint32_t)i * 8u; }
  PSK_D uint32_t dst(int i) const { return base + (uint32_t)i * 8u; }  // bulk-copy destination
  PSK_D int ldy(int i) const {
    int r;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(r) : "r"(base + (uint32_t)i * 8u + 4u));
    return r;
  }
  PSK_D void stx(int i, int x) const {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + (uint32_t)i * 8u), "r"(x) : "memory");
  }
  static PSK_D uint32_t step() { return 8u; }
  PSK_D void st_at(uint32_t c, int x, int y) const {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(c), "r"(x), "r"(y) : "memory");
  }
  // merge step of one source: if its head run ends at e, move on to its next run
  // (predicated load: no branch, no select chain)
  PSK_D void advance_if(int e, uint32_t& c, int& head, uint32_t& val) const {
    asm volatile(
        "{\n .reg .pred q;\n setp.eq.s32 q, %0, %3;\n @q add.u32 %2, %2, 8;\n"
        " @q ld.shared.v2.b32 {%0, %1}, [%2];\n}"
        : "+r"(head), "+r"(val), "+r"(c)
        : "r"(e));
  }
#else
  int2* ptr;
  explicit SmemPairs(void* p) : ptr(reinterpret_cast<int2*>(p)) {}
  uint32_t at(int i) const { return (uint32_t)i; }
  unsigned char* dst(int i) const { return reinterpret_cast<unsigned char*>(ptr + i); }
  int ldy(int i) const { return ptr[i].y; }
  void stx(int i, int x) const { ptr[i].x = x; }
  static uint32_t step() { return 1u; }
  void st_at(uint32_t c, int x, int y) const { ptr[c] = make_int2(x, y); }
  void advance_if(int e, uint32_t& c, int& head, uint32_t& val) const {
    if (head == e) {
      ++c;
      head = ptr[c].x;
      val = (uint32_t)ptr[c].y;
    }
  }
  int2 ld(int i) const { return ptr[i]; }
  int ldx(int i) const { return ptr[i].x; }
  void st(int i, int x, int y) const { ptr[i] = make_int2(x, y); }
#endif
};


// ---- mbarrier + bulk asynchronous copy (cp.async.bulk, the 1-D TMA form) ------------------
// One thread arms the barrier with the byte count of the copies it issues; the copy engine
// completes the phase when the bytes have landed in shared memory; every consumer thread
// waits on the phase parity.  Source, destination and size must be multiples of 16 bytes.
// (Simulator: the copy is a memcpy by the issuing thread; the barrier is a phase counter.)
#ifndef PSK_CPU_SIM
typedef unsigned long long psk_mbar_t;
PSK_D uint32_t smem_addr(const void* p) {
  uint32_t a;
  asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(a) : "l"(p));
  return a;
}
PSK_D void mbar_init(psk_mbar_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
PSK_D void mbar_arrive_expect_tx(psk_mbar_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes)
               : "memory");
}
PSK_D void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, psk_mbar_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
      "l"(src), "r"(bytes), "r"(smem_addr(bar))
      : "memory");
}
PSK_D void mbar_wait(psk_mbar_t* bar, uint32_t parity) {
  asm volatile(
      "{\n .reg .pred q;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 q, [%0], %1;\n"
      " @q bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}" ::"r"(smem_addr(bar)),
      "r"(parity)
      : "memory");
}
#else
struct psk_mbar_t {
  volatile long long pending;
  volatile unsigned phases;  // completed phases
};
inline void mbar_init(psk_mbar_t* bar, int) {
  bar->pending = 0;
  bar->phases = 0;
}
inline void mbar_arrive_expect_tx(psk_mbar_t* bar, uint32_t bytes) {
  if (bytes == 0) {
    __atomic_fetch_add(&bar->phases, 1u, __ATOMIC_SEQ_CST);
  } else {
    __atomic_store_n(&bar->pending, (long long)bytes, __ATOMIC_SEQ_CST);
  }
}
inline void bulk_g2s(unsigned char* dst_smem, const void* src, uint32_t bytes, psk_mbar_t* bar) {
  memcpy(dst_smem, src, bytes);
  if (__atomic_sub_fetch(&bar->pending, (long long)bytes, __ATOMIC_SEQ_CST) == 0)
    __atomic_fetch_add(&bar->phases, 1u, __ATOMIC_SEQ_CST);
}
inline void mbar_wait(psk_mbar_t* bar, uint32_t parity) {
  while ((__atomic_load_n(&bar->phases, __ATOMIC_SEQ_CST) & 1u) == parity) std::this_thread::yield();
}
#endif

PSK_D void sleep_ns(unsigned ns) {
#if defined(__CUDA_ARCH__)
  while (ns > 0) {  // __nanosleep takes at most ~1 ms
    unsigned step = ns > 500000u ? 500000u : ns;
    __nanosleep(step);
    ns -= step;
  }
#else
  (void)ns;
#endif
}

// relaxed device-scope accessors used by the decoupled look-back chain (a word carries both
// status and count, so no ordering with other memory is needed; device scope is enough --
// `volatile` would compile to system-scope accesses)
PSK_D unsigned long long ld_volatile_u64(const unsigned long long* p) {
#if defined(__CUDA_ARCH__)
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
#else
  return *reinterpret_cast<const volatile unsigned long long*>(p);
#endif
}
PSK_D void st_volatile_u64(unsigned long long* p, unsigned long long v) {
#if defined(__CUDA_ARCH__)
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#else
  *reinterpret_cast<volatile unsigned long long*>(p) = v;
#endif
}

}  // namespace psk
