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
#if defined(__CUDA_ARCH__)
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

// volatile / relaxed accessors used by the decoupled look-back chain
PSK_D unsigned long long ld_volatile_u64(const unsigned long long* p) {
  return *reinterpret_cast<const volatile unsigned long long*>(p);
}
PSK_D void st_volatile_u64(unsigned long long* p, unsigned long long v) {
  *reinterpret_cast<volatile unsigned long long*>(p) = v;
}

}  // namespace psk
