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
#define PSK_HD inline
#define PSK_D inline
#else
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#endif
#define PSK_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define PSK_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
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

// volatile / relaxed accessors used by the decoupled look-back chain
PSK_D unsigned long long ld_volatile_u64(const unsigned long long* p) {
  return *reinterpret_cast<const volatile unsigned long long*>(p);
}
PSK_D void st_volatile_u64(unsigned long long* p, unsigned long long v) {
  *reinterpret_cast<volatile unsigned long long*>(p) = v;
}

}  // namespace psk
