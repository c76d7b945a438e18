// psk_jit.h -- run-time specialisation of k_merge_eval for one fused program.
//
// The reference evaluates a fused DAG by interpreting EvalPlan::nodes through function
// pointers once per output span (lang::execute_plan_fixed, include/pyskip/detail/lang.hpp:
// 754-777).  The generic kernel k_merge_eval_interp does the same with op-codes.  For large
// evaluations that interpreter costs several times more issue slots than the merge itself,
// so the library compiles the SAME hand-written kernel template (psk_eval.cuh,
// merge_eval_tile_loop) with the program spliced in as straight-line code: NVRTC -> cubin
// for sm_100a -> cuModuleLoadData, cached by (program shape, K).  Immediates stay run-time
// data (read from the plan), so e.g. all 3x3x3 convolution kernels share one module.
//
// libnvrtc and libcuda are bound with dlopen at first use, so the library still loads on a
// machine without a driver; if either is missing the interpreter kernel is used.
#pragma once

#include <stddef.h>
#include <stdint.h>

#include "../../include/pskb200.h"

namespace psk {

struct JitKernel {
  void* function;  // CUfunction
};

// Returns nullptr when JIT is unavailable/disabled or compilation failed (the reason is
// appended to *why when given).
// T / CAP: threads per CTA and runs per tile the kernel is compiled for (-DPSK_TILE_THREADS / -DPSK_TILE_CAP).
const JitKernel* jit_get_merge_eval(int K, bool typed_float, bool views, int T, int CAP, int regions, int nprog, const psk_node* prog,
                                    const char** why);

// cuLaunchKernel on `stream`; returns 0 on success.
int jit_launch(const JitKernel* k, unsigned grid, unsigned block, size_t smem, void* stream, void* plan_ptr);

// NVRTC stage only (no driver needed): 0 when the specialised source compiles to an sm_100a
// cubin.  Used by the CPU-side tests.
int jit_compile_check(int K, bool typed_float, int T, int CAP, int regions, int nprog, const psk_node* prog, char* log, int loglen);

// number of modules compiled so far (diagnostics / tests)
int jit_compiled_count();

}  // namespace psk
