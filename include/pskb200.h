/* pskb200.h -- C ABI of the B200-native run-length-encoded array evaluation path.
 *
 * This is the drop-in boundary for ONE hot path of turtlesoupy/pyskip: evaluation of
 * run-end encoded arrays (k-way merge of run ends + fused element-wise program per output
 * span + equal-run compaction, through strided / masked N-D views, plus dense<->RLE
 * conversion and reductions).  The reference has no FFI seam of its own: the call that
 * this library replaces is
 *
 *     eval::eval_simple<Ret, Box, StepFn, k>(EvalFn, SimplePool)
 *         include/pyskip/detail/eval.hpp:632-659, called from
 *     lang::execute_plan_fixed            include/pyskip/detail/lang.hpp:749-784
 *
 * whose inputs are an EvalPlan (lang.hpp:602-614: postfix program of SOURCE | MERGE_n) and
 * per source a (store, start, stop, step_fn) tuple (lang.hpp:696-743).  Because a MergeFn
 * is an opaque host function pointer (lang.hpp:56) the program crosses this ABI as
 * op-codes, and because a cyclic::StepFn is a host LUT tree (detail/step.hpp:35-172) a
 * view crosses it as a closed-form strided-box descriptor.
 *
 * Conventions
 *   - plain C types only; no C++ exceptions cross the ABI; no torch types.
 *   - every entry point returns a psk_status; psk_last_error() gives the thread-local
 *     message.  PSK_EINVAL <-> std::invalid_argument (CHECK_ARGUMENT, detail/errors.hpp:7-13),
 *     PSK_ESTATE <-> std::logic_error (CHECK_STATE, errors.hpp:15-21).
 *   - stores are immutable, reference counted, and live in device memory (HBM) as ONE array
 *     of nruns 8-byte pairs { int32 end; uint32 payload }: ends are the strictly increasing
 *     exclusive run ends (the last one == span), the payload is the 32-bit value (int32,
 *     float32 bits, bool 0/1, char sign-extended) -- core::Store (detail/core.hpp:12-62)
 *     interleaved, with the Box (detail/box.hpp:11-127) flattened to its 4 information
 *     bytes, so that a tile of runs is one contiguous piece of memory for the copy engine.
 *     The host-facing calls keep the reference's two-array form (ends[], vals[]).
 *   - all work is issued on one library stream per process; calls that return data to the
 *     host synchronise that stream, everything else is asynchronous.
 *   - there is NO CPU implementation behind this ABI.  Without a CUDA device every compute
 *     entry point fails with PSK_ECUDA.
 */
#ifndef PSKB200_H_
#define PSKB200_H_

#ifdef __CUDACC_RTC__ /* run-time compilation of the kernel templates: no system headers */
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#else
#include <stddef.h>
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define PSK_API
#else
#define PSK_API __attribute__((visibility("default")))
#endif

typedef int32_t psk_status;
enum {
  PSK_OK = 0,
  PSK_EINVAL = 1,  /* bad argument            (reference: std::invalid_argument) */
  PSK_ESTATE = 2,  /* broken invariant        (reference: std::logic_error)      */
  PSK_ECUDA = 3,   /* CUDA runtime / driver error, or no device                  */
  PSK_ENOMEM = 4
};

/* Element types bound by the reference (src/cpp_ext/binds/{int,float,bool,char}_binds.cpp:3-5). */
typedef enum { PSK_I32 = 0, PSK_F32 = 1, PSK_BOOL = 2, PSK_CHAR = 3 } psk_dtype;

/* Equality used by run compaction:
 *   BITWISE: Box::operator== on the 32-bit payload (detail/box.hpp:84-90) -- Array.eval(),
 *            every intermediate materialisation (lang.hpp:890, 902-905);
 *   TYPED  : C++ operator== on the value type (float: NaN != NaN, -0 == +0) -- the final
 *            step of store()/runs()/to_numpy()/get() (lang.hpp:899) and from_buffer
 *            (detail/conv.hpp:21-25).                                                  */
typedef enum { PSK_EQ_BITWISE = 0, PSK_EQ_TYPED = 1 } psk_eq;

typedef struct psk_store_s* psk_store_t; /* opaque, reference counted */

/* ---- views: closed forms of the reference's step functions --------------------------
 * A strided box over a flat array of `shape` (dim 0 fastest, tensor.hpp:102-121) selects
 * positions sel(0) < ... < sel(m-1) (x_d in {start_d, start_d+stride_d, ...} below stop_d).
 *   GATHER : run end E of the source lands at f(E) = |{ j : sel(j) < E }|
 *            (TensorSlice::get_fn tensor.hpp:99-125, Slice::get_fn array.hpp:54-56)
 *   SCATTER: run end E of the (small) source lands at g(E): g(0)=0, g(E)=sel(E), g(m)=n
 *            (TensorSlice::set_fn tensor.hpp:127-169, cyclic::insert_fn step.hpp:583-598)
 *   AFFINE : f(E) = offset + scale*E for 0 < E <= span_in (cyclic::scaled/shift,
 *            step.hpp:319-355); shape[0]=span_in, start[0]=scale, stop[0]=offset.
 * A source carries a chain of up to PSK_MAX_VIEWS views applied innermost (views[0])
 * first (lang::slice_eval, lang.hpp:37-42).                                            */
#define PSK_MAX_DIM 4
#define PSK_MAX_VIEWS 4
#define PSK_MAX_SOURCES 32   /* lang::execute_plan switch, lang.hpp:790-857 */
#define PSK_MAX_PROGRAM 192

typedef enum { PSK_VIEW_GATHER = 1, PSK_VIEW_SCATTER = 2, PSK_VIEW_AFFINE = 3 } psk_view_kind;

typedef struct {
  int32_t kind; /* psk_view_kind */
  int32_t ndim; /* 1..PSK_MAX_DIM */
  int32_t shape[PSK_MAX_DIM];
  int32_t start[PSK_MAX_DIM];
  int32_t stop[PSK_MAX_DIM];
  int32_t stride[PSK_MAX_DIM];
} psk_view;

typedef struct {
  psk_store_t store;
  int32_t nviews;
  psk_view views[PSK_MAX_VIEWS];
} psk_source;

/* ---- fused element-wise program: postfix, one node per EvalPlan::Node (lang.hpp:602-614).
 * Op-codes cover the operator catalogue of include/pyskip/array.hpp:276-338.  The _I
 * family works on int32 payloads (also bool 0/1 and char), the _F family on float32.  */
typedef enum {
  PSK_OP_SRC = 0,   /* push current value of source `arg`                       */
  PSK_OP_CONST = 1, /* push the 32-bit immediate `arg` (a 1-run broadcast scalar,
                       macros.hpp:16-23, folded by the planner)                  */
  /* int32 */
  PSK_OP_ADD_I = 10, PSK_OP_SUB_I, PSK_OP_MUL_I, PSK_OP_DIV_I, PSK_OP_MOD_I, PSK_OP_NEG_I,
  PSK_OP_ABS_I, PSK_OP_BNOT_I, PSK_OP_AND_I, PSK_OP_OR_I, PSK_OP_XOR_I, PSK_OP_SHL_I,
  PSK_OP_SHR_I, PSK_OP_MIN_I, PSK_OP_MAX_I, PSK_OP_POW_I, PSK_OP_SQRT_I, PSK_OP_EXP_I,
  PSK_OP_COALESCE_I, PSK_OP_EQ_I, PSK_OP_NE_I, PSK_OP_LT_I, PSK_OP_GT_I, PSK_OP_LE_I,
  PSK_OP_GE_I, PSK_OP_LNOT_I, PSK_OP_LAND_I, PSK_OP_LOR_I,
  /* float32 (IEEE binary32, round to nearest, no contraction) */
  PSK_OP_ADD_F = 50, PSK_OP_SUB_F, PSK_OP_MUL_F, PSK_OP_DIV_F, PSK_OP_MOD_F, PSK_OP_NEG_F,
  PSK_OP_ABS_F, PSK_OP_MIN_F, PSK_OP_MAX_F, PSK_OP_POW_F, PSK_OP_SQRT_F, PSK_OP_EXP_F,
  PSK_OP_COALESCE_F, PSK_OP_EQ_F, PSK_OP_NE_F, PSK_OP_LT_F, PSK_OP_GT_F, PSK_OP_LE_F,
  PSK_OP_GE_F, PSK_OP_LNOT_F, PSK_OP_LAND_F, PSK_OP_LOR_F,
  /* type agnostic */
  PSK_OP_SELECT = 80, /* (m, a, b) -> m ? a : b   (splat, array.hpp:323-325)   */
  /* casts (array.hpp:277-281) */
  PSK_OP_I2F = 90, PSK_OP_F2I, PSK_OP_I2B, PSK_OP_F2B, PSK_OP_I2C, PSK_OP_F2C
} psk_opcode;

typedef struct {
  int32_t op;   /* psk_opcode */
  uint32_t arg; /* source index or immediate */
} psk_node;

typedef enum { PSK_RED_SUM = 0, PSK_RED_MAX = 1, PSK_RED_MIN = 2, PSK_RED_PROD = 3 } psk_reduce_op;

#ifndef __CUDACC_RTC__ /* the entry points are host functions */
/* ---- library / device ------------------------------------------------------------- */
PSK_API const char* psk_last_error(void);
PSK_API const char* psk_version(void);
/* 0 for the CUDA library.  (The test-only kernel simulator under tests/sim reports 1;
 * the package never loads it.) */
PSK_API int32_t psk_is_simulator(void);
/* Bind this process to CUDA device `device` and create the library stream. */
PSK_API psk_status psk_init(int32_t device);
PSK_API psk_status psk_device_count(int32_t* count);
PSK_API psk_status psk_synchronize(void);
/* Raw cudaStream_t of the library stream (for event timing / interop). */
PSK_API psk_status psk_get_stream(void** stream);
/* Make the library issue all further work on a caller-owned cudaStream_t (e.g. torch's
 * current stream, so NCCL collectives and library kernels share one timeline). */
PSK_API psk_status psk_set_stream(void* stream);
/* Device-side timer on the library stream (CUDA events). */
PSK_API psk_status psk_timer_start(void);
PSK_API psk_status psk_timer_stop(float* elapsed_ms);
/* Number of kernels this library launched since the last reset (bench gpu_launches). */
PSK_API int64_t psk_launch_count(int32_t reset);
/* Bytes currently held by live stores. */
PSK_API int64_t psk_live_bytes(void);
/* Evaluations whose sources hold at least `min_total_runs` runs use a kernel specialised for
 * their program at run time (NVRTC; the generic interpreter kernel otherwise or when NVRTC
 * is unavailable).  Negative: never.  Default 2^18. */
PSK_API psk_status psk_set_jit_threshold(int64_t min_total_runs);
/* Evaluations with at most `max_samples` partition samples run their whole partition step
 * (sample keys, ranking, tile bounds) as one thread block instead of three grid launches,
 * and evaluations that fit one tile skip it altogether.  0: always take the grid path (the
 * tests use this to drive the large-input kernels with small inputs).  Default 2048. */
PSK_API psk_status psk_set_small_partition(int64_t max_samples);
/* Dense <-> RLE conversion of int32 / float32 buffers (psk_store_from_dense*, psk_store_to_dense*)
 * with 16-byte loads and stores when the dense device pointer is 16-byte aligned.  Results are
 * identical; on by default (measured on the B200: 1.2-1.7 x the scalar kernels, profiles/r2_summary.md). */
PSK_API psk_status psk_set_vector_conversion(int32_t enable);
/* NVRTC stage of the specialisation only (needs no GPU): 0 when the kernel specialised for
 * `prog` over k sources compiles to an sm_100a cubin; the compiler log goes to `log`. */
PSK_API int32_t psk_jit_compile_check(int32_t k, int32_t nprog, const psk_node* prog, char* log,
                                      int32_t loglen);
/* Tuning aid: tile geometry of the specialised kernels -- threads per CTA (multiple of 32),
 * runs per tile (multiple of 64) and park depth (finished tiles a CTA may hold back in its
 * global-memory ring while tiles in front of them are still being merged, 1..16).  The
 * partition and the number of resident CTAs follow.  Results do not depend on it.  Also read
 * once from PSKB200_TILE_THREADS / PSKB200_TILE_CAP / PSKB200_PARK_DEPTH. */
PSK_API psk_status psk_set_tile_geometry(int32_t threads, int32_t cap, int32_t park_depth);
/* How many psk_eval calls ran on a specialised kernel. */
PSK_API int64_t psk_jit_launch_count(void);
/* Measurement aid for bench.py's roofline line: when enabled, psk_eval brackets its
 * dominant kernel (k_merge_eval) with CUDA events on the library stream and accumulates
 * the elapsed time, the number of evaluations and the runs read / written. */
PSK_API psk_status psk_profile(int32_t enable);
PSK_API psk_status psk_profile_read(double* merge_eval_ms, int64_t* evals, int64_t* runs_in,
                                    int64_t* runs_out);

/* ---- stores: core::Store (detail/core.hpp:12-62) ----------------------------------- */
/* From host run arrays (core::Store ctor; validated: n>0, ends strictly increasing). */
PSK_API psk_status psk_store_from_runs(psk_dtype dtype, int64_t nruns, const int32_t* ends,
                                       const void* vals, psk_store_t* out);
/* Pipelined host transfers (I32 / F32 stores; host buffers should be pinned): the upload runs
 * on a dedicated copy stream and the first psk_eval that consumes the store waits for it ON
 * THE DEVICE; the download runs on a second copy stream.  With both, step i+1's upload and
 * step i-1's download overlap step i's kernels (PCIe is full duplex).  Validation of the run
 * ends happens on the device and is reported by the consuming psk_eval.
 * psk_copy_synchronize() waits for both copy streams. */
PSK_API psk_status psk_store_from_runs_async(psk_dtype dtype, int64_t nruns, const int32_t* ends,
                                             const void* vals, psk_store_t* out);
PSK_API psk_status psk_store_runs_async(psk_store_t s, int32_t* ends, void* vals);
PSK_API psk_status psk_copy_synchronize(void);
/* Same, from device arrays (copied device-to-device on the library stream). */
PSK_API psk_status psk_store_from_device_runs(psk_dtype dtype, int64_t nruns,
                                              const int32_t* d_ends, const uint32_t* d_vals,
                                              psk_store_t* out);
/* Runs in the store's own pair layout, device to device on the library stream: export into a
 * caller-owned device buffer of nruns pairs (an NCCL send buffer), and a store from nruns
 * validated pairs in device memory (a receive buffer).  The multi-GPU halo exchange moves
 * runs this way: exact sizes, no host copy. */
PSK_API psk_status psk_store_export_pairs(psk_store_t s, void* d_pairs);
PSK_API psk_status psk_store_from_device_pairs(psk_dtype dtype, int64_t nruns, const void* d_pairs,
                                               psk_store_t* out);
/* parts[0] ++ parts[1] ++ ... (spans add up; the result need not be canonical at the seams,
 * which the evaluation path does not require: SURVEY.md A.1).  Used to put halo slices
 * around a z-slab. */
PSK_API psk_status psk_store_concat(int32_t n, const psk_store_t* parts, psk_store_t* out);
/* core::make_store(span, fill) (core.hpp:104-111): one run. */
PSK_API psk_status psk_store_fill(psk_dtype dtype, int32_t span, uint32_t fill_bits,
                                  psk_store_t* out);
/* conv::to_store (detail/conv.hpp:15-44): dense host buffer (element size 4 for
 * I32/F32, 1 for BOOL/CHAR) -> canonical RLE under TYPED equality, encoded on device. */
PSK_API psk_status psk_store_from_dense(psk_dtype dtype, int64_t n, const void* dense,
                                        psk_store_t* out);
PSK_API psk_status psk_store_from_dense_device(psk_dtype dtype, int64_t n,
                                               const void* d_dense, psk_store_t* out);
/* conv::to_buffer (detail/conv.hpp:63-78): expand on device, copy to host. */
PSK_API psk_status psk_store_to_dense(psk_store_t s, void* out_host);
PSK_API psk_status psk_store_to_dense_device(psk_store_t s, void* d_out);
/* Array.runs() (type_binds.hpp:145-151): copy (ends, vals) to host. */
PSK_API psk_status psk_store_runs(psk_store_t s, int32_t* ends, void* vals);
PSK_API int64_t psk_store_nruns(psk_store_t s);
PSK_API int32_t psk_store_span(psk_store_t s);
/* 1 while the store is a queued result (psk_eval_async) whose run count has not reached the
 * host yet: psk_store_nruns would wait for it; psk_eval / psk_reduce / psk_conv_nd /
 * psk_store_boundary_device take it as an input WITHOUT waiting (the count is read on the device). */
PSK_API int32_t psk_store_pending(psk_store_t s);
PSK_API int32_t psk_store_dtype(psk_store_t s);
/* Device pointer of the run array (valid while the store is alive): nruns pairs of
 * { int32 end; uint32 payload }, followed by two sentinel pairs (end = INT32_MAX).  (Behind the
 * padding the same allocation holds the store's compact index -- the ends of every 16th run,
 * filled on first use as a source of psk_eval -- which is private to the library.) */
PSK_API psk_status psk_store_device_pairs(psk_store_t s, const void** d_pairs);
/* The payload of run 0 (used for 1-run stores, i.e. scalars). */
PSK_API psk_status psk_store_first_value(psk_store_t s, uint32_t* bits);
PSK_API void psk_store_retain(psk_store_t s);
PSK_API void psk_store_release(psk_store_t s);

/* ---- the hot path: eval::eval_simple (detail/eval.hpp:632-659) ---------------------- */
/* Evaluate `prog` over `k` sources that share one output span; the result is the
 * canonical RLE under `eq` (equal neighbours merge, the later value is kept:
 * eval.hpp:189-202, :581-588).  out_dtype tags the result. */
PSK_API psk_status psk_eval(int32_t k, const psk_source* sources, int32_t nprog,
                            const psk_node* prog, psk_dtype out_dtype, psk_eq eq,
                            psk_store_t* out);

/* psk_eval without the final wait: the kernels are queued on the library stream and the call
 * returns at once.  The result's run count (and any error of the evaluation) is fetched when
 * the store is first USED -- psk_store_nruns (returns -1 on error), psk_store_runs, a later
 * psk_eval that takes it as a source, ... -- so back-to-back evaluations overlap their host
 * work with the previous one's kernels.  Up to 32 results may be outstanding; more complete
 * the oldest first.  Same arguments and results as psk_eval; the reference has no
 * counterpart (its eval_simple, detail/eval.hpp:632-659, is synchronous). */
PSK_API psk_status psk_eval_async(int32_t k, const psk_source* sources, int32_t nprog,
                                  const psk_node* prog, psk_dtype out_dtype, psk_eq eq,
                                  psk_store_t* out);
/* Output span of a source view chain applied to a store span (lang::slice, lang.hpp:162). */
PSK_API psk_status psk_view_span(int32_t span_in, int32_t nviews, const psk_view* views,
                                 int32_t* span_out);

/* ---- stencil over runs: conv_2d / conv_3d (src/pyskip/convolve.py:5-59) -----------------
 * Cross-correlation (no flip) of an N-D tensor (ndim 2 or 3, shape[] with dim 0 fastest, held
 * as the flat store `s`) with a dense kernel of kshape[], weights[] in the same layout (x
 * fastest: weights[a + KW*(b + KH*c)] = kernel[a, b, c]), after padding every side of dim i
 * with pad[i] elements of `fill_bits`; output extent per dim = shape + 2*pad - kshape + 1.
 * Values are those of the reference's tap loop (convolve.py:52-58: out = 0, then
 * out += kernel[x,y,z] * window, z outer / x inner; int32 wraps, float32 rounds after every
 * multiply and every add); the result is the canonical store under `eq`.  One kernel walks the
 * taps as shifted windows of the same rows; nothing padded is materialised.
 * psk_conv_supported: 1 when the stencil kernel handles this kernel shape (3x3, 3x3x3);
 * other shapes go through psk_eval with gather views, as the reference does. */
PSK_API int32_t psk_conv_supported(int32_t ndim, const int32_t* kshape);
PSK_API psk_status psk_conv_nd(psk_store_t s, int32_t ndim, const int32_t* shape,
                               const int32_t* kshape, const uint32_t* weights, const int32_t* pad,
                               uint32_t fill_bits, psk_eq eq, psk_store_t* out);

/* ---- masks: TensorSlice::mask / mask::stride_mask (tensor.hpp:171-198, mask.hpp:184-200)
 * RLE indicator (BOOL) of the positions a strided box selects, generated on device. */
PSK_API psk_status psk_mask_nd(int32_t ndim, const int32_t* shape, const int32_t* start,
                               const int32_t* stop, const int32_t* stride, psk_store_t* out);

/* ---- reductions (src/pyskip/reduce.py:8-50 collapsed to one pass over runs) ---------
 * SUM: sum(val*len) -- int32 wraps mod 2^32 (result_bits low 32), float32 accumulates in
 * fp64 and returns the fp64 bits; MAX/MIN over run values; PROD = prod(val^len) (int32
 * wrap / fp64). */
PSK_API psk_status psk_reduce(psk_store_t s, psk_reduce_op op, uint64_t* result_bits);

/* Segmented reduction: out[j] = reduction of the flat coordinates [j*seg_len, (j+1)*seg_len)
 * -- reduce(op, t, axis=0) of src/pyskip/reduce.py:8-33 for a tensor whose fastest extent (or
 * product of inner extents) is seg_len; span % seg_len == 0.  One pass over the runs; same
 * arithmetic as psk_reduce (float32 results are the fp64 accumulation rounded once).  The
 * result is a store of span/seg_len elements of the input's dtype (I32 / F32). */
PSK_API psk_status psk_reduce_segments(psk_store_t s, int32_t seg_len, psk_reduce_op op,
                                       psk_store_t* out);

/* ---- multi-GPU helpers (fuse_stores boundary rule, detail/eval.hpp:568-607) ----------
 * A shard's first/last payloads and run count, for the stitch all-gather. */
PSK_API psk_status psk_store_boundary(psk_store_t s, uint32_t* first_bits, uint32_t* last_bits,
                                      int64_t* nruns);
/* Same triple written as 3 x int64 into DEVICE memory on the library stream (no host sync):
 * the send buffer of the all-gather. */
PSK_API psk_status psk_store_boundary_device(psk_store_t s, int64_t* d_out3);
/* Sub-range [start, stop) of a store re-based to 0 (SimpleSource::split, eval.hpp:281-287):
 * two searches on the device and a copy of the runs that cover the range. */
PSK_API psk_status psk_store_slice(psk_store_t s, int32_t start, int32_t stop, psk_store_t* out);
#endif /* !__CUDACC_RTC__ */

#ifdef __cplusplus
}
#endif
#endif /* PSKB200_H_ */
