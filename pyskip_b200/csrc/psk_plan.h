// psk_plan.h -- device-resident description of one evaluation (what psk_eval uploads) and
// the closed-form view maps.  Shared by host and device code.
//
// Reference behaviour restated here:
//   * cyclic::StepFn::operator() / search     include/pyskip/detail/step.hpp:105-160
//   * TensorSlice::get_fn / set_fn            include/pyskip/tensor.hpp:99-169
//   * Slice::get_fn / cyclic::insert_fn       include/pyskip/array.hpp:54-60, step.hpp:583-598
// The reference walks a tree of LUT leaves with a one-entry cache; here every map is a
// closed form so that any thread can evaluate it for any run end.
#pragma once

#include "../../include/pskb200.h"
#include "psk_platform.h"

namespace psk {

struct DView {
  int32_t kind;
  int32_t ndim;
  uint32_t shape[PSK_MAX_DIM];
  uint32_t c0[PSK_MAX_DIM];
  uint32_t c1[PSK_MAX_DIM];
  uint32_t cs[PSK_MAX_DIM];
  uint32_t cnt[PSK_MAX_DIM];    // selected coordinates per dim
  uint32_t scale[PSK_MAX_DIM];  // flat stride of dim d            (prod shape[<d])
  uint32_t inner[PSK_MAX_DIM];  // selected positions per unit of dim d (prod cnt[<d])
  uint32_t n;                   // flat length of the viewed array (span_in for AFFINE)
  uint32_t m;                   // number of selected positions
};

constexpr int kIndexStride = 16;

struct DSource {
  // (end, payload) pairs, 8 bytes per run, ends strictly increasing; pairs [nruns] and
  // [nruns + 1] are sentinels (end = INT_MAX) and the allocation extends 4 pairs beyond nruns,
  // so tile loads may run past the last run without a bounds test
  const int2* runs;
  // run count of a source that is itself a queued (deferred) result: read on the device by
  // k_prepare; nruns then holds the capacity (an upper bound)
  const int32_t* d_count;
  // ends of every kIndexStride-th run, compact: index[i] = runs[kIndexStride (i + 1) - 1].x for
  // i < nruns / kIndexStride (built once per store, on first use as a source); the partition
  // kernel reads its samples here instead of gathering them from the run array
  const int32_t* index;
  int32_t nruns;
  int32_t nviews;
  int32_t a, e;      // raw run-index range [a, e) that can contribute to the output
  int32_t g0;        // first sample: the first run index >= a with (index + 1) % S == 0
  int32_t ns;        // number of partition samples taken from this source
  int32_t blk_off;   // first block of k_partition that ranks samples of this source (blocks never mix sources)
  DView v[PSK_MAX_VIEWS];
};

struct DPlan {
  int32_t k;
  int32_t span;
  int32_t S;  // sample stride (runs)
  int32_t m;  // merged samples per tile
  int32_t eq;
  int32_t nprog;
  int32_t total_samples;
  int32_t n_tiles;
  int32_t max_tiles;
  int32_t out_capacity;
  int32_t any_views;
  int32_t on_device;  // 1: sample / tile counts are computed by k_prepare (views, deferred sources)
  uint32_t epoch;     // tags the look-back words of this evaluation (the state buffer is reused)
  // scratch + output (device pointers)
  int32_t* bounds;                 // [(max_tiles+1) * k] per-tile raw index bounds (total order)
  int2* lookahead;                 // [max_tiles * k] with views: first (mapped) run beyond the tile's hi
  int32_t* tile_hi;                // [max_tiles] inclusive upper coordinate of each tile
  unsigned long long* tile_state;  // [max_tiles] look-back chain
  int32_t* ctrl;                   // [0]=ticket [1]=out_count [2]=error
  unsigned long long* phase;       // [8] cycle counters (PSK_PHASE_TIMING builds only)
  unsigned long long* trace;       // [max_tiles * 8] per-tile globaltimer stamps (same builds)
  int2* park;                      // [grid * kParkDepth * kTileCap] parked spans (no views)
  int2* out;                       // result pairs [out_capacity + 4]
  int32_t* out_count;              // the result's run count, left on the device for consumers
  psk_node prog[PSK_MAX_PROGRAM];
  DSource src[PSK_MAX_SOURCES];
};

enum { CTRL_TICKET = 0, CTRL_COUNT = 1, CTRL_ERROR = 2, CTRL_WORDS = 4 };

// ---- closed-form view maps (all values < 2^31, unsigned 32-bit arithmetic) -------------

// f(E) = |{ j : sel(j) < E }|
PSK_HD uint32_t map_gather(const DView& v, uint32_t E) {
  if (E == 0) return 0;
  if (E >= v.n) return v.m;
  uint32_t out = 0, rem = E;
  for (int d = v.ndim - 1; d >= 0; --d) {
    const uint32_t sc = v.scale[d], c0 = v.c0[d], cs = v.cs[d];
    uint32_t x = rem;
    if (sc != 1) {  // dim 0 has unit stride: no division
      x = rem / sc;
      rem -= x * sc;
    }
    uint32_t below = 0;
    bool sel;
    if (cs == 1) {  // contiguous selection (plain boxes, convolution taps): no division
      if (x > c0) below = x - c0 < v.cnt[d] ? x - c0 : v.cnt[d];
      sel = x >= c0 && x < v.c1[d];
    } else {
      if (x > c0) {
        below = (x - c0 + cs - 1) / cs;
        if (below > v.cnt[d]) below = v.cnt[d];
      }
      sel = x >= c0 && x < v.c1[d] && ((x - c0) % cs) == 0;
    }
    out += below * v.inner[d];
    if (!sel) break;
  }
  return out;
}

// g(0)=0, g(E)=sel(E) for 0<E<m, g(m)=n
PSK_HD uint32_t map_scatter(const DView& v, uint32_t E) {
  if (E == 0) return 0;
  if (E >= v.m) return v.n;
  uint32_t flat = 0, rem = E;
  for (int d = 0; d < v.ndim; ++d) {
    uint32_t j = rem % v.cnt[d];
    rem /= v.cnt[d];
    flat += (v.c0[d] + j * v.cs[d]) * v.scale[d];
  }
  return flat;
}

// f(E) = offset + scale*E on (0, span_in]
PSK_HD uint32_t map_affine(const DView& v, uint32_t E) {
  if (E == 0) return 0;
  if (E > v.n) E = v.n;
  return v.c1[0] + v.c0[0] * E;
}

PSK_HD uint32_t map_view(const DView& v, uint32_t E) {
  if (v.kind == PSK_VIEW_GATHER) return map_gather(v, E);
  if (v.kind == PSK_VIEW_SCATTER) return map_scatter(v, E);
  return map_affine(v, E);
}

PSK_HD int32_t map_chain(const DView* v, int nviews, int32_t E) {
  uint32_t x = (uint32_t)E;
  for (int i = 0; i < nviews; ++i) x = map_view(v[i], x);
  return (int32_t)x;
}

PSK_HD int32_t mapped_end(const DSource& s, int32_t idx) {
  int32_t E = s.runs[idx].x;
  return s.nviews ? map_chain(s.v, s.nviews, E) : E;
}

}  // namespace psk
