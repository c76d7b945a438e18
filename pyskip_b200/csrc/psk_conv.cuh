// psk_conv.cuh -- stencil over runs: N-D cross-correlation of a run-length-encoded tensor
// with a small dense kernel, in one pass over the rows.
//
// Semantics: src/pyskip/convolve.py:5-59 (conv_2d / conv_3d).  The reference pads by
// slice-assigning into a filled tensor (convolve.py:47-49 -> tensor.hpp:291-304: mask store +
// scatter view + select), then accumulates
//     out += kernel[x, y, z] * s[x:stop_x, y:stop_y, z:stop_z]          (convolve.py:52-58)
// tap by tap (z outer, y, x inner), every tap a gather view (tensor.hpp:99-125) merged by
// eval::eval (detail/eval.hpp:177-231).  All taps are shifted windows of ONE padded store:
// KW x-shifts of KH*KD rows.  This kernel merges them directly:
//   * one thread per OUTPUT row (dim 0 = x is the run direction, tensor.hpp:102-121);
//   * the thread walks KW*KH*KD streams -- (input row, x-shift) pairs -- held in registers;
//     a stream's pieces are: fill (left padding), the row's runs clipped to the row, fill
//     (right padding); rows outside the tensor are all fill.  Nothing padded is materialised;
//   * per distinct end: the tap sum in the reference's order and arithmetic (int32 wraps;
//     float32 multiplies and adds separately, no contraction, starting from 0), then the
//     keep-iff-the-next-span-differs rule of the eval kernel (eval.hpp:189-202);
//   * rows are counted first (same sweep, no stores), a block scan + decoupled look-back
//     gives every row its place in the output, and the second sweep writes the spans.
//     The last span of a row is dropped when the next row starts with an equal value
//     (fuse_stores rule, eval.hpp:581-588), so the result is the canonical store.
#pragma once

#include "psk_eval.cuh"

namespace psk {

constexpr int kConvRowThreads = 128;
constexpr int kConvMaxTaps = 27;

struct ConvParams {
  const int2* runs;
  const int32_t* d_count;  // run count of a queued input (else null)
  int32_t nruns;
  int32_t W, H, D;         // input extents (x fastest)
  int32_t px, py, pz;      // fill padding per side
  int32_t OW, OH, OD;      // output extents
  uint32_t fill;
  int32_t typed_float;     // compaction equality: 1 = float operator==, 0 = bitwise
  uint32_t w[kConvMaxTaps];  // weights, tap order (z, y, x): w[(c * KH + b) * KW + a]
  int32_t* row_start;      // [H * D] first run overlapping every input row (k_conv_row_starts)
  int2* out;
  int32_t out_capacity;
  int32_t* out_count;
  unsigned long long* tile_state;
  int32_t* ctrl;
};

// first run whose end is beyond the first element of every input row
__global__ void k_conv_row_starts(ConvParams p) {
  const int32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= p.H * p.D) return;
  int32_t n = p.nruns;
  if (p.d_count) n = ld_nc(p.d_count);
  const int32_t pos = row * p.W;
  int32_t lo = 0, hi = n;
  while (lo < hi) {
    const int32_t mid = lo + ((hi - lo) >> 1);
    if (ld_nc(p.runs + mid).x <= pos) lo = mid + 1; else hi = mid;
  }
  p.row_start[row] = lo;
}

template <int KW, int KH, int KD, bool IS_F>
struct ConvRow {
  static constexpr int NR = KH * KD;
  static constexpr int NS = NR * KW;
  int32_t head[NS];   // output x where the stream's current piece ends
  uint32_t val[NS];
  int32_t cur[NS];    // run index of the current piece
  int32_t rowbase[NR];  // flat coordinate of the input row's first element; -1: row is all fill

  PSK_D void init(const ConvParams& p, int32_t y, int32_t z) {
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const int b = r % KH, c = r / KH;
      const int32_t iy = y + b - p.py, iz = z + c - p.pz;
      const bool valid = iy >= 0 && iy < p.H && iz >= 0 && iz < p.D;
      const int32_t rowid = iz * p.H + iy;
      rowbase[r] = valid ? rowid * p.W : -1;
      const int32_t rs = valid ? p.row_start[rowid] : 0;
#pragma unroll
      for (int a = 0; a < KW; ++a) {
        const int s = r * KW + a;
        const int32_t x0 = a - p.px;  // input x of output x = 0
        val[s] = p.fill;
        head[s] = kSentinel;
        cur[s] = rs - 1;
        if (valid && x0 < p.W) {
          if (x0 < 0) {
            head[s] = -x0;  // left padding
          } else {
            int32_t c_ = rs;
            int2 run = ld_nc(p.runs + c_);
            while (run.x - rowbase[r] <= x0) run = ld_nc(p.runs + (++c_));
            cur[s] = c_;
            val[s] = (uint32_t)run.y;
            head[s] = imin(run.x - rowbase[r], p.W) - x0;
          }
        }
      }
    }
  }

  // Streams whose current piece ends at e move on to their next piece.  Straight-line and
  // predicated: the lanes of a warp (different rows) advance different streams at every step,
  // so a branch per stream would execute nearly all of them serially, each with its own
  // dependent load; here the loads of all streams are issued back to back.
  PSK_D void advance_all(const ConvParams& p, int32_t e) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const int r = s / KW, a = s % KW;
      const int32_t shift = p.px - a;
      const bool adv = head[s] == e;
      const bool fin = adv && e == p.W + shift;  // the row is finished: right padding
      const bool ld = adv && !fin;
      int2 run = make_int2(0, 0);
      if (ld) run = ld_nc(p.runs + cur[s] + 1);
      cur[s] += ld ? 1 : 0;
      val[s] = fin ? p.fill : (ld ? (uint32_t)run.y : val[s]);
      head[s] = fin ? kSentinel : (ld ? imin(run.x - rowbase[r], p.W) + shift : head[s]);
    }
  }

  // the tap sum at the current position: ((0 + w0 * v0) + w1 * v1) + ...  (convolve.py:52-58)
  PSK_D uint32_t eval(const ConvParams& p) const {
    if (IS_F) {
      float acc = 0.0f;
#pragma unroll
      for (int s = 0; s < NS; ++s) acc = acc + u2f(p.w[s]) * u2f(val[s]);
      return f2u(acc);
    }
    uint32_t acc = 0;
#pragma unroll
    for (int s = 0; s < NS; ++s) acc = acc + p.w[s] * val[s];
    return acc;
  }

  // Sweep one output row.  WRITE = false: count the spans kept before the row's last span (the
  // first `stage_cap` of them are also parked at stage[j * stage_stride], so that rows with
  // few spans -- nearly all of them -- need no second sweep); WRITE = true: store them at dst.
  // Returns the count; *fv / *lv = value of the first / last span of the row.
  template <bool WRITE>
  PSK_D int32_t sweep(const ConvParams& p, int32_t y, int32_t z, int2* dst, uint32_t* fv, uint32_t* lv,
                      int2* stage = nullptr, int stage_cap = 0, int stage_stride = 1) {
    init(p, y, z);
    const int32_t rowflat = (z * p.OH + y) * p.OW;
    const bool typed = p.typed_float != 0;
    int32_t n = 0, pe = 0;
    uint32_t pv = 0;
    bool have = false;
    for (;;) {
      int32_t e = head[0];
#pragma unroll
      for (int s = 1; s < NS; ++s) e = imin(e, head[s]);
      if (e > p.OW) e = p.OW;
      const uint32_t v = eval(p);
      if (!have) {
        *fv = v;
      } else if (!payload_equal(pv, v, typed)) {
        if (WRITE) dst[n] = make_int2(rowflat + pe, (int32_t)pv);
        else if (n < stage_cap) stage[n * stage_stride] = make_int2(rowflat + pe, (int32_t)pv);
        ++n;
      }
      pe = e;
      pv = v;
      have = true;
      if (e >= p.OW) break;
      advance_all(p, e);
    }
    *lv = pv;
    return n;
  }

  // value of the row's first span only
  PSK_D uint32_t first_value(const ConvParams& p, int32_t y, int32_t z) {
    init(p, y, z);
    return eval(p);
  }
};

template <int KW, int KH, int KD, bool IS_F>
__global__ void __launch_bounds__(kConvRowThreads) k_conv_rows(ConvParams p) {
  constexpr int T = kConvRowThreads;
  // spans of the rows of this tile, parked during the counting sweep: span j of thread t at
  // [j * T + t] (conflict-free for 8-byte accesses)
  constexpr int kStage = 40;
  __shared__ int2 s_stage[kStage * T];
  __shared__ uint32_t s_fv[T + 1];
  __shared__ int32_t s_warp[T / 32 + 2];
  __shared__ int32_t s_lbsum[LookW<T>::value * (T / 32)];
  __shared__ int32_t s_lbhas[LookW<T>::value * (T / 32)];
  __shared__ int32_t s_tile;
  const int tid = threadIdx.x;
  if (tid == 0) s_tile = atomicAdd(&p.ctrl[CTRL_TICKET], 1);
  __syncthreads();
  const int32_t tile = s_tile;
  const int32_t nrows = p.OH * p.OD;
  const int32_t n_tiles = (nrows + T - 1) / T;
  if (tile >= n_tiles) return;
  const int32_t row = tile * T + tid;
  const bool active = row < nrows;
  const int32_t y = active ? row % p.OH : 0, z = active ? row / p.OH : 0;

  ConvRow<KW, KH, KD, IS_F> cr;
  uint32_t fv = 0, lv = 0;
  int32_t cnt = 0;
  if (active) cnt = cr.template sweep<false>(p, y, z, nullptr, &fv, &lv, s_stage + tid, kStage, T);
  s_fv[tid] = fv;
  if (tid == T - 1) {  // first value of the row behind the tile
    const int32_t nrow = tile * T + T;
    s_fv[T] = nrow < nrows ? cr.first_value(p, nrow % p.OH, nrow / p.OH) : 0u;
  }
  __syncthreads();
  const bool keep_last = active && (row == nrows - 1 || !payload_equal(lv, s_fv[tid + 1], p.typed_float != 0));
  const int32_t n = cnt + (keep_last ? 1 : 0);
  int32_t total;
  const int32_t ofs = block_exclusive_scan<T>(n, s_warp, &total);
  if (tid == 0) publish_aggregate(p.tile_state, tile, total);
  const int32_t gofs = lookback_exclusive_cta<T>(p.tile_state, tile, total, s_lbsum, s_lbhas);
  if (tid == 0 && tile == n_tiles - 1) {
    const int32_t c = gofs + total;
    p.ctrl[CTRL_COUNT] = c;
    *p.out_count = c;
    if (c <= p.out_capacity) {
      p.out[c] = make_int2(kSentinel, 0);
      p.out[c + 1] = make_int2(kSentinel, 0);
    }
  }
  if (gofs + total > p.out_capacity) {
    if (tid == 0) p.ctrl[CTRL_ERROR] = 2;
    return;
  }
  if (active) {
    int2* const dst = p.out + gofs + ofs;
    if (cnt <= kStage) {
      for (int32_t j = 0; j < cnt; ++j) dst[j] = s_stage[j * T + tid];
    } else {
      cr.template sweep<true>(p, y, z, dst, &fv, &lv);  // a row with many spans: sweep it again
    }
    if (keep_last) dst[cnt] = make_int2((z * p.OH + y) * p.OW + p.OW, (int32_t)lv);
  }
}


}  // namespace psk
