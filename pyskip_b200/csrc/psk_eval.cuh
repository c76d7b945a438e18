// psk_eval.cuh -- the hot path: k-way merge of (view-mapped) run ends, one evaluation of the
// fused element-wise program per output span, equal-run compaction.
//
// Replaces eval::eval<sources> (include/pyskip/detail/eval.hpp:177-231: tournament tree +
// dedup hash, one output span at a time), eval::partition_pool (:448-462: uniform split of
// the output span over CPU threads) and eval::fuse_stores (:568-607: concatenation with
// boundary compaction), which the reference drives from lang::execute_plan_fixed
// (detail/lang.hpp:749-784).
//
// B200 design:
//   k_prepare      (only with views) trims every source to the run range that can reach the
//                  output -- the role of lang::make_pool's slice_invert calls (lang.hpp:708-709).
//   k_sample_rank  takes every S-th mapped run end of every source and ranks it among all
//                  samples (k binary searches): a sorted splitter list balanced by RUN COUNT,
//                  not by coordinate (the reference's TODO at eval.hpp:450-451).
//   k_tile_bounds  every m-th splitter closes a tile; one binary search per (tile, source)
//                  gives the tile's run slice.  A tile holds <= (m + 2k) * S runs, so it
//                  always fits in shared memory.
//   k_merge_eval   persistent CTAs take tiles in order (atomic ticket), stage the tile's
//                  run slices in shared memory (coalesced loads, views applied on the fly,
//                  zero-length mapped runs dropped), split the tile's coordinate range over
//                  the threads, and every thread runs the sequential k-way merge over its
//                  sub-range with the heads of all sources in registers.  A span is kept iff
//                  its value differs from the NEXT span's value (the later value wins, as
//                  in eval.hpp:189-202), which is a thread-local decision, so compaction is a
//                  block scan plus a decoupled look-back across tiles -- one pass over HBM.
#pragma once

#include "psk_ops.h"
#include "psk_plan.h"

namespace psk {

constexpr int kTileThreads = 256;
constexpr int kTileCap = 6144;  // runs of all sources staged per tile (excl. look-ahead)
constexpr int kSentinel = 0x7fffffff;

// Development aid (-DPSK_PHASE_TIMING, tools/phase_profile.py): thread 0 of every CTA adds the
// clock64() cycles it spends in each phase of the tile loop to p->phase[].  Not compiled into
// the product library.
#ifdef PSK_PHASE_TIMING
#define PSK_PHASE_DECL long long ph_t = clock64(); long long ph_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define PSK_PHASE(i)                                   \
  do {                                                 \
    if (threadIdx.x == 0) {                            \
      long long now_ = clock64();                      \
      ph_acc[i] += now_ - ph_t;                        \
      ph_t = now_;                                     \
    }                                                  \
  } while (0)
#define PSK_PHASE_FLUSH                                                                      \
  do {                                                                                       \
    if (threadIdx.x == 0)                                                                    \
      for (int q_ = 0; q_ < 8; ++q_) atomicAdd(&p->phase[q_], (unsigned long long)ph_acc[q_]); \
  } while (0)
#else
#define PSK_PHASE_DECL
#define PSK_PHASE(i)
#define PSK_PHASE_FLUSH
#endif

// first index in [lo, hi) whose mapped end is > key   (upper bound)
PSK_D int32_t ub_mapped(const DSource& s, int32_t lo, int32_t hi, int32_t key) {
  while (lo < hi) {
    int32_t mid = lo + ((hi - lo) >> 1);
    if (mapped_end(s, mid) <= key) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// first index in [lo, hi) whose mapped end is >= key  (lower bound)
PSK_D int32_t lb_mapped(const DSource& s, int32_t lo, int32_t hi, int32_t key) {
  while (lo < hi) {
    int32_t mid = lo + ((hi - lo) >> 1);
    if (mapped_end(s, mid) < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// ---------------------------------------------------------------------------------------
// k_prepare: one thread per source.  [a, e) = runs whose mapped end is in (0, span], cut
// after the first run that reaches span (everything later is a zero-length mapped run).
__global__ void k_prepare(DPlan* p) {
  __shared__ int32_t s_ns[PSK_MAX_SOURCES];
  int i = threadIdx.x;
  int k = p->k;
  if (i < k) {
    DSource& s = p->src[i];
    int32_t a = 0, e = s.nruns;
    if (s.nviews) {
      a = ub_mapped(s, 0, s.nruns, 0);
      e = lb_mapped(s, a, s.nruns, p->span) + 1;
      if (e > s.nruns) e = s.nruns;
    }
    s.a = a;
    s.e = e;
    int32_t n = e - a;
    s.ns = (n + p->S - 1) / p->S;
    s_ns[i] = s.ns;
  }
  __syncthreads();
  if (i == 0) {
    int32_t off = 0;
    for (int q = 0; q < k; ++q) {
      p->src[q].samp_off = off;
      off += s_ns[q];
    }
    p->total_samples = off;
    int32_t nt = (off + p->m - 1) / p->m;
    p->n_tiles = nt < 1 ? 1 : nt;
  }
}

PSK_D int32_t sample_key(const DSource& s, int32_t j, int32_t S) {
  int32_t n = s.e - s.a;
  int32_t off = (j + 1) * S;
  if (off > n) off = n;
  return mapped_end(s, s.a + off - 1);
}

// ---------------------------------------------------------------------------------------
// k_sample_rank: merged[rank] = key, where rank orders all samples by (key, source, index).
__global__ void k_sample_rank(DPlan* p) {
  int32_t sid = blockIdx.x * blockDim.x + threadIdx.x;
  if (sid >= p->total_samples) return;
  const int k = p->k, S = p->S;
  int i = 0;
  while (i + 1 < k && p->src[i + 1].samp_off <= sid) ++i;
  const int32_t j = sid - p->src[i].samp_off;
  const int32_t key = sample_key(p->src[i], j, S);
  int32_t rank = j;
  for (int q = 0; q < k; ++q) {
    if (q == i) continue;
    const DSource& s = p->src[q];
    int32_t lo = 0, hi = s.ns;
    while (lo < hi) {
      int32_t mid = lo + ((hi - lo) >> 1);
      int32_t kq = sample_key(s, mid, S);
      bool before = (q < i) ? (kq <= key) : (kq < key);
      if (before) lo = mid + 1; else hi = mid;
    }
    rank += lo;
  }
  p->merged[rank] = key;
}

// ---------------------------------------------------------------------------------------
// k_tile_bounds: bounds[t][i] = first run of source i whose mapped end is > hi(t-1).
__global__ void k_tile_bounds(DPlan* p) {
  int32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = p->k, nt = p->n_tiles, m = p->m;
  if (id >= (nt + 1) * k) return;
  int32_t t = id / k, i = id - t * k;
  const DSource& s = p->src[i];
  int32_t b;
  if (t == 0) {
    b = s.a;
  } else if (t == nt) {
    b = s.e;
  } else {
    b = ub_mapped(s, s.a, s.e, p->merged[t * m - 1]);
  }
  p->bounds[t * k + i] = b;
  if (i == 0 && t > 0) p->tile_hi[t - 1] = (t == nt) ? p->span : p->merged[t * m - 1];
}

// ---------------------------------------------------------------------------------------
// Block-wide exclusive scan of one int per thread (warp shuffles + one shared hop).
template <int T>
PSK_D int32_t block_exclusive_scan(int32_t v, int32_t* s_warp /*[T/32+1]*/, int32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int32_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int32_t o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int32_t w = (lane < T / 32) ? s_warp[lane] : 0;
    int32_t winc = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int32_t o = __shfl_up_sync(0xffffffffu, winc, d);
      if (lane >= d) winc += o;
    }
    if (lane < T / 32) s_warp[lane] = winc - w;  // exclusive warp offsets
    if (lane == 31) s_warp[T / 32] = winc;       // block total
  }
  __syncthreads();
  int32_t excl = s_warp[warp] + inc - v;
  *total = s_warp[T / 32];
  __syncthreads();
  return excl;
}

// Look-back chain word: (status << 32) | count.
enum { ST_INVALID = 0, ST_AGGREGATE = 1, ST_PREFIX = 2 };

// Decoupled look-back, executed by one full warp: lane l inspects tile (j - l); the nearest
// tile that already published an inclusive PREFIX ends the walk, the AGGREGATEs in front of
// it are summed.  Tiles are handed out by an atomic ticket, so every predecessor is resident
// or finished and the spin cannot deadlock.
PSK_D int32_t lookback_exclusive_warp(unsigned long long* state, int32_t tile, int32_t count) {
  const int lane = threadIdx.x & 31;
  if (tile == 0) {
    if (lane == 0) st_volatile_u64(&state[0], ((unsigned long long)ST_PREFIX << 32) | (uint32_t)count);
    return 0;
  }
  if (lane == 0) st_volatile_u64(&state[tile], ((unsigned long long)ST_AGGREGATE << 32) | (uint32_t)count);
  int32_t excl = 0;
  for (int32_t j = tile - 1;; j -= 32) {
    const int32_t idx = j - lane;
    unsigned long long w = (unsigned long long)ST_PREFIX << 32;  // virtual prefix 0 before tile 0
    if (idx >= 0) {
      do {
        w = ld_volatile_u64(&state[idx]);
      } while ((w >> 32) == ST_INVALID);
    }
    const unsigned pmask = __ballot_sync(0xffffffffu, (w >> 32) == ST_PREFIX);
    const int first = pmask ? (__ffs((int)pmask) - 1) : 32;
    int32_t c = (lane <= first) ? (int32_t)(uint32_t)w : 0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    excl += c;
    if (pmask) break;
  }
  if (lane == 0)
    st_volatile_u64(&state[tile], ((unsigned long long)ST_PREFIX << 32) | (uint32_t)(excl + count));
  return excl;
}

// ---------------------------------------------------------------------------------------
// Evaluators: the interpreter walks the postfix program held in shared memory.  (A JIT
// specialisation replaces this type with straight-line code, see psk_jit.)
struct InterpEval {
  const psk_node* prog;
  int nprog;
  template <typename ValArray>
  PSK_D uint32_t operator()(const ValArray& val) const {
    return run_program(prog, nprog, val);
  }
};

// ---------------------------------------------------------------------------------------
// k_merge_eval<K>: K = compile-time bound on the number of sources (heads live in registers).
template <int K, typename Eval>
PSK_D void merge_eval_tile_loop(DPlan* p, unsigned char* smem_raw, const Eval& ev) {
  constexpr int T = kTileThreads;
  constexpr int CAP = kTileCap;
  constexpr int IN_N = CAP + PSK_MAX_SOURCES;  // room for one look-ahead run per source

  int32_t* inE = reinterpret_cast<int32_t*>(smem_raw);
  uint32_t* inV = reinterpret_cast<uint32_t*>(inE + IN_N);
  int32_t* outE = reinterpret_cast<int32_t*>(inV + IN_N);
  uint32_t* outV = reinterpret_cast<uint32_t*>(outE + CAP);

  __shared__ int32_t s_off[PSK_MAX_SOURCES + 1];  // start of each source's slice in inE/inV
  __shared__ int32_t s_b0[PSK_MAX_SOURCES], s_b1[PSK_MAX_SOURCES];
  __shared__ int32_t s_warp[T / 32 + 2];
  __shared__ int32_t s_tile, s_gofs, s_carry, s_fail;

  const int tid = threadIdx.x;
  const int k = p->k;
  const bool typed_float = p->eq != 0;
  const int32_t span = p->span;
  const int32_t n_tiles = p->n_tiles;

  PSK_PHASE_DECL;
  for (;;) {
    if (tid == 0) s_tile = atomicAdd(&p->ctrl[CTRL_TICKET], 1);
    __syncthreads();
    const int32_t tile = s_tile;
    if (tile >= n_tiles) break;
    PSK_PHASE(0);
    const int32_t lo = tile == 0 ? 0 : p->tile_hi[tile - 1];
    const int32_t hi = p->tile_hi[tile];
    if (tid < k) {
      s_b0[tid] = p->bounds[tile * k + tid];
      s_b1[tid] = p->bounds[(tile + 1) * k + tid];
    }
    if (tid == 0) { s_off[0] = 0; s_fail = 0; }
    __syncthreads();

    // ---- stage every source's slice [b0, b1) plus one look-ahead run -------------------
    for (int i = 0; i < k; ++i) {
      const DSource& s = p->src[i];
      const int32_t b0 = s_b0[i], b1 = s_b1[i];
      const int32_t off = s_off[i];
      int32_t kept;
      if (s.nviews == 0) {
        kept = b1 - b0;
        if (off + kept + 1 <= IN_N) {
          for (int32_t j = tid; j < kept; j += T) {
            inE[off + j] = s.ends[b0 + j];
            inV[off + j] = s.vals[b0 + j];
          }
        }
      } else {
        // mapped ends; a run whose mapped end equals its predecessor's is empty -> dropped
        if (tid == 0) s_carry = lo;
        kept = 0;
        for (int32_t base = b0; base < b1; base += T) {
          __syncthreads();
          const int32_t idx = base + tid;
          const bool valid = idx < b1;
          const int32_t me = valid ? map_chain(s.v, s.nviews, s.ends[idx]) : kSentinel;
          // predecessor's mapped end: previous lane, previous warp's last lane, or carry
          int32_t* s_last = s_warp;  // reuse: [T/32] last mapped end of each warp
          int32_t pm = __shfl_up_sync(0xffffffffu, me, 1);
          if ((tid & 31) == 31) s_last[tid >> 5] = me;
          __syncthreads();
          if ((tid & 31) == 0) pm = (tid == 0) ? s_carry : s_last[(tid >> 5) - 1];
          const bool keep = valid && me > pm;
          __syncthreads();  // s_last consumed before the scan reuses s_warp
          int32_t tot;
          const int32_t pos = block_exclusive_scan<T>(keep ? 1 : 0, s_warp, &tot);
          if (keep && off + kept + pos < IN_N) {
            inE[off + kept + pos] = me;
            inV[off + kept + pos] = s.vals[idx];
          }
          kept += tot;
          // carry = mapped end of the last valid run of this chunk
          const int32_t last_valid = (b1 - base < T ? b1 - base : T) - 1;
          if (tid == last_valid) s_carry = me;
        }
        __syncthreads();
      }
      if (tid == 0) {
        if (off + kept + 1 > IN_N) {
          s_fail = 1;
          kept = 0;
        }
        // look-ahead: the first run beyond the tile (its mapped end is > hi) or a sentinel
        if (b1 < s.e) {
          inE[off + kept] = mapped_end(s, b1);
          inV[off + kept] = s.vals[b1];
        } else {
          inE[off + kept] = kSentinel;
          inV[off + kept] = 0;
        }
        s_off[i + 1] = off + kept + 1;
      }
      __syncthreads();
    }
    PSK_PHASE(1);  // staging
    const bool fail = s_fail != 0;
    if (fail && tid == 0) p->ctrl[CTRL_ERROR] = 1;

    // ---- split the tile's coordinate range (lo, hi] over the threads --------------------
    const long long width = (long long)hi - lo;
    const int32_t c0 = lo + (int32_t)((width * tid) / T);
    const int32_t c1 = lo + (int32_t)((width * (tid + 1)) / T);

    int32_t cur[K], head[K];
    uint32_t val[K];
    int32_t base = 0;
#pragma unroll
    for (int i = 0; i < K; ++i) {
      if (i < k && !fail) {
        // first staged run of source i whose end is > c0 (the look-ahead run bounds it)
        int32_t l = s_off[i], h = s_off[i + 1] - 1;
        const int32_t first = l;
        while (l < h) {
          int32_t mid = l + ((h - l) >> 1);
          if (inE[mid] <= c0) l = mid + 1; else h = mid;
        }
        cur[i] = l;
        head[i] = inE[l];
        val[i] = inV[l];
        base += l - first;
      } else {
        cur[i] = 0;
        head[i] = kSentinel;
        val[i] = 0;
      }
    }

    PSK_PHASE(2);  // per-thread binary searches
    // ---- sequential k-way merge over (c0, c1] -------------------------------------------
    int32_t n = 0;
    bool have = false;
    int32_t pe = 0;
    uint32_t pv = 0;
    for (;;) {
      int32_t e = head[0];
#pragma unroll
      for (int i = 1; i < K; ++i) e = imin(e, head[i]);
      if (e > c1) break;
      const uint32_t v = ev(val);  // value of the span that ends at e
      if (have && !payload_equal(pv, v, typed_float)) {
        outE[base + n] = pe;
        outV[base + n] = pv;
        ++n;
      }
      pe = e;
      pv = v;
      have = true;
#pragma unroll
      for (int i = 0; i < K; ++i) {
        if (head[i] == e) {
          ++cur[i];
          head[i] = inE[cur[i]];
          val[i] = inV[cur[i]];
        }
      }
    }
    if (have) {
      // the thread's last span: kept iff it is the last span of the array or the NEXT
      // span (owned by a later thread / tile) evaluates to a different value
      bool keep = (pe == span);
      if (!keep) keep = !payload_equal(pv, ev(val), typed_float);
      if (keep) {
        outE[base + n] = pe;
        outV[base + n] = pv;
        ++n;
      }
    }

    // ---- compaction: block scan, cross-tile look-back, coalesced write -------------------
    PSK_PHASE(3);  // thread 0's own merge loop
    int32_t tile_total;
    const int32_t my_ofs = block_exclusive_scan<T>(n, s_warp, &tile_total);
    PSK_PHASE(4);  // waiting for the slowest thread + scan
    if (tid < 32) {
      const int32_t excl = lookback_exclusive_warp(p->tile_state, tile, tile_total);
      if (tid == 0) {
        s_gofs = excl;
        if (tile == n_tiles - 1) p->ctrl[CTRL_COUNT] = excl + tile_total;
      }
    }
    // inE/inV are dead (every thread passed the scan's barriers): reuse them as the
    // compacted staging area so that the global write is fully coalesced
    for (int32_t j = 0; j < n; ++j) {
      inE[my_ofs + j] = outE[base + j];
      inV[my_ofs + j] = outV[base + j];
    }
    __syncthreads();
    PSK_PHASE(5);  // look-back + smem compaction
    const int32_t gofs = s_gofs;
    if (gofs + tile_total <= p->out_capacity) {
      for (int32_t j = tid; j < tile_total; j += T) {
        p->out_ends[gofs + j] = inE[j];
        p->out_vals[gofs + j] = inV[j];
      }
    } else if (tid == 0) {
      p->ctrl[CTRL_ERROR] = 2;
    }
    __syncthreads();
    PSK_PHASE(6);  // coalesced global write
  }
  PSK_PHASE_FLUSH;
}

template <int K>
__global__ void __launch_bounds__(kTileThreads) k_merge_eval_interp(DPlan* p) {
  PSK_DYN_SMEM(smem_raw);
  __shared__ psk_node s_prog[PSK_MAX_PROGRAM];
  for (int i = threadIdx.x; i < p->nprog; i += blockDim.x) s_prog[i] = p->prog[i];
  __syncthreads();
  InterpEval ev{s_prog, p->nprog};
  merge_eval_tile_loop<K>(p, smem_raw, ev);
}

constexpr unsigned kMergeEvalSmemBytes = (2 * (kTileCap + PSK_MAX_SOURCES) + 2 * kTileCap) * 4;

}  // namespace psk
