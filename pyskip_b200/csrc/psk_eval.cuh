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
//   k_sample_keys  mapped end of every S-th run of every source (the view chain is evaluated
//                  once per sample).
//   k_sample_rank  ranks every sample among all samples: a sorted splitter list balanced by
//                  RUN COUNT, not by coordinate (the reference's TODO at eval.hpp:450-451).
//   k_tile_bounds  every m-th splitter closes a tile; per (tile, source) a search over the
//                  source's sample keys, then inside one sample interval.  A tile holds
//                  <= (m + 2k) * S runs, so it always fits in shared memory.
//   k_partition_small  the four steps above as one block for evaluations with few samples.
//   k_merge_eval   persistent CTAs take tiles in order (atomic ticket).  The tile's raw runs
//                  arrive from registers (loaded one tile earlier, in four bursts) into flat
//                  shared-memory slots; with views the ends are mapped and zero-length mapped
//                  runs squeezed out in place; the tile's coordinate range is split over the
//                  threads (coarse-to-fine cursor search) and every thread merges its slice
//                  sequentially with the heads of all sources in registers, one distinct end
//                  per iteration.  A span is kept iff its value differs from the NEXT span's
//                  value (the later value wins, as in eval.hpp:189-202), which is a
//                  thread-local decision, so compaction is a block scan plus a decoupled
//                  look-back across tiles (deferred by one tile) -- one pass over HBM.
#pragma once

#include "psk_ops.h"
#include "psk_plan.h"

namespace psk {

#ifndef PSK_TILE_THREADS
#define PSK_TILE_THREADS 512
#endif
#ifndef PSK_TILE_CAP
#define PSK_TILE_CAP 8192
#endif
constexpr int kTileThreads = PSK_TILE_THREADS;
constexpr int kTileCap = PSK_TILE_CAP;  // runs of all sources staged per tile (excl. look-ahead)
#ifndef PSK_COARSE_SEARCH
#define PSK_COARSE_SEARCH 1
#endif
constexpr int kSentinel = 0x7fffffff;

// dynamic shared memory of k_merge_eval: in[CAP + 32] | out[CAP] | pending[CAP] (8-byte pairs)
constexpr unsigned kMergeEvalSmemBytes = ((kTileCap + PSK_MAX_SOURCES) + 2 * kTileCap) * 8;
// CTAs that stay resident per SM (227 KB of shared memory, 2048 threads); also the minimum
// the compiler must leave room for in the register file
constexpr int kCtasBySmem = (227 * 1024) / (int)(kMergeEvalSmemBytes + 4096);
constexpr int kCtasByThreads = 2048 / kTileThreads;
constexpr int kCtasPerSm = kCtasBySmem < kCtasByThreads ? (kCtasBySmem < 1 ? 1 : kCtasBySmem) : kCtasByThreads;

// Development aid (-DPSK_PHASE_TIMING, tools/phase_profile.py): thread 0 of every CTA adds the
// clock64() cycles it spends in each phase of the tile loop to p->phase[].  Not compiled into
// the product library.
#ifdef PSK_PHASE_TIMING
#define PSK_PHASE_DECL long long ph_t = clock64(); long long ph_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define PSK_PHASE(i)                                                              \
  do {                                                                            \
    if (threadIdx.x == 0) {                                                       \
      long long now_ = clock64();                                                 \
      ph_acc[i] += now_ - ph_t;                                                   \
      ph_t = now_;                                                                \
      if (p->trace) {                                                             \
        unsigned long long gt_;                                                   \
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));                   \
        p->trace[(long long)tile * 8 + (i)] = gt_;                                \
        if ((i) == 0) p->trace[(long long)tile * 8 + 7] = blockIdx.x;             \
      }                                                                           \
    }                                                                             \
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

// Warp-cooperative form of the two searches above: 32 probes per round, so a source of 2^21
// runs takes 5 rounds of (load + view chain) instead of 21 dependent ones.
// strict = true: first index whose mapped end is > key; false: >= key.
PSK_D int32_t warp_bound_mapped(const DSource& s, int32_t lo, int32_t hi, int32_t key, bool strict) {
  const int lane = threadIdx.x & 31;
  while (hi - lo > 32) {
    // probes at strictly increasing indices inside [lo, hi)
    const int32_t idx = lo + (int32_t)(((long long)(hi - lo) * (lane + 1)) / 33);
    const int32_t me = mapped_end(s, idx);
    const bool before = strict ? (me <= key) : (me < key);
    const int nb = __popc(__ballot_sync(0xffffffffu, before));  // a prefix of the lanes
    const int32_t last_before = __shfl_sync(0xffffffffu, idx, nb > 0 ? nb - 1 : 0);
    const int32_t first_after = __shfl_sync(0xffffffffu, idx, nb < 32 ? nb : 31);
    if (nb < 32) hi = first_after;
    if (nb > 0) lo = last_before + 1;
  }
  const int32_t idx = lo + lane;
  bool before = false;
  if (idx < hi) {
    const int32_t me = mapped_end(s, idx);
    before = strict ? (me <= key) : (me < key);
  }
  return lo + __popc(__ballot_sync(0xffffffffu, before));
}

// ---------------------------------------------------------------------------------------
// k_prepare: one warp per source.  [a, e) = runs whose mapped end is in (0, span], cut
// after the first run that reaches span (everything later is a zero-length mapped run).
PSK_D void prepare_body(DPlan* p, int32_t* s_ns /*[PSK_MAX_SOURCES] shared*/) {
  const int i = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = p->k;
  if (i < k) {
    DSource& s = p->src[i];
    int32_t a = 0, e = s.nruns;
    if (s.nviews) {
      a = warp_bound_mapped(s, 0, s.nruns, 0, true);
      e = warp_bound_mapped(s, a, s.nruns, p->span, false) + 1;
      if (e > s.nruns) e = s.nruns;
    }
    if (lane == 0) {
      s.a = a;
      s.e = e;
      const int32_t n = e - a;
      s.ns = (n + p->S - 1) / p->S;
      s_ns[i] = s.ns;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
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
__global__ void k_prepare(DPlan* p) {
  __shared__ int32_t s_ns[PSK_MAX_SOURCES];
  prepare_body(p, s_ns);
}

PSK_D int32_t sample_key(const DSource& s, int32_t j, int32_t S) {
  int32_t n = s.e - s.a;
  int32_t off = (j + 1) * S;
  if (off > n) off = n;
  return mapped_end(s, s.a + off - 1);
}

// which source owns sample sid: last i with samp_off[i] <= sid (sources without samples
// share their successor's offset and are skipped)
PSK_D int sample_owner(const int32_t* s_off, int k, int32_t sid) {
  int l = 0, h = k - 1;
  while (l < h) {
    const int mid = (l + h + 1) >> 1;
    if (s_off[mid] <= sid) l = mid; else h = mid - 1;
  }
  return l;
}

// ---------------------------------------------------------------------------------------
// k_sample_keys: samp_keys[sid] = mapped end of sample sid.  The view chain of a source is
// evaluated once per sample here, never inside the ranking searches.
PSK_D void sample_key_body(DPlan* p, int32_t sid, const int32_t* s_off, int k) {
  const int i = sample_owner(s_off, k, sid);
  p->samp_keys[sid] = sample_key(p->src[i], sid - s_off[i], p->S);
}
__global__ void k_sample_keys(DPlan* p) {
  __shared__ int32_t s_off[PSK_MAX_SOURCES];
  const int k = p->k;
  if ((int)threadIdx.x < k) s_off[threadIdx.x] = p->src[threadIdx.x].samp_off;
  __syncthreads();
  const int32_t sid = blockIdx.x * blockDim.x + threadIdx.x;
  if (sid < p->total_samples) sample_key_body(p, sid, s_off, k);
}

// ---------------------------------------------------------------------------------------
// k_sample_rank: merged[rank] = key, where rank orders all samples by (key, source, index):
// own index + one binary search over the sample keys of every other source.
PSK_D void sample_rank_body(DPlan* p, int32_t sid, const int32_t* s_off, const int32_t* s_ns, int k) {
  const int32_t* const keys = p->samp_keys;
  const int i = sample_owner(s_off, k, sid);
  const int32_t key = keys[sid];
  int32_t rank = sid - s_off[i];
  // the searches over the other sources advance in lock step, eight at a time: their loads
  // are independent, so the (L2) latencies overlap instead of adding up
  constexpr int W = 8;
  for (int q0 = 0; q0 < k; q0 += W) {
    int32_t pos[W], len[W];
    int32_t longest = 0;
#pragma unroll
    for (int u = 0; u < W; ++u) {
      const int q = q0 + u;
      pos[u] = 0;
      len[u] = (q < k && q != i) ? s_ns[q] : 0;
      longest = imax(longest, len[u]);
    }
    for (; longest > 0; longest >>= 1) {
#pragma unroll
      for (int u = 0; u < W; ++u) {
        const int q = q0 + u;
        if (len[u] > 0) {
          const int32_t half = len[u] >> 1;
          const int32_t v = keys[s_off[q < k ? q : 0] + pos[u] + half];
          const bool before = (q < i) ? (v <= key) : (v < key);
          pos[u] += before ? half + 1 : 0;
          len[u] = before ? len[u] - half - 1 : half;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < W; ++u) rank += pos[u];
  }
  p->merged[rank] = key;
}
// Block-cooperative form for large evaluations.  The 256 samples of a block are (nearly
// always) consecutive samples of one source, so their ranks in another source q fall into a
// short window of q's keys: two threads locate the window in global memory, the block loads
// it into shared memory (coalesced) and every thread searches there.  The per-thread global
// searches of sample_rank_body cost one 32-byte sector per probe and lane once the lanes'
// paths diverge (~120 MB of L2 traffic at config 2); this form reads each key about once.
constexpr int kRankBlock = 256;
constexpr int kRankWindow = 1024;
__global__ void k_sample_rank(DPlan* p) {
  __shared__ int32_t s_off[PSK_MAX_SOURCES];
  __shared__ int32_t s_ns[PSK_MAX_SOURCES];
  __shared__ int32_t s_wlo[PSK_MAX_SOURCES];
  __shared__ int32_t s_whi[PSK_MAX_SOURCES];
  __shared__ int32_t s_win[kRankWindow];
  __shared__ int32_t s_min[kRankBlock / 32], s_max[kRankBlock / 32];
  const int k = p->k, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < k) {
    s_off[tid] = p->src[tid].samp_off;
    s_ns[tid] = p->src[tid].ns;
  }
  __syncthreads();
  const int32_t* const keys = p->samp_keys;
  const int32_t total = p->total_samples;
  const int32_t sid = blockIdx.x * kRankBlock + tid;
  const bool valid = sid < total;
  const int32_t key = valid ? keys[sid] : 0;
  const int i = valid ? sample_owner(s_off, k, sid) : 0;
  // smallest / largest key of the block
  int32_t kmin = valid ? key : kSentinel, kmax = valid ? key : -1;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    kmin = imin(kmin, __shfl_xor_sync(0xffffffffu, kmin, d));
    kmax = imax(kmax, __shfl_xor_sync(0xffffffffu, kmax, d));
  }
  if (lane == 0) {
    s_min[warp] = kmin;
    s_max[warp] = kmax;
  }
  __syncthreads();
#pragma unroll
  for (int w = 0; w < kRankBlock / 32; ++w) {
    kmin = imin(kmin, s_min[w]);
    kmax = imax(kmax, s_max[w]);
  }
  // window of every source: [#keys < kmin, #keys <= kmax)
  if (tid < 2 * k) {
    const int q = tid >> 1;
    const bool upper = (tid & 1) != 0;
    const int32_t* const kq = keys + s_off[q];
    int32_t lo = 0, hi = s_ns[q];
    while (lo < hi) {
      const int32_t mid = lo + ((hi - lo) >> 1);
      const int32_t v = kq[mid];
      if (upper ? (v <= kmax) : (v < kmin)) lo = mid + 1; else hi = mid;
    }
    if (upper) s_whi[q] = lo; else s_wlo[q] = lo;
  }
  __syncthreads();
  int32_t rank = valid ? sid - s_off[i] : 0;
  for (int q = 0; q < k; ++q) {
    const int32_t wlo = s_wlo[q], w = s_whi[q] - wlo;  // uniform over the block
    const int32_t* const kq = keys + s_off[q] + wlo;
    const bool mine = valid && q != i;
    int32_t lo = 0, hi = w;
    if (w <= kRankWindow) {
      for (int32_t t = tid; t < w; t += kRankBlock) s_win[t] = kq[t];
      __syncthreads();
      if (mine) {
        while (lo < hi) {
          const int32_t mid = lo + ((hi - lo) >> 1);
          const int32_t v = s_win[mid];
          if ((q < i) ? (v <= key) : (v < key)) lo = mid + 1; else hi = mid;
        }
      }
      __syncthreads();
    } else if (mine) {  // skewed sources: search the window in global memory
      while (lo < hi) {
        const int32_t mid = lo + ((hi - lo) >> 1);
        const int32_t v = kq[mid];
        if ((q < i) ? (v <= key) : (v < key)) lo = mid + 1; else hi = mid;
      }
    }
    if (mine) rank += wlo + lo;
  }
  if (valid) p->merged[rank] = key;
}

// ---------------------------------------------------------------------------------------
// k_tile_bounds: bounds[t][i] = first run of source i whose mapped end is > hi(t-1).  The
// source's own sample keys narrow the search to one sample interval (<= S runs) before any
// run end is loaded or mapped.
PSK_D void tile_bound_body(DPlan* p, int32_t id) {
  const int k = p->k, nt = p->n_tiles, m = p->m;
  int32_t t = id / k, i = id - t * k;
  const DSource& s = p->src[i];
  int32_t b;
  if (t == 0) {
    b = s.a;
  } else if (t == nt) {
    b = s.e;
  } else {
    const int32_t key = p->merged[t * m - 1];
    const int32_t* const kq = p->samp_keys + s.samp_off;
    int32_t lo = 0, hi = s.ns;  // first sample whose key is > key
    while (lo < hi) {
      const int32_t mid = lo + ((hi - lo) >> 1);
      if (kq[mid] <= key) lo = mid + 1; else hi = mid;
    }
    if (lo == s.ns) {
      b = s.e;  // the last sample is the source's last run
    } else {
      // sample j is run a + min((j + 1) S, n) - 1: every run up to sample lo - 1 is <= key,
      // the run of sample lo is > key
      const int32_t n = s.e - s.a;
      const int32_t r = s.a + imin((lo + 1) * p->S, n) - 1;
      b = ub_mapped(s, s.a + lo * p->S, r, key);
    }
  }
  p->bounds[t * k + i] = b;
  if (i == 0 && t > 0) p->tile_hi[t - 1] = (t == nt) ? p->span : p->merged[t * m - 1];
}
__global__ void k_tile_bounds(DPlan* p) {
  const int32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id < (p->n_tiles + 1) * p->k) tile_bound_body(p, id);
}

// ---------------------------------------------------------------------------------------
// k_partition_small: the four kernels above as ONE block for evaluations with few samples
// (three launches and their dependency gaps cost more than the work).
__global__ void k_partition_small(DPlan* p, int do_prepare) {
  __shared__ int32_t s_off[PSK_MAX_SOURCES];
  __shared__ int32_t s_ns[PSK_MAX_SOURCES];
  const int k = p->k;
  if (do_prepare) {
    prepare_body(p, s_ns);
    __syncthreads();
  }
  if ((int)threadIdx.x < k) {
    s_off[threadIdx.x] = p->src[threadIdx.x].samp_off;
    s_ns[threadIdx.x] = p->src[threadIdx.x].ns;
  }
  __syncthreads();
  const int32_t total = p->total_samples;
  for (int32_t sid = threadIdx.x; sid < total; sid += blockDim.x) sample_key_body(p, sid, s_off, k);
  __syncthreads();
  for (int32_t sid = threadIdx.x; sid < total; sid += blockDim.x) sample_rank_body(p, sid, s_off, s_ns, k);
  __syncthreads();
  const int32_t nb = (p->n_tiles + 1) * k;
  for (int32_t id = threadIdx.x; id < nb; id += blockDim.x) tile_bound_body(p, id);
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

PSK_D void publish_aggregate(unsigned long long* state, int32_t tile, int32_t count) {
  // tile 0 has no predecessor: its aggregate is already its inclusive prefix
  const unsigned long long st = tile == 0 ? ST_PREFIX : ST_AGGREGATE;
  st_volatile_u64(&state[tile], (st << 32) | (uint32_t)count);
}

// Decoupled look-back, executed by the whole CTA: thread r inspects tile (j - r), so one
// round covers T predecessors (with the write deferred by a tile, everything younger than
// about two tile periods has only published an AGGREGATE).  The nearest tile that already
// published an inclusive PREFIX ends the walk.  Tiles are handed out by an atomic ticket, so
// every predecessor is resident or finished and the spin cannot deadlock.
template <int T>
PSK_D int32_t lookback_exclusive_cta(unsigned long long* state, int32_t tile, int32_t count,
                                     unsigned long long early, int32_t* s_sum /*[T/32]*/,
                                     int32_t* s_has /*[T/32]*/) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tile == 0) return 0;
  int32_t excl = 0;
  for (int32_t j = tile - 1;; j -= T) {
    const int32_t idx = j - tid;
    unsigned long long w = (unsigned long long)ST_PREFIX << 32;  // virtual prefix 0 before tile 0
    if (idx >= 0) {
      // first round: the word this thread loaded before the merge (lookback_peek), if it was
      // already published then
      w = (j == tile - 1) ? early : 0ull;
      while ((w >> 32) == ST_INVALID) w = ld_volatile_u64(&state[idx]);
    }
    const unsigned pmask = __ballot_sync(0xffffffffu, (w >> 32) == ST_PREFIX);
    const int first = pmask ? (__ffs((int)pmask) - 1) : 32;
    int32_t c = (lane <= first) ? (int32_t)(uint32_t)w : 0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    __syncthreads();  // s_sum / s_has free (previous round or other users of s_warp)
    if (lane == 0) {
      s_sum[warp] = c;
      s_has[warp] = pmask != 0;
    }
    __syncthreads();
    // every warp folds the warp results in tile order: lane q holds warp q's
    const int32_t ws = (lane < T / 32) ? s_sum[lane] : 0;
    const unsigned hmask = __ballot_sync(0xffffffffu, lane < T / 32 && s_has[lane] != 0);
    const int fw = hmask ? (__ffs((int)hmask) - 1) : 32;
    int32_t part = (lane <= fw) ? ws : 0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    excl += part;
    if (hmask) break;
  }
  __syncthreads();
  if (tid == 0)
    st_volatile_u64(&state[tile], ((unsigned long long)ST_PREFIX << 32) | (uint32_t)(excl + count));
  return excl;
}

// The look-back word a thread will need for `tile`, loaded early (before the merge of the
// next tile) so that its latency is off the critical path; 0 (INVALID) if there is none.
template <int T>
PSK_D unsigned long long lookback_peek(const unsigned long long* state, int32_t tile) {
  const int32_t idx = tile - 1 - (int32_t)threadIdx.x;
  return (tile > 0 && idx >= 0) ? ld_volatile_u64(&state[idx]) : 0ull;
}

// Resolve the global offset of a finished tile whose compacted spans wait at sm[region..] and
// write them out (coalesced).  Called by the whole CTA.
template <int T>
PSK_D void flush_pending(DPlan* p, const SmemPairs& sm, int region, int32_t tile, int32_t total,
                         unsigned long long early, int32_t n_tiles, int32_t* s_sum, int32_t* s_has) {
  const int tid = threadIdx.x;
  const int32_t gofs = lookback_exclusive_cta<T>(p->tile_state, tile, total, early, s_sum, s_has);
  if (tid == 0 && tile == n_tiles - 1) {
    p->ctrl[CTRL_COUNT] = gofs + total;
    p->out_ends[p->out_capacity] = gofs + total;  // spare word: the count stays on the device too
  }
  if (gofs + total <= p->out_capacity) {
    for (int32_t j = tid; j < total; j += T) {
      const int2 r = sm.ld(region + j);
      p->out_ends[gofs + j] = r.x;
      p->out_vals[gofs + j] = (uint32_t)r.y;
    }
  } else if (tid == 0) {
    p->ctrl[CTRL_ERROR] = 2;
  }
  __syncthreads();  // the region may be overwritten now
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
//
// Shared memory per CTA: in[] = the tile's runs as (end, value) pairs, one slice per source
// plus one look-ahead run each; out[] = the spans each thread keeps; pend[] = the compacted
// spans of the previous tile, parked until their global offset is known (one tile later).
struct TileDesc {          // bounds of one tile, fetched one tile ahead
  int32_t tile;
  int32_t lo, hi;
  int32_t b0[PSK_MAX_SOURCES];
  int32_t b1[PSK_MAX_SOURCES];
};

struct SrcDesc {           // the per-source fields the tile loop needs, hoisted into smem once
  const int32_t* ends;
  const uint32_t* vals;
  int32_t nviews;
  int32_t e;
};

// Bounds of a tile held in the registers of one warp while the loads are in flight (lane i
// holds source i's bounds), written to shared memory only when the tile loop is ready for it.
struct DescRegs {
  int32_t tile, lo, hi, b0, b1;
};

// one warp: take the next ticket and start fetching that tile's bounds (no barrier, and no
// use of the loaded values: they stay in flight until store_tile_desc)
PSK_D DescRegs fetch_tile_desc(DPlan* p, int k, int32_t n_tiles) {
  const int lane = threadIdx.x & 31;
  DescRegs r;
  r.tile = 0;
  r.lo = r.hi = r.b0 = r.b1 = 0;
  if (lane == 0) r.tile = atomicAdd(&p->ctrl[CTRL_TICKET], 1);
  r.tile = __shfl_sync(0xffffffffu, r.tile, 0);
  if (r.tile < n_tiles) {
    if (lane < k) {
      r.b0 = p->bounds[r.tile * k + lane];
      r.b1 = p->bounds[(r.tile + 1) * k + lane];
    }
    if (lane == 0) {
      r.lo = r.tile == 0 ? 0 : p->tile_hi[r.tile - 1];
      r.hi = p->tile_hi[r.tile];
    }
  }
  return r;
}

PSK_D void store_tile_desc(const DescRegs& r, TileDesc* d, int k) {
  const int lane = threadIdx.x & 31;
  if (lane == 0) {
    d->tile = r.tile;
    d->lo = r.lo;
    d->hi = r.hi;
  }
  if (lane < k) {
    d->b0[lane] = r.b0;
    d->b1[lane] = r.b1;
  }
}

// Where the RAW runs of a tile go in the shared tile buffer: the sources occupy the flat
// slots [0, flat_total) back to back (slice + one look-ahead run each).  With views the
// staging pass then maps the ends and squeezes out zero-length mapped runs in place.
struct SlotSource {  // ebase[j] / vbase[j] = the run that belongs in flat slot j
  const int32_t* ebase;
  const uint32_t* vbase;
};
struct TileLayout {
  int32_t off[PSK_MAX_SOURCES + 1];
  int32_t len[PSK_MAX_SOURCES];  // staged runs (excl. look-ahead)
  int32_t lim[PSK_MAX_SOURCES];  // slots of the source below lim hold a run, the rest a sentinel
  int32_t flat_total;
  alignas(16) SlotSource ss[PSK_MAX_SOURCES];
};

// one warp: layout of tile d
PSK_D void compute_layout(const TileDesc& d, const SrcDesc* src, int k, TileLayout* L) {
  const int lane = threadIdx.x & 31;
  int32_t len = 0;
  if (lane < k) len = d.b1[lane] - d.b0[lane] + 1;  // + look-ahead
  int32_t inc = len;
#pragma unroll
  for (int dd = 1; dd < 32; dd <<= 1) {
    int32_t o = __shfl_up_sync(0xffffffffu, inc, dd);
    if (lane >= dd) inc += o;
  }
  if (lane < k) {
    const int32_t off = inc - len;
    L->off[lane] = off;
    L->len[lane] = len ? len - 1 : 0;
    // the look-ahead slot holds a run only if the source has one beyond the tile
    L->lim[lane] = off + imin(len, src[lane].e - d.b0[lane]);
    L->ss[lane].ebase = src[lane].ends + (d.b0[lane] - off);
    L->ss[lane].vbase = src[lane].vals + (d.b0[lane] - off);
  }
  if (lane == 31) L->flat_total = inc;
}

// which source owns flat slot j: last i with off[i] <= j (every source has >= 1 slot)
template <int K>
PSK_D int slot_owner(const int32_t* off, int k, int32_t j) {
  int i = 0;
  if (K <= 8) {
#pragma unroll
    for (int q = 1; q < K; ++q) i += (q < k && off[q] <= j) ? 1 : 0;
  } else {
    int l = 0, h = k - 1;
    while (l < h) {
      const int mid = (l + h + 1) >> 1;
      if (off[mid] <= j) l = mid; else h = mid - 1;
    }
    i = l;
  }
  return i;
}

// all threads: issue the global loads of the tile's raw runs into registers (flat
// slot j = tid + u*T).  Nothing waits on them here; they are stored to shared memory one
// tile later, so the HBM latency overlaps the merge of the tile in front.
template <int K, int T, int NB, int U0 = 0, int U1 = NB>
PSK_D void issue_tile_loads(const TileDesc& d, const TileLayout& L, const SrcDesc* src, int k, int32_t limit,
                            int2 (&pr)[NB]) {
  const int32_t used = L.flat_total <= limit ? L.flat_total : 0;
  // branch-free, four slots at a time: owner -> (limit, bases) from shared memory -> two
  // predicated global loads; nothing in here waits for a global load
  int32_t off_[K <= 8 ? K : 1];
  if (K <= 8) {
#pragma unroll
    for (int q = 1; q < K; ++q) off_[q] = (q < k) ? L.off[q] : kSentinel;
  }
  constexpr int G = 4;
#pragma unroll
  for (int u0 = U0; u0 < U1; u0 += G) {
    int i_[G];
    int32_t lim_[G];
    SlotSource ss_[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int32_t j = (int32_t)threadIdx.x + (u0 + g) * T;
      int i = 0;
      if (K <= 8) {
#pragma unroll
        for (int q = 1; q < K; ++q) i += (off_[q] <= j) ? 1 : 0;
      } else {
        i = slot_owner<K>(L.off, k, j);
      }
      i_[g] = i;
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      lim_[g] = L.lim[i_[g]];
      ss_[g] = L.ss[i_[g]];
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      if (u0 + g < U1) {
        const int32_t j = (int32_t)threadIdx.x + (u0 + g) * T;
        // no run in the slot (beyond the tile / the source): sentinel look-ahead
        const bool take = j < used && j < lim_[g];
        pr[u0 + g].x = ld_nc_if(ss_[g].ebase + j, take, kSentinel);
        pr[u0 + g].y = ld_nc_if(reinterpret_cast<const int32_t*>(ss_[g].vbase) + j, take, 0);
      }
    }
  }
}

// TYPED_FLOAT: compaction compares float values with operator== (PSK_EQ_TYPED on F32)
// instead of the 32-bit payloads.
// VIEWS: 1 = some source has a view, 0 = none (the staging pass is compiled out), -1 = ask
// the plan at run time (the interpreter kernels).
template <int K, bool TYPED_FLOAT, int VIEWS, typename Eval>
PSK_D void merge_eval_tile_loop(DPlan* p, const Eval& ev) {
  constexpr int T = kTileThreads;
  constexpr int CAP = kTileCap;
  constexpr int IN_N = CAP + PSK_MAX_SOURCES;  // room for one look-ahead run per source

  // one dynamic shared array, indexed directly (no derived pointers: keeps every access a
  // plain LDS/STS off the static shared base): [0, IN_N) = in, [IN_N, IN_N + CAP) = out
  PSK_DYN_SMEM(smem_raw);
  const SmemPairs sm(smem_raw);

  __shared__ TileDesc s_desc[3];   // tile being merged, tile being loaded, tile being fetched
  __shared__ TileLayout s_layout[2];
  __shared__ SrcDesc s_src[PSK_MAX_SOURCES];
  __shared__ int32_t s_warp[T / 32 + 2];
  __shared__ int32_t s_lbhas[T / 32];
#if PSK_COARSE_SEARCH
  __shared__ unsigned short s_coarse[PSK_MAX_SOURCES * (T / 8 + 1)];  // cursors of every 8th thread
#endif
  __shared__ int32_t s_newoff[PSK_MAX_SOURCES];  // with views: slot layout after the squeeze
  __shared__ int32_t s_newend[PSK_MAX_SOURCES];

  constexpr int PEND = IN_N + CAP;        // third region: compacted spans awaiting their offset
  int32_t pend_tile = -1, pend_total = 0;
  constexpr int NB = (IN_N + T - 1) / T;  // flat slots per thread
  int2 pr[NB];                            // the NEXT tile's runs, in flight

  const int tid = threadIdx.x;
  const int k = p->k;
  constexpr bool typed_float = TYPED_FLOAT;
  const bool any_views = VIEWS < 0 ? (p->any_views != 0) : (VIEWS != 0);
  const int32_t span = p->span;
  const int32_t n_tiles = p->n_tiles;

  if (tid < k) {
    const DSource& s = p->src[tid];
    s_src[tid].ends = s.ends;
    s_src[tid].vals = s.vals;
    s_src[tid].nviews = s.nviews;
    s_src[tid].e = s.e;
  }
  // Persistent CTAs that start together stay in lock step; an optional start delay per CTA
  // index (PSKB200_STAGGER_NS, default 0) de-phases them.
  if (p->stagger_ns > 0 && blockIdx.x > 0) sleep_ns((unsigned)(p->stagger_ns * (int)blockIdx.x));
  // ---- pipeline prologue: descriptors of the first two tiles, data of the first ----------
  if (tid < 32) {
    store_tile_desc(fetch_tile_desc(p, k, n_tiles), &s_desc[0], k);
    store_tile_desc(fetch_tile_desc(p, k, n_tiles), &s_desc[1], k);
  }
  __syncthreads();
  if (tid < 32 && s_desc[0].tile < n_tiles) compute_layout(s_desc[0], s_src, k, &s_layout[0]);
  __syncthreads();
  if (s_desc[0].tile < n_tiles) issue_tile_loads<K, T, NB>(s_desc[0], s_layout[0], s_src, k, IN_N, pr);

  PSK_PHASE_DECL;
  for (int it = 0;; ++it) {
    const TileDesc& d = s_desc[it % 3];
    const TileDesc& dn = s_desc[(it + 1) % 3];
    TileLayout& lay = s_layout[it & 1];
    int32_t* const s_off = lay.off;
    int32_t* const s_len = lay.len;
    const int32_t tile = d.tile;
    if (tile >= n_tiles) break;
    const int32_t lo = d.lo, hi = d.hi;
    const bool next_valid = dn.tile < n_tiles;

    // ---- the tile's runs arrive from registers; meanwhile warp 0 lays out the next tile and
    //      warp 1 starts fetching the bounds of the tile after that -----------------------
    int32_t used = lay.flat_total;
    bool fail = used > IN_N;
    if (!fail) {
#pragma unroll
      for (int u = 0; u < NB; ++u) {
        const int32_t j = tid + u * T;
        if (j < used) {
          int32_t x = pr[u].x;
          if (any_views) {
            // raw end -> coordinate of the expression through the source's view chain
            const int i = slot_owner<K>(s_off, k, j);
            const int nv = s_src[i].nviews;
            if (nv) x = (d.b0[i] + (j - s_off[i]) < s_src[i].e) ? map_chain(p->src[i].v, nv, x) : kSentinel;
          }
          sm.st(j, x, pr[u].y);
        }
      }
    }
    DescRegs fetched;
    fetched.tile = fetched.lo = fetched.hi = fetched.b0 = fetched.b1 = 0;
    if (tid < 32) {
      if (next_valid) compute_layout(dn, s_src, k, &s_layout[(it + 1) & 1]);
    } else if (tid < 64) {
      fetched = fetch_tile_desc(p, k, n_tiles);
    }
    PSK_PHASE(0);
    // ---- with views: squeeze out the zero-length mapped runs (mapped end not beyond its
    //      predecessor's) in place.  Every thread takes nb consecutive slots (nb odd: the
    //      strided shared accesses hit distinct banks); one block scan gives the new slots. --
    if (any_views && !fail) {
      constexpr int NBO = NB | 1;
      __syncthreads();
      const int32_t nb = ((used + T - 1) / T) | 1;
      const int32_t j0 = tid * nb;
      int2 run_[NBO];
      int32_t prev = (j0 > 0 && j0 <= used) ? sm.ldx(j0 - 1) : 0;
#pragma unroll
      for (int u = 0; u < NBO; ++u)
        run_[u] = (u < nb && j0 + u < used) ? sm.ld(j0 + u) : make_int2(-1, 0);
      int i = slot_owner<K>(s_off, k, imin(j0, used - 1));
      int32_t nxt = (i + 1 < k) ? s_off[i + 1] : used;  // first slot of the next source
      int32_t kept = 0;
#pragma unroll
      for (int u = 0; u < NBO; ++u) {
        const int32_t j = j0 + u;
        if (u < nb && j < used) {
          if (j == nxt) {
            ++i;
            nxt = (i + 1 < k) ? s_off[i + 1] : used;
          }
          const int32_t me = run_[u].x;
          // a source's first run is measured against the tile's lower bound; its look-ahead
          // run (last slot) always stays
          const int32_t before = (j == s_off[i]) ? lo : prev;
          const bool keep = s_src[i].nviews == 0 || j == nxt - 1 || me > before;
          prev = me;
          if (keep) ++kept; else run_[u].x = -1;
        }
      }
      int32_t total;
      int32_t pos = block_exclusive_scan<T>(kept, s_warp, &total);
      // (the scan's barriers also order every read above before the writes below)
      i = slot_owner<K>(s_off, k, imin(j0, used - 1));
      nxt = (i + 1 < k) ? s_off[i + 1] : used;
#pragma unroll
      for (int u = 0; u < NBO; ++u) {
        const int32_t j = j0 + u;
        if (u < nb && j < used) {
          if (j == nxt) {
            ++i;
            nxt = (i + 1 < k) ? s_off[i + 1] : used;
          }
          if (j == s_off[i]) s_newoff[i] = pos;
          if (j == nxt - 1) s_newend[i] = pos;  // the look-ahead run is always kept
          if (run_[u].x >= 0) {
            sm.st(pos, run_[u].x, run_[u].y);
            ++pos;
          }
        }
      }
      __syncthreads();
      if (tid < k) {
        s_off[tid] = s_newoff[tid];
        s_len[tid] = s_newend[tid] - s_newoff[tid];
      }
      used = total;
    }
    __syncthreads();
    PSK_PHASE(1);  // staging
    if (fail && tid == 0) p->ctrl[CTRL_ERROR] = 1;
    // the next tile's loads go out now and land while this tile is merged
    // (in four bursts spread over the tile's phases: one burst of the whole tile overruns
    //  what the SM keeps in flight towards HBM and stalls the issuing warps)
    constexpr int NQ = (NB + 3) / 4;  // slots per burst
    const TileLayout& lay_next = s_layout[(it + 1) & 1];
    if (next_valid) issue_tile_loads<K, T, NB, 0, (NQ < NB ? NQ : NB)>(dn, lay_next, s_src, k, IN_N, pr);
    PSK_PHASE(7);  // issuing the next tile's loads

    // ---- split the tile's coordinate range (lo, hi] over the threads --------------------
    const long long width = (long long)hi - lo;
    const int32_t c0 = lo + (int32_t)((width * tid) / T);
    const int32_t c1 = lo + (int32_t)((width * (tid + 1)) / T);

    // first staged run of every source whose end is > c0: branch-free binary searches that
    // advance in lock step over the sources (the look-ahead run bounds each of them).
    // Coarse to fine: the cursors of every CG-th thread are found first (one search per
    // thread, over the whole slice) and bracket the searches of the threads in between --
    // roughly half the shared-memory probes of T * k full searches.
    int32_t cur[K], head[K], rem[K];
    uint32_t val[K];
    int32_t base = 0;
#if PSK_COARSE_SEARCH
    constexpr int CG = 8, NC = T / CG;
    if (!fail) {
      for (int32_t q = tid; q < NC * k; q += T) {
        const int i = q / NC, g = q - i * NC;
        const int32_t cg = lo + (int32_t)((width * (g * CG)) / T);
        int32_t c = s_off[i], r = s_len[i] + 1;
        while (r > 1) {
          const int32_t half = r >> 1;
          c += (sm.ldx(c + half - 1) <= cg) ? half : 0;
          r -= half;
        }
        s_coarse[i * (NC + 1) + g] = (unsigned short)c;
        if (g == 0) s_coarse[i * (NC + 1) + NC] = (unsigned short)(s_off[i] + s_len[i]);
      }
    }
    if (next_valid) issue_tile_loads<K, T, NB, (NQ < NB ? NQ : NB), (2 * NQ < NB ? 2 * NQ : NB)>(dn, lay_next, s_src, k, IN_N, pr);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < K; ++i) {
      if (i < k && !fail) {
        const int32_t a0 = s_coarse[i * (NC + 1) + tid / CG], a1 = s_coarse[i * (NC + 1) + tid / CG + 1];
        cur[i] = a0;
        rem[i] = a1 - a0 + 1;
      } else {
        cur[i] = 0;
        rem[i] = 1;
      }
    }
#else
    if (next_valid) issue_tile_loads<K, T, NB, (NQ < NB ? NQ : NB), (2 * NQ < NB ? 2 * NQ : NB)>(dn, lay_next, s_src, k, IN_N, pr);
#pragma unroll
    for (int i = 0; i < K; ++i) {
      cur[i] = (i < k) ? s_off[i] : 0;
      rem[i] = (i < k && !fail) ? s_len[i] + 1 : 1;
    }
#endif
    {
      int32_t longest = 1;
#pragma unroll
      for (int i = 0; i < K; ++i) longest = imax(longest, rem[i]);
      for (; longest > 1; longest -= longest >> 1) {
#pragma unroll
        for (int i = 0; i < K; ++i) {
          const int32_t half = rem[i] >> 1;
          if (half > 0) {
            const bool right = sm.ldx(cur[i] + half - 1) <= c0;
            cur[i] += right ? half : 0;
            rem[i] -= half;
          }
        }
      }
    }
    uint32_t cur_at[K];  // cursors (shared-window addresses) of the k head runs
#pragma unroll
    for (int i = 0; i < K; ++i) {
      if (i < k && !fail) {
        const int2 pr = sm.ld(cur[i]);
        head[i] = pr.x;
        val[i] = (uint32_t)pr.y;
        base += cur[i] - s_off[i];
      } else {
        head[i] = kSentinel;
        val[i] = 0;
      }
      cur_at[i] = sm.at(cur[i]);
    }
    const uint32_t out_begin = sm.at(IN_N + base);
    uint32_t out_at = out_begin;  // where this thread's next kept span goes

    if (next_valid) issue_tile_loads<K, T, NB, (2 * NQ < NB ? 2 * NQ : NB), (3 * NQ < NB ? 3 * NQ : NB)>(dn, lay_next, s_src, k, IN_N, pr);
    // the look-back word of the tile that is flushed after this merge: in flight meanwhile
    const unsigned long long lb_early = pend_tile >= 0 ? lookback_peek<T>(p->tile_state, pend_tile) : 0ull;
    PSK_PHASE(2);  // per-thread binary searches
    // ---- sequential k-way merge over (c0, c1] -------------------------------------------
    // One distinct end is consumed per iteration: every source whose head run ends there
    // moves on to its next run (one predicated 64-bit shared load per source).
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
        sm.st_at(out_at, pe, (int32_t)pv);
        out_at += sm.step();
      }
      pe = e;
      pv = v;
      have = true;
#pragma unroll
      for (int i = 0; i < K; ++i) sm.advance_if(e, cur_at[i], head[i], val[i]);
    }
    if (have) {
      // the thread's last span: kept iff it is the last span of the array or the NEXT
      // span (owned by a later thread / tile) evaluates to a different value
      bool keep = (pe == span);
      if (!keep) keep = !payload_equal(pv, ev(val), typed_float);
      if (keep) {
        sm.st_at(out_at, pe, (int32_t)pv);
        out_at += sm.step();
      }
    }
    const int32_t n = (int32_t)((out_at - out_begin) / sm.step());  // spans this thread keeps

    if (next_valid) issue_tile_loads<K, T, NB, (3 * NQ < NB ? 3 * NQ : NB), NB>(dn, lay_next, s_src, k, IN_N, pr);
    PSK_PHASE(3);  // thread 0's own merge loop
    // ---- compaction: block scan; the cross-tile look-back and the global write of a tile
    //      are DEFERRED by one tile: its run count is published right after the scan, its
    //      compacted spans wait in the third region of the tile buffer while the CTA merges
    //      the next tile, and by the time it looks back every predecessor has long published
    //      (no waiting on the slowest tile of the wave) ----------------------
    int32_t tile_total;
    const int32_t my_ofs = block_exclusive_scan<T>(n, s_warp, &tile_total);
    PSK_PHASE(4);  // waiting for the slowest thread + scan
    if (tid == 0) publish_aggregate(p->tile_state, tile, tile_total);
    if (pend_tile >= 0) flush_pending<T>(p, sm, PEND, pend_tile, pend_total, lb_early, n_tiles, s_warp, s_lbhas);
    PSK_PHASE(5);  // look-back + global write of the previous tile
    for (int32_t j = 0; j < n; ++j) {
      const int2 r = sm.ld(IN_N + base + j);
      sm.st(PEND + my_ofs + j, r.x, r.y);
    }
    pend_tile = tile;
    pend_total = tile_total;
    if (tid >= 32 && tid < 64) store_tile_desc(fetched, &s_desc[(it + 2) % 3], k);
    __syncthreads();
    PSK_PHASE(6);  // smem compaction
  }
  if (pend_tile >= 0) flush_pending<T>(p, sm, PEND, pend_tile, pend_total, 0ull, n_tiles, s_warp, s_lbhas);
  PSK_PHASE_FLUSH;
}

template <int K>
__global__ void __launch_bounds__(kTileThreads, kCtasPerSm) k_merge_eval_interp(DPlan* p) {
  __shared__ psk_node s_prog[PSK_MAX_PROGRAM];
  for (int i = threadIdx.x; i < p->nprog; i += blockDim.x) s_prog[i] = p->prog[i];
  __syncthreads();
  InterpEval ev{s_prog, p->nprog};
  if (p->eq != 0) merge_eval_tile_loop<K, true, -1>(p, ev);
  else merge_eval_tile_loop<K, false, -1>(p, ev);
}



}  // namespace psk
