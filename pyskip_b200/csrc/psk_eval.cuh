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
//   k_prepare      (only with views / queued sources) trims every source to the run range
//                  that can reach the output -- the role of lang::make_pool's slice_invert
//                  calls (lang.hpp:708-709) -- and fixes the sample / tile counts on the device.
//   k_partition    ONE kernel: every S-th run of every source is a sample; a block ranks its
//                  256 samples among all samples (window of every other source's sample keys
//                  staged in shared memory), and the threads whose sample closes a tile
//                  (every m-th in merged order) refine, per source, the run index where the
//                  tile ends.  Tiles are cut in the TOTAL order (mapped end, source, index),
//                  so a tile holds <= (m + 2k) * S runs whatever the data looks like
//                  (balanced by RUN COUNT, the reference's TODO at eval.hpp:450-451; long
//                  groups of runs that a gather view collapses onto one coordinate are cut
//                  like any other runs).
//   k_merge_eval   persistent CTAs (several per SM) take tiles in order (atomic ticket).  One
//                  elected warp issues a bulk asynchronous copy (cp.async.bulk) per source:
//                  the tile's (end, value) pairs land in shared memory in the layout the
//                  merge reads, signalled by an mbarrier; the NEXT tile's copies are issued as
//                  soon as the merge of the current one is done.  The tile's coordinate range
//                  is split over the threads; every thread merges its slice sequentially
//                  with the heads of all sources in registers, one distinct end per
//                  iteration.  A span is kept iff its value differs from the NEXT span's
//                  value (the later value wins, as in eval.hpp:189-202), which is a
//                  thread-local decision, so compaction is a block scan plus a decoupled
//                  look-back across tiles, and every warp writes its threads' spans out
//                  coalesced -- one pass over HBM.
#pragma once

#include "psk_ops.h"
#include "psk_plan.h"

namespace psk {

#ifndef PSK_TILE_THREADS
#define PSK_TILE_THREADS 128
#endif
#ifndef PSK_TILE_CAP
#define PSK_TILE_CAP 4096
#endif
constexpr int kTileThreads = PSK_TILE_THREADS;
constexpr int kTileCap = PSK_TILE_CAP;  // in-tile runs of all sources together
constexpr int kSentinel = 0x7fffffff;
// Look-back chain word: (status << 32) | count.
enum { ST_INVALID = 0, ST_AGGREGATE = 1, ST_PREFIX = 2 };
constexpr int kStateStride = 4;  // 64-bit words per look-back entry (one 32-byte sector per tile)
// per source beyond its in-tile runs: two look-ahead runs + up to two pairs of 16-byte alignment
constexpr int kTileSlack = 4 * PSK_MAX_SOURCES;
constexpr int kTileRegion = kTileCap + kTileSlack;  // pairs per shared-memory region

// dynamic shared memory of k_merge_eval: regions A | B (| C) of 8-byte pairs
constexpr int kTileRegions = 2;  // A: the tile's runs; B: the spans the threads keep
#ifndef PSK_PARK_DEPTH
#define PSK_PARK_DEPTH 2
#endif
// Finished tiles whose place in the output is not known yet are parked in a per-CTA ring in
// global memory (it stays in L2) and moved to their place a few tiles later, so a CTA never
// waits for the slowest tile in front of it (see merge_eval_tile_loop)
constexpr int kParkDepth = PSK_PARK_DEPTH;
constexpr unsigned kMergeEvalSmemBytes = (unsigned)kTileRegions * kTileRegion * 8u;
// CTAs that stay resident per SM (227 KB of shared memory, 2048 threads)
constexpr int kCtasBySmem = (227 * 1024) / (int)(kMergeEvalSmemBytes + 6144);
#ifndef PSK_COARSE_GROUP
#define PSK_COARSE_GROUP 0
#endif
constexpr int kCtasByThreads = 2048 / kTileThreads;
constexpr int kCtasPerSm = kCtasBySmem < kCtasByThreads ? (kCtasBySmem < 1 ? 1 : kCtasBySmem) : kCtasByThreads;

// Development aid (-DPSK_PHASE_TIMING, tools/phase_profile.py): thread 0 of every CTA adds the
// clock64() cycles it spends in each phase of the tile loop to p->phase[].  Not compiled into
// the product library.
#ifdef PSK_PARTITION_TIMING  // the same for k_partition (tools/phase_profile.py --partition)
#define PSK_PHASE_TIMING
#define PSK_PPHASE_DECL long long pp_t = clock64()
#define PSK_PPHASE(i)                                                     \
  do {                                                                    \
    if (threadIdx.x == 0) {                                               \
      long long now_ = clock64();                                         \
      atomicAdd(&p->phase[i], (unsigned long long)(now_ - pp_t));         \
      pp_t = now_;                                                        \
    }                                                                     \
  } while (0)
#define PSK_PPHASE_ANY(i, since) atomicAdd(&p->phase[i], (unsigned long long)(clock64() - (since)))
#else
#define PSK_PPHASE_DECL
#define PSK_PPHASE(i)
#define PSK_PPHASE_ANY(i, since)
#endif
#if defined(PSK_PHASE_TIMING) && !defined(PSK_PARTITION_TIMING)
#define PSK_PHASE_DECL long long ph_t = clock64(); long long ph_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define PSK_PHASE(i)                                                              \
  do {                                                                            \
    if (threadIdx.x == 0) {                                                       \
      long long now_ = clock64();                                                 \
      ph_acc[i] += now_ - ph_t;                                                   \
      ph_t = now_;                                                                \
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

// Warp-cooperative searches: 32 probes per round, so a source of 2^21 runs takes 5 rounds of
// (load + view chain) instead of 21 dependent ones.
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

// Samples of a source: the runs of [a, e) whose index g has (g + 1) % S == 0 -- aligned to the
// run array, not to a, so that they can be read from the store's compact index -- and the last
// run e - 1.  Between two samples lie fewer than S runs.
PSK_HD int32_t first_sample(int32_t a, int32_t S) { return (a / S + 1) * S - 1; }
PSK_HD int32_t sample_count(int32_t a, int32_t e, int32_t S) {
  if (e <= a) return 0;
  const int32_t g0 = first_sample(a, S);
  return (e - 2 >= g0 ? (e - 2 - g0) / S + 1 : 0) + 1;
}
PSK_D int32_t sample_run(const DSource& s, int32_t j, int32_t S) {
  return j + 1 < s.ns ? s.g0 + j * S : s.e - 1;
}
// the mapped end of run g, read from the compact index where g is on its grid
PSK_D int32_t indexed_end(const DSource& s, int32_t g) {
  int32_t raw;
  if (s.index != nullptr && ((g + 1) & (kIndexStride - 1)) == 0) raw = ld_nc(s.index + ((g + 1) / kIndexStride - 1));
  else raw = ld_nc(s.runs + g).x;
  return s.nviews ? map_chain(s.v, s.nviews, raw) : raw;
}
PSK_D int32_t sample_key(const DSource& s, int32_t j, int32_t S) { return indexed_end(s, sample_run(s, j, S)); }
// runs [lo, hi) between sample pos - 1 and sample pos; run hi is the sample itself
PSK_D void sample_interval(const DSource& s, int32_t pos, int32_t S, int32_t* lo, int32_t* hi) {
  *lo = pos == 0 ? s.a : sample_run(s, pos - 1, S) + 1;
  *hi = sample_run(s, pos, S);
}

// number of samples of s whose key is <= key (le) or < key (!le); warp-cooperative
PSK_D int32_t warp_count_samples(const DSource& s, int32_t ns, int32_t S, int32_t key, bool le) {
  const int lane = threadIdx.x & 31;
  int32_t lo = 0, hi = ns;
  while (hi - lo > 32) {
    const int32_t idx = lo + (int32_t)(((long long)(hi - lo) * (lane + 1)) / 33);
    const int32_t v = sample_key(s, idx, S);
    const bool before = le ? (v <= key) : (v < key);
    const int nb = __popc(__ballot_sync(0xffffffffu, before));
    const int32_t last_before = __shfl_sync(0xffffffffu, idx, nb > 0 ? nb - 1 : 0);
    const int32_t first_after = __shfl_sync(0xffffffffu, idx, nb < 32 ? nb : 31);
    if (nb < 32) hi = first_after;
    if (nb > 0) lo = last_before + 1;
  }
  const int32_t idx = lo + lane;
  bool before = false;
  if (idx < hi) {
    const int32_t v = sample_key(s, idx, S);
    before = le ? (v <= key) : (v < key);
  }
  return lo + __popc(__ballot_sync(0xffffffffu, before));
}

#ifndef __CUDACC_RTC__  // (the run-time specialisation compiles the merge kernel only)
constexpr int kPartBlock = 256;  // threads = samples per block of k_partition

// ---------------------------------------------------------------------------------------
// k_prepare: one warp per source.  [a, e) = runs whose mapped end is in (0, span], cut
// after the first run that reaches span (everything later is a zero-length mapped run).
// Also reads the run counts of sources that are queued results, and fixes the sample and
// tile counts of the evaluation.
__global__ void k_prepare(DPlan* p) {
  __shared__ int32_t s_ns[PSK_MAX_SOURCES];
  const int i = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = p->k;
  if (i < k) {
    DSource& s = p->src[i];
    int32_t n = s.nruns;
    if (s.d_count) n = ld_nc(s.d_count);
    int32_t a = 0, e = n;
    if (s.nviews) {
      // (the searches read s.e / s.a only through the arguments below)
      a = warp_bound_mapped(s, 0, n, 0, true);
      e = warp_bound_mapped(s, a, n, p->span, false) + 1;
      if (e > n) e = n;
    }
    if (lane == 0) {
      s.nruns = n;
      s.a = a;
      s.e = e;
      s.g0 = first_sample(a, p->S);
      s.ns = sample_count(a, e, p->S);
      s_ns[i] = s.ns;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int32_t off = 0, blk = 0;
    for (int q = 0; q < k; ++q) {
      p->src[q].blk_off = blk;
      blk += (s_ns[q] + kPartBlock - 1) / kPartBlock;
      off += s_ns[q];
    }
    p->total_samples = off;
    int32_t nt = (off + p->m - 1) / p->m;
    p->n_tiles = nt < 1 ? 1 : nt;
  }
}

// which source a block of k_partition works on: last i with blk_off[i] <= b (sources without
// samples share their successor's offset and are skipped)
PSK_D int block_owner(const int32_t* s_off, int k, int32_t b) {
  int l = 0, h = k - 1;
  while (l < h) {
    const int mid = (l + h + 1) >> 1;
    if (s_off[mid] <= b) l = mid; else h = mid - 1;
  }
  return l;
}

// samples of source q inside the block's window [wlo, wlo + wn) that come before `key`
// (<= key if le, else < key); the window's keys are staged at pool[wo ...] unless wo < 0
PSK_D int32_t window_count(const DSource& s, const int32_t* pool, int32_t wo, int32_t wlo, int32_t wn, int32_t S,
                           int32_t key, bool le) {
  if (wo >= 0) {
    // staged window: branch-free descent over powers of two (the trip count depends on the
    // window only, so the lanes of a warp stay together); v <= key is v < key + 1
    const int32_t* w = pool + wo - 1;
    const int32_t bound = key + (le ? 1 : 0);
    int32_t pos = 0;
    for (int32_t step = wn > 0 ? 1 << (31 - __clz(wn)) : 0; step > 0; step >>= 1) {
      const int32_t np = pos + step;
      if (np <= wn && w[np] < bound) pos = np;
    }
    return pos;
  }
  int32_t lo = 0, hi = wn;
  while (lo < hi) {
    const int32_t mid = lo + ((hi - lo) >> 1);
    const int32_t v = sample_key(s, wlo + mid, S);
    if (le ? (v <= key) : (v < key)) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// ---------------------------------------------------------------------------------------
// k_partition: samples -> merged ranks -> tile bounds, in one kernel.
//
// Total order of runs: (mapped end, source, run index).  The sample of source i with key x
// that gets merged rank r = (t + 1) m - 1 closes tile t at coordinate x; source q's runs
// before that point are, for q < i, those with mapped end <= x, for q > i those with mapped
// end < x, and for q = i those up to the sample's own run.  bounds[(t + 1) k + q] is the first
// run after it.  With views, lookahead[t k + q] is the first run of q whose mapped end is
// beyond x (it may lie far behind the bound inside a collapsed group of runs), mapped, as a pair.
constexpr int kPartPool = 4096;  // sample keys of the other sources staged per block

#ifndef PSK_PART_MINBLOCKS
#define PSK_PART_MINBLOCKS 5
#endif
__global__ void __launch_bounds__(kPartBlock, PSK_PART_MINBLOCKS) k_partition(DPlan* p) {
  __shared__ int32_t s_off[PSK_MAX_SOURCES + 1];
  __shared__ int32_t s_ns[PSK_MAX_SOURCES];
  __shared__ int32_t s_wlo[PSK_MAX_SOURCES];   // window of source q: first sample
  __shared__ int32_t s_wn[PSK_MAX_SOURCES];    // ... number of samples
  __shared__ int32_t s_woff[PSK_MAX_SOURCES + 1];  // ... offset in s_pool, or -1 (searched in global memory)
  __shared__ int32_t s_pool[kPartPool];
  __shared__ int32_t s_min[kPartBlock / 32], s_max[kPartBlock / 32];
  __shared__ int32_t s_ckey[kPartBlock], s_cj[kPartBlock], s_ci[kPartBlock], s_ct[kPartBlock];  // the block's closing samples
  __shared__ int32_t s_nclose;
#ifdef PSK_PARTITION_TIMING
  unsigned long long gt0_;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0_));
#endif
  const int k = p->k, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int32_t S = p->S, m = p->m;
  const int32_t n_tiles = p->n_tiles;
  const bool any_views = p->any_views != 0;

  // housekeeping that would otherwise be memsets: look-back words, control words, the
  // first and last rows of the bounds table
  for (int32_t t = blockIdx.x * kPartBlock + tid; t < p->max_tiles * kStateStride; t += gridDim.x * kPartBlock)
    p->tile_state[t] = 0ull;
  if (blockIdx.x == 0) {
    if (tid < CTRL_WORDS) p->ctrl[tid] = 0;
    if (tid < k) {
      p->bounds[tid] = p->src[tid].a;
      p->bounds[n_tiles * k + tid] = p->src[tid].e;
      if (any_views) {
        p->lookahead[((n_tiles - 1) * k + tid) * 2] = make_int2(kSentinel, 0);
        p->lookahead[((n_tiles - 1) * k + tid) * 2 + 1] = make_int2(kSentinel, 0);
      }
    }
    if (tid == 0) p->tile_hi[n_tiles - 1] = p->span;
  }
  PSK_PPHASE_DECL;
  if (tid < k) {
    s_off[tid] = p->src[tid].blk_off;
    s_ns[tid] = p->src[tid].ns;
  }
  if (tid == 0) s_nclose = 0;
  __syncthreads();
  // a block ranks up to kPartBlock consecutive samples of ONE source, so the key range it has to
  // stage from the other sources stays narrow
  const int i = block_owner(s_off, k, (int32_t)blockIdx.x);
  const int32_t j = ((int32_t)blockIdx.x - s_off[i]) * kPartBlock + tid;
  const bool valid = j < s_ns[i];
  if (j - tid >= s_ns[i]) return;  // (the grid is sized for the worst case: whole blocks beyond the samples)
  const int32_t key = valid ? sample_key(p->src[i], j, S) : 0;
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
  for (int w = 0; w < kPartBlock / 32; ++w) {
    kmin = imin(kmin, s_min[w]);
    kmax = imax(kmax, s_max[w]);
  }
  PSK_PPHASE(0);
  // window of every source: samples [#keys < kmin, #keys <= kmax); one warp per source
  // (the two searches of a source run in different warps, i.e. at the same time)
  for (int job = warp; job < 2 * k; job += kPartBlock / 32) {
    const int q = job >> 1;
    const bool upper = (job & 1) != 0;
    const int32_t c = warp_count_samples(p->src[q], s_ns[q], S, upper ? kmax : kmin, upper);
    if (lane == 0) {
      if (upper) s_wn[q] = c; else s_wlo[q] = c;
    }
  }
  __syncthreads();
  PSK_PPHASE(1);
  if (tid == 0) {  // pool layout: windows that do not fit are searched in global memory
    int32_t used = 0;
    for (int q = 0; q < k; ++q) s_wn[q] -= s_wlo[q];
    for (int q = 0; q < k; ++q) {
      if (used + s_wn[q] <= kPartPool) {
        s_woff[q] = used;
        used += s_wn[q];
      } else {
        s_woff[q] = -1;
      }
    }
    s_woff[k] = used;
  }
  __syncthreads();
  for (int q = 0; q < k; ++q) {
    const int32_t wo = s_woff[q];
    if (wo < 0) continue;
    const int32_t wlo = s_wlo[q], wn = s_wn[q];
    for (int32_t t = tid; t < wn; t += kPartBlock) s_pool[wo + t] = sample_key(p->src[q], wlo + t, S);
  }
  __syncthreads();
  PSK_PPHASE(2);
  // rank among all samples: own index + samples of every other source that come before
  int32_t rank = j;
  if (valid) {
    for (int q = 0; q < k; ++q) {
      if (q == i) continue;
      rank += s_wlo[q] + window_count(p->src[q], s_pool, s_woff[q], s_wlo[q], s_wn[q], S, key, q < i);
    }
  }
  PSK_PPHASE(3);
  // the samples that close a tile (merged rank = -1 mod m) are collected; the last tile ends at span
  if (valid && (rank + 1) % m == 0 && (rank + 1) / m < n_tiles) {
    const int c = atomicAdd(&s_nclose, 1);
    s_ckey[c] = key;
    s_cj[c] = j;
    s_ci[c] = i;
    s_ct[c] = (rank + 1) / m - 1;
  }
  __syncthreads();
  // Per closing sample and source: the first run after the splitter.  One work item per
  // (closing sample, source), handled by a group of 8 lanes: the sample interval around the
  // splitter (< S runs) is narrowed to fewer than kIndexStride runs by the index entries inside
  // it -- one load per lane -- and those runs are tested in a second round of loads.
  const int nitems = s_nclose * k;
  const int grp = tid >> 3, l8 = tid & 7;
  for (int w0 = 0; w0 < nitems; w0 += kPartBlock / 8) {
    const int w = w0 + grp;
    const bool live = w < nitems;
    const int c = live ? w / k : 0, q = live ? w - c * k : 0;
    const int32_t ckey = s_ckey[c], ct = s_ct[c];
    const int ci = s_ci[c];
    const int32_t cj = s_cj[c];
    const DSource& s = p->src[q];
    const bool le = q < ci;  // runs of a source in front of the sample's own tie with it
    int32_t lo = 0, hi = 0;
    if (live) {
      if (q == ci) {
        lo = hi = sample_run(s, cj, S) + 1;
      } else {
        const int32_t pos = s_wlo[q] + window_count(s, s_pool, s_woff[q], s_wlo[q], s_wn[q], S, ckey, le);
        if (pos >= s_ns[q]) lo = hi = s.e;
        else sample_interval(s, pos, S, &lo, &hi);  // runs below lo are before the splitter, run hi is after it
      }
    }
    {  // (every lane takes part in the shuffles; groups without work pass with lo == hi)
      const int32_t i0 = lo / kIndexStride, n1 = s.index != nullptr ? hi / kIndexStride - i0 : 0;
      int32_t cnt = 0;
      for (int32_t r = l8; r < n1; r += 8) {
        const int32_t raw = ld_nc(s.index + i0 + r);
        const int32_t me = s.nviews ? map_chain(s.v, s.nviews, raw) : raw;
        cnt += (le ? (me <= ckey) : (me < ckey)) ? 1 : 0;
      }
      cnt += __shfl_xor_sync(0xffffffffu, cnt, 1);
      cnt += __shfl_xor_sync(0xffffffffu, cnt, 2);
      cnt += __shfl_xor_sync(0xffffffffu, cnt, 4);
      if (cnt < n1) hi = (i0 + cnt) * kIndexStride + kIndexStride - 1;
      if (cnt > 0) lo = (i0 + cnt) * kIndexStride;
    }
    {
      int32_t cnt = 0;
      for (int32_t r = lo + l8; r < hi; r += 8) {
        const int32_t me = mapped_end(s, r);
        cnt += (le ? (me <= ckey) : (me < ckey)) ? 1 : 0;
      }
      cnt += __shfl_xor_sync(0xffffffffu, cnt, 1);
      cnt += __shfl_xor_sync(0xffffffffu, cnt, 2);
      cnt += __shfl_xor_sync(0xffffffffu, cnt, 4);
      lo += cnt;
    }
    if (live && l8 == 0) {
      if (q == ci) p->tile_hi[ct] = ckey;
      p->bounds[(ct + 1) * k + q] = lo;
      if (any_views) {
        // Look-ahead runs of q behind the tile (mapped, as pairs).  Every run from the bound on
        // has a mapped end >= x.  First: the run that covers the coordinates right behind q's
        // last in-tile run -- the run at the bound unless it maps onto no coordinate (mapped end
        // not beyond its predecessor's: a collapsed group that continues), then the first run
        // beyond x.  Second: the first run beyond x when the first one ends AT x.
        const int32_t b = lo;
        int2 l1 = make_int2(kSentinel, 0), l2 = make_int2(kSentinel, 0);
        if (b < s.e) {
          const int32_t prev = b > s.a ? mapped_end(s, b - 1) : 0;
          int2 r = ld_nc(s.runs + b);
          if (s.nviews) r.x = map_chain(s.v, s.nviews, r.x);
          if (r.x > prev) l1 = r;
          if (r.x <= ckey) {
            // gallop over the runs that end at x
            int32_t step = 1, last = b;
            while (last + step < s.e && mapped_end(s, last + step) <= ckey) {
              last += step;
              step <<= 1;
            }
            const int32_t top = last + step < s.e ? last + step : s.e;
            const int32_t beyond = ub_mapped(s, last + 1, top, ckey);  // first run whose mapped end is > x
            int2 r2 = make_int2(kSentinel, 0);
            if (beyond < s.e) {
              r2 = ld_nc(s.runs + beyond);
              if (s.nviews) r2.x = map_chain(s.v, s.nviews, r2.x);
            }
            if (r.x > prev) l2 = r2; else l1 = r2;
          }
        }
        p->lookahead[(ct * k + q) * 2] = l1;
        p->lookahead[(ct * k + q) * 2 + 1] = l2;
      }
    }
  }
#ifdef PSK_PARTITION_TIMING
  PSK_PPHASE(4);
  if (tid == 0) {
    unsigned long long gt1_;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1_));
    const unsigned long long mask40 = (1ull << 40) - 1;
    atomicAdd(&p->phase[5], gt1_ - gt0_);                          // block durations, ns
    atomicMax(&p->phase[6], gt1_ & mask40);                        // end of the last block ...
    atomicMax(&p->phase[7], (1ull << 40) - (gt0_ & mask40));       // ... and start of the first one
  }
#endif
}

#endif  // !__CUDACC_RTC__

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
  // (no trailing barrier: every caller passes another barrier before s_warp is written again)
  return excl;
}

// One look-back entry per tile, padded to a 32-byte sector: hundreds of resident CTAs read the
// entries of the tiles in front of them at the same time.
PSK_D void publish_aggregate(unsigned long long* state, int32_t tile, int32_t count) {
  // tile 0 has no predecessor: its aggregate is already its inclusive prefix
  const unsigned long long st = tile == 0 ? ST_PREFIX : ST_AGGREGATE;
  st_volatile_u64(&state[(long long)tile * kStateStride], (st << 32) | (uint32_t)count);
}

// Decoupled look-back, executed by the whole CTA.  Thread r loads the entries of the tiles at
// distances r+1, r+1+T, ... in front of `tile` -- kLookW independent loads, one memory round
// trip for T * kLookW predecessors, which normally covers every tile that can still be in
// flight.  The loads are ISSUED early (lookback_issue, right after the tile's own aggregate
// is published) and consumed later (lookback_finish, after the next tile has been searched),
// so their latency never shows; an entry that is still empty then is polled with back-off.
// The block folds in tile order: the nearest tile with an inclusive PREFIX ends the sum,
// everything nearer contributes its AGGREGATE.  Tiles are handed out by an atomic ticket, so
// every predecessor is resident or finished and the wait cannot deadlock.  Status and count
// travel in ONE 64-bit word, so no ordering between separate words is needed.
constexpr int kLookWMax = 8;
template <int T>
struct LookW {  // T * LookW<T>::value >= 1024 predecessors per round
  static constexpr int value = T >= 256 ? 4 : 8;
};

struct LookWords {
  unsigned long long w[kLookWMax];
};

template <int T>
PSK_D LookWords lookback_issue(const unsigned long long* state, int32_t tile, int32_t D0) {
  constexpr int kLookW = LookW<T>::value;
  LookWords lw;
#pragma unroll
  for (int u = 0; u < kLookW; ++u) {
    const int32_t idx = tile - (D0 + 1 + (int32_t)threadIdx.x + T * u);
    lw.w[u] = idx >= 0 ? ld_volatile_u64(&state[(long long)idx * kStateStride]) : 0ull;
  }
  return lw;
}

template <int T>
PSK_D int32_t lookback_finish(unsigned long long* state, int32_t tile, int32_t count, LookWords lw,
                              int32_t* s_sum /*[LookW<T>::value * T/32]*/, int32_t* s_has /*[LookW<T>::value * T/32]*/) {
  constexpr int NW = T / 32;
  constexpr int kLookW = LookW<T>::value;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int32_t excl = 0;
  if (tile > 0) {
    for (int32_t D0 = 0;; D0 += T * kLookW) {
      if (D0 > 0) lw = lookback_issue<T>(state, tile, D0);
      __syncthreads();  // s_sum / s_has free (previous round or other users)
#pragma unroll
      for (int u = 0; u < kLookW; ++u) {
        const int32_t idx = tile - (D0 + 1 + tid + T * u);
        unsigned long long w = lw.w[u];
        if (idx >= 0) {
          while ((w >> 32) == ST_INVALID) {
            sleep_ns(64);
            w = ld_volatile_u64(&state[(long long)idx * kStateStride]);
          }
        }
        // idx == -1: the virtual prefix 0 in front of tile 0
        const bool is_prefix = idx >= 0 ? (w >> 32) == ST_PREFIX : idx == -1;
        const unsigned pmask = __ballot_sync(0xffffffffu, is_prefix);
        const int first = pmask ? (__ffs((int)pmask) - 1) : 32;
        int32_t c = (lane <= first && idx >= 0) ? (int32_t)(uint32_t)w : 0;
#pragma unroll
        for (int dd = 16; dd > 0; dd >>= 1) c += __shfl_xor_sync(0xffffffffu, c, dd);
        if (lane == 0) {
          s_sum[u * NW + warp] = c;
          s_has[u * NW + warp] = pmask != 0;
        }
      }
      __syncthreads();
      // fold in tile order (window, then warp), lane q takes piece q: stop behind the first
      // piece with a PREFIX
      bool done = false;
      for (int q0 = 0; q0 < kLookW * NW && !done; q0 += 32) {
        const int q = q0 + lane;
        const bool in = q < kLookW * NW;
        const unsigned pm = __ballot_sync(0xffffffffu, in && s_has[q] != 0);
        const int first = pm ? (__ffs((int)pm) - 1) : 32;
        int32_t c = (in && lane <= first) ? s_sum[q] : 0;
#pragma unroll
        for (int dd = 16; dd > 0; dd >>= 1) c += __shfl_xor_sync(0xffffffffu, c, dd);
        excl += c;
        done = pm != 0;
      }
      if (done) break;
    }
  }
  if (tid == 0)
    st_volatile_u64(&state[(long long)tile * kStateStride],
                    ((unsigned long long)ST_PREFIX << 32) | (uint32_t)(excl + count));
  return excl;
}

// Non-blocking form: false (and nothing published) when an entry that is needed -- nearer than
// the nearest PREFIX -- has not been published yet, or the window holds no PREFIX at all.
template <int T>
PSK_D bool lookback_try(unsigned long long* state, int32_t tile, int32_t count, const LookWords& lw, int32_t* excl_out,
                        int32_t* s_sum, int32_t* s_has) {
  constexpr int NW = T / 32;
  constexpr int kLookW = LookW<T>::value;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int32_t excl = 0;
  bool ok = true;
  if (tile > 0) {
    // (s_sum / s_has: the caller has passed a barrier since their last readers)
#pragma unroll
    for (int u = 0; u < kLookW; ++u) {
      const int32_t idx = tile - (1 + tid + T * u);
      const unsigned long long w = lw.w[u];
      const bool is_prefix = idx >= 0 ? (w >> 32) == ST_PREFIX : idx == -1;
      const bool is_invalid = idx >= 0 && (w >> 32) == ST_INVALID;
      const unsigned pmask = __ballot_sync(0xffffffffu, is_prefix);
      const unsigned imask = __ballot_sync(0xffffffffu, is_invalid);
      const int first = pmask ? (__ffs((int)pmask) - 1) : 32;
      const unsigned upto = first >= 31 ? 0xffffffffu : ((2u << first) - 1u);  // lanes 0..first
      int32_t c = (lane <= first && idx >= 0) ? (int32_t)(uint32_t)w : 0;
#pragma unroll
      for (int dd = 16; dd > 0; dd >>= 1) c += __shfl_xor_sync(0xffffffffu, c, dd);
      if (lane == 0) {
        s_sum[u * NW + warp] = c;
        // bit 0: piece holds a PREFIX; bit 1: an entry in front of it is still empty
        s_has[u * NW + warp] = (pmask != 0 ? 1 : 0) | ((imask & upto) != 0 ? 2 : 0);
      }
    }
    __syncthreads();
    // fold in tile order (window, then warp), lane q takes piece q
    bool done = false;
    for (int q0 = 0; q0 < kLookW * NW && !done; q0 += 32) {
      const int q = q0 + lane;
      const bool in = q < kLookW * NW;
      const int32_t h = in ? s_has[q] : 0;
      const unsigned pm = __ballot_sync(0xffffffffu, (h & 1) != 0);
      const int first = pm ? (__ffs((int)pm) - 1) : 32;
      if (__ballot_sync(0xffffffffu, (h & 2) != 0 && lane <= first)) ok = false;
      int32_t c = (in && lane <= first) ? s_sum[q] : 0;
#pragma unroll
      for (int dd = 16; dd > 0; dd >>= 1) c += __shfl_xor_sync(0xffffffffu, c, dd);
      excl += c;
      done = pm != 0;
    }
    if (!done) ok = false;
  }
  if (ok && tid == 0)
    st_volatile_u64(&state[(long long)tile * kStateStride],
                    ((unsigned long long)ST_PREFIX << 32) | (uint32_t)(excl + count));
  *excl_out = excl;
  return ok;
}

// both steps at once (callers without anything to overlap)
template <int T>
PSK_D int32_t lookback_exclusive_cta(unsigned long long* state, int32_t tile, int32_t count, int32_t* s_sum,
                                     int32_t* s_has) {
  return lookback_finish<T>(state, tile, count, lookback_issue<T>(state, tile, 0), s_sum, s_has);
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
struct TileDesc {  // bounds of one tile, fetched one tile ahead
  int32_t tile;
  int32_t lo, hi;
  int32_t b0[PSK_MAX_SOURCES];
  int32_t b1[PSK_MAX_SOURCES];
  int2 la[2][PSK_MAX_SOURCES];  // with views: the two (mapped) look-ahead runs of every source
};

struct TileLayout {
  int32_t start[PSK_MAX_SOURCES];  // pair index of the source's first in-tile run
  int32_t len[PSK_MAX_SOURCES];    // in-tile runs (look-ahead runs follow)
  int32_t flat[PSK_MAX_SOURCES + 1];  // exclusive scan of len
  int32_t ok;                      // 0: the tile does not fit (partition bound violated)
};

// Bounds of a tile held in the registers of one warp while the loads are in flight (lane i
// holds source i's bounds), written to shared memory only when the tile loop is ready for it.
struct DescRegs {
  int32_t tile, lo, hi, b0, b1;
  int2 la0, la1;
};

// one warp: start fetching the bounds of tile `ticket` (lane 0's value; no barrier, and no use
// of the loaded values: they stay in flight until store_tile_desc).  The ticket itself is
// taken earlier (take_ticket, at the top of the tile loop), so the atomic's round trip hides
// behind the wait for the tile's runs and the cursor search.
PSK_D int32_t take_ticket(DPlan* p) {
  int32_t t = 0;
  if ((threadIdx.x & 31) == 0) t = atomicAdd(&p->ctrl[CTRL_TICKET], 1);
  return t;
}
PSK_D DescRegs fetch_tile_desc(DPlan* p, int k, int32_t n_tiles, bool views, int32_t ticket) {
  const int lane = threadIdx.x & 31;
  DescRegs r;
  r.lo = r.hi = r.b0 = r.b1 = 0;
  r.la0 = r.la1 = make_int2(kSentinel, 0);
  r.tile = __shfl_sync(0xffffffffu, ticket, 0);
  if (r.tile < n_tiles) {
    if (lane < k) {
      r.b0 = p->bounds[r.tile * k + lane];
      r.b1 = p->bounds[(r.tile + 1) * k + lane];
      if (views) {
        r.la0 = p->lookahead[(r.tile * k + lane) * 2];
        r.la1 = p->lookahead[(r.tile * k + lane) * 2 + 1];
      }
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
    d->la[0][lane] = r.la0;
    d->la[1][lane] = r.la1;
  }
}

// one warp: lay the tile out in a shared region and issue one bulk copy per source (lane q
// copies source q).  A source's slice [b0, b1) plus two look-ahead runs is copied from the
// even pair index at or below b0 (16-byte alignment of the copy engine); the stores keep
// sentinel pairs behind their last run, so the copy never needs a bounds test.
PSK_D void issue_tile_copies(const DPlan* p, const TileDesc& d, int k, const SmemPairs& sm, int region,
                             TileLayout* L, psk_mbar_t* bar) {
  const int lane = threadIdx.x & 31;
  int32_t cnt = 0, seg = 0, odd = 0;
  if (lane < k) {
    odd = d.b0[lane] & 1;
    cnt = d.b1[lane] - d.b0[lane];
    seg = (odd + cnt + 2 + 1) & ~1;
  }
  int32_t inc = seg, cinc = cnt;
#pragma unroll
  for (int dd = 1; dd < 32; dd <<= 1) {
    const int32_t o = __shfl_up_sync(0xffffffffu, inc, dd);
    const int32_t c = __shfl_up_sync(0xffffffffu, cinc, dd);
    if (lane >= dd) {
      inc += o;
      cinc += c;
    }
  }
  const int32_t total = __shfl_sync(0xffffffffu, inc, 31);
  const int32_t ctotal = __shfl_sync(0xffffffffu, cinc, 31);
  const bool ok = total <= kTileRegion && ctotal <= kTileCap;
  if (lane < k) {
    L->start[lane] = inc - seg + odd;
    L->len[lane] = cnt;
    L->flat[lane] = cinc - cnt;
  }
  if (lane == 0) {
    L->ok = ok ? 1 : 0;
    L->flat[k] = ctotal;
    mbar_arrive_expect_tx(bar, ok ? (uint32_t)total * 8u : 0u);
  }
  __syncwarp();
  if (ok && lane < k)
    bulk_g2s(sm.dst(region + inc - seg), p->src[lane].runs + (d.b0[lane] - odd), (uint32_t)seg * 8u, bar);
}

// One step of the cursor search: first index in [c, c + r) whose end is > key, where the last
// candidate is known to qualify.  4-ary while the range allows it: three INDEPENDENT probes per
// step, so a range of n takes log4(n) dependent shared-memory round trips instead of log2(n).
// The range shrinks by the same amount whatever the outcome (branch-free).
PSK_D void search_step4(const SmemPairs& sm, int32_t& c, int32_t& r, int32_t key) {
  if (r >= 4) {
    const int32_t q = r >> 2;
    const bool b1 = sm.ldx(c + q - 1) <= key;
    const bool b2 = sm.ldx(c + 2 * q - 1) <= key;
    const bool b3 = sm.ldx(c + 3 * q - 1) <= key;
    c += b3 ? 3 * q : b2 ? 2 * q : b1 ? q : 0;
    r -= 3 * q;
  } else {
    const int32_t half = r >> 1;
    c += (sm.ldx(c + half - 1) <= key) ? half : 0;
    r -= half;
  }
}

// which source owns flat in-tile slot f: last q with flat[q] <= f
PSK_D int flat_owner(const int32_t* flat, int k, int32_t f) {
  int l = 0, h = k - 1;
  while (l < h) {
    const int mid = (l + h + 1) >> 1;
    if (flat[mid] <= f) l = mid; else h = mid - 1;
  }
  return l;
}

// TYPED_FLOAT: compaction compares float values with operator== (PSK_EQ_TYPED on F32)
// instead of the 32-bit payloads.
// VIEWS: 1 = some source has a view, 0 = none (the staging pass is compiled out), -1 = ask
// the plan at run time (the interpreter kernels).
template <int K, bool TYPED_FLOAT, int VIEWS, typename Eval>
PSK_D void merge_eval_tile_loop(DPlan* p, const Eval& ev) {
  constexpr int T = kTileThreads;
  constexpr int RA = 0;                // region A: the bulk copies land here
  constexpr int RB = kTileRegion;      // region B: the spans the threads keep

  PSK_DYN_SMEM(smem_raw);
  const SmemPairs sm(smem_raw);

  __shared__ TileDesc s_desc[2];
  __shared__ TileLayout s_layout[2];
  __shared__ int32_t s_warp[T / 32 + 2];
  __shared__ int32_t s_lbsum[LookW<T>::value * (T / 32)];
  __shared__ int32_t s_lbhas[LookW<T>::value * (T / 32)];
  __shared__ int32_t s_newstart[PSK_MAX_SOURCES];
  // cursors of every kCoarse-th thread (first level of the cursor search)
  constexpr int kCoarse = PSK_COARSE_GROUP > 0 ? PSK_COARSE_GROUP : 8;
  __shared__ int32_t s_coarse[K * (T / kCoarse + 1)];
  __shared__ psk_mbar_t s_bar;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int k = p->k;
  constexpr bool typed_float = TYPED_FLOAT;
  const bool views = VIEWS < 0 ? (p->any_views != 0) : (VIEWS != 0);
  const int32_t span = p->span;
  const int32_t n_tiles = p->n_tiles;
  // Without views the merge reads region A and keeps its spans in B (or C); they stay there
  // while the CTA already works on the next tile -- until that tile has been searched (two
  // regions) or merged (three regions) -- so the look-back of a tile is off the critical path.
  // With views the squeeze moves the mapped runs A -> B, the merge keeps its spans in A and
  // they are written out at once.
  const int RD = views ? RB : RA;

  if (tid == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  if (warp == 0) {
    const int32_t t0 = take_ticket(p);
    store_tile_desc(fetch_tile_desc(p, k, n_tiles, views, t0), &s_desc[0], k);
    __syncwarp();
    if (s_desc[0].tile < n_tiles) issue_tile_copies(p, s_desc[0], k, sm, RA, &s_layout[0], &s_bar);
  }
  __syncthreads();

  // ---- parked tiles (no views): ring of kParkDepth slots of this CTA in global memory --------
  __shared__ int32_t s_pk_tile[kParkDepth > 0 ? kParkDepth : 1], s_pk_total[kParkDepth > 0 ? kParkDepth : 1];
  int pk_head = 0, pk_count = 0;   // oldest slot, parked tiles (uniform over the CTA)
  int32_t look_tile = -1;          // tile whose look-back entries were requested last
  LookWords look;
#pragma unroll
  for (int u = 0; u < kLookWMax; ++u) look.w[u] = 0ull;
  int2* const park = views ? nullptr : p->park + (long long)blockIdx.x * kParkDepth * kTileCap;
  // plan fields of the tile loop, read once (the plan lives in global memory)
  unsigned long long* const tile_state = p->tile_state;
  int2* const out = p->out;
  const int32_t out_capacity = p->out_capacity;

  PSK_PHASE_DECL;
  // the last tile also leaves the run count and the sentinel pairs behind
  auto finish_output = [&](int32_t ftile, int32_t cnt) {
    if (tid == 0 && ftile == n_tiles - 1) {
      p->ctrl[CTRL_COUNT] = cnt;
      *p->out_count = cnt;  // the count stays on the device too (queued consumers)
      if (cnt <= out_capacity) {
        out[cnt] = make_int2(kSentinel, 0);
        out[cnt + 1] = make_int2(kSentinel, 0);
      }
    }
  };
  // every thread copies its own spans (shared region -> global), 16 bytes per store
  auto store_spans = [&](int2* base_ptr, int32_t o, int src, int32_t fn) {
    int2* const dst = base_ptr + o;
    int32_t j = 0;
    if (fn > 0 && (o & 1)) {  // 16-byte alignment of the vector stores
      dst[0] = sm.ld(src);
      j = 1;
    }
    for (; j + 1 < fn; j += 2) {
      const int2 a = sm.ld(src + j), b = sm.ld(src + j + 1);
      *reinterpret_cast<uint4*>(dst + j) = make_uint4((unsigned)a.x, (unsigned)a.y, (unsigned)b.x, (unsigned)b.y);
    }
    if (j < fn) dst[j] = sm.ld(src + j);
  };
  // with views: global offset by a blocking look-back, spans straight to their place
  auto flush_direct = [&](int32_t ftile, int32_t ftotal, int32_t fn, int32_t fbase, int32_t fofs, int fregion) {
    const int32_t gofs = lookback_exclusive_cta<T>(tile_state, ftile, ftotal, s_lbsum, s_lbhas);
    finish_output(ftile, gofs + ftotal);
    if (gofs + ftotal <= out_capacity) store_spans(out, gofs + fofs, fregion + fbase, fn);
    else if (tid == 0) p->ctrl[CTRL_ERROR] = 2;
    __syncthreads();  // the span region may be overwritten now
  };
  // the oldest parked tile moves to its place in the output if every tile in front of it has
  // published its count (`block`: wait for that); coalesced 8-byte copies through L2
  auto unpark = [&](bool block) -> bool {
    const int32_t ftile = s_pk_tile[pk_head], ftotal = s_pk_total[pk_head];
    int32_t gofs = 0;
    if (block) __syncthreads();  // (the fold arrays may still be read from a previous attempt)
    bool ok = look_tile == ftile && lookback_try<T>(tile_state, ftile, ftotal, look, &gofs, s_lbsum, s_lbhas);
    while (!ok && block) {
      look = lookback_issue<T>(tile_state, ftile, 0);
      look_tile = ftile;
      __syncthreads();
      ok = lookback_try<T>(tile_state, ftile, ftotal, look, &gofs, s_lbsum, s_lbhas);
      if (!ok) sleep_ns(200);
    }
    look_tile = -1;
    if (!ok) return false;
    finish_output(ftile, gofs + ftotal);
    if (gofs + ftotal <= out_capacity) {
      const int2* const src = park + (long long)pk_head * kTileCap;
      int2* const dst = out + gofs;
      // eight loads in flight per thread (every iteration of a plain copy loop would pay the
      // L2 round trip again)
      for (int32_t j0 = tid; j0 < ftotal; j0 += 8 * T) {
        int2 r_[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (j0 + u * T < ftotal) r_[u] = ld_cg(src + j0 + u * T);
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (j0 + u * T < ftotal) dst[j0 + u * T] = r_[u];
      }
    } else if (tid == 0) {
      p->ctrl[CTRL_ERROR] = 2;
    }
    pk_head = (pk_head + 1) % (kParkDepth > 0 ? kParkDepth : 1);
    --pk_count;
    // every thread must have read its part before the slot is parked into again: the caller of a
    // blocking call writes at once, everybody else passes the next tile's search barrier first
    if (block) __syncthreads();
    return true;
  };

  for (int it = 0;; ++it) {
    const TileDesc& d = s_desc[it & 1];
    TileLayout& lay = s_layout[it & 1];
    const int32_t tile = d.tile;
    if (tile >= n_tiles) break;
    const int32_t lo = d.lo, hi = d.hi;
    const bool fail = lay.ok == 0;

    // ---- warp 0 takes the ticket of the next tile (the atomic's result is not needed before
    //      the search is done); everybody waits for this tile's runs to land; the look-back
    //      entries for the spans that are still parked are requested now and read later ------
    int32_t ticket = 0;
    if (warp == 0) ticket = take_ticket(p);
    mbar_wait(&s_bar, (uint32_t)(it & 1));
    if (pk_count > 0) {
      look_tile = s_pk_tile[pk_head];
      look = lookback_issue<T>(tile_state, look_tile, 0);
    }
    PSK_PHASE(0);

    // ---- with views: map the raw ends through the view chains and squeeze out the runs
    //      that map onto no output coordinate (mapped end not beyond the predecessor's, or
    //      not beyond the tile's lower bound), A -> B; every source gets its two look-ahead
    //      runs (fetched with the tile's bounds) behind its kept runs ----------------------
    if (views && !fail) {
      const int32_t F = lay.flat[k];
      const int32_t per = (F + T - 1) / T;
      const int32_t f0 = imin(F, tid * per), f1 = imin(F, f0 + per);
      int q = 0;
      int32_t prev = lo;
      if (f0 < f1) {
        q = flat_owner(lay.flat, k, f0);
        if (f0 > lay.flat[q]) {  // mapped end of the run in front of the chunk
          const int32_t e0 = sm.ldx(RA + lay.start[q] + (f0 - lay.flat[q]) - 1);
          prev = imax(lo, p->src[q].nviews ? map_chain(p->src[q].v, p->src[q].nviews, e0) : e0);
        }
      }
      __syncthreads();  // every chunk has read its predecessor before ends are rewritten
      int32_t kept = 0;
      {
        int qq = q;
        int32_t nxt = f0 < f1 ? lay.flat[qq + 1] : 0;
        for (int32_t f = f0; f < f1; ++f) {
          while (f >= nxt) {  // next source (sources without in-tile runs are skipped)
            ++qq;
            nxt = lay.flat[qq + 1];
            prev = lo;
          }
          const int32_t slot = RA + lay.start[qq] + (f - lay.flat[qq]);
          const int32_t e0 = sm.ldx(slot);
          const int nv = p->src[qq].nviews;
          const int32_t me = nv ? map_chain(p->src[qq].v, nv, e0) : e0;
          const bool keep = me > prev;
          sm.stx(slot, keep ? me : -1);
          if (keep) {
            prev = me;
            ++kept;
          }
        }
      }
      int32_t total_kept;
      int32_t pos = block_exclusive_scan<T>(kept, s_warp, &total_kept);
      // new start of every source = kept runs in front of its first flat slot (+ two
      // look-ahead slots per earlier source)
      if (tid < k) s_newstart[tid] = -1;
      __syncthreads();
      {
        int qq = q;
        int32_t nxt = f0 < f1 ? lay.flat[qq + 1] : 0;
        if (f0 < f1 && f0 == lay.flat[qq]) s_newstart[qq] = pos;
        for (int32_t f = f0; f < f1; ++f) {
          while (f >= nxt) {
            ++qq;
            nxt = lay.flat[qq + 1];
            if (f == lay.flat[qq]) s_newstart[qq] = pos;
          }
          const int32_t slot = RA + lay.start[qq] + (f - lay.flat[qq]);
          const int2 r = sm.ld(slot);
          if (r.x >= 0) {
            sm.st(RB + pos + 2 * qq, r.x, r.y);
            ++pos;
          }
        }
      }
      __syncthreads();
      if (tid == 0) {
        // sources without in-tile runs start where their successor starts
        int32_t nxt = total_kept;
        for (int qq = k - 1; qq >= 0; --qq) {
          if (s_newstart[qq] < 0) s_newstart[qq] = nxt;
          const int32_t st = s_newstart[qq];
          lay.start[qq] = st + 2 * qq;
          lay.len[qq] = nxt - st;
          sm.st(RB + nxt + 2 * qq, d.la[0][qq].x, d.la[0][qq].y);  // look-ahead runs
          sm.st(RB + nxt + 2 * qq + 1, d.la[1][qq].x, d.la[1][qq].y);
          nxt = st;
        }
      }
      __syncthreads();
    }
    PSK_PHASE(1);
    if (fail && tid == 0) p->ctrl[CTRL_ERROR] = 1;

    // ---- split the tile's coordinate range (lo, hi] over the threads --------------------
    const long long width = (long long)hi - lo;
    const int32_t c0 = lo + (int32_t)((width * tid) / T);
    const int32_t c1 = lo + (int32_t)((width * (tid + 1)) / T);

    // first staged run of every source whose end is > c0.  Two levels: the cursors of every
    // 8th thread are found first -- k * T/8 full binary searches spread over the threads, one
    // each -- and bracket the searches of the threads in between (branch-free, in lock step
    // over the sources; the look-ahead run bounds each of them).
    int32_t cur[K], head[K], rem[K];
    uint32_t val[K];
    int32_t base = 0;
    constexpr int CG = kCoarse, NC = T / CG;
    if (!fail) {
      for (int32_t q = tid; q < NC * k; q += T) {
        const int i = q / NC, g_ = q - i * NC;
        const int32_t cg = lo + (int32_t)((width * (g_ * CG)) / T);
        // candidates: the in-tile runs and BOTH look-ahead runs -- the first one may end exactly at
        // the tile's upper bound (a tie with the splitter), which is not beyond c0 in an empty
        // tile (lo == hi); the second one always is
        int32_t c = RD + lay.start[i], r = lay.len[i] + 2;
        while (r > 1) search_step4(sm, c, r, cg);
        s_coarse[i * (NC + 1) + g_] = c;
        if (g_ == 0) s_coarse[i * (NC + 1) + NC] = RD + lay.start[i] + lay.len[i] + 1;
      }
    }
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
    {
      int32_t longest = 1;
#pragma unroll
      for (int i = 0; i < K; ++i) longest = imax(longest, rem[i]);
      while (longest > 1) {
        longest = 1;
#pragma unroll
        for (int i = 0; i < K; ++i) {
          if (rem[i] > 1) search_step4(sm, cur[i], rem[i], c0);
          longest = imax(longest, rem[i]);
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
        base += cur[i] - (RD + lay.start[i]);
      } else {
        head[i] = kSentinel;
        val[i] = 0;
      }
      cur_at[i] = sm.at(cur[i]);
    }
    PSK_PHASE(2);
    // the next tile's bounds: requested now, in flight during the merge
    DescRegs fetched;
    fetched.tile = 0;
    if (warp == 0) fetched = fetch_tile_desc(p, k, n_tiles, views, ticket);

    const int WR = views ? RA : RB;
    const uint32_t out_begin = sm.at(WR + base);
    uint32_t out_at = out_begin;  // where this thread's next kept span goes

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
    PSK_PHASE(3);

    // ---- the next tile: its bounds (in warp 0's registers since the top of the loop) go to
    //      shared memory now; its runs are requested as soon as nobody reads region A any
    //      more (without views: after the merge; with views: after the spans left A) -------
    if (warp == 0) store_tile_desc(fetched, &s_desc[(it + 1) & 1], k);
    int32_t tile_total;
    const int32_t my_ofs = block_exclusive_scan<T>(n, s_warp, &tile_total);
    PSK_PHASE(4);
    if (tid == 0) publish_aggregate(tile_state, tile, tile_total);
    const bool next_valid = s_desc[(it + 1) & 1].tile < n_tiles;
    if (!views) {
      // park the tile's spans (dense, in tile order) in this CTA's ring; a full ring first
      // waits for its oldest tile
      if (pk_count == kParkDepth) unpark(true);
      const int slot = (pk_head + pk_count) % kParkDepth;
      store_spans(park + (long long)slot * kTileCap, my_ofs, WR + base, n);
      if (tid == 0) {
        s_pk_tile[slot] = tile;
        s_pk_total[slot] = tile_total;
      }
      ++pk_count;
      if (warp == 0 && next_valid)
        issue_tile_copies(p, s_desc[(it + 1) & 1], k, sm, RA, &s_layout[(it + 1) & 1], &s_bar);
      __syncthreads();  // region B is free, the parked spans and the next layout are visible
      PSK_PHASE(6);
      // the oldest parked tile leaves if its place is known by now (entries requested at the
      // top of this iteration); never waits here
      if (look_tile >= 0 && look_tile == s_pk_tile[pk_head]) unpark(false);
      PSK_PHASE(5);
    } else {
      flush_direct(tile, tile_total, n, base, my_ofs, WR);
      if (warp == 0 && next_valid)
        issue_tile_copies(p, s_desc[(it + 1) & 1], k, sm, RA, &s_layout[(it + 1) & 1], &s_bar);
      __syncthreads();
      PSK_PHASE(6);
    }
  }
  while (pk_count > 0) unpark(true);
  PSK_PHASE_FLUSH;
}

#ifndef __CUDACC_RTC__
template <int K>
__global__ void __launch_bounds__(kTileThreads, kCtasPerSm) k_merge_eval_interp(DPlan* p) {
  __shared__ psk_node s_prog[PSK_MAX_PROGRAM];
  for (int i = threadIdx.x; i < p->nprog; i += blockDim.x) s_prog[i] = p->prog[i];
  __syncthreads();
  InterpEval ev{s_prog, p->nprog};
  if (p->eq != 0) merge_eval_tile_loop<K, true, -1>(p, ev);
  else merge_eval_tile_loop<K, false, -1>(p, ev);
}

#endif  // !__CUDACC_RTC__

}  // namespace psk
