// psk_convert.cuh -- dense <-> RLE conversion, mask generation and reductions over runs.
//
//   encode  : conv::to_store   include/pyskip/detail/conv.hpp:15-44  (two passes over the
//             dense buffer on one CPU thread) -> count / scan / write kernels.
//   expand  : conv::to_buffer  detail/conv.hpp:63-78                -> one kernel.
//   mask    : mask::build      detail/mask.hpp:132-182 driven by TensorSlice::mask
//             (tensor.hpp:171-198) -> one kernel emits the raw (segment) run list, the eval
//             kernel canonicalises it.
//   reduce  : src/pyskip/reduce.py:8-50 (log2(n) strided 2-source evals) collapses to one
//             pass over the runs: sum(val * len), max, min, prod.
#pragma once

#include "psk_ops.h"
#include "psk_plan.h"

namespace psk {

constexpr int kConvThreads = 256;
constexpr int kConvItems = 8;
constexpr int kConvChunk = kConvThreads * kConvItems;  // dense elements per block

// dense element -> 32-bit payload
template <int DT>
PSK_D uint32_t load_payload(const void* dense, long long i) {
  if (DT == PSK_I32 || DT == PSK_F32) return reinterpret_cast<const uint32_t*>(dense)[i];
  if (DT == PSK_BOOL) return reinterpret_cast<const uint8_t*>(dense)[i] ? 1u : 0u;
  return (uint32_t)(int32_t) reinterpret_cast<const int8_t*>(dense)[i];
}
template <int DT>
PSK_D void store_payload(void* dense, long long i, uint32_t v) {
  if (DT == PSK_I32 || DT == PSK_F32) reinterpret_cast<uint32_t*>(dense)[i] = v;
  else reinterpret_cast<uint8_t*>(dense)[i] = (uint8_t)v;
}
// conv.hpp:21-25: typed `buffer[i] != buffer[i-1]`
template <int DT>
PSK_D bool typed_differs(uint32_t a, uint32_t b) {
  if (DT == PSK_F32) return u2f(a) != u2f(b);
  return a != b;
}

// element i closes a run iff it is the last element or differs from element i+1
template <int DT>
__global__ void __launch_bounds__(kConvThreads) k_encode_count(const void* dense, long long n,
                                                               int32_t* block_count) {
  __shared__ int32_t s_warp[kConvThreads / 32];
  const long long base = (long long)blockIdx.x * kConvChunk;
  int32_t c = 0;
#pragma unroll
  for (int it = 0; it < kConvItems; ++it) {
    long long i = base + it * kConvThreads + threadIdx.x;
    if (i < n) {
      bool b = (i == n - 1) || typed_differs<DT>(load_payload<DT>(dense, i), load_payload<DT>(dense, i + 1));
      c += b ? 1 : 0;
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) c += __shfl_down_sync(0xffffffffu, c, d);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int32_t t = 0;
    for (int w = 0; w < kConvThreads / 32; ++w) t += s_warp[w];
    block_count[blockIdx.x] = t;
  }
}

// single-block exclusive scan of nb counts; total -> *total_out
__global__ void __launch_bounds__(1024) k_scan_counts(int32_t* counts, int32_t nb, int32_t* total_out) {
  __shared__ int32_t s_warp[33];
  __shared__ int32_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int32_t base = 0; base < nb; base += 1024) {
    int32_t i = base + threadIdx.x;
    int32_t v = i < nb ? counts[i] : 0;
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
      int32_t w = s_warp[lane];
      int32_t winc = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int32_t o = __shfl_up_sync(0xffffffffu, winc, d);
        if (lane >= d) winc += o;
      }
      s_warp[lane] = winc - w;
      if (lane == 31) s_warp[32] = winc;
    }
    __syncthreads();
    const int32_t carry = s_carry;
    if (i < nb) counts[i] = carry + s_warp[warp] + inc - v;
    __syncthreads();
    if (threadIdx.x == 0) s_carry = carry + s_warp[32];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = s_carry;
}

template <int DT>
__global__ void __launch_bounds__(kConvThreads) k_encode_write(const void* dense, long long n,
                                                               const int32_t* block_off,
                                                               int2* runs) {
  __shared__ int32_t s_warp[kConvThreads / 32 + 1];
  const long long base = (long long)blockIdx.x * kConvChunk;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int32_t running = block_off[blockIdx.x];
  for (int it = 0; it < kConvItems; ++it) {
    long long i = base + it * kConvThreads + threadIdx.x;
    bool b = false;
    uint32_t v = 0;
    if (i < n) {
      v = load_payload<DT>(dense, i);
      b = (i == n - 1) || typed_differs<DT>(v, load_payload<DT>(dense, i + 1));
    }
    unsigned bal = __ballot_sync(0xffffffffu, b);
    int32_t wcount = __popc(bal);
    int32_t wpos = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) s_warp[warp] = wcount;
    __syncthreads();
    int32_t woff = 0, tot = 0;
    for (int w = 0; w < kConvThreads / 32; ++w) {
      int32_t cw = s_warp[w];
      if (w < warp) woff += cw;
      tot += cw;
    }
    if (b) runs[running + woff + wpos] = make_int2((int32_t)(i + 1), (int32_t)v);
    running += tot;
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// Streaming forms for 4-byte element types and 16-byte aligned dense buffers (the default;
// psk_set_vector_conversion(0) selects the scalar kernels above, which also serve the 1-byte
// types).  Every thread moves 16 bytes per access; a vector's right neighbour comes from the
// next lane by shuffle.

// elements [i, i + 4) of a 4-byte dense buffer (i a multiple of 4; elements beyond n read 0)
PSK_D uint4 load_vec4(const void* dense, long long i, long long n) {
  const uint32_t* p = reinterpret_cast<const uint32_t*>(dense);
  if (i + 3 < n) return *reinterpret_cast<const uint4*>(p + i);
  uint4 r = make_uint4(0, 0, 0, 0);
  if (i < n) r.x = p[i];
  if (i + 1 < n) r.y = p[i + 1];
  if (i + 2 < n) r.z = p[i + 2];
  return r;
}

// bit j set: element i + j exists and closes a run (last element, or differs from i + j + 1)
// (edge: the element behind the warp's 32 vectors, loaded by lane 31 together with the vectors)
template <int DT>
PSK_D unsigned close_flags4_edge(long long i, long long n, const uint4& v, uint32_t edge) {
  uint32_t nx = __shfl_down_sync(0xffffffffu, v.x, 1);
  if ((threadIdx.x & 31) == 31) nx = edge;
  unsigned f = 0;
  if (i < n && (i == n - 1 || typed_differs<DT>(v.x, v.y))) f |= 1u;
  if (i + 1 < n && (i + 1 == n - 1 || typed_differs<DT>(v.y, v.z))) f |= 2u;
  if (i + 2 < n && (i + 2 == n - 1 || typed_differs<DT>(v.z, v.w))) f |= 4u;
  if (i + 3 < n && (i + 3 == n - 1 || typed_differs<DT>(v.w, nx))) f |= 8u;
  return f;
}


// ---------------------------------------------------------------------------------------
// One-pass encode (4-byte elements, 16-byte aligned buffer): the dense buffer is read ONCE.
// Block b encodes chunk b (kEncChunk elements): it counts the runs it closes, gets its output
// offset by a decoupled look-back over the chunks in front of it (one warp polls a window of
// 32 predecessors; status and count travel in one 64-bit word) and writes its runs.  Chunks are
// taken in blockIdx order -- blocks of a 1-D grid are dispatched in that order, so every
// predecessor of a running block is running or done (the assumption of CUB's DeviceScan).
// The output is allocated for `capacity` runs; runs beyond it are counted but not written and
// the caller repeats the pass with the exact size (dense data only).
constexpr unsigned long long kEncAggregate = 1ull, kEncPrefix = 2ull;
#ifndef PSK_ENC_LOOK
#define PSK_ENC_LOOK 8
#endif
#ifndef PSK_ENC_ITEMS
#define PSK_ENC_ITEMS 8
#endif
#ifndef PSK_ENC_MINBLOCKS
#define PSK_ENC_MINBLOCKS 2
#endif
constexpr int kEncLook = PSK_ENC_LOOK;                                 // look-back: kEncLook windows of 32 chunks per round
constexpr int kEncStride = 4;                               // 64-bit words per look-back entry (one 32-byte sector each)
#ifndef PSK_ENC_THREADS
#define PSK_ENC_THREADS 512
#endif
constexpr int kEncThreads = PSK_ENC_THREADS;
constexpr int kEncItems = PSK_ENC_ITEMS;                                // 16-byte vectors per thread
constexpr int kEncChunk = kEncThreads * 4 * kEncItems;     // dense elements per block

// bit j set: element j of the vector closes a run; `full`: the vector and its right neighbour exist
template <int DT>
PSK_D unsigned close_flags4_fast(const uint4& v, uint32_t edge) {
  uint32_t nx = __shfl_down_sync(0xffffffffu, v.x, 1);
  if ((threadIdx.x & 31) == 31) nx = edge;
  return (typed_differs<DT>(v.x, v.y) ? 1u : 0u) | (typed_differs<DT>(v.y, v.z) ? 2u : 0u) |
         (typed_differs<DT>(v.z, v.w) ? 4u : 0u) | (typed_differs<DT>(v.w, nx) ? 8u : 0u);
}

template <int DT>
__global__ void __launch_bounds__(kEncThreads, PSK_ENC_MINBLOCKS) k_encode_onepass(const void* dense, long long n, int32_t nb,
                                                                    unsigned long long* state, int32_t* total_out,
                                                                    int2* runs, int32_t capacity) {
  constexpr int W = kEncThreads / 32;
  __shared__ int32_t s_tot[kEncItems * W + 1];  // runs closed by warp w in item it; then their exclusive offsets
  __shared__ int32_t s_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int32_t tile = (int32_t)blockIdx.x;
  const long long base = (long long)tile * kEncChunk;
  const bool interior = base + kEncChunk + 4 <= n;  // no bounds tests needed (and no last element here)
  // Only the flags stay in registers: the payloads of the (few) elements that close a run are
  // read again from the cache when the runs are written, which keeps enough blocks resident
  // to cover the look-back with other blocks' loads.
  unsigned f[kEncItems];
  int32_t before[kEncItems];  // runs closed by lower lanes of this warp in the same item
  const uint32_t* const p32 = reinterpret_cast<const uint32_t*>(dense);
  {
    uint4 v[kEncItems];
    uint32_t edge[kEncItems];  // lane 31: the element behind the warp's vectors (all loads issued up front)
    if (interior) {
#pragma unroll
      for (int it = 0; it < kEncItems; ++it) {
        const long long i = base + ((long long)it * kEncThreads + threadIdx.x) * 4;
        v[it] = *reinterpret_cast<const uint4*>(p32 + i);
        edge[it] = lane == 31 ? p32[i + 4] : 0u;
      }
#pragma unroll
      for (int it = 0; it < kEncItems; ++it) f[it] = close_flags4_fast<DT>(v[it], edge[it]);
    } else {
#pragma unroll
      for (int it = 0; it < kEncItems; ++it) {
        const long long i = base + ((long long)it * kEncThreads + threadIdx.x) * 4;
        v[it] = load_vec4(dense, i, n);
        edge[it] = (lane == 31 && i + 4 < n) ? p32[i + 4] : 0u;
      }
#pragma unroll
      for (int it = 0; it < kEncItems; ++it) {
        const long long i = base + ((long long)it * kEncThreads + threadIdx.x) * 4;
        f[it] = close_flags4_edge<DT>(i, n, v[it], edge[it]);
      }
    }
  }
#pragma unroll
  for (int it = 0; it < kEncItems; ++it) {
    const int32_t c = __popc(f[it]);
    int32_t inc = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int32_t o = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += o;
    }
    before[it] = inc - c;
    if (lane == 31) s_tot[it * W + warp] = inc;
  }
  __syncthreads();
  if (warp == 0) {
    // exclusive offsets of the (item, warp) pieces in output order, the chunk's run count, the look-back
    constexpr int kPieces = kEncItems * W, kPer = (kPieces + 31) / 32;  // consecutive pieces per lane
    int32_t c[kPer], sum = 0;
#pragma unroll
    for (int u = 0; u < kPer; ++u) {
      c[u] = lane * kPer + u < kPieces ? s_tot[lane * kPer + u] : 0;
      sum += c[u];
    }
    int32_t inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int32_t o = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += o;
    }
    const int32_t total = __shfl_sync(0xffffffffu, inc, 31);
    int32_t run = inc - sum;
#pragma unroll
    for (int u = 0; u < kPer; ++u) {
      if (lane * kPer + u < kPieces) s_tot[lane * kPer + u] = run;
      run += c[u];
    }
    int32_t excl = 0;
    if (tile == 0) {
      if (lane == 0) st_volatile_u64(&state[0], (kEncPrefix << 32) | (uint32_t)total);
    } else {
      if (lane == 0) st_volatile_u64(&state[(long long)tile * kEncStride], (kEncAggregate << 32) | (uint32_t)total);
      // kEncLook windows of 32 predecessors are requested together (one memory round trip) and
      // folded from the nearest one outwards; a window with a chunk that has not published yet is
      // polled again on its own.  The rate at which chunks can resolve their offsets is about
      // (window * chunk bytes) per round trip, so the window has to be wide: the blocks of a wave
      // start together and the nearest inclusive prefix is often a whole wave away.
      int32_t idx = tile - 1;
      bool done = false;
      while (!done) {
        unsigned long long w[kEncLook];
#pragma unroll
        for (int u = 0; u < kEncLook; ++u) {
          const int32_t my = idx - (u * 32 + lane);
          w[u] = my >= 0 ? ld_volatile_u64(&state[(long long)my * kEncStride]) : (kEncPrefix << 32);
        }
#pragma unroll
        for (int u = 0; u < kEncLook; ++u) {
          if (done) continue;
          while (true) {
            const unsigned st = (unsigned)(w[u] >> 32);
            const unsigned empty = __ballot_sync(0xffffffffu, st == 0u);
            const unsigned pre = __ballot_sync(0xffffffffu, st == (unsigned)kEncPrefix);
            const int first_pre = pre ? __ffs((int)pre) - 1 : 32;  // nearest chunk with an inclusive prefix
            const unsigned need = first_pre >= 31 ? 0xffffffffu : ((2u << first_pre) - 1u);
            if ((empty & need) == 0u) {
              int32_t part = lane <= first_pre ? (int32_t)(uint32_t)w[u] : 0;
#pragma unroll
              for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
              excl += part;
              if (first_pre < 32) done = true;
              break;
            }
            sleep_ns(64);  // a chunk in between has not published yet
            const int32_t my = idx - (u * 32 + lane);
            w[u] = my >= 0 ? ld_volatile_u64(&state[(long long)my * kEncStride]) : (kEncPrefix << 32);
          }
        }
        idx -= kEncLook * 32;
      }
      if (lane == 0) st_volatile_u64(&state[(long long)tile * kEncStride], (kEncPrefix << 32) | (uint32_t)(excl + total));
    }
    if (lane == 0) {
      s_base = excl;
      if (tile == nb - 1) *total_out = excl + total;
    }
  }
  __syncthreads();
  // output order = element order: items, then warps, then lanes, then the 4 elements
  const int32_t pos = s_base;
#pragma unroll
  for (int it = 0; it < kEncItems; ++it) {
    int32_t mine = pos + s_tot[it * W + warp] + before[it];
    const long long i = base + ((long long)it * kEncThreads + threadIdx.x) * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (f[it] & (1u << j)) {
        if (mine < capacity) runs[mine] = make_int2((int32_t)(i + j) + 1, (int32_t)p32[i + j]);
        ++mine;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Streaming expand (4-byte elements, 16-byte aligned buffer), two kernels:
//   k_expand_table   first run of every chunk of kExpChunk elements (one search per chunk, all
//                    chunks at once);
//   k_expand_chunks  a persistent grid walks the chunks in grid-stride order (the chunks being
//                    written at any time are neighbours in memory): stage the chunk's runs, find
//                    the run of every 4-element group in shared memory, store 16 bytes per thread.
constexpr int kExpChunk = kConvThreads * 8;  // dense elements per chunk (a chunk has at most kExpChunk + 1 runs)

__global__ void k_expand_table(const int2* runs, int32_t nruns, int32_t nchunks, int32_t* table /*[nchunks + 1]*/) {
  const int32_t c = (int32_t)(blockIdx.x * blockDim.x + threadIdx.x);
  if (c > nchunks) return;
  if (c == nchunks) {
    table[c] = nruns - 1;
    return;
  }
  // first run whose end is > c * kExpChunk (core::Store::index, core.hpp:49-52)
  const int32_t pos = c * kExpChunk;
  int32_t lo = 0, hi = nruns;
  while (lo < hi) {
    const int32_t mid = lo + ((hi - lo) >> 1);
    if (ld_nc(runs + mid).x <= pos) lo = mid + 1; else hi = mid;
  }
  table[c] = lo;
}

template <int DT>
__global__ void __launch_bounds__(kConvThreads) k_expand_chunks(const int2* runs, int32_t nruns, long long n,
                                                                const int32_t* table, int32_t nchunks, void* dense) {
  constexpr int kWords = kExpChunk / 32;
  __shared__ uint32_t s_vals[kExpChunk + 2];
  __shared__ uint32_t s_bits[kWords];   // bit x: element b0 + x is the last one of its run
  __shared__ int32_t s_pref[kWords];    // runs that end in front of word w
  uint32_t* const out = reinterpret_cast<uint32_t*>(dense);
  int32_t c = (int32_t)blockIdx.x;
  if (c >= nchunks) return;
  int32_t r0 = ld_nc(table + c), r1 = ld_nc(table + c + 1);
  for (; c < nchunks; c += (int32_t)gridDim.x) {
    const long long b0 = (long long)c * kExpChunk;
    const long long b1 = b0 + kExpChunk < n ? b0 + kExpChunk : n;
    if (threadIdx.x < kWords) s_bits[threadIdx.x] = 0u;
    __syncthreads();
    // runs r0 .. r1 cover the chunk (r1 = first run of the next chunk; it may reach into this one)
    const int32_t nr = r1 - r0 + 1;
    for (int32_t j = threadIdx.x; j < nr; j += kConvThreads) {
      const int2 pr = ld_nc(runs + r0 + j);
      s_vals[j] = (uint32_t)pr.y;
      const long long last = (long long)pr.x - 1 - b0;  // the run's last element, relative to the chunk
      if (last < (long long)kExpChunk) atomicOr(&s_bits[last >> 5], 1u << (last & 31));
    }
    // the next chunk's table entries, in flight while this chunk is written
    const int32_t cn = c + (int32_t)gridDim.x;
    int32_t nr0 = 0, nr1 = 0;
    if (cn < nchunks) {
      nr0 = ld_nc(table + cn);
      nr1 = ld_nc(table + cn + 1);
    }
    __syncthreads();
    if (threadIdx.x < 32) {  // exclusive prefix of the words' bit counts (kWords = 64: two per lane)
      static_assert(kWords == 64, "two words per lane");
      const int32_t c0 = __popc(s_bits[2 * threadIdx.x]), c1 = __popc(s_bits[2 * threadIdx.x + 1]);
      int32_t inc = c0 + c1;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int32_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if ((int)threadIdx.x >= d) inc += o;
      }
      s_pref[2 * threadIdx.x] = inc - c0 - c1;
      s_pref[2 * threadIdx.x + 1] = inc - c1;
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < kExpChunk / (kConvThreads * 4); ++it) {
      const int32_t x = (it * kConvThreads + (int)threadIdx.x) * 4;  // offset in the chunk
      const long long i = b0 + x;
      if (i >= b1) continue;
      // run of element x (relative to r0) = number of runs that end in front of it
      const uint32_t word = s_bits[x >> 5];
      const int bit = x & 31;
      int32_t k0 = s_pref[x >> 5] + __popc(word & ((1u << bit) - 1u));
      const uint32_t nib = (word >> bit) & 7u;  // elements x, x+1, x+2 that end a run
      uint4 o;
      o.x = s_vals[k0];
      if (nib == 0u) {
        o.y = o.z = o.w = o.x;
      } else {
        k0 += nib & 1u;
        o.y = s_vals[k0];
        k0 += (nib >> 1) & 1u;
        o.z = s_vals[k0];
        k0 += (nib >> 2) & 1u;
        o.w = s_vals[k0];
      }
      if (i + 3 < b1) {
        *reinterpret_cast<uint4*>(out + i) = o;
      } else {
        const uint32_t xs[4] = {o.x, o.y, o.z, o.w};
        for (int j = 0; j < 4 && i + j < b1; ++j) out[i + j] = xs[j];
      }
    }
    r0 = nr0;
    r1 = nr1;
    // (the next iteration's first barrier separates these reads from its writes of s_vals / s_pref;
    //  s_bits is cleared before that barrier, so it needs its own)
    __syncthreads();
  }
}

// expand: each block writes kConvChunk consecutive dense elements
template <int DT>
__global__ void __launch_bounds__(kConvThreads) k_expand(const int2* runs,
                                                         int32_t nruns, long long n, void* dense) {
  __shared__ int32_t s_ends[kConvChunk + 1];
  __shared__ int32_t s_r0, s_nr;
  const long long b0 = (long long)blockIdx.x * kConvChunk;
  long long b1 = b0 + kConvChunk;
  if (b1 > n) b1 = n;
  if (threadIdx.x < 2) {
    // first run whose end is > pos (core::Store::index, core.hpp:49-52)
    int32_t pos = threadIdx.x == 0 ? (int32_t)b0 : (int32_t)(b1 - 1);
    int32_t lo = 0, hi = nruns;
    while (lo < hi) {
      int32_t mid = lo + ((hi - lo) >> 1);
      if (runs[mid].x <= pos) lo = mid + 1; else hi = mid;
    }
    if (threadIdx.x == 0) s_r0 = lo; else s_nr = lo;
  }
  __syncthreads();
  const int32_t r0 = s_r0;
  const int32_t nr = s_nr - r0 + 1;  // <= elements in the chunk
  for (int32_t j = threadIdx.x; j < nr; j += kConvThreads) s_ends[j] = runs[r0 + j].x;
  __syncthreads();
  for (int it = 0; it < kConvItems; ++it) {
    long long i = b0 + it * kConvThreads + threadIdx.x;
    if (i < b1) {
      int32_t lo = 0, hi = nr - 1;
      while (lo < hi) {
        int32_t mid = (lo + hi) >> 1;
        if (s_ends[mid] <= (int32_t)i) lo = mid + 1; else hi = mid;
      }
      store_payload<DT>(dense, i, (uint32_t)runs[r0 + lo].y);
    }
  }
}

// position of the j-th selected element of a strided box
PSK_HD uint32_t sel_pos(const DView& v, uint32_t j) {
  uint32_t flat = 0, rem = j;
  for (int d = 0; d < v.ndim; ++d) {
    uint32_t q = rem % v.cnt[d];
    rem /= v.cnt[d];
    flat += (v.c0[d] + q * v.cs[d]) * v.scale[d];
  }
  return flat;
}

// raw run list of a box mask: per segment (a dim-0 row when stride_0 == 1, else one
// element) an excluded run ending at its start and an included run ending at its end
__global__ void k_mask_segments(DView v, uint32_t seg_len, int32_t nseg, int2* runs) {
  int32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < nseg) {
    uint32_t start = sel_pos(v, (uint32_t)s * seg_len);
    runs[2 * s] = make_int2((int32_t)start, 0);
    runs[2 * s + 1] = make_int2((int32_t)(start + seg_len), 1);
  }
  if (s == 0) {
    runs[2 * nseg] = make_int2((int32_t)v.n, 0);
    runs[2 * nseg + 1] = make_int2(0x7fffffff, 0);  // sentinel pairs behind the last run
    runs[2 * nseg + 2] = make_int2(0x7fffffff, 0);
  }
}

// ---- reductions over runs ---------------------------------------------------------------
constexpr int kRedThreads = 256;
constexpr int kRedBlocks = 1184;  // 8 x 148 SMs

struct RedAcc {
  uint32_t isum;  // int32 sum / product, wraps mod 2^32
  double fsum;    // float sum / product in fp64
  uint32_t mx, mn;
  int32_t any;
};

PSK_D uint32_t ipow_wrap(uint32_t b, uint32_t e) {
  uint32_t r = 1;
  while (e) {
    if (e & 1) r *= b;
    b *= b;
    e >>= 1;
  }
  return r;
}

template <bool IS_F>
PSK_D void red_merge(RedAcc& a, const RedAcc& b, int op) {
  if (!b.any) return;
  if (!a.any) { a = b; return; }
  if (op == PSK_RED_PROD) { a.isum *= b.isum; a.fsum *= b.fsum; }
  else { a.isum += b.isum; a.fsum += b.fsum; }
  if (IS_F) {
    if (u2f(a.mx) < u2f(b.mx)) a.mx = b.mx;
    if (u2f(b.mn) < u2f(a.mn)) a.mn = b.mn;
  } else {
    if ((int32_t)a.mx < (int32_t)b.mx) a.mx = b.mx;
    if ((int32_t)b.mn < (int32_t)a.mn) a.mn = b.mn;
  }
}

template <bool IS_F>
__global__ void __launch_bounds__(kRedThreads) k_reduce_partial(const int2* runs, int32_t nruns,
                                                                const int32_t* d_count, int op, RedAcc* partial) {
  __shared__ RedAcc s_acc[kRedThreads];
  RedAcc acc;
  acc.any = 0; acc.isum = 0; acc.fsum = 0; acc.mx = 0; acc.mn = 0;
  if (d_count) nruns = *d_count;  // a queued result: its run count is still on the device
  for (int32_t i = blockIdx.x * kRedThreads + threadIdx.x; i < nruns; i += gridDim.x * kRedThreads) {
    const int2 r_ = runs[i];
    uint32_t len = (uint32_t)(r_.x - (i ? runs[i - 1].x : 0));
    uint32_t v = (uint32_t)r_.y;
    RedAcc r;
    r.any = 1; r.mx = v; r.mn = v;
    if (op == PSK_RED_PROD) {
      r.isum = ipow_wrap(v, len);
      r.fsum = IS_F ? pow((double)u2f(v), (double)len) : 0.0;
    } else {
      r.isum = v * len;
      r.fsum = IS_F ? (double)u2f(v) * (double)len : 0.0;
    }
    red_merge<IS_F>(acc, r, op);
  }
  s_acc[threadIdx.x] = acc;
  __syncthreads();
  for (int d = kRedThreads / 2; d > 0; d >>= 1) {
    if ((int)threadIdx.x < d) {
      RedAcc a = s_acc[threadIdx.x];
      red_merge<IS_F>(a, s_acc[threadIdx.x + d], op);
      s_acc[threadIdx.x] = a;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = s_acc[0];
}

template <bool IS_F>
__global__ void __launch_bounds__(kRedThreads) k_reduce_final(const RedAcc* partial, int32_t nparts, int op,
                                                              RedAcc* out) {
  __shared__ RedAcc s_acc[kRedThreads];
  RedAcc acc;
  acc.any = 0; acc.isum = 0; acc.fsum = 0; acc.mx = 0; acc.mn = 0;
  // fixed association order -> bit-reproducible fp64 sums
  for (int32_t i = threadIdx.x; i < nparts; i += kRedThreads) red_merge<IS_F>(acc, partial[i], op);
  s_acc[threadIdx.x] = acc;
  __syncthreads();
  for (int d = kRedThreads / 2; d > 0; d >>= 1) {
    if ((int)threadIdx.x < d) {
      RedAcc a = s_acc[threadIdx.x];
      red_merge<IS_F>(a, s_acc[threadIdx.x + d], op);
      s_acc[threadIdx.x] = a;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = s_acc[0];
}


// ---- segmented reductions over runs (src/pyskip/reduce.py:8-33 with axis = 0) -------------
// Segment j = flat coordinates [j L, (j + 1) L): the reduction of the fastest axis (or of all
// inner axes at once) of a tensor.  One warp per segment: lane 0/1 find the first and last
// run of the segment, the lanes stride over those runs with the piece of every run that lies
// inside the segment (sum(val * len) with int32 wrap-around or fp64 accumulation, max, min,
// prod(val^len)), a shuffle tree in fixed order folds the lanes.  One pass over the runs
// (+ one search per segment) instead of log2(L) strided two-source evaluations.
constexpr int kSegThreads = 256;

template <bool IS_F>
__global__ void __launch_bounds__(kSegThreads) k_reduce_segments(const int2* runs, int32_t nruns,
                                                                 const int32_t* d_count, int32_t seg_len,
                                                                 int32_t nseg, int op, uint32_t* out) {
  const int lane = threadIdx.x & 31;
  const int32_t j = (int32_t)((blockIdx.x * (long long)kSegThreads + threadIdx.x) >> 5);
  if (j >= nseg) return;
  if (d_count) nruns = ld_nc(d_count);
  const int32_t s0 = j * seg_len, s1 = s0 + seg_len;
  // first run whose end is > s0 (lane 0) / first run whose end is >= s1 (lane 1)
  int32_t lo = 0, hi = nruns;
  const int32_t key = lane == 0 ? s0 : s1 - 1;
  if (lane < 2) {
    while (lo < hi) {
      const int32_t mid = lo + ((hi - lo) >> 1);
      if (ld_nc(runs + mid).x <= key) lo = mid + 1; else hi = mid;
    }
  }
  const int32_t r0 = __shfl_sync(0xffffffffu, lo, 0), r1 = __shfl_sync(0xffffffffu, lo, 1);
  RedAcc acc;
  acc.any = 0; acc.isum = 0; acc.fsum = 0; acc.mx = 0; acc.mn = 0;
  for (int32_t i = r0 + lane; i <= r1; i += 32) {
    const int2 r_ = ld_nc(runs + i);
    const int32_t b = i > r0 ? ld_nc(runs + i - 1).x : s0;  // piece [max(begin, s0), min(end, s1))
    const uint32_t len = (uint32_t)(imin(r_.x, s1) - imax(b, s0));
    const uint32_t v = (uint32_t)r_.y;
    RedAcc r;
    r.any = 1; r.mx = v; r.mn = v;
    if (op == PSK_RED_PROD) {
      r.isum = ipow_wrap(v, len);
      r.fsum = IS_F ? pow((double)u2f(v), (double)len) : 0.0;
    } else {
      r.isum = v * len;
      r.fsum = IS_F ? (double)u2f(v) * (double)len : 0.0;
    }
    red_merge<IS_F>(acc, r, op);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    RedAcc o;
    o.any = __shfl_down_sync(0xffffffffu, acc.any, d);
    o.isum = __shfl_down_sync(0xffffffffu, acc.isum, d);
    o.fsum = __shfl_down_sync(0xffffffffu, acc.fsum, d);
    o.mx = __shfl_down_sync(0xffffffffu, acc.mx, d);
    o.mn = __shfl_down_sync(0xffffffffu, acc.mn, d);
    if (lane + d < 32) red_merge<IS_F>(acc, o, op);
  }
  if (lane == 0) {
    uint32_t res;
    if (op == PSK_RED_MAX) res = acc.mx;
    else if (op == PSK_RED_MIN) res = acc.mn;
    else if (IS_F) res = f2u((float)acc.fsum);
    else res = acc.isum;
    out[j] = res;
  }
}

}  // namespace psk
