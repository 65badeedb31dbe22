// Shared device helpers for the fgnn sm_100a kernels: Philox4x32-10, the
// chunk-chained single-pass scan, cache-hinted loads/stores, launch helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fgnn_kernels.h"

namespace fgnn {

constexpr uint32_t kEmpty = 0xFFFFFFFFu;
constexpr int kBlock = 256;          // threads per CTA for all tiled kernels
constexpr int kMaxChainCtas = 4096;  // FGNN_CHAIN_WS_BYTES = 16 + 8*4096
// Tickets up to which chain_scan sums the lower aggregates directly instead of looking back.  Measured on B200
// (r1_q c5): with 688-888 tickets the direct sum is SLOWER than the look-back (ht_compact 24 -> 39 us: every
// CTA polls all lower tickets), so it is only used for tiny grids where the look-back has nothing to amortise.
constexpr unsigned kChainDirectMax = 32;

// ---------------------------------------------------------------------------
// launch bookkeeping
// ---------------------------------------------------------------------------
extern unsigned long long g_launch_count;  // api.cu
int sm_count();                            // api.cu (cached per device)
inline void note_launch(int n = 1) { g_launch_count += (unsigned long long)n; }

// Event trace of the launches (profiling aid, off by default; api.cu).  With fgnn_k_trace_enable(1) every C-ABI
// entry that takes part in a mini-batch records a CUDA event after each of its launches; fgnn_k_trace_dump turns
// them into (label, stream, milliseconds since enable) so that the timeline of an OVERLAPPED loop — which kernel
// ran when on which stream — can be read without nsys (not in the image).  Labels: FGNN_TRACE_* in the header.
extern bool g_trace_on;
void trace_mark_slow(cudaStream_t st, int label);
inline void trace_mark(cudaStream_t st, int label) {
  if (g_trace_on) trace_mark_slow(st, label);
}

inline int check_last() {
  cudaError_t e = cudaGetLastError();
  return (int)e;
}

// resident CTAs per SM for a kernel (cached by the caller in a static)
template <typename Kern>
inline int occupancy(Kern kern, int block, size_t smem) {
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, block, smem) != cudaSuccess || occ < 1)
    occ = 1;
  return occ;
}

// persistent grid: enough CTAs to cover n_max at `min_items` per CTA, capped
// at sm_count * ctas_per_sm (and the chain limit when chained).
int grid_share_div();  // api.cu: FGNN_GRID_DIV (>= 1), read once
inline int persistent_grid(uint64_t n_max, uint32_t min_items, int ctas_per_sm,
                           bool chained, bool whole_gpu = false) {
  uint64_t want = (n_max + min_items - 1) / min_items;
  uint64_t cap = (uint64_t)sm_count() * (uint64_t)ctas_per_sm;
  // sampling-side kernels of several slots run next to each other on different streams: with
  // FGNN_GRID_DIV = d each launch takes at most 1/d of the resident CTA slots, so d launches co-reside
  if (!whole_gpu) cap = (cap + grid_share_div() - 1) / grid_share_div();
  if (chained && cap > (uint64_t)kMaxChainCtas) cap = kMaxChainCtas;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}

// ---------------------------------------------------------------------------
// Philox4x32-10 — identical stream layout to oracle/fgnn_oracle.c
//   key = (seed.lo, seed.hi ^ batch_key.hi)
//   ctr = (draw >> 2, item, tag, batch_key.lo); word = draw & 3
// ---------------------------------------------------------------------------
struct RngKey {
  uint32_t k0, k1, c2, c3;
};
__host__ __device__ inline RngKey make_rng_key(const fgnn_rng &r) {
  RngKey k;
  k.k0 = (uint32_t)r.seed;
  k.k1 = (uint32_t)(r.seed >> 32) ^ (uint32_t)(r.batch_key >> 32);
  k.c2 = r.tag;
  k.c3 = (uint32_t)r.batch_key;
  return k;
}

__device__ __forceinline__ uint4 philox_block(const RngKey &key, uint32_t item,
                                              uint32_t block) {
  uint32_t c0 = block, c1 = item, c2 = key.c2, c3 = key.c3;
  uint32_t k0 = key.k0, k1 = key.k1;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0;
    const uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ uint32_t pick_word(const uint4 &b, uint32_t w) {
  return w == 0 ? b.x : (w == 1 ? b.y : (w == 2 ? b.z : b.w));
}

__device__ __forceinline__ uint32_t rand_u32(const RngKey &key, uint32_t item,
                                             uint32_t draw) {
  return pick_word(philox_block(key, item, draw >> 2), draw & 3u);
}

// curand_uniform.h:69-72 (float in (0,1])
__device__ __forceinline__ float uniform_f32(uint32_t x) {
  return __fmaf_rn((float)x, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
}
// curand_uniform.h:101-106
__device__ __forceinline__ double uniform_f64(uint32_t x, uint32_t y) {
  unsigned long long z = (unsigned long long)x ^ ((unsigned long long)y << 21);
  return (double)z * 1.1102230246251565e-16 + (1.1102230246251565e-16 / 2.0);
}

// ---------------------------------------------------------------------------
// cache-hinted memory ops
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ldg_u32(const uint32_t *p) { return __ldg(p); }

__device__ __forceinline__ uint4 ld_nc_na_v4(const void *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_na_v4(void *p, const uint4 &v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(
    const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p,
                                               unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---------------------------------------------------------------------------
// block primitives (blockDim.x == kBlock)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, d);
    if ((threadIdx.x & 31) >= d) v += t;
  }
  return v;
}

// exclusive scan over the block; returns this thread's exclusive prefix and
// the block total through *total.  `s_warp` = NT/32 + 1 words of smem.
template <int NT = kBlock>
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *s_warp,
                                                    uint32_t *total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = warp_incl_scan(v);
  __syncthreads();  // protect s_warp reuse
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < (NT / 32) ? s_warp[lane] : 0u;
    uint32_t wi = warp_incl_scan(w);
    if (lane < (NT / 32)) s_warp[lane] = wi - w;
    if (lane == (NT / 32) - 1) s_warp[NT / 32] = wi;
  }
  __syncthreads();
  *total = s_warp[NT / 32];
  return s_warp[warp] + incl - v;
}

template <int NT = kBlock>
__device__ __forceinline__ unsigned long long block_sum_u64(
    unsigned long long v, unsigned long long *s_warp64) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, d);
  __syncthreads();
  if (lane == 0) s_warp64[warp] = v;
  __syncthreads();
  unsigned long long t = 0;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) t += s_warp64[w];
  return t;
}

// ---------------------------------------------------------------------------
// chunk-chained single-pass scan (decoupled look-back).
//
// A persistent grid of P <= kMaxChainCtas CTAs splits [0,n) into P contiguous
// chunks in *ticket* order.  Each CTA publishes the aggregate of its chunk, then
// warp 0 looks back over the lower tickets 32 at a time, summing aggregates until
// it meets a ticket that has already published its inclusive prefix, and finally
// publishes its own inclusive prefix.  Lower tickets belong to CTAs that started
// earlier and publish before they wait, so this cannot deadlock even when P
// exceeds residency.  Only ONE warp per CTA ever polls, with a nanosleep
// back-off: the first version let every thread spin on every lower ticket, which
// is harmless when the kernel has the GPU to itself but saturated the L2 request
// path once several sampling slots ran next to the HBM-bound gather (r1_g/h:
// 0.3 ms -> 1-4 ms per step).  The last CTA to leave re-zeroes the workspace, so
// the same buffer serves the next launch on the stream.
// ---------------------------------------------------------------------------
struct ChainWs {
  unsigned int ticket;
  unsigned int done;
  unsigned int pad[2];
  unsigned long long agg[kMaxChainCtas];  // [63:62] 0 = empty, 1 = aggregate, 2 = inclusive prefix
};
static_assert(sizeof(ChainWs) == FGNN_CHAIN_WS_BYTES, "chain ws size");

struct ChainSmem {
  uint32_t ticket;
  uint32_t last;
  unsigned long long excl;
  unsigned long long warp64[kBlock / 32];
};

__device__ __forceinline__ uint32_t chain_ticket(ChainWs *ws, ChainSmem *sm) {
  if (threadIdx.x == 0) sm->ticket = atomicAdd(&ws->ticket, 1u);
  __syncthreads();
  return sm->ticket;
}

// all threads pass their partial; returns exclusive prefix over lower tickets
// and publishes this chunk's aggregate / inclusive prefix.
template <int NT = kBlock>
__device__ __forceinline__ unsigned long long chain_scan(
    ChainWs *ws, ChainSmem *sm, uint32_t p, unsigned long long thread_partial,
    unsigned long long *chunk_total) {
  constexpr unsigned long long kAgg = 1ull << 62, kPre = 2ull << 62, kVal = (1ull << 62) - 1;
  const unsigned long long total = block_sum_u64<NT>(thread_partial, sm->warp64);
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    unsigned long long excl = 0;
    if (gridDim.x <= kChainDirectMax) {
      // Direct sum: every ticket publishes its aggregate right after its own first pass and nobody waits for a
      // PREFIX, so the scan has no serial chain at all: warp 0 loads the (<= 1024) lower aggregates, 32 at a
      // time, all independent.  The look-back below costs ~P/32 dependent L2 round trips (ncu r1_q: 20-25 us
      // for the 688-888-ticket kernels of a GraphSAGE batch, most of their duration).
      if (lane == 0) st_relaxed_u64(&ws->agg[p], kAgg | total);
      unsigned long long c = 0;
      for (int t = lane; t < (int)p; t += 32) {
        unsigned long long v = ld_relaxed_u64(&ws->agg[t]);
        while (!(v >> 62)) {
          __nanosleep(64);
          v = ld_relaxed_u64(&ws->agg[t]);
        }
        c += v & kVal;
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) c += __shfl_down_sync(0xFFFFFFFFu, c, d);
      excl = c;  // valid in lane 0
    } else {
    if (lane == 0) st_relaxed_u64(&ws->agg[p], (p == 0 ? kPre : kAgg) | total);
    if (p > 0) {
      int newest = (int)p - 1;  // window = tickets newest, newest-1, ..., newest-31 (lane order)
      while (true) {
        const int t = newest - lane;
        unsigned long long v = kPre;  // lanes past ticket 0 read as a zero prefix
        if (t >= 0) {
          v = ld_relaxed_u64(&ws->agg[t]);
          while (!(v >> 62)) {
            __nanosleep(64);
            v = ld_relaxed_u64(&ws->agg[t]);
          }
        }
        const unsigned pre = __ballot_sync(0xFFFFFFFFu, (v >> 62) == 2ull);
        const int stop = pre ? __ffs(pre) - 1 : 31;  // nearest ticket that already holds a prefix
        unsigned long long c = lane <= stop ? (v & kVal) : 0ull;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) c += __shfl_down_sync(0xFFFFFFFFu, c, d);
        excl += c;  // valid in lane 0
        if (pre) break;
        newest -= 32;
      }
      if (lane == 0) st_relaxed_u64(&ws->agg[p], kPre | ((excl + total) & kVal));
    }
    }
    if (lane == 0) sm->excl = excl;
  }
  __syncthreads();
  *chunk_total = total;
  return sm->excl;
}

template <int NT = kBlock>
__device__ __forceinline__ void chain_finish(ChainWs *ws, ChainSmem *sm) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int prev = atomicAdd(&ws->done, 1u);
    sm->last = (prev == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (sm->last) {
    for (uint32_t t = threadIdx.x; t < gridDim.x; t += NT) ws->agg[t] = 0ull;
    if (threadIdx.x == 0) {
      ws->ticket = 0u;
      ws->done = 0u;
    }
  }
}

// contiguous chunk of ticket p out of P over n items, tile-aligned
__device__ __forceinline__ void chunk_range(uint32_t n, uint32_t p, uint32_t P,
                                            uint32_t tile, uint32_t *begin,
                                            uint32_t *end) {
  const uint32_t tiles = (n + tile - 1) / tile;
  const uint32_t per = (tiles + P - 1) / P;
  const unsigned long long b = (unsigned long long)p * per * tile;
  const unsigned long long e = b + (unsigned long long)per * tile;
  *begin = b < n ? (uint32_t)b : n;
  *end = e < n ? (uint32_t)e : n;
}

__device__ __forceinline__ uint32_t load_count(uint32_t n_max, const uint32_t *d_n) {
  if (d_n == nullptr) return n_max;
  const uint32_t v = __ldg(d_n);
  return v < n_max ? v : n_max;
}

}  // namespace fgnn
