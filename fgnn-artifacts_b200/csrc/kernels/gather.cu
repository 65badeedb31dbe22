// Feature / label extraction kernels.
//
// fgnn_k_row_copy      : dst[dst_index?[i]] = src[src_index?[i] & mask]
//                        (GPUExtract cuda_extraction.cu:31-49, combine_miss_data /
//                        combine_cache_data dist_cache_manager_device.cu:37-82,
//                        extract_miss_data dist_cache_manager_host.cc:38-56 when
//                        `src` is pinned host memory read through UVA)
// fgnn_k_gather_cached : the reference's five-step trainer-side extraction
//                        (dist_loops.cc:713-846) as one kernel; cache shards may
//                        live on NVLink peers.
//
// The reference maps one thread to one 4-byte column element.  Here rows are
// cut into 16-byte chunks (or the widest power of two that divides the row and
// the base alignment) and chunks are spread flat over a persistent grid, four
// independent 16-byte loads in flight per thread, streaming cache hints on both
// sides (rows are touched once per batch).
#include "common.cuh"

#include <stdlib.h>
#include <string.h>

#include <algorithm>

namespace fgnn {
namespace {

template <typename V> struct VecIO;
template <> struct VecIO<uint4> {
  static __device__ __forceinline__ uint4 ld(const void *p) { return ld_nc_na_v4(p); }
  static __device__ __forceinline__ void st(void *p, const uint4 &v) { st_na_v4(p, v); }
};
template <> struct VecIO<uint2> {
  static __device__ __forceinline__ uint2 ld(const void *p) { return __ldg((const uint2 *)p); }
  static __device__ __forceinline__ void st(void *p, const uint2 &v) { *(uint2 *)p = v; }
};
template <> struct VecIO<uint32_t> {
  static __device__ __forceinline__ uint32_t ld(const void *p) { return __ldg((const uint32_t *)p); }
  static __device__ __forceinline__ void st(void *p, const uint32_t &v) { *(uint32_t *)p = v; }
};
template <> struct VecIO<uint8_t> {
  static __device__ __forceinline__ uint8_t ld(const void *p) { return __ldg((const uint8_t *)p); }
  static __device__ __forceinline__ void st(void *p, const uint8_t &v) { *(uint8_t *)p = v; }
};

constexpr int kUnroll = 4;

template <typename V>
__global__ void __launch_bounds__(kBlock)
row_copy_kernel(char *__restrict__ dst, const uint32_t *__restrict__ dst_index,
                const char *__restrict__ src, const uint32_t *__restrict__ src_index,
                uint64_t src_mask, uint32_t n_max, const uint32_t *__restrict__ d_n,
                size_t row_bytes, uint32_t cpr /* chunks per row */) {
  const uint32_t n = load_count(n_max, d_n);
  const uint64_t total = (uint64_t)n * cpr;
  const uint64_t stride = (uint64_t)gridDim.x * kBlock;
  for (uint64_t c0 = (uint64_t)blockIdx.x * kBlock + threadIdx.x; c0 < total;
       c0 += stride * kUnroll) {
    V v[kUnroll];
    char *dp[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint64_t c = c0 + u * stride;
      dp[u] = nullptr;
      if (c < total) {
        const uint32_t row = (uint32_t)(c / cpr);
        const uint32_t col = (uint32_t)(c - (uint64_t)row * cpr);
        const uint64_t s = src_index ? ((uint64_t)__ldg(src_index + row) & src_mask) : row;
        const uint64_t d = dst_index ? (uint64_t)__ldg(dst_index + row) : row;
        v[u] = VecIO<V>::ld(src + s * row_bytes + (size_t)col * sizeof(V));
        dp[u] = dst + d * row_bytes + (size_t)col * sizeof(V);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      if (dp[u]) VecIO<V>::st(dp[u], v[u]);
  }
}

// (gather_cached_kernel<V>, the flat variant for rows that are not 16-byte multiples, follows RowSrc below)

// ---------------------------------------------------------------------------
// Warp-group gather (16-byte rows, production path of the LDG family).
//
// ncu r1_b on the flat kernel above: every 16-byte chunk pays the dependent
// chain nodes[row] -> table[node] -> row load, so only a third of the time a
// thread is in flight carries payload.  Here a warp owns a *group* of G
// consecutive output rows: lane l < G resolves row l's source pointer
// (coalesced id load, one table lookup per ROW instead of per chunk), the
// pointers are broadcast by shuffle and all 32 lanes stream the group's chunks,
// 8 x 16 B in flight per thread.  The two index loads are software-pipelined two
// and one groups ahead, so the copy loop never waits on them.
// ---------------------------------------------------------------------------
constexpr int kGroupUnroll = 8;
constexpr uint32_t kGroupChunks = 32 * kGroupUnroll;  // chunks of one pass

// Where a row lives.  Cache slot s (= position in the hotness ranking):
//   s <  num_replicated : on THIS GPU, row s of `replica` (the hottest rows are replicated on every trainer)
//   s >= num_replicated : striped, owner = (s - R) mod T, local row = (s - R) div T of that owner's shard
//                         (own shard: HBM; a peer's: loads over NVLink through the IPC mapping)
//   EMPTY               : the pinned host feature table (UVA, host link)
struct RowSrc {
  const void *const *shards;
  const char *shard0;
  uint32_t num_shards;
  const char *miss_src;
  uint64_t miss_mask;
  size_t row_bytes;
  const char *replica;
  uint32_t num_replicated, self_shard;
  unsigned long long *d_remote;  // optional: rows read from peer shards
  // optional deferred pass for peer rows: defer_cnt[0] = number of listed rows, entries {src pointer, dst row}
  unsigned int *defer_cnt;
  ulonglong2 *defer_list;
  __device__ __forceinline__ const char *resolve(uint32_t node, uint32_t slot) const {
    if (slot != kEmpty) {
      if (slot < num_replicated) return replica + (size_t)slot * row_bytes;
      const uint32_t s = slot - num_replicated;
      if (num_shards == 1) return shard0 + (size_t)s * row_bytes;
      const uint32_t owner = s % num_shards, lrow = s / num_shards;
      return (const char *)__ldg((const unsigned long long *)shards + owner) + (size_t)lrow * row_bytes;
    }
    return miss_src + ((uint64_t)node & miss_mask) * row_bytes;
  }
  __device__ __forceinline__ bool is_remote(uint32_t slot) const {
    return slot != kEmpty && slot >= num_replicated && num_shards > 1 &&
           (slot - num_replicated) % num_shards != self_shard;
  }
  // warp-reduced counters -> d_stats[0] hits, [1] misses; *d_remote rows read from peer shards
  __device__ __forceinline__ void report(unsigned long long *d_stats, uint32_t hits, uint32_t misses,
                                         uint32_t remote) const {
    if (!d_stats && !d_remote) return;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      hits += __shfl_down_sync(0xFFFFFFFFu, hits, d);
      misses += __shfl_down_sync(0xFFFFFFFFu, misses, d);
      remote += __shfl_down_sync(0xFFFFFFFFu, remote, d);
    }
    if ((threadIdx.x & 31) == 0) {
      if (d_stats && hits) atomicAdd(d_stats + 0, (unsigned long long)hits);
      if (d_stats && misses) atomicAdd(d_stats + 1, (unsigned long long)misses);
      if (d_remote && remote) atomicAdd(d_remote, (unsigned long long)remote);
    }
  }
};

template <typename V>
__global__ void __launch_bounds__(kBlock)
gather_cached_kernel(char *__restrict__ out, const uint32_t *__restrict__ nodes, uint32_t n_max,
                     const uint32_t *__restrict__ d_n, const uint32_t *__restrict__ table, RowSrc rs,
                     uint32_t cpr, unsigned long long *d_stats) {
  rs.shard0 = (const char *)__ldg((const unsigned long long *)rs.shards);
  const uint32_t n = load_count(n_max, d_n);
  const uint64_t total = (uint64_t)n * cpr;
  const uint64_t stride = (uint64_t)gridDim.x * kBlock;
  uint32_t hits = 0, misses = 0, remote = 0;
  for (uint64_t c0 = (uint64_t)blockIdx.x * kBlock + threadIdx.x; c0 < total;
       c0 += stride * kUnroll) {
    V v[kUnroll];
    char *dp[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint64_t c = c0 + u * stride;
      dp[u] = nullptr;
      if (c < total) {
        const uint32_t row = (uint32_t)(c / cpr);
        const uint32_t col = (uint32_t)(c - (uint64_t)row * cpr);
        const uint32_t node = __ldg(nodes + row);
        const uint32_t slot = __ldg(table + node);
        const char *sp = rs.resolve(node, slot);
        if (col == 0) {
          if (slot != kEmpty) ++hits; else ++misses;
          if (rs.is_remote(slot)) ++remote;
        }
        v[u] = VecIO<V>::ld(sp + (size_t)col * sizeof(V));
        dp[u] = out + (size_t)row * rs.row_bytes + (size_t)col * sizeof(V);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      if (dp[u]) VecIO<V>::st(dp[u], v[u]);
  }
  rs.report(d_stats, hits, misses, remote);
}

__device__ __forceinline__ const char *shfl_ptr(const char *p, int src) {
  unsigned long long v = (unsigned long long)p;
  v = __shfl_sync(0xFFFFFFFFu, v, src);
  return (const char *)v;
}

__global__ void __launch_bounds__(kBlock)
gather_group_kernel(char *__restrict__ out, const uint32_t *__restrict__ nodes, uint32_t n_max,
                    const uint32_t *__restrict__ d_n, const uint32_t *__restrict__ table,
                    RowSrc rs, uint32_t cpr, uint32_t G, unsigned long long *d_stats) {
  rs.shard0 = (const char *)__ldg((const unsigned long long *)rs.shards);
  const uint32_t n = load_count(n_max, d_n);
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t W = gridDim.x * (kBlock / 32);
  uint32_t g = blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
  const uint32_t groups = (n + G - 1) / G;
  const uint32_t CH = G * cpr;  // chunks per group; <= kGroupChunks unless G == 1
  // loop-invariant chunk map of this lane: (row in group) << 24 | byte offset in row
  uint32_t cmap[kGroupUnroll];
#pragma unroll
  for (int k = 0; k < kGroupUnroll; ++k) {
    const uint32_t c = lane + 32u * k;
    cmap[k] = kEmpty;
    if (c < CH && G > 1) {
      const uint32_t r = c / cpr;
      cmap[k] = (r << 24) | ((c - r * cpr) * 16u);
    }
  }
  uint32_t hits = 0, misses = 0, remote = 0;
  auto load_node = [&](uint32_t gg) -> uint32_t {
    const uint64_t row = (uint64_t)gg * G + lane;
    return (gg < groups && lane < G && row < n) ? __ldg(nodes + row) : kEmpty;
  };
  // pipeline registers: node/slot of the current group, node of the next one
  uint32_t node_c = load_node(g);
  uint32_t node_n = load_node(g + W);
  uint32_t slot_c = node_c != kEmpty ? __ldg(table + node_c) : kEmpty;
  for (; g < groups; g += W) {
    const char *sp = nullptr;
    if (node_c != kEmpty) {
      sp = rs.resolve(node_c, slot_c);
      if (slot_c != kEmpty) ++hits; else ++misses;
      if (rs.is_remote(slot_c)) ++remote;
    }
    // stage the next groups' index loads before streaming this one
    const uint32_t slot_n = node_n != kEmpty ? __ldg(table + node_n) : kEmpty;
    const uint32_t node_nn = load_node(g + 2 * W);
    const uint64_t row0 = (uint64_t)g * G;
    const uint32_t rows_here = (uint32_t)(n - row0 < G ? n - row0 : G);
    char *obase = out + row0 * rs.row_bytes;
    if (G > 1) {
      uint4 v[kGroupUnroll];
#pragma unroll
      for (int k = 0; k < kGroupUnroll; ++k) {
        const uint32_t r = cmap[k] == kEmpty ? 0u : (cmap[k] >> 24);
        const char *p = shfl_ptr(sp, (int)r);
        if (cmap[k] != kEmpty && r < rows_here) v[k] = ld_nc_na_v4(p + (cmap[k] & 0xFFFFFFu));
      }
#pragma unroll
      for (int k = 0; k < kGroupUnroll; ++k) {
        const uint32_t r = cmap[k] >> 24;
        if (cmap[k] != kEmpty && r < rows_here)
          st_na_v4(obase + (size_t)r * rs.row_bytes + (cmap[k] & 0xFFFFFFu), v[k]);
      }
    } else {  // long rows: one row per group, several passes
      const char *p = shfl_ptr(sp, 0);
      for (uint32_t c0 = lane; c0 < cpr; c0 += kGroupChunks) {
        uint4 v[kGroupUnroll];
#pragma unroll
        for (int k = 0; k < kGroupUnroll; ++k)
          if (c0 + 32u * k < cpr) v[k] = ld_nc_na_v4(p + (size_t)(c0 + 32u * k) * 16u);
#pragma unroll
        for (int k = 0; k < kGroupUnroll; ++k)
          if (c0 + 32u * k < cpr) st_na_v4(obase + (size_t)(c0 + 32u * k) * 16u, v[k]);
      }
    }
    node_c = node_n;
    slot_c = slot_n;
    node_n = node_nn;
  }
  rs.report(d_stats, hits, misses, remote);
}

// ---------------------------------------------------------------------------
// Bulk-copy gather (TMA engine, no tensor map: rows are plain byte ranges).
//
// A warp walks super-groups of 32 consecutive output rows (lane l resolves row
// l: one coalesced id load + one table lookup per ROW, software-pipelined one
// and two super-groups ahead).  A super-group is cut into sub-groups of G rows;
// each sub-group goes through one stage of the warp's shared-memory ring:
//   - lanes of the sub-group issue `cp.async.bulk` global->shared for their
//     (hit) row, completion counted on the stage's mbarrier;
//   - miss rows (pinned host memory over the host link) are copied by the whole
//     warp with 16-byte loads into the same stage;
//   - the G output rows are consecutive, so the stage leaves the SM as ONE bulk
//     store of G*row_bytes.
// Payload never touches the register file; S-2 sub-groups of loads and two
// stores are in flight per warp.
// ---------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
// L2 policy for data that is touched once per batch (feature rows in, feature tensor out): evict-first, so
// the 566 MB a gather streams through the 126 MB L2 do not displace the samplers' hash tables and the hot
// part of the CSR that the batches in flight on the other streams live on.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes,
                                         uint32_t bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst_smem),
      "l"(src), "r"(bytes), "r"(bar), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src_smem, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst),
               "r"(src_smem), "r"(bytes), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes,
                                         uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
               "r"(src_smem), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

template <int S, int NW>
__global__ void __launch_bounds__(NW * 32)
gather_bulk_kernel(char *__restrict__ out, const uint32_t *__restrict__ nodes, uint32_t n_max,
                   const uint32_t *__restrict__ d_n, const uint32_t *__restrict__ table,
                   RowSrc rs, uint32_t G, uint32_t stage_bytes, int miss_by_ldg,
                   unsigned long long *d_stats) {
  rs.shard0 = (const char *)__ldg((const unsigned long long *)rs.shards);
  const bool hint = (miss_by_ldg & 2) != 0;  // bit 1 of the mode word: stream through L2 with evict-first
  const bool peer_ldg = (miss_by_ldg & 4) != 0;  // bit 2: rows of NVLink-peer stripes by warp loads, not by the bulk engine
  miss_by_ldg &= 1;
  const uint64_t pol = l2_evict_first_policy();
  constexpr int A = S - 2;  // sub-groups of loads in flight ahead of the store
  extern __shared__ __align__(128) unsigned char s_raw[];
  __shared__ __align__(8) unsigned long long s_bar[NW][S];
  const uint32_t n = load_count(n_max, d_n);
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // balanced contiguous partition: warp w owns sub-groups [w*T/W, (w+1)*T/W) of G rows each
  const uint64_t W = (uint64_t)gridDim.x * NW, w = (uint64_t)blockIdx.x * NW + warp;
  const uint64_t T = ((uint64_t)n + G - 1) / G;
  const uint64_t jb = w * T / W, je = (w + 1) * T / W;
  const uint32_t total = (uint32_t)(je - jb);
  const uint64_t r0 = jb * G;                                   // first row of this warp
  const uint64_t r_end = je * G < n ? je * G : n;               // one past its last row
  const uint32_t spg = 32 / G;                                  // sub-groups per 32-row super-group
  const uint32_t row_bytes = (uint32_t)rs.row_bytes;
  unsigned char *stage0 = s_raw + (size_t)warp * S * stage_bytes;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) mbar_init(smem_u32(&s_bar[warp][s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t hits = 0, misses = 0, remote = 0;
  auto load_node = [&](uint32_t m) -> uint32_t {  // super-group m of this warp: rows r0+32m+lane
    const uint64_t row = r0 + (uint64_t)m * 32 + lane;
    return row < r_end ? __ldg(nodes + row) : kEmpty;
  };

  // index pipeline: (node,slot) of the super-group being issued, the next one, and the node after
  uint32_t node_i = load_node(0), node_n = load_node(1);
  uint32_t slot_i = node_i != kEmpty ? __ldg(table + node_i) : kEmpty;
  uint32_t slot_n = node_n != kEmpty ? __ldg(table + node_n) : kEmpty;
  uint32_t node_nn = load_node(2);
  const char *sp = node_i != kEmpty ? rs.resolve(node_i, slot_i) : nullptr;
  if (node_i != kEmpty) { if (slot_i != kEmpty) ++hits; else ++misses; if (rs.is_remote(slot_i)) ++remote; }

  // issue the loads of this warp's sub-group number j (all lanes call)
  auto issue = [&](uint32_t j) {
    const uint32_t sub = j % spg;
    if (j != 0 && sub == 0) {  // entering the next super-group: rotate the index pipeline
      node_i = node_n; slot_i = slot_n;
      node_n = node_nn;
      slot_n = node_n != kEmpty ? __ldg(table + node_n) : kEmpty;
      node_nn = load_node(j / spg + 2);
      sp = node_i != kEmpty ? rs.resolve(node_i, slot_i) : nullptr;
      if (node_i != kEmpty) { if (slot_i != kEmpty) ++hits; else ++misses; if (rs.is_remote(slot_i)) ++remote; }
    }
    const uint32_t s = j % S;
    const uint32_t bar = smem_u32(&s_bar[warp][s]);
    unsigned char *st = stage0 + (size_t)s * stage_bytes;
    const bool mine = (lane / G) == sub && node_i != kEmpty;
    const bool remote_row = mine && rs.is_remote(slot_i);
    const bool deferred = remote_row && rs.defer_list != nullptr;
    if (deferred) {  // listed for the second pass (its slot of the stage is stored as is and overwritten later)
      const unsigned int at = atomicAdd(rs.defer_cnt, 1u);
      rs.defer_list[at] = make_ulonglong2((unsigned long long)sp, r0 + (uint64_t)(j / spg) * 32 + lane);
    }
    const bool by_bulk = mine && !deferred && (slot_i != kEmpty ? !(peer_ldg && remote_row) : !miss_by_ldg);
    const uint32_t nbulk = __popc(__ballot_sync(0xFFFFFFFFu, by_bulk));
    if (lane == 0) mbar_expect_tx(bar, nbulk * row_bytes);
    __syncwarp();
    if (by_bulk) {
      if (hint) bulk_g2s(smem_u32(st + (size_t)(lane - sub * G) * row_bytes), sp, row_bytes, bar, pol);
      else bulk_g2s(smem_u32(st + (size_t)(lane - sub * G) * row_bytes), sp, row_bytes, bar);
    }
    uint32_t ldg_rows = __ballot_sync(0xFFFFFFFFu, mine && !by_bulk && !deferred);
    while (ldg_rows) {  // host-resident rows: warp-wide 16-byte loads into the stage
      const int r = __ffs(ldg_rows) - 1;
      ldg_rows &= ldg_rows - 1;
      const char *p = shfl_ptr(sp, r);
      for (uint32_t c = lane * 16u; c < row_bytes; c += 32u * 16u)
        *reinterpret_cast<uint4 *>(st + (size_t)(r - sub * G) * row_bytes + c) = ld_nc_na_v4(p + c);
    }
  };

  for (uint32_t j = 0; j < (uint32_t)A && j < total; ++j) issue(j);
  for (uint32_t j = 0; j < total; ++j) {
    if (j + A < total) {
      // stage (j+A)%S was last stored at iteration j-2: allow one younger store to still read
      if (lane == 0) bulk_wait_read<1>();
      __syncwarp();
      issue(j + A);
    }
    const uint32_t s = j % S;
    mbar_wait(smem_u32(&s_bar[warp][s]), (j / S) & 1u);
    const uint64_t row0 = r0 + (uint64_t)j * G;
    const uint32_t rows_here = (uint32_t)(r_end - row0 < G ? r_end - row0 : G);
    fence_proxy_async();  // generic-proxy stage writes (miss rows) -> async proxy
    __syncwarp();
    if (lane == 0) {
      if (hint) bulk_s2g(out + row0 * row_bytes, smem_u32(stage0 + (size_t)s * stage_bytes), rows_here * row_bytes, pol);
      else bulk_s2g(out + row0 * row_bytes, smem_u32(stage0 + (size_t)s * stage_bytes), rows_here * row_bytes);
      bulk_commit();
    }
  }
  if (lane == 0) bulk_wait_read<0>();
  rs.report(d_stats, hits, misses, remote);
}


// Second pass of the gather: the peer rows the main kernel listed.  Every warp keeps kDeferStages rows in flight
// through the bulk-copy engine (global -> shared -> global, one elected lane; 16-byte warp loads from a peer turn
// into 16-byte NVLink reads and reach only ~30 GB/s, r2_partition_diag_n4.txt), so with ~4700 warps all listed
// rows are in flight at once and the pass costs about one NVLink round trip.  The last CTA re-arms the counter.
constexpr int kDeferStages = 8;
constexpr int kDeferWarps = 8;

__global__ void __launch_bounds__(kDeferWarps * 32)
gather_deferred_kernel(char *__restrict__ out, unsigned int *cnt, const ulonglong2 *__restrict__ list,
                       uint32_t row_bytes) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  __shared__ __align__(8) unsigned long long s_bar[kDeferWarps][kDeferStages];
  const uint32_t n = *((volatile unsigned int *)cnt);
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t W = gridDim.x * kDeferWarps, w = blockIdx.x * kDeferWarps + warp;
  unsigned char *stage0 = s_raw + (size_t)warp * kDeferStages * row_bytes;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kDeferStages; ++s) mbar_init(smem_u32(&s_bar[warp][s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (lane == 0 && w < n) {
    const uint32_t total = (n - w + W - 1) / W;  // rows of this warp: w, w + W, ...
    auto issue = [&](uint32_t k) {
      const ulonglong2 e = list[w + (size_t)k * W];
      const uint32_t s = k % kDeferStages;
      const uint32_t bar = smem_u32(&s_bar[warp][s]);
      mbar_expect_tx(bar, row_bytes);
      bulk_g2s(smem_u32(stage0 + (size_t)s * row_bytes), (const void *)e.x, row_bytes, bar);
    };
    for (uint32_t k = 0; k < (uint32_t)kDeferStages - 1 && k < total; ++k) issue(k);
    for (uint32_t k = 0; k < total; ++k) {
      if (k + kDeferStages - 1 < total) {
        bulk_wait_read<0>();  // the stage about to be refilled was stored at iteration k - 1
        issue(k + kDeferStages - 1);
      }
      const uint32_t s = k % kDeferStages;
      mbar_wait(smem_u32(&s_bar[warp][s]), (k / kDeferStages) & 1u);
      const ulonglong2 e = list[w + (size_t)k * W];
      bulk_s2g(out + (size_t)e.y * row_bytes, smem_u32(stage0 + (size_t)s * row_bytes), row_bytes);
      bulk_commit();
    }
    bulk_wait_read<0>();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int prev = atomicAdd(cnt + 1, 1u);
    if (prev == gridDim.x - 1) {  // everybody has read the count
      cnt[0] = 0u;
      __threadfence();
      cnt[1] = 0u;
    }
  }
}

inline int vec_width(size_t row_bytes, const void *a, const void *b, const void *c = nullptr) {
  const uintptr_t bits = (uintptr_t)row_bytes | (uintptr_t)a | (uintptr_t)b | (uintptr_t)c;
  if ((bits & 15) == 0) return 16;
  if ((bits & 7) == 0) return 8;
  if ((bits & 3) == 0) return 4;
  return 1;
}

// exactly one resident wave (ncu r1_a: a fixed 8 CTAs/SM grid ran 1.33 waves
// because only 6 CTAs of the 40-register uint4 kernel fit -> 25 % tail)
inline int copy_grid(uint64_t chunks, int occ) {
  return persistent_grid(chunks, kBlock * kUnroll, occ, false, true);
}

}  // namespace
}  // namespace fgnn

using namespace fgnn;

extern "C" int fgnn_k_row_copy(void *dst, const uint32_t *dst_index, const void *src,
                               const uint32_t *src_index, uint64_t src_mask, uint32_t n_max,
                               const uint32_t *d_n, size_t row_bytes, fgnn_stream_t stream) {
  if (n_max == 0 || row_bytes == 0) return 0;
  if (!dst || !src) return FGNN_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int w = vec_width(row_bytes, dst, src);
  const uint32_t cpr = (uint32_t)(row_bytes / w);
#define FGNN_RC(V)                                                                              \
  do {                                                                                          \
    static const int occ = occupancy(row_copy_kernel<V>, kBlock, 0);                            \
    const int grid = copy_grid((uint64_t)n_max * cpr, occ);                                     \
    row_copy_kernel<V><<<grid, kBlock, 0, st>>>((char *)dst, dst_index, (const char *)src,      \
                                                src_index, src_mask, n_max, d_n, row_bytes, cpr); \
  } while (0)
  if (w == 16) FGNN_RC(uint4);
  else if (w == 8) FGNN_RC(uint2);
  else if (w == 4) FGNN_RC(uint32_t);
  else FGNN_RC(uint8_t);
#undef FGNN_RC
  note_launch();
  return check_last();
}

namespace fgnn {
namespace {
// A/B switches for profiling (read once): FGNN_GATHER_IMPL = flat | group | bulk
struct GatherTuning {
  int impl;         // 0 flat, 1 group, 2 bulk
  int stages;       // bulk: ring depth per warp (6, or 3 for long rows)
  uint32_t stage_cap;  // bulk: max bytes per stage
  int miss_ldg;     // bulk: host-resident rows by warp loads (1) or by the bulk engine (0)
  int peer_ldg;     // bulk: rows of NVLink-peer stripes by warp loads (1) or by the bulk engine (0)
  int l2_hint;      // bulk: evict-first L2 policy on the streamed rows
  uint32_t group_rows; // group: rows per warp group (0 = auto)
  int ctas_per_sm;  // 0 = occupancy
};
inline int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}
GatherTuning read_tuning() {
  GatherTuning g;
  const char *v = getenv("FGNN_GATHER_IMPL");
  g.impl = 2;
  if (v && !strcmp(v, "flat")) g.impl = 0;
  if (v && !strcmp(v, "group")) g.impl = 1;
  if (v && !strcmp(v, "bulk")) g.impl = 2;
  // round-1 sweeps (profiles/r1_d_gather_sweep*.txt, r1_q_overlap_sweep.txt) settled on 16 warps x 6 stages of
  // <= 2 KB; the other warp/stage shapes were dropped from the binary in round 2
  g.stages = env_int("FGNN_BULK_STAGES", 6);
  g.stage_cap = (uint32_t)env_int("FGNN_BULK_STAGE_BYTES", 2048);
  g.miss_ldg = env_int("FGNN_BULK_MISS_LDG", 0);
  g.peer_ldg = env_int("FGNN_BULK_PEER_LDG", 0);
  g.l2_hint = env_int("FGNN_GATHER_L2HINT", 1);
  g.group_rows = (uint32_t)env_int("FGNN_GROUP_ROWS", 0);
  g.ctas_per_sm = env_int("FGNN_GATHER_CTAS_PER_SM", 0);
  return g;
}
const GatherTuning &tuning() {
  static GatherTuning t = read_tuning();
  if (env_int("FGNN_TUNING_DYNAMIC", 0) != 0) t = read_tuning();  // sweeps/tests: re-read per call
  return t;
}

constexpr int kBulkWarps = 16;

template <int S>
int launch_bulk(char *out, const uint32_t *nodes, uint32_t n_max, const uint32_t *d_n,
                const uint32_t *table, const RowSrc &rs, uint32_t G, uint32_t stage_bytes,
                unsigned long long *d_stats, cudaStream_t st) {
  auto kern = gather_bulk_kernel<S, kBulkWarps>;
  const size_t smem = (size_t)kBulkWarps * S * stage_bytes;
  // the opt-in is per device and cheap: set it on every launch (a process may drive several GPUs)
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  int occ = tuning().ctas_per_sm ? tuning().ctas_per_sm : occupancy(kern, kBulkWarps * 32, smem);
  const uint64_t subs = ((uint64_t)n_max + G - 1) / G;
  const int grid = persistent_grid(subs, kBulkWarps * 4, occ, false, true);  // >= 4 sub-groups per warp
  kern<<<grid, kBulkWarps * 32, smem, st>>>(out, nodes, n_max, d_n, table, rs, G, stage_bytes,
                                            (tuning().miss_ldg ? 1 : 0) | (tuning().l2_hint ? 2 : 0) |
                                                (tuning().peer_ldg ? 4 : 0), d_stats);
  return 0;
}
}  // namespace
}  // namespace fgnn

extern "C" size_t fgnn_k_gather_defer_workspace_bytes(uint32_t n_max) { return 256 + (size_t)n_max * sizeof(ulonglong2); }

extern "C" int fgnn_k_gather_cached_layout(void *out, const uint32_t *nodes, uint32_t n_max,
                                           const uint32_t *d_n, const fgnn_cache_layout *lay,
                                           unsigned long long *d_stats, unsigned long long *d_remote,
                                           fgnn_stream_t stream) {
  if (!lay) return FGNN_ERR_BAD_ARG;
  const size_t row_bytes = lay->row_bytes;
  if (n_max == 0 || row_bytes == 0) return 0;
  if (!out || !nodes || !lay->table || !lay->shards || lay->num_shards == 0 || !lay->miss_src) return FGNN_ERR_BAD_ARG;
  if (lay->num_replicated > 0 && !lay->replica) return FGNN_ERR_BAD_ARG;
  if (lay->self_shard >= lay->num_shards) return FGNN_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  trace_mark(st, FGNN_TRACE_GATHER_BEGIN);
  // shard bases are cudaMalloc'ed (256-B aligned); only out/miss_src/replica/row_bytes decide
  const int w = vec_width(row_bytes, out, lay->miss_src, lay->replica);
  const uint32_t cpr = (uint32_t)(row_bytes / w);
  const GatherTuning &tn = tuning();
  RowSrc rs;
  rs.shards = lay->shards;
  rs.shard0 = nullptr;  // fetched from shards[0] by the kernel
  rs.num_shards = lay->num_shards;
  rs.miss_src = (const char *)lay->miss_src;
  rs.miss_mask = lay->miss_mask;
  rs.row_bytes = row_bytes;
  rs.replica = (const char *)lay->replica;
  rs.num_replicated = lay->num_replicated;
  rs.self_shard = lay->self_shard;
  rs.d_remote = d_remote;
  rs.defer_cnt = nullptr;
  rs.defer_list = nullptr;
  const uint32_t *table = lay->table;
  int rc = 0;
  bool done = false;
  if (w == 16 && tn.impl >= 2) {
    // stage = the most rows (power of two, <= 32) that fit the stage cap; long rows get the 3-stage ring so that
    // 16 warps * stages * stage_bytes stays within shared memory, longer ones the warp-group kernel
    uint32_t G = 32;
    while (G > 1 && (size_t)G * row_bytes > tn.stage_cap) G >>= 1;
    const uint32_t stage_bytes = (uint32_t)(G * row_bytes);
    const size_t kSmemBudget = 200 * 1024;
    int stages = tn.stages >= 6 ? 6 : 3;
    if ((size_t)kBulkWarps * stages * stage_bytes > kSmemBudget) stages = 3;
    if ((size_t)kBulkWarps * stages * stage_bytes <= kSmemBudget) {
      const bool defer = lay->defer_ws != nullptr && lay->num_shards > 1 && env_int("FGNN_GATHER_DEFER", 0) != 0 &&
                         (size_t)kDeferWarps * kDeferStages * row_bytes <= kSmemBudget;
      if (defer) {
        rs.defer_cnt = (unsigned int *)lay->defer_ws;
        rs.defer_list = (ulonglong2 *)((char *)lay->defer_ws + 256);
      }
      rc = stages == 6 ? launch_bulk<6>((char *)out, nodes, n_max, d_n, table, rs, G, stage_bytes, d_stats, st)
                       : launch_bulk<3>((char *)out, nodes, n_max, d_n, table, rs, G, stage_bytes, d_stats, st);
      if (rc) return rc;
      if (defer) {
        note_launch();
        const size_t dsm = (size_t)kDeferWarps * kDeferStages * row_bytes;
        cudaError_t e2 = cudaFuncSetAttribute(gather_deferred_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm);
        if (e2 != cudaSuccess) return (int)e2;
        const int dgrid = sm_count() * (int)std::max<size_t>(1, std::min<size_t>(4, (200 * 1024) / dsm));
        gather_deferred_kernel<<<dgrid, kDeferWarps * 32, dsm, st>>>((char *)out, rs.defer_cnt, rs.defer_list,
                                                                    (uint32_t)row_bytes);
        rs.defer_cnt = nullptr;
        rs.defer_list = nullptr;
      }
      done = true;
    }
  }
  if (!done && w == 16 && tn.impl != 0) {
    uint32_t G = tn.group_rows ? tn.group_rows : kGroupChunks / cpr;
    if (G > 32) G = 32;
    if (G < 1 || G * cpr > kGroupChunks) G = 1;
    static const int occ_g = occupancy(gather_group_kernel, kBlock, 0);
    const int occ = tn.ctas_per_sm ? tn.ctas_per_sm : occ_g;
    const uint64_t groups = ((uint64_t)n_max + G - 1) / G;
    const int grid = persistent_grid(groups, kBlock / 32, occ, false, true);
    gather_group_kernel<<<grid, kBlock, 0, st>>>((char *)out, nodes, n_max, d_n, table, rs, cpr, G, d_stats);
    done = true;
  }
  if (!done) {
#define FGNN_GC(V)                                                                               \
  do {                                                                                           \
    static const int occ = occupancy(gather_cached_kernel<V>, kBlock, 0);                        \
    const int grid = copy_grid((uint64_t)n_max * cpr, occ);                                      \
    gather_cached_kernel<V><<<grid, kBlock, 0, st>>>((char *)out, nodes, n_max, d_n, table, rs, cpr, d_stats); \
  } while (0)
    if (w == 16) FGNN_GC(uint4);
    else if (w == 8) FGNN_GC(uint2);
    else if (w == 4) FGNN_GC(uint32_t);
    else FGNN_GC(uint8_t);
#undef FGNN_GC
  }
  note_launch();
  rc = check_last();
  trace_mark(st, FGNN_TRACE_GATHER_END);
  return rc;
}

extern "C" int fgnn_k_gather_cached(void *out, const uint32_t *nodes, uint32_t n_max,
                                    const uint32_t *d_n, const uint32_t *table,
                                    const void *const *shards, uint32_t num_shards,
                                    const void *miss_src, uint64_t miss_mask, size_t row_bytes,
                                    unsigned long long *d_stats, fgnn_stream_t stream) {
  fgnn_cache_layout lay;
  memset(&lay, 0, sizeof(lay));
  lay.table = table;
  lay.shards = shards;
  lay.num_shards = num_shards;
  lay.miss_src = miss_src;
  lay.miss_mask = miss_mask;
  lay.row_bytes = row_bytes;
  return fgnn_k_gather_cached_layout(out, nodes, n_max, d_n, &lay, d_stats, nullptr, stream);
}
