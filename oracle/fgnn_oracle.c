/* TEST INFRASTRUCTURE — see fgnn_oracle.h.  Plain C11 restatement of the
 * reference's hot path.  Every function cites the reference lines it follows
 * (relative to /root/reference/samgraph/common/). */
#include "fgnn_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int g_threads = 0;
void fgo_set_threads(int n) {
  g_threads = n;
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#endif
}

/* ======================================================================== */
/* Philox4x32-10 (Salmon et al., SC'11; constants as in Random123 / cuRAND)  */
/* ======================================================================== */
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

void fgo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2],
                       uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
    uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += PHILOX_W0;
    k1 += PHILOX_W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Stream layout shared with the CUDA kernels (csrc/kernels/philox.cuh):
 *   key = (seed.lo, seed.hi ^ batch_key.hi)
 *   ctr = (draw >> 2, item, tag, batch_key.lo);  result word = draw & 3      */
uint32_t fgo_rand_u32(uint64_t seed, uint64_t batch_key, uint32_t tag,
                      uint32_t item, uint32_t draw) {
  uint32_t key[2] = {(uint32_t)seed,
                     (uint32_t)(seed >> 32) ^ (uint32_t)(batch_key >> 32)};
  uint32_t ctr[4] = {draw >> 2, item, tag, (uint32_t)batch_key};
  uint32_t out[4];
  fgo_philox4x32_10(ctr, key, out);
  return out[draw & 3];
}

/* curand_uniform.h:69-72: x * 2^-32 + 2^-33, evaluated in float, in (0,1] */
float fgo_uniform_f32(uint32_t x) {
  return fmaf((float)x, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
}
/* curand_uniform.h:101-106 */
double fgo_uniform_f64(uint32_t x, uint32_t y) {
  uint64_t z = (uint64_t)x ^ ((uint64_t)y << (53 - 32));
  return (double)z * 1.1102230246251565e-16 + (1.1102230246251565e-16 / 2.0);
}

/* ======================================================================== */
/* sizing                                                                     */
/* ======================================================================== */
size_t fgo_predict_num_nodes(size_t batch, const size_t *fanout,
                             size_t num_fanout) {
  size_t count = batch;
  for (size_t i = num_fanout; i-- > 0;) count += count * fanout[i];
  return count;
}

size_t fgo_table_size(size_t num, size_t scale) {
  /* 1 << (size_t)(1 + log2(num >> 1)) then << scale */
  size_t half = num >> 1;
  size_t lg = 0;
  while ((half >> (lg + 1)) != 0) ++lg; /* floor(log2(half)), half >= 1 */
  return ((size_t)1 << (1 + lg)) << scale;
}

/* ======================================================================== */
/* epoch shuffle                                                              */
/* ======================================================================== */
typedef struct { uint32_t key; uint32_t idx; } key_idx;
static int cmp_key_idx(const void *a, const void *b) {
  const key_idx *x = a, *y = b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  if (x->idx != y->idx) return x->idx < y->idx ? -1 : 1;
  return 0;
}
void fgo_shuffle(const uint32_t *train_set, size_t n, uint64_t seed,
                 uint64_t epoch, uint32_t *out) {
  key_idx *k = malloc(sizeof(key_idx) * (n + 1));
  for (size_t i = 0; i < n; ++i) {
    k[i].key = fgo_rand_u32(seed, epoch, FGO_SHUFFLE_TAG, (uint32_t)i, 0);
    k[i].idx = (uint32_t)i;
  }
  qsort(k, n, sizeof(key_idx), cmp_key_idx);
  for (size_t i = 0; i < n; ++i) out[i] = train_set[k[i].idx];
  free(k);
}

/* ======================================================================== */
/* samplers                                                                   */
/* ======================================================================== */
/* compaction of a padded [num_input x fanout] COO whose valid entries form a
 * prefix of every row: cuda_sampling_khop0.cu:128-173 */
static size_t compact_rows(const uint32_t *tmp_src, const uint32_t *tmp_dst,
                           size_t num_input, size_t fanout, uint32_t *out_src,
                           uint32_t *out_dst) {
  size_t n = 0;
  for (size_t i = 0; i < num_input; ++i) {
    for (size_t j = 0; j < fanout; ++j) {
      if (tmp_src[i * fanout + j] == FGO_EMPTY) break;
      out_src[n] = tmp_src[i * fanout + j];
      out_dst[n] = tmp_dst[i * fanout + j];
      ++n;
    }
  }
  return n;
}

size_t fgo_sample_khop0(const uint32_t *indptr, const uint32_t *indices,
                        const uint32_t *input, size_t num_input, size_t fanout,
                        uint64_t seed, uint64_t batch_key, uint32_t tag,
                        uint32_t *out_src, uint32_t *out_dst) {
  uint32_t *tmp_src = malloc(sizeof(uint32_t) * (num_input * fanout + 1));
  uint32_t *tmp_dst = malloc(sizeof(uint32_t) * (num_input * fanout + 1));
#pragma omp parallel for schedule(dynamic, 64)
  for (size_t i = 0; i < num_input; ++i) {
    const uint32_t rid = input[i];
    const uint32_t off = indptr[rid];
    const uint32_t len = indptr[rid + 1] - off;
    uint32_t *s = tmp_src + i * fanout, *d = tmp_dst + i * fanout;
    if (len <= fanout) { /* khop0.cu:61-71 */
      size_t j = 0;
      for (; j < len; ++j) { s[j] = rid; d[j] = indices[off + j]; }
      for (; j < fanout; ++j) { s[j] = FGO_EMPTY; d[j] = FGO_EMPTY; }
    } else { /* khop0.cu:72-84, curand()%(j+1) -> Philox draw j-fanout */
      for (size_t j = 0; j < fanout; ++j) { s[j] = rid; d[j] = indices[off + j]; }
      for (size_t j = fanout; j < len; ++j) {
        uint32_t r = fgo_rand_u32(seed, batch_key, tag, (uint32_t)i,
                                  (uint32_t)(j - fanout));
        size_t k = r % (uint32_t)(j + 1);
        if (k < fanout) d[k] = indices[off + j];
      }
    }
  }
  size_t n = compact_rows(tmp_src, tmp_dst, num_input, fanout, out_src, out_dst);
  free(tmp_src);
  free(tmp_dst);
  return n;
}

size_t fgo_sample_khop2(const uint32_t *indptr, const uint32_t *indices,
                        const uint32_t *input, size_t num_input, size_t fanout,
                        uint64_t seed, uint64_t batch_key, uint32_t tag,
                        uint32_t *out_src, uint32_t *out_dst) {
  uint32_t *tmp_src = malloc(sizeof(uint32_t) * (num_input * fanout + 1));
  uint32_t *tmp_dst = malloc(sizeof(uint32_t) * (num_input * fanout + 1));
#pragma omp parallel for schedule(dynamic, 64)
  for (size_t i = 0; i < num_input; ++i) {
    const uint32_t rid = input[i];
    const uint32_t off = indptr[rid];
    const uint32_t len = indptr[rid + 1] - off;
    uint32_t *s = tmp_src + i * fanout, *d = tmp_dst + i * fanout;
    if (len <= fanout) { /* khop2.cu:61-71 */
      size_t j = 0;
      for (; j < len; ++j) { s[j] = rid; d[j] = indices[off + j]; }
      for (; j < fanout; ++j) { s[j] = FGO_EMPTY; d[j] = FGO_EMPTY; }
    } else {
      /* khop2.cu:72-83.  The reference swaps inside `indices`; here the
       * swaps live in a sparse virtual copy (key -> current position
       * content) so the CSR stays immutable.  Slot len-j-1 is never read
       * again after step j, so only the write to slot k must be kept. */
      uint32_t *mk = malloc(sizeof(uint32_t) * fanout);
      uint32_t *mv = malloc(sizeof(uint32_t) * fanout);
      for (size_t j = 0; j < fanout; ++j) {
        uint32_t r = fgo_rand_u32(seed, batch_key, tag, (uint32_t)i, (uint32_t)j);
        uint32_t k = r % (uint32_t)(len - j);
        uint32_t last = (uint32_t)(len - j - 1);
        uint32_t vk = k, vlast = last;
        for (size_t t = 0; t < j; ++t) { /* latest write wins */
          if (mk[t] == k) vk = mv[t];
          if (mk[t] == last) vlast = mv[t];
        }
        s[j] = rid;
        d[j] = indices[off + vk];
        mk[j] = k;
        mv[j] = vlast;
      }
      free(mk);
      free(mv);
    }
  }
  size_t n = compact_rows(tmp_src, tmp_dst, num_input, fanout, out_src, out_dst);
  free(tmp_src);
  free(tmp_dst);
  return n;
}

/* order of seeds after the reference's stable radix sort on src id
 * (cuda_sampling_khop1.cu:169-178): ascending id, ties by position. */
typedef struct { uint32_t id; uint32_t pos; } id_pos;
static int cmp_id_pos(const void *a, const void *b) {
  const id_pos *x = a, *y = b;
  if (x->id != y->id) return x->id < y->id ? -1 : 1;
  if (x->pos != y->pos) return x->pos < y->pos ? -1 : 1;
  return 0;
}

/* sort-by-src + "differs from successor" compaction shared by khop1 and the
 * weighted samplers: cuda_sampling_khop1.cu:74-127 */
static size_t sort_dedup_rows(const uint32_t *input, const uint32_t *tmp_src,
                              const uint32_t *tmp_dst, size_t num_input,
                              size_t fanout, uint32_t *out_src,
                              uint32_t *out_dst) {
  id_pos *order = malloc(sizeof(id_pos) * (num_input + 1));
  size_t m = 0;
  for (size_t i = 0; i < num_input; ++i) {
    /* rows whose src is EMPTY (len == 0) sort last and are dropped */
    if (fanout > 0 && tmp_src[i * fanout] == FGO_EMPTY) continue;
    order[m].id = input[i];
    order[m].pos = (uint32_t)i;
    ++m;
  }
  qsort(order, m, sizeof(id_pos), cmp_id_pos);
  /* flatten in sorted order, then compare each entry with its successor */
  size_t total = m * fanout, n = 0;
  for (size_t e = 0; e < total; ++e) {
    size_t i = order[e / fanout].pos, j = e % fanout;
    uint32_t s = tmp_src[i * fanout + j], d = tmp_dst[i * fanout + j];
    int keep = 1;
    if (e + 1 < total) {
      size_t i2 = order[(e + 1) / fanout].pos, j2 = (e + 1) % fanout;
      uint32_t s2 = tmp_src[i2 * fanout + j2], d2 = tmp_dst[i2 * fanout + j2];
      keep = (s != s2) || (d != d2);
    }
    if (keep) { out_src[n] = s; out_dst[n] = d; ++n; }
  }
  free(order);
  return n;
}

size_t fgo_sample_khop1(const uint32_t *indptr, const uint32_t *indices,
                        const uint32_t *input, size_t num_input, size_t fanout,
                        uint64_t seed, uint64_t batch_key, uint32_t tag,
                        uint32_t *out_src, uint32_t *out_dst) {
  uint32_t *tmp_src = malloc(sizeof(uint32_t) * (num_input * fanout + 1));
  uint32_t *tmp_dst = malloc(sizeof(uint32_t) * (num_input * fanout + 1));
#pragma omp parallel for schedule(static)
  for (size_t t = 0; t < num_input * fanout; ++t) {
    size_t i = t / fanout, j = t % fanout;
    const uint32_t rid = input[i];
    const uint32_t off = indptr[rid];
    const uint32_t len = indptr[rid + 1] - off;
    if (len == 0) { /* khop1.cu:61-63 */
      tmp_src[t] = FGO_EMPTY; tmp_dst[t] = FGO_EMPTY;
    } else { /* khop1.cu:64-68 */
      uint32_t k = fgo_rand_u32(seed, batch_key, tag, (uint32_t)i, (uint32_t)j) % len;
      tmp_src[t] = rid; tmp_dst[t] = indices[off + k];
    }
  }
  size_t n = sort_dedup_rows(input, tmp_src, tmp_dst, num_input, fanout, out_src, out_dst);
  free(tmp_src); free(tmp_dst);
  return n;
}

size_t fgo_sample_weighted_khop(const uint32_t *indptr, const uint32_t *indices,
                                const float *prob_table,
                                const uint32_t *alias_table,
                                const uint32_t *input, size_t num_input,
                                size_t fanout, uint64_t seed,
                                uint64_t batch_key, uint32_t tag,
                                uint32_t *out_src, uint32_t *out_dst) {
  uint32_t *tmp_src = malloc(sizeof(uint32_t) * (num_input * fanout + 1));
  uint32_t *tmp_dst = malloc(sizeof(uint32_t) * (num_input * fanout + 1));
#pragma omp parallel for schedule(static)
  for (size_t t = 0; t < num_input * fanout; ++t) {
    size_t i = t / fanout, j = t % fanout;
    const uint32_t rid = input[i];
    const uint32_t off = indptr[rid];
    const uint32_t len = indptr[rid + 1] - off;
    if (len == 0) { /* weighted_khop.cu:60-62 */
      tmp_src[t] = FGO_EMPTY; tmp_dst[t] = FGO_EMPTY;
    } else { /* weighted_khop.cu:63-72 */
      uint32_t k = fgo_rand_u32(seed, batch_key, tag, (uint32_t)i, (uint32_t)(2 * j)) % len;
      float r = fgo_uniform_f32(
          fgo_rand_u32(seed, batch_key, tag, (uint32_t)i, (uint32_t)(2 * j + 1)));
      tmp_src[t] = rid;
      tmp_dst[t] = (r < prob_table[off + k]) ? indices[off + k] : alias_table[off + k];
    }
  }
  size_t n = sort_dedup_rows(input, tmp_src, tmp_dst, num_input, fanout, out_src, out_dst);
  free(tmp_src); free(tmp_dst);
  return n;
}

size_t fgo_sample_weighted_khop_prefix(
    const uint32_t *indptr, const uint32_t *indices,
    const float *prob_prefix_table, const uint32_t *input, size_t num_input,
    size_t fanout, uint64_t seed, uint64_t batch_key, uint32_t tag,
    uint32_t *out_src, uint32_t *out_dst) {
  uint32_t *tmp_src = malloc(sizeof(uint32_t) * (num_input * fanout + 1));
  uint32_t *tmp_dst = malloc(sizeof(uint32_t) * (num_input * fanout + 1));
#pragma omp parallel for schedule(static)
  for (size_t t = 0; t < num_input * fanout; ++t) {
    size_t i = t / fanout, j = t % fanout;
    const uint32_t rid = input[i];
    const uint32_t off = indptr[rid];
    const uint32_t len = indptr[rid + 1] - off;
    if (len == 0) { /* prefix.cu:61-62 (the :59 read happens after the check here) */
      tmp_src[t] = FGO_EMPTY; tmp_dst[t] = FGO_EMPTY;
    } else { /* prefix.cu:59,64-87 */
      const float up = prob_prefix_table[off + len - 1];
      float rand_x = fgo_uniform_f32(fgo_rand_u32(seed, batch_key, tag,
                                                  (uint32_t)i, (uint32_t)j)) * up;
      tmp_src[t] = rid;
      if (rand_x <= prob_prefix_table[off]) {
        tmp_dst[t] = indices[off];
      } else {
        size_t lo = off, hi = (size_t)off + len - 1;
        while (hi - lo >= 2) {
          size_t mid = (lo + hi) >> 1;
          if (prob_prefix_table[mid] >= rand_x) hi = mid; else lo = mid;
        }
        tmp_dst[t] = indices[hi];
      }
    }
  }
  size_t n = sort_dedup_rows(input, tmp_src, tmp_dst, num_input, fanout, out_src, out_dst);
  free(tmp_src); free(tmp_dst);
  return n;
}

#define FGO_HASH_DEDUP_MAX_DRAWS 4096u

size_t fgo_sample_weighted_khop_hash_dedup(
    const uint32_t *indptr, const uint32_t *indices, const float *prob_table,
    const uint32_t *alias_table, const uint32_t *input, size_t num_input,
    size_t fanout, uint64_t seed, uint64_t batch_key, uint32_t tag,
    uint32_t *out_src, uint32_t *out_dst) {
  uint32_t *tmp_src = malloc(sizeof(uint32_t) * (num_input * fanout + 1));
  uint32_t *tmp_dst = malloc(sizeof(uint32_t) * (num_input * fanout + 1));
#pragma omp parallel for schedule(dynamic, 64)
  for (size_t i = 0; i < num_input; ++i) {
    const uint32_t rid = input[i];
    const uint32_t off = indptr[rid];
    const uint32_t len = indptr[rid + 1] - off;
    uint32_t *s = tmp_src + i * fanout, *d = tmp_dst + i * fanout;
    if (len <= fanout) { /* hash_dedup.cu:86-96 */
      size_t j = 0;
      for (; j < len; ++j) { s[j] = rid; d[j] = indices[off + j]; }
      for (; j < fanout; ++j) { s[j] = FGO_EMPTY; d[j] = FGO_EMPTY; }
    } else { /* hash_dedup.cu:97-111; the 50-slot local table is a set */
      size_t got = 0;
      for (uint32_t t = 0; got < fanout; ++t) {
        uint32_t k = fgo_rand_u32(seed, batch_key, tag, (uint32_t)i, 2 * t) % len;
        float r = fgo_uniform_f32(fgo_rand_u32(seed, batch_key, tag, (uint32_t)i, 2 * t + 1));
        uint32_t cand = indices[off + k];
        if (r > prob_table[off + k]) cand = alias_table[off + k];
        int dup = 0;
        for (size_t q = 0; q < got; ++q) if (d[q] == cand) { dup = 1; break; }
        /* termination guard (ours): past MAX_DRAWS duplicates are accepted */
        if (dup && t < FGO_HASH_DEDUP_MAX_DRAWS) continue;
        s[got] = rid; d[got] = cand; ++got;
      }
    }
  }
  size_t n = compact_rows(tmp_src, tmp_dst, num_input, fanout, out_src, out_dst);
  free(tmp_src); free(tmp_dst);
  return n;
}

void fgo_random_walk(const uint32_t *indptr, const uint32_t *indices,
                     const uint32_t *input, size_t num_input, size_t walk_len,
                     double restart_prob, size_t num_walk, uint64_t seed,
                     uint64_t batch_key, uint32_t tag, uint32_t *tmp_src,
                     uint32_t *tmp_dst) {
#pragma omp parallel for schedule(static)
  for (size_t t = 0; t < num_input * num_walk; ++t) {
    size_t n = t / num_walk, w = t % num_walk;
    const uint32_t start = input[n];
    uint32_t node = start;
    for (size_t s = 0; s < walk_len; ++s) {
      /* random_walk.cu:77-78 */
      size_t pos = n * num_walk * walk_len + s * num_walk + w;
      if (node == FGO_EMPTY) { /* :79-80 */
        tmp_src[pos] = FGO_EMPTY; tmp_dst[pos] = FGO_EMPTY;
        continue;
      }
      const uint32_t off = indptr[node];
      const uint32_t len = indptr[node + 1] - off;
      if (len == 0) { /* :85-87 */
        tmp_src[pos] = FGO_EMPTY; tmp_dst[pos] = FGO_EMPTY;
        node = FGO_EMPTY;
      } else { /* :88-97 */
        uint32_t k = fgo_rand_u32(seed, batch_key, tag, (uint32_t)t, (uint32_t)(3 * s)) % len;
        uint32_t x = fgo_rand_u32(seed, batch_key, tag, (uint32_t)t, (uint32_t)(3 * s + 1));
        uint32_t y = fgo_rand_u32(seed, batch_key, tag, (uint32_t)t, (uint32_t)(3 * s + 2));
        tmp_src[pos] = start;
        tmp_dst[pos] = indices[off + k];
        node = indices[off + k];
        if (fgo_uniform_f64(x, y) < restart_prob) node = FGO_EMPTY;
      }
    }
  }
}

size_t fgo_topk(const uint32_t *tmp_src, const uint32_t *tmp_dst,
                const uint32_t *input, size_t num_input, size_t edges_per_node,
                size_t K, uint32_t *out_src, uint32_t *out_dst,
                uint32_t *out_data) {
  size_t n_out = 0;
  uint32_t *u_dst = malloc(sizeof(uint32_t) * (edges_per_node + 1));
  uint32_t *u_cnt = malloc(sizeof(uint32_t) * (edges_per_node + 1));
  for (size_t n = 0; n < num_input; ++n) {
    /* count_frequency_revised (:361-401): unique (start,dst) pairs in
     * first-occurrence order with multiplicities */
    size_t nu = 0;
    for (size_t p = 0; p < edges_per_node; ++p) {
      size_t idx = n * edges_per_node + p;
      if (tmp_src[idx] == FGO_EMPTY) continue;
      size_t q = 0;
      for (; q < nu; ++q) if (u_dst[q] == tmp_dst[idx]) break;
      if (q == nu) { u_dst[nu] = tmp_dst[idx]; u_cnt[nu] = 1; ++nu; }
      else ++u_cnt[q];
    }
    /* SortPairsDescending on ((num_node-node_idx)<<32 | count) (:502-504,
     * :1241-1249) is stable -> within a node: count desc, ties by first
     * occurrence.  Keep min(K, nu) (:585-607), emit (start, dst, count)
     * (:644-676). */
    size_t keep = nu < K ? nu : K;
    for (size_t r = 0; r < keep; ++r) {
      size_t best = (size_t)-1;
      for (size_t q = 0; q < nu; ++q) {
        if (u_cnt[q] == 0) continue;
        if (best == (size_t)-1 || u_cnt[q] > u_cnt[best]) best = q;
      }
      out_src[n_out] = input[n];
      out_dst[n_out] = u_dst[best];
      out_data[n_out] = u_cnt[best];
      u_cnt[best] = 0;
      ++n_out;
    }
  }
  free(u_dst); free(u_cnt);
  return n_out;
}

/* ======================================================================== */
/* ordered hash table                                                         */
/* ======================================================================== */
struct fgo_hashtable {
  uint32_t *keys;   /* open addressing, FGO_EMPTY = free */
  uint32_t *vals;   /* local id */
  size_t cap;       /* power of two */
  uint32_t *n2o;    /* local -> global */
  size_t n2o_cap;
  size_t num_items;
};

static size_t ht_hash(const fgo_hashtable *t, uint32_t id) {
  return (size_t)((id * 0x9E3779B1u) >> 7) & (t->cap - 1);
}

fgo_hashtable *fgo_hashtable_new(size_t max_items) {
  fgo_hashtable *t = calloc(1, sizeof(*t));
  size_t cap = 16;
  while (cap < 2 * max_items + 2) cap <<= 1;
  t->cap = cap;
  t->keys = malloc(sizeof(uint32_t) * cap);
  t->vals = malloc(sizeof(uint32_t) * cap);
  t->n2o_cap = max_items + 1;
  t->n2o = malloc(sizeof(uint32_t) * t->n2o_cap);
  fgo_hashtable_reset(t);
  return t;
}
void fgo_hashtable_free(fgo_hashtable *t) {
  if (!t) return;
  free(t->keys); free(t->vals); free(t->n2o); free(t);
}
void fgo_hashtable_reset(fgo_hashtable *t) { /* cuda_hashtable.cu:714-723 */
  memset(t->keys, 0xFF, sizeof(uint32_t) * t->cap);
  t->num_items = 0;
}
size_t fgo_hashtable_num_items(const fgo_hashtable *t) { return t->num_items; }

static void ht_insert(fgo_hashtable *t, uint32_t id) {
  size_t pos = ht_hash(t, id);
  while (t->keys[pos] != FGO_EMPTY) {
    if (t->keys[pos] == id) return; /* already has a local id: keep it */
    pos = (pos + 1) & (t->cap - 1);
  }
  if (t->num_items >= t->n2o_cap) abort();
  t->keys[pos] = id;
  t->vals[pos] = (uint32_t)t->num_items;
  t->n2o[t->num_items++] = id;
}
/* cuda_hashtable.cu:1017-1037: local = offset + index */
void fgo_hashtable_fill_unique(fgo_hashtable *t, const uint32_t *input, size_t n) {
  for (size_t i = 0; i < n; ++i) ht_insert(t, input[i]);
}
/* cuda_hashtable.cu:725-807 with the canonical (first-occurrence) winner,
 * i.e. exactly cpu_hashtable0.cc:37-47 */
void fgo_hashtable_fill_duplicates(fgo_hashtable *t, const uint32_t *input, size_t n) {
  for (size_t i = 0; i < n; ++i) ht_insert(t, input[i]);
}
void fgo_hashtable_unique(const fgo_hashtable *t, uint32_t *out, size_t n) {
  memcpy(out, t->n2o, sizeof(uint32_t) * n);
}
int fgo_hashtable_map(const fgo_hashtable *t, const uint32_t *in, size_t n,
                      uint32_t *out) {
  int rc = 0;
  for (size_t i = 0; i < n; ++i) {
    size_t pos = ht_hash(t, in[i]);
    while (t->keys[pos] != in[i]) {
      if (t->keys[pos] == FGO_EMPTY) { rc = -1; break; }
      pos = (pos + 1) & (t->cap - 1);
    }
    out[i] = (t->keys[pos] == in[i]) ? t->vals[pos] : FGO_EMPTY;
  }
  return rc;
}

/* ======================================================================== */
/* cache                                                                      */
/* ======================================================================== */
size_t fgo_num_cached(size_t num_nodes, double cache_percentage) {
  /* dist_cache_manager_host.cc:66: size_t(num_nodes * cache_percentage) */
  return (size_t)((double)num_nodes * cache_percentage);
}

void fgo_cache_table_build(const uint32_t *ranking_nodes, size_t num_nodes,
                           size_t num_cached, uint32_t *table) {
  for (size_t i = 0; i < num_nodes; ++i) table[i] = FGO_EMPTY;
  for (size_t i = 0; i < num_cached; ++i) table[ranking_nodes[i]] = (uint32_t)i;
}

void fgo_cache_split(const uint32_t *table, const uint32_t *nodes, size_t n,
                     uint32_t *miss_src, uint32_t *miss_dst, size_t *num_miss,
                     uint32_t *cache_src, uint32_t *cache_dst,
                     size_t *num_cache) {
  size_t nm = 0, nc = 0;
  for (size_t i = 0; i < n; ++i) {
    uint32_t slot = table[nodes[i]];
    if (slot == FGO_EMPTY) { /* cuda_cache.cu:96-102 */
      miss_dst[nm] = (uint32_t)i; miss_src[nm] = nodes[i]; ++nm;
    } else {                 /* cuda_cache.cu:140-146 */
      cache_dst[nc] = (uint32_t)i; cache_src[nc] = slot; ++nc;
    }
  }
  *num_miss = nm;
  *num_cache = nc;
}

void fgo_row_copy(void *dst, const uint32_t *dst_index, const void *src,
                  const uint32_t *src_index, size_t n, size_t row_bytes,
                  uint64_t src_index_mask) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i) {
    size_t s = src_index ? ((size_t)src_index[i] & src_index_mask) : i;
    size_t d = dst_index ? (size_t)dst_index[i] : i;
    memcpy((char *)dst + d * row_bytes, (const char *)src + s * row_bytes, row_bytes);
  }
}

/* ======================================================================== */
/* PreSC                                                                      */
/* ======================================================================== */
void fgo_freq_count(uint32_t *freq, const uint32_t *nodes, size_t n) {
  for (size_t i = 0; i < n; ++i) freq[nodes[i]] += 1; /* pre_sampler.cc:84-88 */
}
static int cmp_u64_desc(const void *a, const void *b) {
  uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
  return x > y ? -1 : (x < y ? 1 : 0);
}
void fgo_presc_rank(const uint32_t *freq, size_t num_nodes, uint32_t *rank) {
  /* pre_sampler.cc:44-49 (lo32 = id, hi32 = freq), :97-99 sort greater */
  uint64_t *keys = malloc(sizeof(uint64_t) * (num_nodes + 1));
  for (size_t i = 0; i < num_nodes; ++i)
    keys[i] = ((uint64_t)freq[i] << 32) | (uint64_t)i;
  qsort(keys, num_nodes, sizeof(uint64_t), cmp_u64_desc);
  for (size_t i = 0; i < num_nodes; ++i) rank[i] = (uint32_t)keys[i];
  free(keys);
}

/* ======================================================================== */
/* weight tables                                                              */
/* ======================================================================== */
void fgo_build_alias_table(const uint32_t *indptr, const uint32_t *indices,
                           size_t num_nodes, const float *weights,
                           float *prob_table, uint32_t *alias_table) {
#pragma omp parallel for schedule(dynamic, 256)
  for (size_t v = 0; v < num_nodes; ++v) {
    const uint32_t off = indptr[v];
    const uint32_t len = indptr[v + 1] - off;
    if (len == 0) continue;
    float *w = malloc(sizeof(float) * len);
    uint32_t *smalls = malloc(sizeof(uint32_t) * 2 * len);
    uint32_t *larges = malloc(sizeof(uint32_t) * 2 * len);
    size_t sh = 0, st = 0, lh = 0, lt = 0; /* FIFO queues (std::queue) */
    float sum = 0.0f;
    for (uint32_t i = 0; i < len; ++i) { w[i] = weights[off + i]; sum += w[i]; }
    for (uint32_t i = 0; i < len; ++i) { w[i] /= sum; w[i] *= (float)len; }
    for (uint32_t i = 0; i < len; ++i) {
      alias_table[off + i] = 0; /* std::vector<uint32_t> alias_table(num_edges): stays 0 where prob ends at 1 */
      if (w[i] < 1.0) smalls[st++] = i; else larges[lt++] = i;
    }
    while (sh < st && lh < lt) { /* create_alias_table.cc:145-161 */
      uint32_t s = smalls[sh++], l = larges[lh++];
      prob_table[off + s] = w[s];
      alias_table[off + s] = indices[off + l];
      w[l] -= (1 - w[s]);
      if (w[l] < 1.0) smalls[st++] = l; else larges[lt++] = l;
    }
    while (lh < lt) prob_table[off + larges[lh++]] = 1;
    while (sh < st) prob_table[off + smalls[sh++]] = 1;
    free(w); free(smalls); free(larges);
  }
}

void fgo_build_prefix_table(const uint32_t *indptr, size_t num_nodes,
                            const float *weights, float *prefix_table) {
#pragma omp parallel for schedule(dynamic, 256)
  for (size_t v = 0; v < num_nodes; ++v) {
    const uint32_t off = indptr[v];
    const uint32_t len = indptr[v + 1] - off;
    float sum = 0.0f;
    for (uint32_t i = 0; i < len; ++i) {
      sum += weights[off + i];
      prefix_table[off + i] = sum;
    }
  }
}
