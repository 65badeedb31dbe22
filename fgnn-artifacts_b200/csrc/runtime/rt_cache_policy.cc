// Host-side builders of the degree_hop and fake_optimal cache rankings (include/fgnn_dataset_tools.h).  The
// reference computes both with offline host tools (utility/data-process/toolkit/cache/cache_by_degree_hop.cc,
// cache_by_fake_optimal.cc) and its engine only loads their files (engine.cc:233-244); here the same rankings can
// also be built at data_init when the file is absent.  Dataset preparation, not the per-batch path.
#include "fgnn_dataset_tools.h"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <functional>
#include <thread>
#include <utility>
#include <vector>

namespace {

int ResolveThreads(int n) {
  if (n > 0) return n;
  const unsigned hc = std::thread::hardware_concurrency();
  return (int)std::max(1u, std::min(hc ? hc : 1u, 64u));
}

// fn(thread, begin, end) over [0, n) in contiguous blocks
template <typename Fn>
void ParallelFor(size_t n, int threads, Fn fn) {
  threads = (int)std::min<size_t>((size_t)threads, std::max<size_t>(n / 4096, 1));
  if (threads <= 1) {
    fn(0, (size_t)0, n);
    return;
  }
  std::vector<std::thread> pool;
  const size_t per = (n + threads - 1) / threads;
  for (int t = 0; t < threads; ++t) {
    const size_t b = std::min(n, (size_t)t * per), e = std::min(n, b + per);
    pool.emplace_back([=] { fn(t, b, e); });
  }
  for (auto &th : pool) th.join();
}

// chunk-wise std::sort + rounds of pairwise merges; `less` defines the final (descending) order
template <typename T, typename Less>
void ParallelSort(std::vector<T> &v, int threads, Less less) {
  const size_t n = v.size();
  int parts = 1;
  while (parts * 2 <= threads && n / (parts * 2) >= 65536) parts *= 2;
  if (parts == 1) {
    std::sort(v.begin(), v.end(), less);
    return;
  }
  std::vector<size_t> cut(parts + 1);
  for (int i = 0; i <= parts; ++i) cut[i] = n * (size_t)i / parts;
  {
    std::vector<std::thread> pool;
    for (int i = 0; i < parts; ++i)
      pool.emplace_back([&, i] { std::sort(v.begin() + cut[i], v.begin() + cut[i + 1], less); });
    for (auto &th : pool) th.join();
  }
  for (int width = 1; width < parts; width *= 2) {
    std::vector<std::thread> pool;
    for (int i = 0; i + width < parts; i += 2 * width) {
      const size_t a = cut[i], m = cut[i + width], b = cut[std::min(parts, i + 2 * width)];
      pool.emplace_back([&, a, m, b] { std::inplace_merge(v.begin() + a, v.begin() + m, v.begin() + b, less); });
    }
    for (auto &th : pool) th.join();
  }
}

inline void AtomicInc(uint32_t *p) { __atomic_fetch_add(p, 1u, __ATOMIC_RELAXED); }

// occurrences of every vertex in the adjacency rows selected by `row_mask` (nullptr = all rows): the "out degree"
// of the reference's CSC view (common/graph_loader.cc:109-147)
void OutDegrees(const uint32_t *indptr, const uint32_t *indices, size_t V, const uint8_t *row_mask, int threads,
                uint32_t *deg) {
  std::memset(deg, 0, V * sizeof(uint32_t));
  ParallelFor(V, threads, [&](int, size_t b, size_t e) {
    for (size_t v = b; v < e; ++v) {
      if (row_mask && !row_mask[v]) continue;
      for (uint32_t j = indptr[v]; j < indptr[v + 1]; ++j) AtomicInc(&deg[indices[j]]);
    }
  });
}

void RankByKeyDescending(const uint32_t *key, size_t V, int threads, uint32_t *rank) {
  std::vector<uint64_t> packed(V);
  ParallelFor(V, threads, [&](int, size_t b, size_t e) {
    for (size_t v = b; v < e; ++v) packed[v] = ((uint64_t)key[v] << 32) | (uint64_t)v;
  });
  ParallelSort(packed, threads, std::greater<uint64_t>());
  ParallelFor(V, threads, [&](int, size_t b, size_t e) {
    for (size_t i = b; i < e; ++i) rank[i] = (uint32_t)packed[i];
  });
}

// vertex -> dense index of the vertices one training node touches (open addressing, cleared through the list)
class TouchMap {
 public:
  TouchMap() { Resize(1u << 12); }
  // index of v, inserting it (h1 = h2 = 1) when new
  uint32_t Get(uint32_t v) {
    if ((nodes.size() + 1) * 2 > table_.size()) Grow();
    uint32_t p = Hash(v);
    for (;;) {
      const uint32_t s = table_[p];
      if (s == kFree) break;
      if (nodes[s] == v) return s;
      p = (p + 1) & mask_;
    }
    const uint32_t idx = (uint32_t)nodes.size();
    table_[p] = idx;
    nodes.push_back(v);
    h1.push_back(1.0);
    h2.push_back(1.0);
    return idx;
  }
  void Clear() {
    if (nodes.size() * 8 > table_.size()) {
      std::fill(table_.begin(), table_.end(), kFree);
    } else {
      for (uint32_t v : nodes) {  // remove exactly the probes' home runs: every slot holding one of our indices
        uint32_t p = Hash(v);
        while (table_[p] != kFree) {
          table_[p] = kFree;
          p = (p + 1) & mask_;
        }
      }
    }
    nodes.clear();
    h1.clear();
    h2.clear();
  }
  std::vector<uint32_t> nodes;
  std::vector<double> h1, h2;  // hop1_miss_prob_table / hop2_miss_prob_table of the touched vertices

 private:
  static constexpr uint32_t kFree = 0xFFFFFFFFu;
  uint32_t Hash(uint32_t v) const { return (v * 2654435761u) & mask_; }
  void Resize(size_t cap) {
    table_.assign(cap, kFree);
    mask_ = (uint32_t)cap - 1;
  }
  void Grow() {
    Resize(table_.size() * 2);
    for (uint32_t i = 0; i < nodes.size(); ++i) {
      uint32_t p = Hash(nodes[i]);
      while (table_[p] != kFree) p = (p + 1) & mask_;
      table_[p] = i;
    }
  }
  std::vector<uint32_t> table_;
  uint32_t mask_ = 0;
};

struct Contribution {
  uint32_t node;
  double value;
};

// procBatchTrainNode for ONE training node (the tool's main() uses batch_size = 1, cache_by_fake_optimal.cc:173)
void FakeOptimalOneSeed(const uint32_t *indptr, const uint32_t *indices, uint32_t t, double fanout0, double fanout1,
                        uint32_t order_threads, TouchMap &m, std::vector<uint32_t> &compact, std::vector<uint32_t> &cnt,
                        std::vector<Contribution> *out) {
  m.Clear();
  const uint32_t it = m.Get(t);
  {  // hop 1: every edge of the seed multiplies its endpoint's miss probability (:70-84)
    const uint32_t b = indptr[t], e = indptr[t + 1];
    if (e > b) {
      double miss = 1 - fanout1 / static_cast<double>(e - b);
      miss = std::max(0.0, miss);
      for (uint32_t j = b; j < e; ++j) {
        const uint32_t i = m.Get(indices[j]);
        m.h1[i] *= miss;
      }
    }
  }
  m.h1[it] = 0.0;  // :86-90
  // TouchedNodeCtx::compact(): vertices bucketed by id % threads, first-touch order inside a bucket (:44-60)
  const uint32_t n1 = (uint32_t)m.nodes.size();
  compact.resize(n1);
  if (order_threads <= 1) {
    for (uint32_t i = 0; i < n1; ++i) compact[i] = i;
  } else {
    cnt.assign(order_threads + 1, 0);
    for (uint32_t i = 0; i < n1; ++i) ++cnt[m.nodes[i] % order_threads + 1];
    for (uint32_t k = 0; k < order_threads; ++k) cnt[k + 1] += cnt[k];
    for (uint32_t i = 0; i < n1; ++i) compact[cnt[m.nodes[i] % order_threads]++] = i;
  }
  for (uint32_t c = 0; c < n1; ++c) {  // hop 2 (:96-112)
    const uint32_t ih = compact[c];
    const uint32_t h = m.nodes[ih];
    const uint32_t b = indptr[h], e = indptr[h + 1];
    if (e == b) continue;
    const double b1_hit = 1 - m.h1[ih];
    const double b2_hit = std::min(1.0, fanout0 / static_cast<double>(e - b));
    const double path_miss = 1 - b1_hit * b2_hit;
    for (uint32_t k = b; k < e; ++k) {
      const uint32_t i = m.Get(indices[k]);
      m.h2[i] *= path_miss;
    }
  }
  m.h2[it] = 0.0;  // :113-117
  for (uint32_t i = 0; i < m.nodes.size(); ++i) {  // :121-127
    if (m.h1[i] == 1 && m.h2[i] == 1) continue;
    out->push_back(Contribution{m.nodes[i], 1 - m.h1[i] * m.h2[i]});
  }
}

}  // namespace

extern "C" int fgnn_rt_rank_degree_hop(const uint32_t *indptr, const uint32_t *indices, size_t V,
                                       const uint32_t *train_set, size_t num_train, int hops, int num_threads,
                                       uint32_t *rank) {
  if (!indptr || !rank || (num_train && !train_set) || hops < 0 || V >= 0xFFFFFFFFull) return -1;
  if (V == 0) return 0;
  const int threads = ResolveThreads(num_threads);
  // hopNodes (:31-82): 2 = frontier, 1 = visited earlier, 0 = untouched
  std::vector<uint8_t> before(V, 0), after(V, 0);
  for (size_t i = 0; i < num_train; ++i) {
    if (train_set[i] >= V) return -1;
    before[train_set[i]] = 2;
  }
  for (int hop = 0; hop < hops; ++hop) {
    ParallelFor(V, threads, [&](int, size_t b, size_t e) {
      for (size_t v = b; v < e; ++v) {
        if (before[v] != 2) continue;
        for (uint32_t j = indptr[v]; j < indptr[v + 1]; ++j)
          __atomic_store_n(&after[indices[j]], (uint8_t)1, __ATOMIC_RELAXED);
      }
    });
    ParallelFor(V, threads, [&](int, size_t b, size_t e) {
      for (size_t v = b; v < e; ++v) {
        if (after[v] == 0) {
          if (before[v]) before[v] = 1;
        } else {
          before[v] = before[v] ? 1 : 2;
          after[v] = 0;
        }
      }
    });
  }
  // gen_khop_graph + GetDegrees + merge_degree_info (:85-131): touched vertices carry their out-degree inside the
  // sub-graph of the touched vertices' rows, flagged with bit 30; the others their out-degree in the whole graph
  std::vector<uint32_t> whole(V), sub(V);
  OutDegrees(indptr, indices, V, nullptr, threads, whole.data());
  OutDegrees(indptr, indices, V, before.data(), threads, sub.data());
  ParallelFor(V, threads, [&](int, size_t b, size_t e) {
    for (size_t v = b; v < e; ++v)
      if (before[v]) whole[v] = sub[v] | 0x40000000u;
  });
  RankByKeyDescending(whole.data(), V, threads, rank);  // randkingNodesToFile (:133-165)
  return 0;
}

extern "C" int fgnn_rt_rank_fake_optimal(const uint32_t *indptr, const uint32_t *indices, size_t V,
                                         const uint32_t *train_set, size_t num_train, int fanout0, int fanout1,
                                         int order_threads, int num_threads, uint32_t *rank) {
  if (!indptr || !rank || (num_train && !train_set) || fanout0 <= 0 || fanout1 <= 0 || order_threads < 1 ||
      V >= 0xFFFFFFFFull)
    return -1;
  if (V == 0) return 0;
  for (size_t i = 0; i < num_train; ++i)
    if (train_set[i] >= V) return -1;
  const int threads = ResolveThreads(num_threads);
  std::vector<double> expectation(V, 0.0);
  // Training nodes are independent (batch_size = 1) except for the `+=` into the expectation table, whose order
  // decides the rounding: workers compute the contributions of a block of training nodes, one thread then adds
  // them in training-set order.
  const size_t block = (size_t)threads * 256;
  std::vector<std::vector<Contribution>> contrib(block);
  std::vector<TouchMap> maps(threads);
  for (size_t base = 0; base < num_train; base += block) {
    const size_t n = std::min(block, num_train - base);
    std::atomic<size_t> next{0};
    auto work = [&](int th) {
      std::vector<uint32_t> compact, cnt;
      for (;;) {
        const size_t i = next.fetch_add(1, std::memory_order_relaxed);
        if (i >= n) return;
        contrib[i].clear();
        FakeOptimalOneSeed(indptr, indices, train_set[base + i], (double)fanout0, (double)fanout1,
                           (uint32_t)order_threads, maps[th], compact, cnt, &contrib[i]);
      }
    };
    if (threads == 1 || n < 8) {
      work(0);
    } else {
      std::vector<std::thread> pool;
      for (int th = 0; th < threads; ++th) pool.emplace_back(work, th);
      for (auto &t : pool) t.join();
    }
    for (size_t i = 0; i < n; ++i)
      for (const Contribution &c : contrib[i]) expectation[c.node] += c.value;
  }
  // randkingNodesToFile (:133-165): std::greater on {expectation, id}
  std::vector<std::pair<double, uint32_t>> order(V);
  ParallelFor(V, threads, [&](int, size_t b, size_t e) {
    for (size_t v = b; v < e; ++v) order[v] = {expectation[v], (uint32_t)v};
  });
  ParallelSort(order, threads, std::greater<std::pair<double, uint32_t>>());
  ParallelFor(V, threads, [&](int, size_t b, size_t e) {
    for (size_t i = b; i < e; ++i) rank[i] = order[i].second;
  });
  return 0;
}
