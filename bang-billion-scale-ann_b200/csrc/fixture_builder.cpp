// Host-side Vamana (R-bounded RobustPrune) graph builder used to create test/bench fixtures in the
// reference's `_disk.bin` format.  The reference does not build indices: it consumes DiskANN's
// `build_disk_index` output (README.md:46-58, "-R 64 -L 200"), which is not available offline, so the
// fixtures are produced here.  This is fixture tooling (host only, no CUDA); it is not on the search
// path.  Algorithm: Vamana as published (greedy search from the medoid, RobustPrune(alpha), reverse
// edge insertion with re-prune), two passes (alpha = 1 then alpha), batch-parallel with OpenMP.
//
// Output contract (what BANG reads, bang_preprocess.py:81-110): per node a degree in [1, R] and the
// first `degree` neighbour ids sorted ascending.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <random>
#include <utility>
#include <vector>
#include <omp.h>

namespace {

struct Builder {
  const float* x;  // [N][D]
  uint32_t N, D, R, L;
  float alpha;
  uint32_t medoid;
  uint32_t slackR;
  std::vector<std::vector<uint32_t>> adj;
  std::vector<std::atomic_flag> locks;

  Builder(const float* x_, uint32_t N_, uint32_t D_, uint32_t R_, uint32_t L_)
      : x(x_), N(N_), D(D_), R(R_), L(L_), alpha(1.f), medoid(0), slackR((uint32_t)(R_ * 1.3f)), adj(N_), locks(N_) {
    for (auto& l : locks) l.clear();
    for (auto& a : adj) a.reserve(slackR + 1);
  }

  inline float dist(const float* a, const float* b) const {
    float s = 0.f;
    for (uint32_t j = 0; j < D; ++j) {
      float d = a[j] - b[j];
      s += d * d;
    }
    return s;
  }
  inline const float* vec(uint32_t i) const { return x + (size_t)i * D; }
  void lock(uint32_t i) { while (locks[i].test_and_set(std::memory_order_acquire)) {} }
  void unlock(uint32_t i) { locks[i].clear(std::memory_order_release); }

  void find_medoid() {
    std::vector<double> c(D, 0.0);
    for (uint32_t i = 0; i < N; ++i)
      for (uint32_t j = 0; j < D; ++j) c[j] += vec(i)[j];
    std::vector<float> cf(D);
    for (uint32_t j = 0; j < D; ++j) cf[j] = (float)(c[j] / N);
    float best = INFINITY;
    uint32_t bi = 0;
#pragma omp parallel
    {
      float lb = INFINITY;
      uint32_t li = 0;
#pragma omp for nowait
      for (int64_t i = 0; i < (int64_t)N; ++i) {
        float d = dist(vec((uint32_t)i), cf.data());
        if (d < lb) { lb = d; li = (uint32_t)i; }
      }
#pragma omp critical
      if (lb < best || (lb == best && li < bi)) { best = lb; bi = li; }
    }
    medoid = bi;
  }

  struct Scratch {
    std::vector<uint32_t> stamp;
    uint32_t epoch = 0;
    std::vector<std::pair<float, uint32_t>> beam;  // sorted ascending, size <= L
    std::vector<uint8_t> expanded;
    std::vector<std::pair<float, uint32_t>> visited;
    std::vector<uint32_t> nbuf;
  };

  // Greedy (best-first) search from the medoid; collects every expanded node with its distance.
  void greedy(const float* q, Scratch& s) {
    if (s.stamp.size() != N) s.stamp.assign(N, 0);
    if (++s.epoch == 0) { std::fill(s.stamp.begin(), s.stamp.end(), 0); s.epoch = 1; }
    s.beam.clear(); s.expanded.clear(); s.visited.clear();
    s.beam.emplace_back(dist(q, vec(medoid)), medoid);
    s.expanded.push_back(0);
    s.stamp[medoid] = s.epoch;
    for (;;) {
      size_t pi = 0;
      while (pi < s.beam.size() && s.expanded[pi]) ++pi;
      if (pi == s.beam.size()) break;
      s.expanded[pi] = 1;
      uint32_t p = s.beam[pi].second;
      s.visited.push_back(s.beam[pi]);
      lock(p);
      s.nbuf.assign(adj[p].begin(), adj[p].end());
      unlock(p);
      for (uint32_t nb : s.nbuf) {
        if (s.stamp[nb] == s.epoch) continue;
        s.stamp[nb] = s.epoch;
        float d = dist(q, vec(nb));
        if (s.beam.size() == L && d >= s.beam.back().first) continue;
        std::pair<float, uint32_t> e(d, nb);
        size_t pos = std::upper_bound(s.beam.begin(), s.beam.end(), e) - s.beam.begin();
        s.beam.insert(s.beam.begin() + pos, e);
        s.expanded.insert(s.expanded.begin() + pos, 0);
        if (s.beam.size() > L) { s.beam.pop_back(); s.expanded.pop_back(); }
      }
    }
  }

  // RobustPrune(p, cand, alpha, R): cand holds (dist to p, id), duplicates and p itself allowed.
  void robust_prune(uint32_t p, std::vector<std::pair<float, uint32_t>>& cand, std::vector<uint32_t>& out) {
    std::sort(cand.begin(), cand.end());
    cand.erase(std::unique(cand.begin(), cand.end(), [](auto& a, auto& b) { return a.second == b.second; }), cand.end());
    out.clear();
    std::vector<uint8_t> dead(cand.size(), 0);
    for (size_t i = 0; i < cand.size() && out.size() < R; ++i) {
      if (dead[i] || cand[i].second == p) continue;
      uint32_t s = cand[i].second;
      out.push_back(s);
      for (size_t j = i + 1; j < cand.size(); ++j) {
        if (dead[j]) continue;
        float dsj = dist(vec(s), vec(cand[j].second));
        if (alpha * dsj <= cand[j].first) dead[j] = 1;
      }
    }
  }

  void insert_pass(const std::vector<uint32_t>& order) {
#pragma omp parallel
    {
      Scratch s;
      std::vector<std::pair<float, uint32_t>> cand;
      std::vector<uint32_t> pruned, tmp;
#pragma omp for schedule(dynamic, 64)
      for (int64_t oi = 0; oi < (int64_t)order.size(); ++oi) {
        uint32_t p = order[oi];
        greedy(vec(p), s);
        cand = s.visited;
        lock(p);
        for (uint32_t nb : adj[p]) cand.emplace_back(dist(vec(p), vec(nb)), nb);
        unlock(p);
        robust_prune(p, cand, pruned);
        lock(p);
        adj[p] = pruned;
        unlock(p);
        for (uint32_t j : pruned) {
          bool need_prune = false;
          lock(j);
          if (std::find(adj[j].begin(), adj[j].end(), p) == adj[j].end()) {
            adj[j].push_back(p);
            if (adj[j].size() > slackR) { need_prune = true; tmp = adj[j]; }
          }
          unlock(j);
          if (need_prune) {
            cand.clear();
            for (uint32_t nb : tmp) cand.emplace_back(dist(vec(j), vec(nb)), nb);
            std::vector<uint32_t> pj;
            robust_prune(j, cand, pj);
            lock(j);
            adj[j] = pj;
            unlock(j);
          }
        }
      }
    }
  }

  void finalize() {
#pragma omp parallel
    {
      std::vector<std::pair<float, uint32_t>> cand;
      std::vector<uint32_t> pj;
#pragma omp for schedule(dynamic, 256)
      for (int64_t i = 0; i < (int64_t)N; ++i) {
        auto& a = adj[i];
        if (a.size() > R) {
          cand.clear();
          for (uint32_t nb : a) cand.emplace_back(dist(vec((uint32_t)i), vec(nb)), nb);
          robust_prune((uint32_t)i, cand, pj);
          a = pj;
        }
        if (a.empty()) a.push_back((uint32_t)i == medoid ? (medoid + 1) % N : medoid);
        std::sort(a.begin(), a.end());
      }
    }
  }
};

}  // namespace

extern "C" int bang_fixture_build_vamana(const void* data, int dtype, uint32_t N, uint32_t D, uint32_t R, uint32_t L,
                                         float alpha, uint32_t seed, int nthreads, int passes, uint32_t* out_deg,
                                         uint32_t* out_nbrs, uint64_t* out_medoid) {
  if (N < 2 || R == 0) return -1;
  if (nthreads > 0) omp_set_num_threads(nthreads);
  std::vector<float> xf((size_t)N * D);
  if (dtype == 2) {
    std::memcpy(xf.data(), data, xf.size() * sizeof(float));
  } else if (dtype == 1) {
    const uint8_t* p = (const uint8_t*)data;
    for (size_t i = 0; i < xf.size(); ++i) xf[i] = (float)p[i];
  } else {
    const int8_t* p = (const int8_t*)data;
    for (size_t i = 0; i < xf.size(); ++i) xf[i] = (float)p[i];
  }
  Builder b(xf.data(), N, D, R, L);
  b.find_medoid();
  std::vector<uint32_t> order(N);
  std::iota(order.begin(), order.end(), 0u);
  std::mt19937 rng(seed);
  for (int pass = 0; pass < passes; ++pass) {
    b.alpha = (pass == passes - 1) ? alpha : 1.0f;
    std::shuffle(order.begin(), order.end(), rng);
    b.insert_pass(order);
  }
  b.finalize();
  for (uint32_t i = 0; i < N; ++i) {
    out_deg[i] = (uint32_t)b.adj[i].size();
    uint32_t* row = out_nbrs + (size_t)i * R;
    std::memset(row, 0, sizeof(uint32_t) * R);
    std::copy(b.adj[i].begin(), b.adj[i].end(), row);
  }
  *out_medoid = b.medoid;
  return 0;
}
