// bang_search — CLI driver with the reference's argument list, modes and report line (BANG_Base/test_driver.cpp):
//   bang_search <index_prefix> <query.bin> <groundtruth.bin> <num_queries> <k> <uint8|int8|float> <l2|mips> [auto | L ...]
//   bang_search <query.bin> <num_queries>            MIPS query normaliser (test_driver.cpp:280-336,566-571)
// Exactly 7 arguments = the reference's interactive mode: the worklist length is read from stdin before every
// round and "Try Next run ? [y|n]" after it (test_driver.cpp:384-401,535-543; here the loop also ends at end of
// input).  Any 8th argument = auto mode, L = k, k+12, ... <= MAX_L (:404-420); if the extra arguments are numbers
// they are taken as the list of worklist lengths instead (an addition).  Five timed runs per L, wall clock around
// bang_query only (:424-439), the same output table `L  Time  QPS  k-r@k` (:402-403,526).
// Mode: BANG_B200_MODE=base|inmemory|exact.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <set>
#include <string>
#include <vector>

#include "bang.h"

// k-recall@k with distance ties in the ground truth, as calculate_recall (test_driver.cpp:43-93)
static double recall_at_k(unsigned nq, const uint32_t* gt_ids, const float* gt_d, unsigned dim_gs, const result_ann_t* res,
                          unsigned k) {
  double total = 0;
  for (unsigned i = 0; i < nq; ++i) {
    unsigned t = k;
    if (gt_d) {
      t = k - 1;
      while (t < dim_gs && gt_d[(size_t)i * dim_gs + t] == gt_d[(size_t)i * dim_gs + k - 1]) ++t;
    }
    std::set<uint32_t> gt(gt_ids + (size_t)i * dim_gs, gt_ids + (size_t)i * dim_gs + t);
    std::set<uint32_t> rs;
    for (unsigned j = 0; j < k; ++j) rs.insert((uint32_t)res[(size_t)i * k + j]);
    for (uint32_t v : gt) total += rs.count(v);
  }
  return total / nq * (100.0 / k);
}

static bool load_truth(const char* path, std::vector<uint32_t>* ids, std::vector<float>* d, unsigned* nq, unsigned* dim) {
  std::ifstream in(path, std::ios::binary | std::ios::ate);
  if (!in.is_open()) return false;
  const size_t size = in.tellg();
  in.seekg(0);
  int32_t hdr[2];
  in.read((char*)hdr, 8);
  *nq = hdr[0];
  *dim = hdr[1];
  if (size != 8 + 8ull * *nq * *dim) {
    fprintf(stderr, "Error. Truthset file size mismatch\n");
    return false;
  }
  ids->resize((size_t)*nq * *dim);
  d->resize((size_t)*nq * *dim);
  in.read((char*)ids->data(), ids->size() * 4);
  in.read((char*)d->data(), d->size() * 4);
  return true;
}

template <typename T>
static int run_anns(int argc, char** argv) {
  BANGSearch<T> bang;
  if (!bang.bang_load(argv[1])) {
    printf("Error: Bang_load failed\n");
    return -1;
  }
  const int numQueries = atoi(argv[4]);
  std::ifstream in(argv[2], std::ios::binary);
  if (!in.is_open()) {
    printf("Error.. Could not open the Query File: %s\n", argv[2]);
    return -1;
  }
  int npts = 0, dim = 0;
  in.read((char*)&npts, 4);
  in.read((char*)&dim, 4);
  if (numQueries > npts) {
    printf("Error.. query file holds only %d queries\n", npts);
    return -1;
  }
  std::vector<T> queries((size_t)numQueries * dim);
  in.read((char*)queries.data(), sizeof(T) * queries.size());
  const int k = atoi(argv[5]);
  const DistFunc dist = !strcmp(argv[7], "mips") ? ENUM_DIST_MIPS : ENUM_DIST_L2;
  std::vector<uint32_t> gt_ids;
  std::vector<float> gt_d;
  unsigned gt_n = 0, gt_dim = 0;
  if (!load_truth(argv[3], &gt_ids, &gt_d, &gt_n, &gt_dim)) {
    printf("Groundtruth file could not be loaded:%s\n", argv[3]);
    return -1;
  }
  std::vector<int> Ls;
  for (int i = 8; i < argc; ++i)
    if (atoi(argv[i]) > 0) Ls.push_back(atoi(argv[i]));
  const bool interactive = argc == 8;
  if (!interactive && Ls.empty())
    for (int L = k; L <= MAX_L; L += 12) Ls.push_back(L);
  std::vector<result_ann_t> ids((size_t)numQueries * k);
  std::vector<float> dists((size_t)numQueries * k);
  auto header = [&] { printf("L\tTime \tQPS\t\t%d-r@%d\n--\t---- \t---\t\t------\n", k, k); };
  auto one_round = [&](int L) {
    bang.bang_set_searchparams(k, L, dist);
    bang.bang_alloc(numQueries);
    for (int run = 0; run < 5; ++run) {
      bang.bang_init(numQueries);
      auto t0 = std::chrono::high_resolution_clock::now();
      bang.bang_query(queries.data(), numQueries, ids.data(), dists.data());
      auto t1 = std::chrono::high_resolution_clock::now();
      const double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
      const double rec = recall_at_k(numQueries, gt_ids.data(), gt_d.data(), gt_dim, ids.data(), k);
      printf("%d\t%.2f\t%.2f\t%.2f\n", L, ms, numQueries * 1000.0 / ms, rec);
    }
    bang.bang_free();
    fflush(stdout);
  };
  if (interactive) {
    for (;;) {
      printf("Enter value of WorkList Length\n");
      fflush(stdout);
      int L = 0;
      if (!(std::cin >> L)) break;
      if (L < k) {
        printf(" Error: WorkList Length must be at least recall_at\n");
        continue;
      }
      if (L > MAX_L) {
        printf(" Error: WorkList Length must be at most %d\n", MAX_L);
        continue;
      }
      header();
      one_round(L);
      printf("Try Next run ? [y|n]\n");
      fflush(stdout);
      char c = 'n';
      if (!(std::cin >> c) || c == 'n') break;
    }
  } else {
    header();
    for (int L : Ls) {
      if (L < k) {
        printf(" Error: WorkList Length must be at least recall_at\n");
        continue;
      }
      one_round(L);
    }
  }
  bang.bang_unload();
  return 0;
}

// x -> x / |x| with one zero dimension appended, written as float to <file>_transformed (the query side of the
// MIPS -> L2 reduction; the reference does this for float queries only, test_driver.cpp:280-336)
static int normalise_queries(const char* path, int numQueries) {
  std::ifstream in(path, std::ios::binary);
  if (!in.is_open()) {
    printf("Error.. Could not open the Query File: %s\n", path);
    return -1;
  }
  int npts = 0, dim = 0;
  in.read((char*)&npts, 4);
  in.read((char*)&dim, 4);
  if (numQueries <= 0 || numQueries > npts || dim <= 0) {
    printf("Error.. query file holds %d queries of %d dimensions\n", npts, dim);
    return -1;
  }
  std::vector<float> q((size_t)numQueries * dim), out((size_t)numQueries * (dim + 1));
  in.read((char*)q.data(), sizeof(float) * q.size());
  if (!in) {
    printf("Error.. query file is shorter than its header says\n");
    return -1;
  }
  for (int i = 0; i < numQueries; ++i) {
    const float* x = q.data() + (size_t)i * dim;
    float* y = out.data() + (size_t)i * (dim + 1);
    float norm = 0.f;
    for (int j = 0; j < dim; ++j) norm += x[j] * x[j];   // float accumulation in element order, as the reference
    norm = std::sqrt(norm);
    for (int j = 0; j < dim; ++j) y[j] = x[j] / norm;
    y[dim] = 0.f;
  }
  const std::string dst = std::string(path) + "_transformed";
  std::ofstream w(dst, std::ios::binary);
  if (!w.is_open()) {
    printf("Error.. Could not write %s\n", dst.c_str());
    return -1;
  }
  const int d1 = dim + 1;
  w.write((const char*)&numQueries, 4);
  w.write((const char*)&d1, 4);
  w.write((const char*)out.data(), sizeof(float) * out.size());
  printf("Writing bin: %s\nbin: #pts = %d, #dims = %d, size = %zuB\nFinished writing bin.\n", dst.c_str(), numQueries, d1,
         out.size() * sizeof(float) + 8);
  return 0;
}

int main(int argc, char** argv) {
  if (argc == 3) return normalise_queries(argv[1], atoi(argv[2]));
  if (argc < 8) {
    printf("Usage: %s <index_prefix> <query.bin> <groundtruth.bin> <num_queries> <k> <uint8|int8|float> <l2|mips> [auto | L ...]\n"
           "       %s <query.bin> <num_queries>     (writes <query.bin>_transformed for MIPS)\n", argv[0], argv[0]);
    return 1;
  }
  const std::string dt(argv[6]);
  if (dt == "uint8") return run_anns<uint8_t>(argc, argv);
  if (dt == "int8") return run_anns<int8_t>(argc, argv);
  return run_anns<float>(argc, argv);
}
