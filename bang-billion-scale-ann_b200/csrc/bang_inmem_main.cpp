// bang — CLI with the argument set of the reference's Inmemory / Exactdistance forks (BANG_Inmemory/README.md:63-69,
// parANN.cu:79-93):
//   bang <pq_pivots.bin> <pq_compressed.bin> <disk.bin> <query.bin> <chunk_offsets.bin> <centroid.bin> <gt.bin>
//        <num_queries> <tb_parent> <tb_pqtable> <tb_dist> <tb_filter> <k> <omp_threads> <flags> [medoid [L]]
// The four thread-block sizes, the OpenMP thread count and the capability bit flags tuned the reference's separate
// kernels; they are accepted and ignored (one fused kernel, geometry chosen by the library).  The reference takes N, D,
// MEDOID and L from parANN.h at compile time (parANN.h:38-158); here N and the chunk count come from the
// pq_compressed header (exact mode: N from the graph file size), D from the query file, the element type from
// BANG_B200_DTYPE (uint8 | int8 | float, default uint8), and medoid / L from the two optional trailing arguments or
// BANG_B200_MEDOID / BANG_B200_L.  BANG_B200_MODE=exact selects the Exactdistance fork's search (default inmemory).
// Output follows parANN.cu:108,821-822,854,880,898: "<medoid>\t<Q>", wall clock (microsec), throughput, the
// `Ls  Recall@k` table, then "Try Next run ? [y|n]" read from stdin (end of input = n).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <set>
#include <string>
#include <vector>

#include "bang_b200.h"

static bool read_header(const char* path, int32_t* a, int32_t* b) {
  std::ifstream in(path, std::ios::binary);
  if (!in.is_open()) return false;
  in.read((char*)a, 4);
  in.read((char*)b, 4);
  return (bool)in;
}

static double recall_at_k(unsigned nq, const uint32_t* gt_ids, const float* gt_d, unsigned dim_gs, const uint64_t* res, unsigned k) {
  double total = 0;  // k-recall@k with distance ties in the ground truth (calculate_recall, parANN.cu / test_driver.cpp:43-93)
  for (unsigned i = 0; i < nq; ++i) {
    unsigned t = k - 1;
    while (t < dim_gs && gt_d[(size_t)i * dim_gs + t] == gt_d[(size_t)i * dim_gs + k - 1]) ++t;
    std::set<uint32_t> gt(gt_ids + (size_t)i * dim_gs, gt_ids + (size_t)i * dim_gs + t), rs;
    for (unsigned j = 0; j < k; ++j) rs.insert((uint32_t)res[(size_t)i * k + j]);
    for (uint32_t v : gt) total += rs.count(v);
  }
  return total / nq * (100.0 / k);
}

int main(int argc, char** argv) {
  if (argc < 16) {
    printf("Usage: %s <pq_pivots.bin> <pq_compressed.bin> <disk.bin> <query.bin> <chunk_offsets.bin> <centroid.bin> <groundtruth.bin> "
           "<num_queries> <tb_parent> <tb_pqtable> <tb_dist> <tb_filter> <k> <omp_threads> <flags> [medoid [L]]\n", argv[0]);
    return 1;
  }
  const char* mode_env = getenv("BANG_B200_MODE");
  const bool exact = mode_env && !strncmp(mode_env, "exact", 5);
  const int Q = atoi(argv[8]), k = atoi(argv[13]);
  const char* med_s = argc > 16 ? argv[16] : getenv("BANG_B200_MEDOID");
  const char* l_s = argc > 17 ? argv[17] : getenv("BANG_B200_L");
  if (!med_s) {
    printf("Error.. the medoid id is needed (16th argument or BANG_B200_MEDOID; the reference compiles it in, parANN.h MEDOID)\n");
    return 1;
  }
  const uint64_t medoid = strtoull(med_s, nullptr, 10);
  const int L = l_s ? atoi(l_s) : 152;  // parANN.h:94-105 default for SIFT1B
  int32_t qn = 0, D = 0, N = 0, m = 0, gn = 0, gdim = 0;
  if (!read_header(argv[4], &qn, &D)) { printf("Error.. Could not open the file4.."); return 1; }
  if (!exact && !read_header(argv[2], &N, &m)) { printf("Error.. Could not open the file2.."); return 1; }
  if (Q <= 0 || Q > qn || k <= 0 || L < k) { printf("Error.. num_queries / k / L out of range (queries in file: %d)\n", qn); return 1; }
  const std::string dt = getenv("BANG_B200_DTYPE") ? getenv("BANG_B200_DTYPE") : "uint8";  // the reference: typedef in parANN.h
  const bang_dtype_t dtype = dt == "float" ? BANG_DT_FLOAT : dt == "int8" ? BANG_DT_INT8 : BANG_DT_UINT8;
  const size_t esz = dtype == BANG_DT_FLOAT ? 4 : 1;
  if (exact) {  // no PQ files in this mode: N = graph file size / entry length (vector + degree + 64 neighbour slots)
    std::ifstream g(argv[3], std::ios::binary | std::ios::ate);
    if (!g.is_open()) { printf("Error.. Could not open the file.."); return 1; }
    N = (int32_t)((uint64_t)g.tellg() / ((uint64_t)D * esz + 4 + 4 * BANG_B200_MAX_R));
  }

  std::vector<unsigned char> queries((size_t)Q * D * esz);
  {
    std::ifstream in(argv[4], std::ios::binary);
    in.seekg(8);
    in.read((char*)queries.data(), queries.size());
  }
  std::vector<uint32_t> gt_ids;
  std::vector<float> gt_d;
  bool have_gt = read_header(argv[7], &gn, &gdim) && gn >= Q && gdim >= k;
  if (have_gt) {
    std::ifstream in(argv[7], std::ios::binary);
    gt_ids.resize((size_t)gn * gdim);
    gt_d.resize((size_t)gn * gdim);
    in.seekg(8);
    in.read((char*)gt_ids.data(), gt_ids.size() * 4);
    in.read((char*)gt_d.data(), gt_d.size() * 4);
    have_gt = (bool)in;
  }

  bang_handle_t h = nullptr;
  if (bang_b200_create(&h, dtype, exact ? BANG_MODE_EXACTDISTANCE : BANG_MODE_INMEMORY, -1) != BANG_OK) {
    printf("Error.. %s\n", bang_b200_last_error());
    return 2;
  }
  printf("%llu\t%d\n", (unsigned long long)medoid, Q);
  if (bang_b200_load_files(h, exact ? nullptr : argv[1], exact ? nullptr : argv[2], argv[3], exact ? nullptr : argv[5],
                           exact ? nullptr : argv[6], (uint64_t)N, (uint32_t)D, medoid) != BANG_OK) {
    printf("Error.. %s\n", bang_b200_last_error());
    return 2;
  }
  bang_b200_set_dists_layout(h, BANG_DISTS_QUERY_MAJOR);
  if (bang_b200_set_searchparams(h, k, L, BANG_DIST_L2) != BANG_OK || bang_b200_alloc(h, Q) != BANG_OK) {
    printf("Error.. %s\n", bang_b200_last_error());
    return 2;
  }
  std::vector<uint64_t> ids((size_t)Q * k);
  std::vector<float> dists((size_t)Q * k);
  for (;;) {
    bang_b200_init(h, Q);
    auto t0 = std::chrono::high_resolution_clock::now();
    if (bang_b200_query(h, queries.data(), Q, ids.data(), dists.data()) != BANG_OK) {
      printf("Error.. %s\n", bang_b200_last_error());
      return 2;
    }
    auto t1 = std::chrono::high_resolution_clock::now();
    const double us = std::chrono::duration<double, std::micro>(t1 - t0).count();
    printf("Wall Clock Time = %.0f microsec\n", us);
    printf("Throughput = %.2f QPS\n", Q * 1e6 / us);
    printf("Ls\tRecall@%d\n", k);
    printf("%d\t%.2f\n", L, have_gt ? recall_at_k(Q, gt_ids.data(), gt_d.data(), gdim, ids.data(), k) : 0.0);
    printf("Try Next run ? [y|n]\n");
    fflush(stdout);
    char c = 'n';
    if (!(std::cin >> c) || c != 'y') break;
  }
  bang_b200_free(h);
  bang_b200_unload(h);
  bang_b200_destroy(h);
  return 0;
}
