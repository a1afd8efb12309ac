// ref_driver — thin driver over the UNMODIFIED reference's public API (BANG_Base/bang.h:36-87).
// Test infrastructure: it is compiled against /root/reference/BANG_Base/bang.h and linked with
// oracle/_ref/libbang.so (see build_ref.sh).  It runs one search at a fixed worklist length through
// the reference's own call sequence (test_driver.cpp:342,421-435,535,553) and dumps the returned ids
// so the oracle and the sm_100a path can be compared with the real reference on a B200.
//
// usage: ref_driver <index_prefix> <query.bin> <Q> <k> <L> <uint8|int8|float> <out_ids.bin> [runs]
// out:   u64 ids[Q][k] of the LAST run; prints "RUN <i> <ms>" per run on stdout.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>
#include "bang.h"

template <typename T>
static int run(int argc, char** argv) {
  const int Q = atoi(argv[3]), k = atoi(argv[4]), L = atoi(argv[5]);
  const int runs = argc > 8 ? atoi(argv[8]) : 1;
  BANGSearch<T> bang;
  if (!bang.bang_load(argv[1])) { fprintf(stderr, "bang_load failed\n"); return 2; }
  std::ifstream in(argv[2], std::ios::binary);
  if (!in.is_open()) { fprintf(stderr, "cannot open query file\n"); return 2; }
  int npts = 0, dim = 0;
  in.read((char*)&npts, 4);
  in.read((char*)&dim, 4);
  if (Q > npts) { fprintf(stderr, "query file holds %d < %d queries\n", npts, Q); return 2; }
  std::vector<T> queries((size_t)Q * dim);
  in.read((char*)queries.data(), sizeof(T) * queries.size());
  std::vector<result_ann_t> ids((size_t)Q * k);
  std::vector<float> dists((size_t)Q * k);
  bang.bang_set_searchparams(k, L, ENUM_DIST_L2);
  bang.bang_alloc(Q);
  for (int r = 0; r < runs; ++r) {
    bang.bang_init(Q);
    auto t0 = std::chrono::high_resolution_clock::now();
    bang.bang_query(queries.data(), Q, ids.data(), dists.data());
    auto t1 = std::chrono::high_resolution_clock::now();
    printf("RUN %d %.3f\n", r, std::chrono::duration<double, std::milli>(t1 - t0).count());
  }
  bang.bang_free();
  bang.bang_unload();
  FILE* f = fopen(argv[7], "wb");
  if (!f) { fprintf(stderr, "cannot write %s\n", argv[7]); return 2; }
  std::vector<unsigned long long> out(ids.begin(), ids.end());
  fwrite(out.data(), sizeof(unsigned long long), out.size(), f);
  fclose(f);
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 8) {
    fprintf(stderr, "usage: %s <prefix> <query.bin> <Q> <k> <L> <uint8|int8|float> <out_ids.bin> [runs]\n", argv[0]);
    return 1;
  }
  std::string dt(argv[6]);
  if (dt == "uint8") return run<uint8_t>(argc, argv);
  if (dt == "int8") return run<int8_t>(argc, argv);
  return run<float>(argc, argv);
}
