// bang.h — C++ face of the B200-native BANG search path.
//
// Same public surface as the reference's header (BANG_Base/bang.h:20-87): `result_ann_t`, `DistFunc`,
// `MAX_L`, `MIPS_EXTRA_DIM` and the pimpl class `BANGSearch<T>` with the seven life-cycle methods, so a
// program written against the reference (e.g. its test_driver.cpp) compiles and links unchanged against
// libbang_b200.so.  Every method forwards to the C ABI in bang_b200.h.  Differences from the reference:
//   * the storage/search mode (Base / Inmemory / Exactdistance — three separate programs in the
//     reference) is chosen per instance with bang_set_mode() or the BANG_B200_MODE environment
//     variable (base | inmemory | exact), default Base;
//   * CUDA or file errors never call exit(): bang_load returns false, the other methods report through
//     bang_last_error() and leave the outputs untouched.
#ifndef BANG_H_
#define BANG_H_

#include <cstdint>

#define MAX_L 512  // upper bound of the search worklist length

typedef unsigned long result_ann_t;  // ids are returned as 64-bit integers (big-ann-benchmarks convention)

typedef enum _DistFunc {
  ENUM_DIST_L2 = 0,  // squared Euclidean distance
  ENUM_DIST_MIPS,    // maximum inner product, searched as L2 over one extra dimension
} DistFunc;
#define MIPS_EXTRA_DIM (1)  // the index carries one more dimension than a MIPS query

typedef enum _BangMode { ENUM_MODE_BASE = 0, ENUM_MODE_INMEMORY = 1, ENUM_MODE_EXACTDISTANCE = 2 } BangMode;

template <typename T>
class BANGSearch {
  void* m_pImpl;

 public:
  BANGSearch();
  virtual ~BANGSearch();

  // Loads <prefix>_pq_pivots.bin, _pq_compressed.bin, _disk.bin and _disk_metadata.bin into HBM.
  bool bang_load(char* indexfile_path_prefix);
  void bang_alloc(int numQueries);
  void bang_init(int numQueries);
  void bang_set_searchparams(int recall, int worklist_length, DistFunc nDistFunc = ENUM_DIST_L2);
  // query_array: host T[num_queries][D]; nearestNeighbours: host [num_queries][recall] ids, nearest first;
  // nearestNeighbours_dist: host float[recall][num_queries] squared distances (rank-major, as the reference).
  void bang_query(T* query_array, int num_queries, result_ann_t* nearestNeighbours, float* nearestNeighbours_dist);
  void bang_free();
  void bang_unload();

  // extensions
  bool bang_set_mode(BangMode mode);  // before bang_load
  const char* bang_last_error() const;
  void* bang_c_handle() const;        // the underlying bang_handle_t
};

extern template class BANGSearch<float>;
extern template class BANGSearch<uint8_t>;
extern template class BANGSearch<int8_t>;

#endif  // BANG_H_
