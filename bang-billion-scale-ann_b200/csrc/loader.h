// loader.h — host-side readers for the reference's index file formats.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace bang {

struct GraphMeta {        // `_disk_metadata.bin`, packed 32 bytes (bang_search.cuh:42-50)
  uint64_t medoid = 0;
  uint64_t entry_len = 0;
  int32_t dtype = 0;
  uint32_t D = 0, R = 0, N = 0;
};

struct PQHost {
  std::vector<float> pivots;       // [256][D] as stored in the file
  std::vector<float> centroid;     // [D]
  std::vector<uint32_t> chunk_off; // [m+1]
};

// All return false and fill `err` on failure (the reference returns false from bang_load, or exits).
bool file_size(const std::string& path, uint64_t* size, std::string* err);
bool read_graph_meta(const std::string& path, GraphMeta* out, std::string* err);
// bin header: int32 npts, int32 dim; checks size == 8 + npts*dim*elem (load_bin_impl, bang_search.cuh:299-311)
bool read_bin_header(const std::string& path, uint32_t elem_size, uint32_t* npts, uint32_t* dim, std::string* err);
// new 4-section `_pq_pivots.bin` (bang_search.cu:246-296)
bool read_pq_pivots_new(const std::string& path, uint32_t D, uint32_t m, PQHost* out, std::string* err);
// old three-file layout (BANG_Inmemory/parANN.cu:146-147,216,221)
bool read_pq_pivots_old(const std::string& pivots, const std::string& centroid, const std::string& chunk_offsets,
                        uint32_t D, PQHost* out, std::string* err);

}  // namespace bang
