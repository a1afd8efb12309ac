// loader.cpp — see loader.h.  Plain stdio; no CUDA.
#include "loader.h"

#include <cstdio>
#include <cstring>
#include <sys/stat.h>

namespace bang {

namespace {
struct File {
  FILE* f = nullptr;
  explicit File(const std::string& p) : f(fopen(p.c_str(), "rb")) {}
  ~File() { if (f) fclose(f); }
  bool read_at(uint64_t off, void* dst, size_t n) {
    if (fseeko(f, (off_t)off, SEEK_SET) != 0) return false;
    return fread(dst, 1, n, f) == n;
  }
};
bool fail(std::string* err, const std::string& msg) {
  if (err) *err = msg;
  return false;
}
}  // namespace

bool file_size(const std::string& path, uint64_t* size, std::string* err) {
  struct stat st;
  if (stat(path.c_str(), &st) != 0) return fail(err, "cannot stat " + path);
  *size = (uint64_t)st.st_size;
  return true;
}

bool read_graph_meta(const std::string& path, GraphMeta* out, std::string* err) {
  File f(path);
  if (!f.f) return fail(err, "Could not open the Metadata File: " + path);
  uint8_t raw[32];
  if (!f.read_at(0, raw, 32)) return fail(err, "Metadata file shorter than 32 bytes: " + path);
  memcpy(&out->medoid, raw + 0, 8);
  memcpy(&out->entry_len, raw + 8, 8);
  memcpy(&out->dtype, raw + 16, 4);
  memcpy(&out->D, raw + 20, 4);
  memcpy(&out->R, raw + 24, 4);
  memcpy(&out->N, raw + 28, 4);
  return true;
}

bool read_bin_header(const std::string& path, uint32_t elem_size, uint32_t* npts, uint32_t* dim, std::string* err) {
  uint64_t sz = 0;
  if (!file_size(path, &sz, err)) return false;
  File f(path);
  if (!f.f) return fail(err, "Could not open " + path);
  int32_t hdr[2];
  if (!f.read_at(0, hdr, 8)) return fail(err, "short bin header: " + path);
  *npts = (uint32_t)hdr[0];
  *dim = (uint32_t)hdr[1];
  const uint64_t expect = (uint64_t)*npts * *dim * elem_size + 8;
  if (sz != expect)
    return fail(err, "File size mismatch for " + path + ": actual " + std::to_string(sz) + " expected " + std::to_string(expect));
  return true;
}

bool read_pq_pivots_new(const std::string& path, uint32_t D, uint32_t m, PQHost* out, std::string* err) {
  File f(path);
  if (!f.f) return fail(err, "Could not open the PQ Pivots File: " + path);
  uint32_t nsec = 0;
  if (!f.read_at(0, &nsec, 4)) return fail(err, "short PQ pivots file: " + path);
  if (nsec != 4) return fail(err, "PQ Pivots File does not contain the required # of sub-sections");
  uint64_t off[4];
  if (!f.read_at(8, off, 32)) return fail(err, "short PQ pivots offset table: " + path);
  out->pivots.resize((size_t)256 * D);
  out->centroid.resize(D);
  out->chunk_off.resize(m + 1);
  if (!f.read_at(off[0] + 8, out->pivots.data(), out->pivots.size() * 4)) return fail(err, "short pivots section: " + path);
  if (!f.read_at(off[1] + 8, out->centroid.data(), (size_t)D * 4)) return fail(err, "short centroid section: " + path);
  if (!f.read_at(off[2] + 8, out->chunk_off.data(), (size_t)(m + 1) * 4)) return fail(err, "short chunk offsets section: " + path);
  return true;
}

bool read_pq_pivots_old(const std::string& pivots, const std::string& centroid, const std::string& chunk_offsets,
                        uint32_t D, PQHost* out, std::string* err) {
  uint32_t n = 0, d = 0;
  if (!read_bin_header(pivots, 4, &n, &d, err)) return false;
  if (n != 256 || d != D) return fail(err, "pivots bin is not 256 x D: " + pivots);
  out->pivots.resize((size_t)256 * D);
  { File f(pivots); if (!f.f || !f.read_at(8, out->pivots.data(), out->pivots.size() * 4)) return fail(err, "cannot read " + pivots); }
  if (!read_bin_header(centroid, 4, &n, &d, err)) return false;
  if ((uint64_t)n * d != D) return fail(err, "centroid bin does not hold D floats: " + centroid);
  out->centroid.resize(D);
  { File f(centroid); if (!f.f || !f.read_at(8, out->centroid.data(), (size_t)D * 4)) return fail(err, "cannot read " + centroid); }
  if (!read_bin_header(chunk_offsets, 4, &n, &d, err)) return false;
  const uint64_t cnt = (uint64_t)n * d;
  if (cnt < 2) return fail(err, "chunk offsets bin too small: " + chunk_offsets);
  out->chunk_off.resize(cnt);
  { File f(chunk_offsets); if (!f.f || !f.read_at(8, out->chunk_off.data(), cnt * 4)) return fail(err, "cannot read " + chunk_offsets); }
  return true;
}

}  // namespace bang
