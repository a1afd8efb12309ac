// bang_preprocess — DiskANN `<X>_disk.index` -> BANG `<X>_disk.bin` + `<X>_disk_metadata.bin`.
//
// Host-only replacement of BANG_Base/bang_preprocess.py (which issues one read() per byte), same arguments:
//   bang_preprocess <X_disk.index> <X_disk.bin> <dimension> <datatype: 0 int8, 1 uint8, 2 float> <degree R> [sector bytes = 4096]
//
// What the script does, and this does too (bang_preprocess.py:26-116):
//   * sector 0 of the .index file is DiskANN's metadata block: two int32 (skipped), then uint64 npts, ndims, medoid,
//     max_node_len, nnodes_per_sector, three uint64 that are skipped, and the file size  (:28-63);
//   * every further sector holds nnodes_per_sector node entries back to back from the start of the sector
//     (:73-79): vector T[dim], uint32 degree, uint32 neighbours[];
//   * an entry is written as vector, degree, the `degree` neighbour ids SORTED ASCENDING (:95-97), then the unused
//     neighbour slots up to R copied as they are (:99-102);  degree 0 or > R aborts the conversion (:86-89);
//   * the metadata file gets medoid (u64), max_node_len (u64) — both copied from the .index header —, datatype,
//     dimension and R from the command line (u32 each) and, last, the number of nodes converted (u32) (:41-51,110);
//     its name is the output name with "_metadata" inserted before the 4-character extension (:24).
// Beyond the script: entries are located at the .index file's own stride (max_node_len), so an index built with a
// larger degree bound than the R asked for converts correctly (the script would read garbage), a dimension that
// contradicts the header is reported, and short files are errors instead of silent truncation.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

extern "C" {
// 0 = ok; < 0 = error, text in bang_preprocess_last_error().  nodes_out (may be NULL) receives the node count.
int bang_preprocess_index(const char* index_path, const char* out_bin_path, uint32_t dim, uint32_t datatype, uint32_t degree,
                          uint32_t sector_len, uint64_t* nodes_out);
const char* bang_preprocess_last_error(void);
}

static thread_local std::string g_err;
static int fail(const std::string& m) { g_err = m; return -1; }
const char* bang_preprocess_last_error(void) { return g_err.c_str(); }

int bang_preprocess_index(const char* index_path, const char* out_bin_path, uint32_t dim, uint32_t datatype, uint32_t degree,
                          uint32_t sector_len, uint64_t* nodes_out) {
  if (!index_path || !out_bin_path) return fail("null path");
  if (datatype > 2) return fail("datatype must be 0 (int8), 1 (uint8) or 2 (float)");
  if (dim == 0 || degree == 0) return fail("dimension and degree must be positive");
  if (sector_len == 0) sector_len = 4096;  // SECTORLEN, bang_preprocess.py:20
  const std::string out(out_bin_path);
  if (out.size() < 4) return fail("output name needs a 4-character extension (e.g. .bin)");
  const std::string meta = out.substr(0, out.size() - 4) + "_metadata" + out.substr(out.size() - 4);
  const size_t esz = datatype == 2 ? 4 : 1;
  const size_t vec_bytes = (size_t)dim * esz, entry_out = vec_bytes + 4 + (size_t)degree * 4;

  FILE* f = fopen(index_path, "rb");
  if (!f) return fail(std::string("cannot open ") + index_path);
  struct Closer { FILE* p; ~Closer() { if (p) fclose(p); } } cf{f};
  uint64_t h[9];
  unsigned char skip[8];
  if (fread(skip, 1, 8, f) != 8 || fread(h, 8, 9, f) != 9) return fail("index file shorter than its metadata block");
  const uint64_t npts = h[0], ndims = h[1], medoid = h[2], max_node_len = h[3], per_sector = h[4], file_size = h[8];
  if (ndims != dim)
    fprintf(stderr, "bang_preprocess: warning: the index header says %llu dimensions, the command line %u\n", (unsigned long long)ndims, dim);
  if (per_sector == 0) return fail("index stores less than one node per sector (multi-sector nodes are not supported, as in the reference script)");
  if (max_node_len < entry_out)
    return fail("index entries (" + std::to_string(max_node_len) + " B) are shorter than dimension*size + 4 + 4*degree = " + std::to_string(entry_out));
  if (per_sector * max_node_len > sector_len) return fail("nodes per sector x entry length exceeds the sector length");
  const uint64_t sectors = file_size / sector_len;
  if (sectors < 1) return fail("file size field smaller than one sector");

  FILE* w = fopen(out_bin_path, "wb");
  if (!w) return fail(std::string("cannot create ") + out_bin_path);
  Closer cw{w};
  FILE* wm = fopen(meta.c_str(), "wb");
  if (!wm) return fail("cannot create " + meta);
  Closer cm{wm};

  std::vector<unsigned char> sector(sector_len), outbuf;
  outbuf.reserve((size_t)per_sector * entry_out);
  uint64_t nodes = 0;
  for (uint64_t s = 1; s < sectors && nodes < npts; ++s) {
    if (fseeko(f, (off_t)(s * sector_len), SEEK_SET) != 0) return fail("seek failed");
    const size_t got = fread(sector.data(), 1, sector_len, f);
    outbuf.clear();
    for (uint64_t j = 0; j < per_sector && nodes < npts; ++j) {
      const size_t off = (size_t)j * max_node_len;
      if (off + entry_out > got) return fail("index file ends inside node " + std::to_string(nodes));
      const unsigned char* e = sector.data() + off;
      uint32_t d;
      memcpy(&d, e + vec_bytes, 4);
      if (d == 0 || d > degree)
        return fail("node " + std::to_string(nodes) + " has degree " + std::to_string(d) + " (must be 1.." + std::to_string(degree) + ")");
      const size_t o = outbuf.size();
      outbuf.insert(outbuf.end(), e, e + entry_out);
      uint32_t* nb = reinterpret_cast<uint32_t*>(outbuf.data() + o + vec_bytes + 4);
      std::sort(nb, nb + d);
      ++nodes;
    }
    if (!outbuf.empty() && fwrite(outbuf.data(), 1, outbuf.size(), w) != outbuf.size()) return fail("write failed");
  }
  const uint32_t tail[4] = {datatype, dim, degree, (uint32_t)nodes};
  if (fwrite(&medoid, 8, 1, wm) != 1 || fwrite(&max_node_len, 8, 1, wm) != 1 || fwrite(tail, 4, 4, wm) != 4) return fail("write failed");
  if (nodes_out) *nodes_out = nodes;
  return 0;
}

#ifdef BANG_PREPROCESS_MAIN
int main(int argc, char** argv) {
  if (argc != 6 && argc != 7) {
    printf("Usage : %s <path to DiskANN graph index file (.index)> <path to store the o/p file for use by BANG search (.bin)> "
           "<dataset dimension> <dataset datatype: 0 -> int8, 1 -> uint8, 2 -> float> <degree (i.e. R) of the DiskANN graph index> "
           "[sector length, default 4096]\n", argv[0]);
    return 1;
  }
  uint64_t nodes = 0;
  const int rc = bang_preprocess_index(argv[1], argv[2], (uint32_t)atoi(argv[3]), (uint32_t)atoi(argv[4]), (uint32_t)atoi(argv[5]),
                                       argc == 7 ? (uint32_t)atoi(argv[6]) : 4096u, &nodes);
  if (rc != 0) {
    printf("Error: %s\n", bang_preprocess_last_error());
    return 2;
  }
  printf("Total # of Nodes Discovered = %llu\n", (unsigned long long)nodes);
  return 0;
}
#endif
