#!/bin/bash
# Builds the reference's Inmemory and Exactdistance forks for ONE fixture: the id-level pin of those two storage
# modes (tests/golden/make_ref_forks_golden.py runs the binaries on a GPU box; tests/test_oracle.py compares).
#
# The forks compile N, D, MEDOID, INDEX_ENTRY_LEN, the element type, L and the chunk count in (parANN.h) and do not
# link as shipped, so — unlike oracle/build_ref.sh, which compiles BANG_Base where it lies — this works
# on a TEMPORARY COPY under /tmp (never inside the repository): the header gets one extra dataset block for the
# fixture (the Exactdistance fork's own compile.sh fills its header skeleton with sed in the same way), the three BFS
# hooks the Inmemory fork calls but never defines become empty functions in a separate file, and a one-line
# boost/dynamic_bitset.hpp stands in for the unused Boost include.  The forks print only a recall figure, so the copy
# also gets ONE added statement: right after the program's own device-to-host copy of the result ids
# (BANG_Inmemory/parANN.cu:733-734, BANG_Exactdistance/parANN.cu:824-825) the host array `nearestNeighbours`
# ([k][Q], unsigned) is written to the file named by $BANG_DUMP_IDS.  Nothing on the search path is touched.
# Only binaries go to oracle/_ref/.
#
# usage: oracle/build_ref_forks.sh <tag> <uint8_t|int8_t|float> <N> <D> <medoid> <L> <chunks>
#   e.g. oracle/build_ref_forks.sh fx_u8 uint8_t 3000 32 1424 32 8
# run (on a GPU box), same 15 arguments as `bang`:
#   oracle/_ref/bang_inmem_<tag>_L<L> <pq_pivots> <pq_compressed> <disk.bin> <query> <chunk_offsets> <centroid> <gt> <Q> 1 256 512 256 <k> 64 0
set -euo pipefail
REF=${BANG_REFERENCE:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
[ $# -eq 7 ] || { sed -n 2,18p "$0"; exit 1; }
TAG=$1; TYPE=$2; N=$3; D=$4; MEDOID=$5; L=$6; CHUNKS=$7
[ -d "$REF/BANG_Inmemory" ] || { echo "reference not mounted at $REF"; exit 0; }
case $TYPE in float) ESZ=4;; *) ESZ=1;; esac
ENTRY=$((D * ESZ + 4 + 256))
mkdir -p "$OUT"
BLOCK="#ifdef BANGFIXTURE\ntypedef $TYPE datatype_t;\n#define INDEX_ENTRY_LEN ($ENTRY)\n#define D $D\n#define MEDOID $MEDOID\n#define N $N\n#define NUMTHREADS_COMPUTEPARENT 1\n#endif\n"

build_fork() {  # <fork dir> <output name> <header source>
  local W=/tmp/bang_ref_forks/$2
  rm -rf "$W"; mkdir -p "$W/boost"
  cp -r "$REF/$1/main.cu" "$REF/$1/parANN.cu" "$REF/$1/utils" "$W/"
  cp "$3" "$W/parANN.h"
  echo '#include <list>' > "$W/boost/dynamic_bitset.hpp"
  ( cd "$W"
    # dataset selection: the first "#define <DATASET>", "#define L ..." and "#define CHUNKS ..." lines of the header
    python3 - "$L" "$CHUNKS" "$BLOCK" <<'EOF'
import re, sys
L, chunks, block = sys.argv[1], sys.argv[2], sys.argv[3].replace("\\n", "\n")
s = open("parANN.h").read()
m = re.search(r"^#define\s+(DATABASE_PLACE_HOLDER|[A-Z0-9]+)\s*\n#define L\s+\S+.*\n#define CHUNKS\s+\S+.*\n", s, re.M)
assert m, "dataset selection lines not found"
s = s[:m.start()] + f"#define BANGFIXTURE\n#define L {L}\n#define CHUNKS {chunks}\n" + block + s[m.end():]
open("parANN.h", "w").write(s)
c = open("parANN.cu").read()
m = re.search(r"gpuErrchk\(cudaMemcpy\(nearestNeighbours, d_nearestNeighbours,[^;]*;\n", c)
assert m, "result copy not found"
dump = ('\t{ const char* bang_dump = getenv("BANG_DUMP_IDS"); if (bang_dump) { FILE* bang_fp = fopen(bang_dump, "wb"); '
        'fwrite(nearestNeighbours, sizeof(unsigned), (size_t)recall_at * numQueries, bang_fp); fclose(bang_fp); } }\n')
open("parANN.cu", "w").write(c[:m.end()] + dump + c[m.end():])
EOF
    EXTRA=""
    if grep -q "SetupBFS" parANN.cu && ! grep -q "^void SetupBFS" parANN.cu; then
      printf '#include <iostream>\n#include <fstream>\n#include <string>\n#include <string.h>\n#include <set>\n#include <sstream>\n#include <map>\n#include <vector>\n#include <cstdint>\n#include <assert.h>\n#include "utils/timer.h"\n#include "parANN.h"\nvoid SetupBFS(NodeIDMap&) {}\nvoid ExitBFS(NodeIDMap&) {}\nvoid bfs(unsigned, const unsigned, unsigned&, NodeIDMap&, uint8_t*) {}\n' > bfs_stubs.cu
      EXTRA=bfs_stubs.cu
    fi
    nvcc main.cu parANN.cu $EXTRA -Xcompiler -fopenmp -std=c++14 -I. -I./utils -arch=sm_100a -O3 -w -o "$OUT/$2" )
  echo "built $OUT/$2"
}

build_fork BANG_Exactdistance "bang_exact_${TAG}_L${L}" "$REF/BANG_Exactdistance/code/parANN_skeleton.h"
build_fork BANG_Inmemory "bang_inmem_${TAG}_L${L}" "$REF/BANG_Inmemory/parANN.h"
