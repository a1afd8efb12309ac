#!/usr/bin/env bash
# Builds the UNMODIFIED reference (BANG_Base) from the sources where they lie under /root/reference
# into oracle/_ref/ (git-ignored, NOT gpurun-ignored: the binaries travel to the GPU box).
# No reference source is copied into this repository.  The reference's own build is cmake + legacy
# FindCUDA (BANG_Base/CMakeLists.txt:16-32); this is the equivalent two-command recipe with its flags
# plus an explicit sm_100a target (the reference passes no -arch at all).
#
# Outputs:  oracle/_ref/libbang.so     reference library (bang_search.cu)
#           oracle/_ref/ref_driver     our thin driver over the reference's public BANGSearch<T> API
#           oracle/_ref/bang_search    the reference's own CLI driver (test_driver.cpp)
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF=/root/reference/BANG_Base
OUT="$HERE/_ref"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
mkdir -p "$OUT"
"$NVCC" -Xcompiler -fopenmp -std=c++14 --compiler-options -fPIC -O3 \
    -gencode arch=compute_100a,code=sm_100a -shared "$REF/bang_search.cu" -o "$OUT/libbang.so" -lgomp
/usr/bin/g++ -O2 -std=c++14 -fopenmp -I"$REF" "$HERE/ref_driver.cpp" -o "$OUT/ref_driver" \
    -L"$OUT" -lbang -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,'$ORIGIN' -Wl,-rpath,/usr/local/cuda/lib64
/usr/bin/g++ -O2 -std=c++14 -fopenmp -I"$REF" "$REF/test_driver.cpp" -o "$OUT/bang_search" \
    -L"$OUT" -lbang -L/usr/local/cuda/lib64 -lcudart -ldl -Wl,-rpath,'$ORIGIN' -Wl,-rpath,/usr/local/cuda/lib64
echo "reference built into $OUT"
# Drop-in proof: the reference's OWN driver source compiled against THIS repo's include/bang.h and linked
# with libbang_b200.so instead of the reference library (no reference code in the link except the driver).
ROOT="$(dirname "$HERE")"
PKG="$ROOT/bang-billion-scale-ann_b200"
if [ -f "$PKG/libbang_b200.so" ]; then
  /usr/bin/g++ -O2 -std=c++14 -fopenmp -I"$ROOT/include" "$REF/test_driver.cpp" -o "$OUT/bang_search_dropin" \
      -L"$PKG" -lbang_b200 -ldl -Wl,-rpath,'$ORIGIN/../../bang-billion-scale-ann_b200'
  echo "drop-in driver built: $OUT/bang_search_dropin"
fi
