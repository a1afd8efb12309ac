#!/bin/bash
# Round 2, GPU call J (1 GPU): the rewritten 32-chunk ADC path (padded [256][32][4] pivot table, residual slices in
# registers per hop, one PRMT + IMAD per table address, packed fp32x2 subtractions): parity suite + timings on the C2
# shape and on C4's shape at 10^7 points, against the scalar-subtraction build.
mkdir -p gpurun_out
echo "== GPU parity tests"
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pq_shapes.py tests/test_gpu_device_paths.py -m gpu -q -x --timeout 600 2>&1 | tail -5
for sh in "sift1m 176" "deep10m 36"; do set -- $sh
  echo "== $1 L=$2: default build"; timeout 600 python profiles/prof_search.py $2 inmemory 4 $1 2>&1 | grep -E "^run|Error" | tail -3
  for v in $VARIANTS; do
    echo "== $1 L=$2: $v"; BANG_B200_LIB=$PWD/bang-billion-scale-ann_b200/libbang_b200_$v.so timeout 600 python profiles/prof_search.py $2 inmemory 4 $1 2>&1 | grep -E "^run|Error" | tail -3
  done
done
