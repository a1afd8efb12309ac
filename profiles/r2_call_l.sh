#!/bin/bash
# Round 2, GPU call L (1 GPU): parity suite + timings with run-time toggles.  usage: r2_call_l.sh "<env settings>;<env settings>;..."
mkdir -p gpurun_out
echo "== GPU parity tests"
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pq_shapes.py tests/test_gpu_device_paths.py -m gpu -q -x --timeout 600 2>&1 | tail -4
IFS=';' read -ra SETS <<< "$1"
for sh in "sift1m 176" "deep10m 36"; do set -- $sh
  for e in "${SETS[@]}"; do
    echo "== $1 L=$2: [$e]"; env $e timeout 600 python profiles/prof_search.py $2 inmemory 4 $1 2>&1 | grep -E "^run|Error" | tail -2
  done
done
