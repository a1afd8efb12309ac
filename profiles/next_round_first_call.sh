#!/bin/bash
# First 1-GPU call of the next round (≈ 4-5 GPU-minutes).  Build the variants HERE first (no GPU needed, the .so files
# travel with the snapshot):
#   python -c "import bang_b200; from bang_b200 import build as b; b.build_variant('eager', ['BANG_EAGER_EXACT']); b.build_variant('tma', ['BANG_TMA_ROWS']); b.build_variant('plain', ['BANG_PLAIN_ROW_LOADS'])"
#   gpurun --timeout 900 -- 'bash profiles/next_round_first_call.sh > gpurun_out/first_call.log 2>&1; tail -40 gpurun_out/first_call.log'
P=$PWD/bang-billion-scale-ann_b200
echo "== opt-in cases on the default library"
BANG_B200_UNVERIFIED_TESTS=1 timeout 600 python -m pytest tests/test_gpu_pq_shapes.py tests/test_gpu_device_paths.py tests/test_gpu_parity.py -k 'general_chunking or device or degenerate or builder or inmemory_cli' -q -m gpu 2>&1 | tail -15
for v in eager tma plain; do
  [ -f $P/libbang_b200_$v.so ] || { echo "variant $v not built"; continue; }
  echo "== parity with libbang_b200_$v.so"
  BANG_B200_LIB=$P/libbang_b200_$v.so BANG_B200_UNVERIFIED_TESTS=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_pq_shapes.py tests/test_gpu_device_paths.py -q -m gpu 2>&1 | tail -8
  echo "== bench with libbang_b200_$v.so"
  BANG_B200_LIB=$P/libbang_b200_$v.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-cuda 2>/dev/null | cut -c1-700
done
echo "== bench, default library"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-cuda 2>/dev/null | cut -c1-700
