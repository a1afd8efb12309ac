#!/bin/bash
# Round 2, GPU call A (1 GPU): box probe, the whole GPU suite with the former opt-in cases on by default, the
# reference forks run on the fixtures (id-level goldens), parity + bench of the experimental builds.
#   gpurun --timeout 1500 -- 'bash profiles/r2_call_a.sh > gpurun_out/r2a.log 2>&1; tail -60 gpurun_out/r2a.log'
P=$PWD/bang-billion-scale-ann_b200
mkdir -p gpurun_out
echo "== box"; nproc; free -g | head -2; df -h /tmp /dev/shm . | cat; nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
echo "== reference forks -> goldens"
timeout 600 python tests/golden/make_ref_forks_golden.py run 2>&1 | tail -40
echo "== GPU suite, default library"
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -25
for v in eager tma plain; do
  [ -f $P/libbang_b200_$v.so ] || { echo "variant $v not built"; continue; }
  echo "== parity with libbang_b200_$v.so"
  BANG_B200_LIB=$P/libbang_b200_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pq_shapes.py -q -m gpu -k 'bit_exact or spill or general' 2>&1 | tail -6
  echo "== bench with libbang_b200_$v.so"
  BANG_B200_LIB=$P/libbang_b200_$v.so timeout 300 python bench.py --workload sift1m --steps 5 --warmup 3 --no-cpu-baseline --no-ref-cuda 2>/dev/null | cut -c1-600
done
echo "== bench, default library"
timeout 300 python bench.py --workload sift1m --steps 5 --warmup 3 --no-cpu-baseline --no-ref-cuda 2>/dev/null | cut -c1-600
