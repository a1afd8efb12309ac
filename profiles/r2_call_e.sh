#!/bin/bash
# Round 2, GPU call E (1 GPU): the GPU suite with the data-preparation kernels, then the two bench arms exactly as the
# driver runs them (reference arm first: it prepares the C4 index in a child process; then this repo's arm on the cache).
mkdir -p gpurun_out
echo "== GPU suite"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -15
echo "== bench --impl reference (default workload)"
S=$SECONDS; timeout 1500 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2e_ref.json 2> gpurun_out/r2e_ref.err
echo "wall $((SECONDS-S)) s"; grep "\[bench\]" gpurun_out/r2e_ref.err | cut -c1-400; cut -c1-900 gpurun_out/r2e_ref.json
echo "== bench (default workload)"
S=$SECONDS; timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2e_b200.json 2> gpurun_out/r2e_b200.err
echo "wall $((SECONDS-S)) s"; tail -3 gpurun_out/r2e_b200.err | cut -c1-300; grep "\[bench" gpurun_out/r2e_b200.err | cut -c1-300; cut -c1-3000 gpurun_out/r2e_b200.json
df -h /dev/shm | tail -1
