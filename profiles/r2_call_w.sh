#!/bin/bash
# Round 2, GPU call W (1 GPU): evidence for the round's final kernel: GPU suite, compute-sanitizer memcheck on two parity
# cases, both bench arms as the driver runs them (reference first), one ncu pass over the bench command with per-launch
# duration and DRAM bytes (launch list + traffic of the search kernel), one `ncu --set full` capture of the C4 kernel.
mkdir -p gpurun_out
echo "== GPU suite"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -5
echo "== compute-sanitizer memcheck (C1-shape parity, both row layouts)"
timeout 600 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 9 python -m pytest tests/test_gpu_slot_block.py -m gpu -q -x -k "equal_plain_rows and inmemory and 20" --timeout 500 > gpurun_out/r2w_memcheck.log 2>&1; echo "exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/r2w_memcheck.log | tail -5
echo "== C2: code prefetch off (default) / on"
for cp in 0 1; do BANG_B200_CODE_PREFETCH=$cp timeout 300 python bench.py --workload sift1m --steps 10 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        j=json.loads(ln); print('   %.3f ms  %.0f QPS | r95 %.3f ms' % (j['ms_per_step'], j['value'], j['at_recall_95']['ms_per_step']))
"; done
echo "== bench --impl reference (default workload; prepares the C4 cache)"; S=$SECONDS
timeout 1500 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2w_ref.json 2> gpurun_out/r2w_ref.err
echo "exit $? wall $((SECONDS-S)) s"; cut -c1-400 gpurun_out/r2w_ref.json
echo "== bench N = 1 (full line with other_configs)"; S=$SECONDS
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2w_b200_1gpu.json 2> gpurun_out/r2w_b200_1gpu.err
echo "exit $? wall $((SECONDS-S)) s"; grep -E "Error|Traceback" gpurun_out/r2w_b200_1gpu.err | head; cut -c1-1400 gpurun_out/r2w_b200_1gpu.json
echo "== ncu: per-launch duration + DRAM bytes of the bench command"; S=$SECONDS
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2w_launches.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2w_launch_run.log 2>&1
echo "wall $((SECONDS-S)) s"; wc -l gpurun_out/r2w_launches.csv; grep bang_search gpurun_out/r2w_launches.csv | tail -6 | cut -c1-60,200-420
echo "== ncu --set full, C4 kernel (default workload)"; S=$SECONDS
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base mangled -k regex:bang_search_kernelIfLi1ELi4E -s 6 -c 1 -o gpurun_out/r2w_c4 -f python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2w_ncu.log 2>&1; echo "wall $((SECONDS-S)) s"; grep -E "PROF|Error" gpurun_out/r2w_ncu.log | tail -3
ls -la gpurun_out | grep r2s
