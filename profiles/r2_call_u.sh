#!/bin/bash
# Round 2, GPU call U (1 GPU): where do the DRAM bytes of the big-index kernel come from?  DEEP 10^7 shape (6.4 GB of rows and
# codes, nothing L2-resident but the filters): L2 fetch granularity 32 / 64 / 128 B (cudaLimitMaxL2FetchGranularity) crossed with
# the speculative code prefetch on / off; time (bench) and DRAM bytes of one launch (ncu).  C2 and C3 timings for the granularity.
mkdir -p gpurun_out
python - <<'PY'
import ctypes
rt = ctypes.CDLL("libcudart.so")
v = ctypes.c_size_t(0); print("cudaDeviceGetLimit(MaxL2FetchGranularity) ->", rt.cudaDeviceGetLimit(ctypes.byref(v), 5), v.value)
PY
short() { python -c "
import json,sys
for ln in sys.stdin:
    if not ln.startswith('{'): continue
    j=json.loads(ln)
    print('   %.3f ms  %.0f QPS  e2e %.0f  recall %s L %s | r95 %.3f ms' % (j['ms_per_step'], j['value'], j['e2e']['value'], j['config']['recall_at_10'], j['config']['L_at_recall_90'], j['at_recall_95']['ms_per_step']))
"; }
D="python bench.py --workload deep100m --n 10000000 --steps 10 --warmup 3 --no-cpu-baseline --no-extra"
for g in default 32 128; do for cp in 1 0; do
  if [ $g = default ]; then unset BANG_B200_L2_FETCH; else export BANG_B200_L2_FETCH=$g; fi
  export BANG_B200_CODE_PREFETCH=$cp
  echo "== deep10m  L2 fetch $g  code prefetch $cp"; timeout 600 $D 2>>gpurun_out/r2u_err.log | short
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --kernel-name-base mangled -k regex:bang_search_kernelIfLi1ELi4E -s 6 -c 1 --csv --log-file gpurun_out/r2u_dram_${g}_$cp.csv $D --steps 2 > /dev/null 2>&1
  grep bang_search gpurun_out/r2u_dram_${g}_$cp.csv | awk -F'","' '{print "     ", $(NF-2), $(NF-1), $NF}'
done; done
unset BANG_B200_CODE_PREFETCH
for g in default 32; do
  if [ $g = default ]; then unset BANG_B200_L2_FETCH; else export BANG_B200_L2_FETCH=$g; fi
  echo "== sift1m  L2 fetch $g"; timeout 300 python bench.py --workload sift1m --steps 10 --warmup 3 --no-cpu-baseline --no-extra 2>>gpurun_out/r2u_err.log | short
  echo "== gist1m  L2 fetch $g"; timeout 300 python bench.py --workload gist1m --steps 10 --warmup 3 --no-cpu-baseline --no-extra 2>>gpurun_out/r2u_err.log | short
done
tail -2 gpurun_out/r2u_err.log
