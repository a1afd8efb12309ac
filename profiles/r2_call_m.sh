#!/bin/bash
# Round 2, GPU call M (1 GPU): throughput against the number of resident query warps per SM, batches of 10 000 and 40 000
# queries (the larger batch shows the steady state without the tail), C2 shape and C4's shape at 10^7 points.
mkdir -p gpurun_out
short() { python -c "
import json,sys
for ln in sys.stdin:
    if not ln.startswith('{'): continue
    j=json.loads(ln); r=j['roofline']
    print('   value %.0f QPS  %.3f ms  e2e %.0f  recall %s  L %s grid %dx%d smem %d  frac %.4f B/q %.0f' % (j['value'], j['ms_per_step'], j['e2e']['value'], j['config']['recall_at_10'], j['config']['L_at_recall_90'], r['grid'], r['block'], r['smem_bytes'], r['frac'], r['bytes_per_query']))
"; }
B="timeout 300 python bench.py --workload sift1m --steps 5 --warmup 3 --no-cpu-baseline --no-extra"
D="timeout 600 python bench.py --workload deep100m --n 10000000 --steps 5 --warmup 3 --no-cpu-baseline --no-extra"
for w in $1; do
  echo "== sift1m, $w warps, 10k"; BANG_B200_WARPS_PER_SM=$w $B 2>>gpurun_out/r2m_err.log | short
  echo "== sift1m, $w warps, 40k"; BANG_B200_WARPS_PER_SM=$w $B --q 40000 2>>gpurun_out/r2m_err.log | short
done
for w in $2; do
  echo "== deep10m, $w warps, 10k"; BANG_B200_WARPS_PER_SM=$w $D 2>>gpurun_out/r2m_err.log | short
  echo "== deep10m, $w warps, 40k"; BANG_B200_WARPS_PER_SM=$w $D --q 40000 2>>gpurun_out/r2m_err.log | short
done
tail -3 gpurun_out/r2m_err.log
