#!/bin/bash
# Round 2, GPU call C (1 GPU): ncu of the v5 kernel at 16 and 32 query warps per SM (C2, L = 176), the 24-warp / 80-register
# variant, and the code-prefetch A/B where the PQ codes do not fit in L2 (DEEP shape, 10^7 points: 320 MB of codes).
mkdir -p gpurun_out
python profiles/prof_search.py 176 inmemory 1 > /dev/null 2>&1   # builds the C2 index once (/tmp/bang_prof)
for w in 16 32; do
  BANG_B200_WARPS_PER_SM=$w timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base mangled \
    -k regex:bang_search_kernelIhLi1ELi4E -s 1 -c 1 -o gpurun_out/r2c_w$w -f python profiles/prof_search.py 176 inmemory 3 > gpurun_out/r2c_ncu_w$w.log 2>&1
  tail -2 gpurun_out/r2c_ncu_w$w.log
done
B="timeout 300 python bench.py --workload sift1m --steps 5 --warmup 3 --no-cpu-baseline --no-extra --L 176 --L95 256"
short() { python -c "
import json,sys
for ln in sys.stdin:
    if not ln.startswith('{'): continue
    j=json.loads(ln); r=j['roofline']
    print('   value %.0f QPS  %.3f ms  e2e %.0f  recall %s  L %s grid %dx%d smem %d  frac %.4f B/q %.0f | r95 L %s %.0f QPS %.3f ms' % (j['value'], j['ms_per_step'], j['e2e']['value'], j['config']['recall_at_10'], j['config']['L_at_recall_90'], r['grid'], r['block'], r['smem_bytes'], r['frac'], r['bytes_per_query'], j['at_recall_95']['L'], j['at_recall_95']['value'], j['at_recall_95']['ms_per_step']))
"; }
for w in 16 24 32; do echo "== sift1m, $w query warps per SM"; BANG_B200_WARPS_PER_SM=$w $B 2>gpurun_out/r2c_err.log | short; done
D="timeout 600 python bench.py --workload deep100m --n 10000000 --steps 5 --warmup 3 --no-cpu-baseline --no-extra"
for w in 16 24 32; do
  for pf in 1 0; do echo "== deep 10M, $w warps, code prefetch $pf"; BANG_B200_WARPS_PER_SM=$w BANG_B200_CODE_PREFETCH=$pf $D 2>>gpurun_out/r2c_err.log | short; done
done
tail -5 gpurun_out/r2c_err.log
