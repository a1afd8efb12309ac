#!/bin/bash
# Round 2, GPU call F (1 GPU): GPU suite on the slimmed kernel (3.5k instructions, no spills at 16/32 warps) + the
# BANG_B200_TIMERS=2 build; C2 and DEEP-10M timings against the number of resident query warps per SM.
mkdir -p gpurun_out
echo "== GPU suite"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -8
short() { python -c "
import json,sys
for ln in sys.stdin:
    if not ln.startswith('{'): continue
    j=json.loads(ln); r=j['roofline']
    print('   value %.0f QPS  %.3f ms  e2e %.0f  recall %s  L %s grid %dx%d smem %d  frac %.4f B/q %.0f | r95 L %s %.0f QPS %.3f ms' % (j['value'], j['ms_per_step'], j['e2e']['value'], j['config']['recall_at_10'], j['config']['L_at_recall_90'], r['grid'], r['block'], r['smem_bytes'], r['frac'], r['bytes_per_query'], j['at_recall_95']['L'], j['at_recall_95']['value'], j['at_recall_95']['ms_per_step']))
"; }
B="timeout 300 python bench.py --workload sift1m --steps 5 --warmup 3 --no-cpu-baseline --no-extra"
for w in 16 24 32; do echo "== sift1m, $w query warps per SM"; BANG_B200_WARPS_PER_SM=$w $B 2>gpurun_out/r2f_err.log | short; done
echo "== sift1m, 16 warps, code prefetch off"; BANG_B200_WARPS_PER_SM=16 BANG_B200_CODE_PREFETCH=0 $B 2>>gpurun_out/r2f_err.log | short
for q in 40000; do echo "== sift1m, 16 warps, batch of $q queries"; BANG_B200_WARPS_PER_SM=16 $B --q $q 2>>gpurun_out/r2f_err.log | short; done
D="timeout 600 python bench.py --workload deep100m --n 10000000 --steps 5 --warmup 3 --no-cpu-baseline --no-extra"
for w in 16 24 32; do echo "== deep 10M, $w warps"; BANG_B200_WARPS_PER_SM=$w $D 2>>gpurun_out/r2f_err.log | short; done
echo "== deep 10M, 32 warps, code prefetch off"; BANG_B200_WARPS_PER_SM=32 BANG_B200_CODE_PREFETCH=0 $D 2>>gpurun_out/r2f_err.log | short
echo "== gist1m (C3)"; timeout 300 python bench.py --workload gist1m --steps 5 --warmup 3 --no-cpu-baseline --no-extra 2>>gpurun_out/r2f_err.log | short
echo "== ncu, sift1m 16 warps"
BANG_B200_WARPS_PER_SM=16 timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base mangled -k regex:bang_search_kernelIhLi1ELi4E -s 1 -c 1 -o gpurun_out/r2f_w16 -f python profiles/prof_search.py 176 inmemory 3 > gpurun_out/r2f_ncu.log 2>&1; tail -2 gpurun_out/r2f_ncu.log
tail -3 gpurun_out/r2f_err.log
