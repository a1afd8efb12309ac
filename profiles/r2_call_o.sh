#!/bin/bash
# Round 2, GPU call O (1 GPU): the filter/expand instruction diet (32-bit filter offsets from the kernel parameter, spill checks behind one
# warp vote, medoid slots from the host, statistics in registers, code-row addresses by one multiply-add, loop-free neighbour scan):
# parity suite, then A/B against the previous kernel (libbang_b200_r2n.so = commit 10d3dff) on the same box, then one ncu --set full capture.
mkdir -p gpurun_out
echo "== GPU suite (new kernel)"
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -6
short() { python -c "
import json,sys
for ln in sys.stdin:
    if not ln.startswith('{'): continue
    j=json.loads(ln); r=j['roofline']
    print('   value %.0f QPS  %.3f ms  e2e %.0f  recall %s  L %s grid %dx%d smem %d  frac %.4f B/q %.0f | r95 L %s %.3f ms' % (j['value'], j['ms_per_step'], j['e2e']['value'], j['config']['recall_at_10'], j['config']['L_at_recall_90'], r['grid'], r['block'], r['smem_bytes'], r['frac'], r['bytes_per_query'], j['at_recall_95']['L'], j['at_recall_95']['ms_per_step']))
"; }
B="timeout 300 python bench.py --workload sift1m --steps 10 --warmup 3 --no-cpu-baseline --no-extra"
D="timeout 600 python bench.py --workload deep100m --n 10000000 --steps 10 --warmup 3 --no-cpu-baseline --no-extra"
OLD=$PWD/bang-billion-scale-ann_b200/libbang_b200_r2n.so
echo "== sift1m new";            $B 2>gpurun_out/r2o_err.log | short
echo "== sift1m old (10d3dff)";  BANG_B200_LIB=$OLD $B 2>>gpurun_out/r2o_err.log | short
echo "== sift1m new, 32 warps";  BANG_B200_WARPS_PER_SM=32 $B 2>>gpurun_out/r2o_err.log | short
echo "== sift1m new, 40k queries"; $B --q 40000 2>>gpurun_out/r2o_err.log | short
echo "== deep10m new";           $D 2>>gpurun_out/r2o_err.log | short
echo "== deep10m old (10d3dff)"; BANG_B200_LIB=$OLD $D 2>>gpurun_out/r2o_err.log | short
echo "== deep10m new, 32 warps"; BANG_B200_WARPS_PER_SM=32 $D 2>>gpurun_out/r2o_err.log | short
echo "== ncu --set full, sift1m (C2), new kernel, default warps"
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base mangled -k regex:bang_search_kernelIhLi1ELi4E -s 1 -c 1 -o gpurun_out/r2o_c2 -f python profiles/prof_search.py 176 inmemory 3 > gpurun_out/r2o_ncu.log 2>&1; tail -3 gpurun_out/r2o_ncu.log
tail -3 gpurun_out/r2o_err.log
