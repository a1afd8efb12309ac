#!/bin/bash
# Round 2, GPU call P (1 GPU): rows with the slot block (precomputed visited-filter slots, `PH` kernels): GPU suite incl.
# tests/test_gpu_slot_block.py, A/B of BANG_B200_PREHASH=1 / 0 on the same box, one ncu --set full capture of the PH kernel.
mkdir -p gpurun_out
echo "== GPU suite"
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -6
short() { python -c "
import json,sys
for ln in sys.stdin:
    if not ln.startswith('{'): continue
    j=json.loads(ln); r=j['roofline']
    print('   value %.0f QPS  %.3f ms  e2e %.0f  recall %s  L %s grid %dx%d smem %d  frac %.4f B/q %.0f row %s | r95 L %s %.3f ms' % (j['value'], j['ms_per_step'], j['e2e']['value'], j['config']['recall_at_10'], j['config']['L_at_recall_90'], r['grid'], r['block'], r['smem_bytes'], r['frac'], r['bytes_per_query'], j['config'].get('row_bytes'), j['at_recall_95']['L'], j['at_recall_95']['ms_per_step']))
"; }
B="timeout 300 python bench.py --workload sift1m --steps 10 --warmup 3 --no-cpu-baseline --no-extra"
D="timeout 600 python bench.py --workload deep100m --n 10000000 --steps 10 --warmup 3 --no-cpu-baseline --no-extra"
echo "== sift1m, slot block (default)";  $B 2>gpurun_out/r2p_err.log | short
echo "== sift1m, plain rows";            BANG_B200_PREHASH=0 $B 2>>gpurun_out/r2p_err.log | short
echo "== sift1m, slot block, 32 warps";  BANG_B200_WARPS_PER_SM=32 $B 2>>gpurun_out/r2p_err.log | short
echo "== sift1m, slot block, 40k queries"; $B --q 40000 2>>gpurun_out/r2p_err.log | short
echo "== deep10m, slot block (default)"; $D 2>>gpurun_out/r2p_err.log | short
echo "== deep10m, plain rows";           BANG_B200_PREHASH=0 $D 2>>gpurun_out/r2p_err.log | short
echo "== deep10m, slot block, 32 warps"; BANG_B200_WARPS_PER_SM=32 $D 2>>gpurun_out/r2p_err.log | short
echo "== ncu --set full, sift1m (C2), slot-block kernel, default warps"
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base mangled -k regex:bang_search_kernelIhLi1ELi4E -s 1 -c 1 -o gpurun_out/r2p_c2 -f python profiles/prof_search.py 176 inmemory 3 > gpurun_out/r2p_ncu.log 2>&1; tail -3 gpurun_out/r2p_ncu.log
tail -3 gpurun_out/r2p_err.log
