#!/bin/bash
# Round 2, GPU call B (1 GPU): whole GPU suite on the v5 kernel (8-byte filter blocks, up to 32 query warps per SM,
# expanded-node log in global memory), then C2 timings against resident warps per SM, batch size and the code prefetch,
# and one ncu capture.
#   gpurun --timeout 1500 -- 'bash profiles/r2_call_b.sh > gpurun_out/r2b.log 2>&1; tail -80 gpurun_out/r2b.log'
mkdir -p gpurun_out
echo "== GPU suite"
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -30
B="timeout 300 python bench.py --workload sift1m --steps 5 --warmup 3 --no-cpu-baseline --no-ref-cuda --L 176 --L95 256"
short() { python -c "
import json,sys
for ln in sys.stdin:
    if not ln.startswith('{'): continue
    j=json.loads(ln); r=j['roofline']
    print('   value %.0f QPS  %.3f ms  e2e %.0f  recall %s  grid %dx%d smem %d  frac %.4f  | r95 %.0f QPS %.3f ms' % (j['value'], j['ms_per_step'], j['e2e']['value'], j['config']['recall_at_10'], r['grid'], r['block'], r['smem_bytes'], r['frac'], j['at_recall_95']['value'], j['at_recall_95']['ms_per_step']))
"; }
for w in 16 20 24 28 32; do echo "== sift1m, $w query warps per SM"; BANG_B200_WARPS_PER_SM=$w $B 2>/dev/null | short; done
echo "== sift1m, code prefetch off"; BANG_B200_CODE_PREFETCH=0 $B 2>/dev/null | short
for q in 20000 40000; do echo "== sift1m, batch of $q queries"; $B --q $q 2>/dev/null | short; done
echo "== ncu (one launch, full set)"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:bang_search_kernel -s 2 -c 1 -o gpurun_out/r2b_full -f python profiles/prof_search.py 176 inmemory 4 > gpurun_out/r2b_ncu.log 2>&1
tail -3 gpurun_out/r2b_ncu.log
