#!/bin/bash
# Round 2, GPU call Q (1 GPU): speculative L2 prefetch of the next node's row one hop ahead (BANG_B200_ROW_PREFETCH), crossed with
# the slot block (BANG_B200_PREHASH), C2 and DEEP 10^7 shapes, same box; parity suite first.
mkdir -p gpurun_out
echo "== GPU suite"
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -4
short() { python -c "
import json,sys
for ln in sys.stdin:
    if not ln.startswith('{'): continue
    j=json.loads(ln); r=j['roofline']
    print('   value %.0f QPS  %.3f ms  e2e %.0f  recall %s  L %s  frac %.4f row %s | r95 L %s %.3f ms' % (j['value'], j['ms_per_step'], j['e2e']['value'], j['config']['recall_at_10'], j['config']['L_at_recall_90'], r['frac'], j['config'].get('row_bytes'), j['at_recall_95']['L'], j['at_recall_95']['ms_per_step']))
"; }
B="timeout 300 python bench.py --workload sift1m --steps 10 --warmup 3 --no-cpu-baseline --no-extra"
D="timeout 600 python bench.py --workload deep100m --n 10000000 --steps 10 --warmup 3 --no-cpu-baseline --no-extra"
for ph in 1 0; do for rp in 1 0; do
  echo "== sift1m  slot block $ph  row prefetch $rp"; BANG_B200_PREHASH=$ph BANG_B200_ROW_PREFETCH=$rp $B 2>>gpurun_out/r2q_err.log | short
done; done
for ph in 1 0; do for rp in 1 0; do
  echo "== deep10m slot block $ph  row prefetch $rp"; BANG_B200_PREHASH=$ph BANG_B200_ROW_PREFETCH=$rp $D 2>>gpurun_out/r2q_err.log | short
done; done
echo "== gist1m (C3, Exactdistance) row prefetch 1 / 0"
for rp in 1 0; do BANG_B200_ROW_PREFETCH=$rp timeout 300 python bench.py --workload gist1m --steps 10 --warmup 3 --no-cpu-baseline --no-extra 2>>gpurun_out/r2q_err.log | short; done
echo "== ncu --set full, sift1m (C2), plain rows + row prefetch"
BANG_B200_PREHASH=0 timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base mangled -k regex:bang_search_kernelIhLi1ELi4E -s 1 -c 1 -o gpurun_out/r2q_c2 -f python profiles/prof_search.py 176 inmemory 3 > gpurun_out/r2q_ncu.log 2>&1; tail -2 gpurun_out/r2q_ncu.log
tail -3 gpurun_out/r2q_err.log
