#!/bin/bash
# Round 2, GPU call AA (1 GPU): the next 16 candidates' code words requested before the current batch is evaluated (lists longer
# than 16 are the rule on the DEEP shape): parity suite on the new build, A/B against the kernel of call W on one box.
mkdir -p gpurun_out
echo "== GPU suite (new build)"
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -3
short() { python -c "
import json,sys
for ln in sys.stdin:
    if not ln.startswith('{'): continue
    j=json.loads(ln)
    print('   %.3f ms  %.0f QPS  e2e %.0f  recall %s L %s | r95 %.3f ms' % (j['ms_per_step'], j['value'], j['e2e']['value'], j['config']['recall_at_10'], j['config']['L_at_recall_90'], j['at_recall_95']['ms_per_step']))
"; }
B="timeout 200 python bench.py --workload sift1m --steps 10 --warmup 3 --no-cpu-baseline --no-extra"
D="timeout 300 python bench.py --workload deep100m --n 10000000 --steps 10 --warmup 3 --no-cpu-baseline --no-extra"
for v in w default; do
  if [ $v = default ]; then unset BANG_B200_LIB; else export BANG_B200_LIB=$PWD/bang-billion-scale-ann_b200/libbang_b200_$v.so; fi
  echo "==== build: $v"
  echo "  deep10m"; $D 2>>gpurun_out/r2aa_err.log | short
  echo "  sift1m";  $B 2>>gpurun_out/r2aa_err.log | short
done
tail -2 gpurun_out/r2aa_err.log | cut -c1-200
