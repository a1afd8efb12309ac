#!/bin/bash
# Round 2, GPU call R (1 GPU): two reorderings inside a hop, as separate builds of the library on one box:
#   ce   -DBANG_COMMIT_EARLY  the filter bytes are stored while the first code words travel (instead of after the distances)
#   ap   -DBANG_ADC_PAIR      two candidates per lane group and pass in the ADC loop (two independent chains)
#   ceap both
# Parity (bit-exact suites) for every build, then C2 / DEEP 10^7 timings with plain rows and with the slot block.
mkdir -p gpurun_out
short() { python -c "
import json,sys
for ln in sys.stdin:
    if not ln.startswith('{'): continue
    j=json.loads(ln); r=j['roofline']
    print('   %.3f ms  %.0f QPS  e2e %.0f  recall %s L %s row %s | r95 %.3f ms' % (j['ms_per_step'], j['value'], j['e2e']['value'], j['config']['recall_at_10'], j['config']['L_at_recall_90'], j['config'].get('row_bytes'), j['at_recall_95']['ms_per_step']))
"; }
B="timeout 300 python bench.py --workload sift1m --steps 10 --warmup 3 --no-cpu-baseline --no-extra"
D="timeout 600 python bench.py --workload deep100m --n 10000000 --steps 10 --warmup 3 --no-cpu-baseline --no-extra"
for v in default ce ap ceap; do
  if [ $v = default ]; then unset BANG_B200_LIB; else export BANG_B200_LIB=$PWD/bang-billion-scale-ann_b200/libbang_b200_$v.so; fi
  echo "==== build: $v"
  timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slot_block.py tests/test_gpu_pq_shapes.py -m gpu -q -x --timeout 600 2>&1 | tail -2
  for ph in 0 1; do
    echo "  sift1m  slot block $ph"; BANG_B200_PREHASH=$ph $B 2>>gpurun_out/r2r_err.log | short
    echo "  deep10m slot block $ph"; BANG_B200_PREHASH=$ph $D 2>>gpurun_out/r2r_err.log | short
  done
done
tail -2 gpurun_out/r2r_err.log
