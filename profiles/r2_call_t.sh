#!/bin/bash
# Round 2, GPU call T (1 GPU): L2 residency of the visited filters.  Three builds on one box:
#   s        the kernel of call S (filter stores / atomics with the default L2 policy, filter loads also for unused neighbour slots)
#   plainst  filter loads only for real neighbours
#   default  + evict-last policy on the filter's clear, reservations and byte stores
# Parity suite on the default, timings and DRAM bytes (one ncu pass each) on the C2 and DEEP 10^7 shapes.
mkdir -p gpurun_out
echo "== GPU suite (default build)"
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -4
short() { python -c "
import json,sys
for ln in sys.stdin:
    if not ln.startswith('{'): continue
    j=json.loads(ln)
    print('   %.3f ms  %.0f QPS  e2e %.0f  recall %s L %s | r95 %.3f ms' % (j['ms_per_step'], j['value'], j['e2e']['value'], j['config']['recall_at_10'], j['config']['L_at_recall_90'], j['at_recall_95']['ms_per_step']))
"; }
B="timeout 300 python bench.py --workload sift1m --steps 10 --warmup 3 --no-cpu-baseline --no-extra"
D="timeout 600 python bench.py --workload deep100m --n 10000000 --steps 10 --warmup 3 --no-cpu-baseline --no-extra"
for v in s plainst default; do
  if [ $v = default ]; then unset BANG_B200_LIB; else export BANG_B200_LIB=$PWD/bang-billion-scale-ann_b200/libbang_b200_$v.so; fi
  echo "==== build: $v"
  echo "  sift1m";  $B 2>>gpurun_out/r2t_err.log | short
  echo "  deep10m"; $D 2>>gpurun_out/r2t_err.log | short
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read_evict_last_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read_evict_last_lookup_miss.sum,gpu__time_duration.sum --clock-control none --kernel-name-base mangled -k regex:bang_search_kernelIfLi1ELi4E -s 6 -c 1 --csv --log-file gpurun_out/r2t_dram_$v.csv python bench.py --workload deep100m --n 10000000 --steps 2 --warmup 3 --no-cpu-baseline --no-extra > /dev/null 2>&1
  grep bang_search gpurun_out/r2t_dram_$v.csv | awk -F'","' '{print "     ", $(NF-2), $(NF-1), $NF}'
done
tail -2 gpurun_out/r2t_err.log
