#!/bin/bash
# GPU sweep of the fast bulk kernel: resident CTAs per SM x register cap variants (round 1)
B="python bench.py --records 200000000 --steps 3 --no-cpu --no-e2e"
for v in 2 3 4; do
  L=$PWD/te_counter_b200/libtecount_c$v.so; [ $v = 2 ] && L=$PWD/te_counter_b200/libtecount.so
  TEC_LIB=$L $B --opt ctas_per_sm=$v > gpurun_out/sw_c$v.log 2>&1
  echo "ctas=$v $(tail -1 gpurun_out/sw_c$v.log | grep -o '"value": [0-9.]*' | head -1) $(tail -1 gpurun_out/sw_c$v.log | grep -o '"kernel_ms": [0-9.]*')"
done
