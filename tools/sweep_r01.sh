#!/bin/bash
# GPU sweep of the fast bulk kernel: resident CTAs per SM x register cap variants (round 1)
set -x
B="python bench.py --records 200000000 --steps 3 --no-cpu --no-e2e"
$B --opt ctas_per_sm=2 > gpurun_out/sw_c2.log 2>&1
$B --opt ctas_per_sm=1 > gpurun_out/sw_c1.log 2>&1
TEC_LIB=$PWD/te_counter_b200/libtecount_c3.so $B --opt ctas_per_sm=3 > gpurun_out/sw_c3.log 2>&1
TEC_LIB=$PWD/te_counter_b200/libtecount_c4.so $B --opt ctas_per_sm=4 > gpurun_out/sw_c4.log 2>&1
$B --opt ctas_per_sm=2 --opt stab_shift=10 > gpurun_out/sw_c2_s10.log 2>&1
$B --workload bulk_se --opt ctas_per_sm=2 > gpurun_out/sw_se_c2.log 2>&1
$B --opt bulk_algo=0 --records 100000000 > gpurun_out/sw_exact.log 2>&1
tail -n 1 gpurun_out/sw_*.log | cut -c 1-400
