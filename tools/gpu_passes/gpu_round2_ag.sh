#!/bin/bash
O=gpurun_out/r02ag
mkdir -p $O
Q="--no-cpu --no-e2e --file-records 0 --sc-parity-records 0 --opt second_parts=4"
C1="python bench.py --workload bulk_se --steps 1 --warmup 3 $Q"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bulk2_pair -s 6 -c 1 -o $O/prof_pair_se -f $C1 > $O/ncu_full.log 2>&1
tail -2 $O/ncu_full.log
