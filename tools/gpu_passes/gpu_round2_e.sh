#!/bin/bash
# profiles of the three legs: launch lists (serialised, cold-cache) + ncu --set full of the bulk kernels
set -x
O=gpurun_out/r02e
mkdir -p $O
Q="--no-cpu --no-e2e --file-records 0 --sc-parity-records 0"
for wl in bulk_pe bulk_se; do
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:bulk -s 12 -c 9 --csv --log-file $O/launches_$wl.csv python bench.py --workload $wl --steps 3 --warmup 3 $Q > $O/ncu_launches_$wl.log 2>&1
done
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2000 --csv --log-file $O/launches_sc.csv python bench.py --workload sc --steps 1 --warmup 3 $Q > $O/ncu_launches_sc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bulk2_fast -s 4 -c 1 -o $O/prof_fast_pe -f python bench.py --workload bulk_pe --steps 2 --warmup 3 $Q > $O/ncu_full_pe.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bulk2_ -s 8 -c 2 -o $O/prof_se -f python bench.py --workload bulk_se --steps 2 --warmup 3 $Q > $O/ncu_full_se.log 2>&1
ls -la $O
