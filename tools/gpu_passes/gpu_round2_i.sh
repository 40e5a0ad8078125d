#!/bin/bash
set -x
O=gpurun_out/r02i
mkdir -p $O
timeout 900 python tools/bulk_sweep.py --workload bulk_pe --configs "bulk_mode=1;bulk_mode=9;bulk_mode=13;bulk_mode=5;bulk_mode=11" > $O/sweep_pe.jsonl 2> $O/sweep_pe.err
cut -c1-330 $O/sweep_pe.jsonl
timeout 900 python tools/bulk_sweep.py --workload bulk_se --configs "bulk_mode=1;bulk_mode=9;bulk_mode=13" > $O/sweep_se.jsonl 2> $O/sweep_se.err
cut -c1-330 $O/sweep_se.jsonl
M=gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 300 ncu --metrics $M --clock-control none -k regex:bulk -s 9 -c 3 --csv --log-file $O/launches_pe_deep.csv python tools/bulk_sweep.py --workload bulk_pe --steps 2 --configs "bulk_mode=9" > $O/ncu_pe.log 2>&1
timeout 300 ncu --metrics $M --clock-control none -k regex:bulk -s 18 -c 6 --csv --log-file $O/launches_se_deep.csv python tools/bulk_sweep.py --workload bulk_se --steps 2 --configs "bulk_mode=9" > $O/ncu_se.log 2>&1
