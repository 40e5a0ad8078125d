#!/bin/bash
# how the fast kernel's time depends on the share of the cell table a launch touches (reads of 1/w of the genome)
set -x
O=gpurun_out/r02f
mkdir -p $O
for sh in 0,1 0,2 1,2 0,3 0,4 0,8; do
timeout 300 python tools/bulk_sweep.py --workload bulk_pe --records 250000000 --steps 5 --shard $sh --configs "bulk_mode=1" >> $O/shard_pe.jsonl 2>> $O/shard_pe.err
done
cut -c1-300 $O/shard_pe.jsonl
for sh in 0,1 0,2 0,4; do
timeout 300 python tools/bulk_sweep.py --workload bulk_se --records 250000000 --steps 5 --shard $sh --configs "bulk_mode=1" >> $O/shard_se.jsonl 2>> $O/shard_se.err
done
cut -c1-300 $O/shard_se.jsonl
M=gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
for sh in 0,2 0,4; do
timeout 300 ncu --metrics $M --clock-control none -k regex:bulk2 -s 6 -c 2 --csv --log-file $O/ncu_shard_${sh/,/_}.csv python tools/bulk_sweep.py --workload bulk_pe --records 250000000 --steps 2 --shard $sh --configs "bulk_mode=1" > $O/ncu_${sh/,/_}.log 2>&1
done
