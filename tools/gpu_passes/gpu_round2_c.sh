#!/bin/bash
set -x
O=gpurun_out/r02c
mkdir -p $O
./tools/microbench2 > $O/microbench2.txt 2>&1
cat $O/microbench2.txt
timeout 900 python tools/bulk_sweep.py --workload bulk_pe --configs "bulk_mode=1;bulk_mode=5;stab_shift=11,bulk_mode=1;stab_shift=11,bulk_mode=5" > $O/sweep_pe.jsonl 2> $O/sweep_pe.err
cut -c1-330 $O/sweep_pe.jsonl
timeout 900 python tools/bulk_sweep.py --workload bulk_se --configs "bulk_mode=1;bulk_mode=5" > $O/sweep_se.jsonl 2> $O/sweep_se.err
cut -c1-330 $O/sweep_se.jsonl
M=gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct,smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct,smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct
for mode in 1 5; do
timeout 600 ncu --metrics $M --clock-control none -k regex:bulk -s 9 -c 3 --csv --log-file $O/ncu_bulk_mode$mode.csv python tools/bulk_sweep.py --workload bulk_pe --records 200000000 --steps 2 --configs "bulk_mode=$mode" > $O/ncu_mode$mode.log 2>&1
done
( time timeout 1200 python bench.py --steps 20 --warmup 5 > $O/bench_all.json 2> $O/bench_all.err ) 2> $O/bench_all.time
tail -3 $O/bench_all.err; cat $O/bench_all.time
head -c 600 $O/bench_all.json
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err ) 2> $O/bench_ref.time
cat $O/bench_ref.time; head -c 400 $O/bench_ref.json
ls -la $O
