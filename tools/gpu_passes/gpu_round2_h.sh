#!/bin/bash
# hit-queue tally: parity tests + A/B against the direct tally + launch list
set -x
O=gpurun_out/r02h
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_bulk.py tests/test_gpu_edge_indices.py tests/test_abi_symbols.py -x -q -m gpu > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
timeout 900 python tools/bulk_sweep.py --workload bulk_pe --configs "bulk_mode=1;bulk_mode=5;bulk_mode=1,second_parts=1;bulk_mode=1,second_parts=4;bulk_algo=1" > $O/sweep_pe.jsonl 2> $O/sweep_pe.err
cut -c1-330 $O/sweep_pe.jsonl
timeout 900 python tools/bulk_sweep.py --workload bulk_se --configs "bulk_mode=1;bulk_mode=5;bulk_algo=1" > $O/sweep_se.jsonl 2> $O/sweep_se.err
cut -c1-330 $O/sweep_se.jsonl
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
for wl in bulk_pe bulk_se; do
timeout 300 ncu --metrics $M --clock-control none -k regex:bulk -s 9 -c 6 --csv --log-file $O/launches_$wl.csv python tools/bulk_sweep.py --workload $wl --steps 2 --configs "bulk_mode=1" > $O/ncu_$wl.log 2>&1
done
