#!/bin/bash
set -x
O=gpurun_out/r02b
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_bulk.py tests/test_gpu_edge_indices.py -x -q -m gpu > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 900 python tools/bulk_sweep.py --workload bulk_pe --configs "stab_shift=10,bulk_mode=0;stab_shift=10,bulk_mode=1;stab_shift=10,bulk_mode=2;stab_shift=10,bulk_mode=3;stab_shift=11,bulk_mode=0;stab_shift=11,bulk_mode=1;stab_shift=11,bulk_mode=2;stab_shift=11,bulk_mode=3;stab_shift=10,bulk_mode=3,second_parts=1;stab_shift=10,bulk_mode=3,second_parts=4;bulk_algo=1" > $O/sweep_pe.jsonl 2> $O/sweep_pe.err
cat $O/sweep_pe.jsonl | cut -c1-330
timeout 900 python tools/bulk_sweep.py --workload bulk_se --configs "stab_shift=10,bulk_mode=3;stab_shift=11,bulk_mode=3;stab_shift=10,bulk_mode=0" > $O/sweep_se.jsonl 2> $O/sweep_se.err
cat $O/sweep_se.jsonl | cut -c1-330
for sh in 10 11; do
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:bulk -s 9 -c 6 --csv --log-file $O/ncu_bulk_sh$sh.csv python tools/bulk_sweep.py --workload bulk_pe --records 200000000 --steps 2 --configs "stab_shift=$sh,bulk_mode=3" > $O/ncu_sh$sh.log 2>&1
done
ls -la $O
