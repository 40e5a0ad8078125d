#!/bin/bash
# bulk single end on the timed configuration: launch list with DRAM bytes, then ncu --set full of one launch of each kernel
set -x
O=gpurun_out/r02aa
mkdir -p $O
Q="--no-cpu --no-e2e --file-records 0 --sc-parity-records 0"
C1="python bench.py --workload bulk_se --steps 2 --warmup 3 $Q"
timeout 300 $C1 > $O/plain_se.json 2> $O/plain_se.err || exit 1
head -c 300 $O/plain_se.json
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/launches_se.csv $C1 > $O/ncu_se.log 2>&1
tail -2 $O/ncu_se.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bulk -s 24 -c 3 -o $O/prof_bulk_se -f $C1 > $O/ncu_full_se.log 2>&1
tail -2 $O/ncu_full_se.log
ls -la $O
