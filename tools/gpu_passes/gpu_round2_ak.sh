#!/bin/bash
# final ncu --set full of the bulk paired-end step (default options): one launch of each of the four kernels
O=gpurun_out/r02ak
mkdir -p $O
Q="--no-cpu --no-e2e --file-records 0 --sc-parity-records 0"
C1="python bench.py --workload bulk_pe --steps 2 --warmup 3 $Q"
timeout 300 $C1 > $O/plain_pe.json 2> $O/plain_pe.err || exit 1
head -c 250 $O/plain_pe.json; echo
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bulk -s 16 -c 4 -o $O/prof_bulk_pe -f $C1 > $O/ncu_full_pe.log 2>&1
tail -2 $O/ncu_full_pe.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/launches_pe.csv $C1 > $O/ncu_list.log 2>&1
