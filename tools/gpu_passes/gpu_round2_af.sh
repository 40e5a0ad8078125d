#!/bin/bash
# per-kernel times of the bulk step with the two-sector kernel (SE and PE), ncu launch lists
O=gpurun_out/r02af
mkdir -p $O
Q="--no-cpu --no-e2e --file-records 0 --sc-parity-records 0 --opt second_parts=4"
for wl in bulk_se bulk_pe; do
C1="python bench.py --workload $wl --steps 2 --warmup 3 $Q"
timeout 300 $C1 > $O/plain_$wl.json 2> $O/plain_$wl.err || exit 1
head -c 200 $O/plain_$wl.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --csv --log-file $O/launches_$wl.csv $C1 > $O/ncu_$wl.log 2>&1
tail -1 $O/ncu_$wl.log | head -c 200; echo
done
