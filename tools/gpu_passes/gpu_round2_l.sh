#!/bin/bash
set -x
O=gpurun_out/r02l
mkdir -p $O
CMD="python bench.py --workload sc --steps 1 --warmup 1 --no-cpu --no-e2e --file-records 0 --sc-parity-records 0"
timeout 600 $CMD > $O/plain.json 2> $O/plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/launches_sc_1b.csv $CMD > $O/ncu.log 2>&1
tail -3 $O/ncu.log
wc -l $O/launches_sc_1b.csv
