#!/bin/bash
# single end, coordinate-sorted arrival order (the order of a sorted BAM) next to the random order of the default leg
O=gpurun_out/r02ai
mkdir -p $O
Q="--no-cpu --no-e2e --file-records 0 --sc-parity-records 0"
timeout 600 python bench.py --workload bulk_se --sorted --steps 5 --warmup 3 $Q > $O/bench_se_sorted.json 2> $O/bench_se_sorted.err
head -c 300 $O/bench_se_sorted.json; echo; tail -2 $O/bench_se_sorted.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --csv --log-file $O/launches_se_sorted.csv python bench.py --workload bulk_se --sorted --steps 1 --warmup 3 $Q > $O/ncu.log 2>&1
tail -1 $O/ncu.log | head -c 200
