#!/bin/bash
# round-2 evidence for the timed configuration: tests of the changed single-cell path, ncu --set full of the bulk kernels
# (default options, 500 M records, bench.py itself) and the launch list of a 1 B-record single-cell step
set -x
O=gpurun_out/r02t
mkdir -p $O
Q="--no-cpu --no-e2e --file-records 0 --sc-parity-records 0"
timeout 900 python -m pytest tests/test_gpu_sc.py tests/test_gpu_sc_dist.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
C1="python bench.py --workload bulk_pe --steps 2 --warmup 3 $Q"
timeout 300 $C1 > $O/plain_pe.json 2> $O/plain_pe.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bulk -s 12 -c 3 -o $O/prof_bulk_pe -f $C1 > $O/ncu_full_pe.log 2>&1
tail -2 $O/ncu_full_pe.log
C2="python bench.py --workload sc --steps 1 --warmup 1 $Q"
timeout 600 $C2 > $O/plain_sc.json 2> $O/plain_sc.err && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/launches_sc_1b.csv $C2 > $O/ncu_sc.log 2>&1
tail -2 $O/ncu_sc.log
head -c 300 $O/plain_sc.json
