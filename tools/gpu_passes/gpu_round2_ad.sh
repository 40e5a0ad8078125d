#!/bin/bash
# triples sorted by csrc/radix.cuh (keys only) instead of the library sort; single-cell step again
O=gpurun_out/r02ad
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sc.py tests/test_gpu_comm.py tests/test_gpu_sc_dist.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -4 $O/pytest.log
timeout 900 python bench.py --workload sc --steps 3 --warmup 3 --no-cpu --no-e2e --file-records 0 --sc-parity-records 0 > $O/bench_sc.json 2> $O/bench_sc.err
head -c 300 $O/bench_sc.json; echo; tail -2 $O/bench_sc.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/launches_sc_1b.csv python bench.py --workload sc --steps 1 --warmup 1 --no-cpu --no-e2e --file-records 0 --sc-parity-records 0 > $O/ncu_sc.log 2>&1
tail -2 $O/ncu_sc.log
