#!/bin/bash
O=gpurun_out/r02v
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sc.py tests/test_gpu_comm.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -8 $O/pytest.log
for p in 1 0; do
timeout 900 python bench.py --workload sc --steps 3 --warmup 3 --no-cpu --no-e2e --file-records 0 --sc-parity-records 0 --opt sc_prev_partition=$p > $O/bench_sc_p$p.json 2> $O/bench_sc_p$p.err
head -c 260 $O/bench_sc_p$p.json; echo; tail -2 $O/bench_sc_p$p.err
done
