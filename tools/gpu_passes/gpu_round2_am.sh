#!/bin/bash
# winner list by one selection pass; the reference's CLI with the class swapped
O=gpurun_out/r02am
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sc.py tests/test_gpu_comm.py tests/test_gpu_sc_dist.py tests/test_gpu_reference_cli.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -4 $O/pytest.log
timeout 900 python bench.py --workload sc --steps 3 --warmup 3 --no-cpu --no-e2e --file-records 0 --sc-parity-records 0 > $O/bench_sc.json 2> $O/bench_sc.err
head -c 300 $O/bench_sc.json; echo; tail -2 $O/bench_sc.err
