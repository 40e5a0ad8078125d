#!/bin/bash
set -x
O=gpurun_out/r02k
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_sc.py tests/test_gpu_sc_dist.py tests/test_gpu_bam.py -x -q -m gpu > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rdx_scatter -s 20 -c 1 -o $O/prof_rdx -f ./tools/radix_test 100000000 > $O/ncu_rdx.log 2>&1
tail -3 $O/ncu_rdx.log
timeout 900 python bench.py --workload sc --steps 3 --warmup 3 --no-cpu --no-e2e --file-records 0 > $O/bench_sc.json 2> $O/bench_sc.err
head -c 600 $O/bench_sc.json; tail -3 $O/bench_sc.err
