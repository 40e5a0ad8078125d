#!/bin/bash
# first GPU pass of round 2: parity of the two-pass bulk kernels + A/B timing + one ncu capture
set -x
O=gpurun_out/r02a
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
timeout 900 python -m pytest tests/test_gpu_bulk.py tests/test_gpu_edge_indices.py tests/test_abi_symbols.py -x -q -m gpu > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
for algo in 2 1; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --file-records 0 --opt bulk_algo=$algo > $O/bench_pe_algo$algo.json 2> $O/bench_pe_algo$algo.err
  timeout 600 python bench.py --workload bulk_se --steps 10 --warmup 3 --no-cpu --no-e2e --file-records 0 --opt bulk_algo=$algo > $O/bench_se_algo$algo.json 2> $O/bench_se_algo$algo.err
done
cat $O/bench_pe_algo2.json | head -c 1500
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bulk2_fast -s 3 -c 1 -o $O/prof_fast2 -f python bench.py --records 200000000 --steps 2 --warmup 3 --no-cpu --no-e2e --file-records 0 > $O/ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --file-records 0 > $O/ncu_launches.log 2>&1
ls -la $O
