#!/bin/bash
# second bulk pass with prefetch (second_mode 0/1/2 x second_parts) on PE and SE; one-GPU bundle search by chunk counts (sc)
O=gpurun_out/r02ab
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_bulk.py tests/test_gpu_sc.py tests/test_gpu_edge_indices.py tests/test_gpu_extensions.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
CFG="second_mode=0,second_parts=2;second_mode=0,second_parts=3;second_mode=1,second_parts=3;second_mode=1,second_parts=5;second_mode=1,second_parts=8;second_mode=2,second_parts=5;second_mode=2,second_parts=8"
timeout 600 python tools/bulk_sweep.py --workload bulk_pe --steps 8 --configs "$CFG" > $O/sweep_pe.jsonl 2> $O/sweep_pe.err
cut -c 1-200 $O/sweep_pe.jsonl; tail -2 $O/sweep_pe.err
timeout 600 python tools/bulk_sweep.py --workload bulk_se --steps 8 --configs "$CFG" > $O/sweep_se.jsonl 2> $O/sweep_se.err
cut -c 1-200 $O/sweep_se.jsonl; tail -2 $O/sweep_se.err
timeout 900 python bench.py --workload sc --steps 3 --warmup 3 --no-cpu --no-e2e --file-records 0 --sc-parity-records 0 > $O/bench_sc.json 2> $O/bench_sc.err
head -c 300 $O/bench_sc.json; echo; tail -2 $O/bench_sc.err
