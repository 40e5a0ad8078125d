#!/bin/bash
# re-entry pass: full gpu tests, bulk mode sweep, whole bench line, reference arm
set -x
O=gpurun_out/r02d
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
nproc > $O/nproc.txt
( time timeout 1200 python -m pytest tests -x -q -m gpu > $O/pytest.log 2>&1 ; echo "pytest rc=$?" >> $O/pytest.log ) 2> $O/pytest.time
tail -5 $O/pytest.log
timeout 900 python tools/bulk_sweep.py --workload bulk_pe --configs "bulk_algo=2;bulk_mode=0;bulk_mode=1;bulk_mode=3;bulk_mode=5;stab_shift=11,bulk_mode=1;stab_shift=11,bulk_mode=5;bulk_algo=1" > $O/sweep_pe.jsonl 2> $O/sweep_pe.err
cut -c1-330 $O/sweep_pe.jsonl
timeout 900 python tools/bulk_sweep.py --workload bulk_se --configs "bulk_algo=2;bulk_mode=1;bulk_mode=5;bulk_algo=1" > $O/sweep_se.jsonl 2> $O/sweep_se.err
cut -c1-330 $O/sweep_se.jsonl
( time timeout 1200 python bench.py > $O/bench_all.json 2> $O/bench_all.err ) 2> $O/bench_all.time
tail -3 $O/bench_all.err; cat $O/bench_all.time
head -c 3000 $O/bench_all.json
( time timeout 900 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err ) 2> $O/bench_ref.time
cat $O/bench_ref.time; head -c 1000 $O/bench_ref.json
ls -la $O
