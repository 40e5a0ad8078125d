#!/bin/bash
O=gpurun_out/r02q
mkdir -p $O
for c in 1 2; do
timeout 900 python bench.py --workload sc --steps 3 --warmup 3 --no-cpu --no-e2e --file-records 0 --opt sc_sort_chunk=$c > $O/bench_sc_c$c.json 2> $O/bench_sc_c$c.err
head -c 260 $O/bench_sc_c$c.json; echo; tail -2 $O/bench_sc_c$c.err
done
