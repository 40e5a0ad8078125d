#!/bin/bash
O=gpurun_out/r02n
mkdir -p $O
timeout 300 ./tools/radix_test 400000000 > $O/radix_test.txt 2>&1; echo "rc=$?" >> $O/radix_test.txt
tail -12 $O/radix_test.txt
for c in 2 8; do echo "chunk_tiles=$c"; RDX_CHUNK=$c RDX_BIG_ONLY=1 timeout 120 ./tools/radix_test 400000000; done > $O/radix_chunk.txt 2>&1
cat $O/radix_chunk.txt
timeout 900 python -m pytest tests/test_gpu_sc.py tests/test_gpu_sc_dist.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 900 python bench.py --workload sc --steps 3 --warmup 3 --no-cpu --no-e2e --file-records 0 > $O/bench_sc.json 2> $O/bench_sc.err
head -c 300 $O/bench_sc.json; tail -3 $O/bench_sc.err
