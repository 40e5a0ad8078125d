#!/bin/bash
# bulk tests + default bench line with bulk_mode 77 as the default
O=gpurun_out/r02aq
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_bulk.py tests/test_gpu_edge_indices.py tests/test_gpu_extensions.py tests/test_gpu_comm.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
( time timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err ) 2> $O/bench.time
head -c 300 $O/bench_default.json; echo; tail -3 $O/bench_default.err; cat $O/bench.time
