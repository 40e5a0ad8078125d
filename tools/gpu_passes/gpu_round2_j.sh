#!/bin/bash
set -x
O=gpurun_out/r02j
mkdir -p $O
timeout 600 ./tools/radix_test > $O/radix_test.txt 2>&1
cat $O/radix_test.txt
timeout 900 python -m pytest tests/test_gpu_bulk.py tests/test_gpu_edge_indices.py -x -q -m gpu > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
