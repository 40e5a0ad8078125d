#!/bin/bash
O=gpurun_out/r02y
mkdir -p $O
( time timeout 1500 python -m pytest tests/test_gpu_sharded_file.py tests/test_gpu_comm.py tests/test_gpu_bam.py -x -q -m gpu --durations=8 > $O/pytest.log 2>&1 ) 2> $O/pytest.time; echo "pytest rc=$?" >> $O/pytest.log
tail -25 $O/pytest.log; cat $O/pytest.time
