#!/bin/bash
O=gpurun_out/r02x
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_sharded_file.py tests/test_gpu_bam.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -30 $O/pytest.log
