#!/bin/bash
# two-sector kernel in front of the second bulk pass (second_mode 2)
O=gpurun_out/r02ae
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_bulk.py tests/test_gpu_edge_indices.py -x -q -m gpu -k "not bam_file" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
CFG="second_mode=1;second_mode=2;second_mode=2,second_parts=2;second_mode=2,second_parts=4"
timeout 600 python tools/bulk_sweep.py --workload bulk_pe --steps 8 --configs "$CFG" > $O/sweep_pe.jsonl 2> $O/sweep_pe.err
cut -c 1-200 $O/sweep_pe.jsonl; tail -2 $O/sweep_pe.err
timeout 600 python tools/bulk_sweep.py --workload bulk_se --steps 8 --configs "$CFG" > $O/sweep_se.jsonl 2> $O/sweep_se.err
cut -c 1-200 $O/sweep_se.jsonl; tail -2 $O/sweep_se.err
