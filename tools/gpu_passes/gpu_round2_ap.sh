#!/bin/bash
# ballot queue drained once per tile (bulk_mode bit 6)
O=gpurun_out/r02ap
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_bulk.py -x -q -m gpu -k "variants" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
CFG="bulk_mode=13;bulk_mode=77;bulk_mode=13;bulk_mode=77"
timeout 600 python tools/bulk_sweep.py --workload bulk_pe --steps 10 --configs "$CFG" > $O/sweep_pe.jsonl 2> $O/sweep_pe.err
cut -c 1-160 $O/sweep_pe.jsonl; tail -2 $O/sweep_pe.err
timeout 600 python tools/bulk_sweep.py --workload bulk_se --steps 10 --configs "bulk_mode=13;bulk_mode=77" > $O/sweep_se.jsonl 2> $O/sweep_se.err
cut -c 1-160 $O/sweep_se.jsonl; tail -2 $O/sweep_se.err
