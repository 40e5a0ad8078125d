#!/bin/bash
# scan queue of the fast kernel (bulk_mode bit 4) and the shifted-in register set of the second pass (second_mode 1)
O=gpurun_out/r02ac
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_bulk.py -x -q -m gpu -k "variants or dense or placement or seeded" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
CFG="bulk_mode=13,second_mode=0;bulk_mode=13,second_mode=1;bulk_mode=29,second_mode=1;bulk_mode=21,second_mode=1;bulk_mode=29,second_mode=1,second_parts=2;bulk_mode=29,second_mode=1,second_parts=4"
timeout 600 python tools/bulk_sweep.py --workload bulk_pe --steps 8 --configs "$CFG" > $O/sweep_pe.jsonl 2> $O/sweep_pe.err
cut -c 1-200 $O/sweep_pe.jsonl; tail -2 $O/sweep_pe.err
timeout 600 python tools/bulk_sweep.py --workload bulk_se --steps 8 --configs "$CFG" > $O/sweep_se.jsonl 2> $O/sweep_se.err
cut -c 1-200 $O/sweep_se.jsonl; tail -2 $O/sweep_se.err
