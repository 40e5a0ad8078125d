#!/bin/bash
# microbenchmarks + compute-sanitizer (memcheck, racecheck) over the smoke path and a subset of the GPU tests
set -x
O=gpurun_out/r02g
mkdir -p $O
./tools/microbench2 > $O/microbench2.txt 2>&1
cat $O/microbench2.txt
CS=/usr/local/cuda/bin/compute-sanitizer
( time timeout 900 $CS --tool memcheck --print-limit 20 python __graft_entry__.py --smoke > $O/memcheck_smoke.log 2>&1 ) 2> $O/memcheck_smoke.time
tail -4 $O/memcheck_smoke.log
( time timeout 900 $CS --tool racecheck --print-limit 20 python __graft_entry__.py --smoke > $O/racecheck_smoke.log 2>&1 ) 2> $O/racecheck_smoke.time
tail -4 $O/racecheck_smoke.log
( time timeout 900 $CS --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_bam.py -x -q -m gpu -k "golden or errors" > $O/memcheck_bam.log 2>&1 ) 2> $O/memcheck_bam.time
tail -4 $O/memcheck_bam.log
( time timeout 900 $CS --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_bam.py tests/test_gpu_bulk.py -x -q -m gpu -k "golden" > $O/racecheck_tests.log 2>&1 ) 2> $O/racecheck_tests.time
tail -4 $O/racecheck_tests.log
( time timeout 600 $CS --tool synccheck --print-limit 20 python __graft_entry__.py --smoke > $O/synccheck_smoke.log 2>&1 ) 2> $O/synccheck_smoke.time
tail -4 $O/synccheck_smoke.log
