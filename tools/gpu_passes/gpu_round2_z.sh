#!/bin/bash
# full GPU suite + the default bench line (what the driver runs at round end)
O=gpurun_out/r02z
mkdir -p $O
( time timeout 1500 python -m pytest tests -x -q -m gpu --durations=10 > $O/pytest.log 2>&1 ) 2> $O/pytest.time; echo "pytest rc=$?" >> $O/pytest.log
tail -18 $O/pytest.log; cat $O/pytest.time
( time timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err ) 2> $O/bench.time
head -c 400 $O/bench_default.json; echo; tail -3 $O/bench_default.err; cat $O/bench.time
