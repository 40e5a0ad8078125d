#!/bin/bash
# what the driver runs at round end: GPU suite, smoke(), the default bench line (and the reference arm)
O=gpurun_out/r02final
mkdir -p $O
( time timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 > $O/pytest.log 2>&1 ) 2> $O/pytest.time; echo "pytest rc=$?" >> $O/pytest.log
tail -14 $O/pytest.log; cat $O/pytest.time
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
( time timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err ) 2> $O/bench.time
head -c 300 $O/bench_default.json; echo; tail -3 $O/bench_default.err; cat $O/bench.time
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err ) 2> $O/ref.time
head -c 300 $O/bench_reference.json; echo; cat $O/ref.time
