#!/bin/bash
O=gpurun_out/r02n2
mkdir -p $O
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 > $O/bench_n2.json 2> $O/bench_n2.err ) 2> $O/bench.time
head -c 300 $O/bench_n2.json; echo; tail -3 $O/bench_n2.err; cat $O/bench.time
