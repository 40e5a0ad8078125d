#!/bin/bash
O=gpurun_out/r02r
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err ) 2> $O/bench_n2.time
echo "rc=$?"; tail -c 1500 $O/bench_n2.json | head -c 1500; echo; tail -5 $O/bench_n2.err; cat $O/bench_n2.time
