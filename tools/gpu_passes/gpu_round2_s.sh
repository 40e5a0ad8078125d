#!/bin/bash
O=gpurun_out/r02s
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
nproc > $O/nproc.txt; cat /sys/devices/system/node/node*/cpulist >> $O/nproc.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/h2d_scaling.py > $O/h2d_scaling.json 2> $O/h2d_scaling.err
echo "h2d rc=$?"; head -c 1200 $O/h2d_scaling.json; echo; tail -3 $O/h2d_scaling.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 5 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err ) 2> $O/bench_n8.time
echo "bench rc=$?"; head -c 400 $O/bench_n8.json; echo; tail -3 $O/bench_n8.err; cat $O/bench_n8.time
