#!/bin/bash
O=gpurun_out/r02u
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_comm.py tests/test_gpu_sc_dist.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -15 $O/pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 tools/sc_dist_parity.py --records-per-rank 20000000 > $O/sc_dist_parity_lib.json 2> $O/sc_dist_parity_lib.err; echo "parity rc=$?"
tail -2 $O/sc_dist_parity_lib.json; tail -3 $O/sc_dist_parity_lib.err
TEC_DIST_CALLBACKS=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 tools/sc_dist_parity.py --records-per-rank 20000000 > $O/sc_dist_parity_cb.json 2> $O/sc_dist_parity_cb.err; echo "parity cb rc=$?"
tail -2 $O/sc_dist_parity_cb.json
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err ) 2> $O/bench_n2.time
echo "bench rc=$?"; head -c 300 $O/bench_n2.json; echo; tail -3 $O/bench_n2.err; cat $O/bench_n2.time
