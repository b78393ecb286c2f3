#!/bin/bash
O=gpurun_out/r3e; mkdir -p $O
run() { name=$1; shift; ( "$@" ) > $O/$name.log 2>&1; echo "== $name rc=$?" >> $O/summary.log; }
: > $O/summary.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run t_multi timeout 240 python -m pytest tests/test_gpu_multi.py -q -m gpu
run bench_n2 timeout 300 $TR --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 5
cat $O/summary.log; tail -3 $O/t_multi.log | cut -c1-300; tail -1 $O/bench_n2.log | cut -c1-700
