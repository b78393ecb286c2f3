#!/bin/bash
O=gpurun_out/r2z2; mkdir -p $O
run() { name=$1; shift; ( "$@" ) > $O/$name.log 2>&1; echo "== $name rc=$?" >> $O/summary.log; }
: > $O/summary.log
nvidia-smi topo -m > $O/topo.log 2>&1
run t_multi timeout 200 python -m pytest tests/test_gpu_multi.py -q -m gpu
run bench_n8 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 5 --quick
run bench_n4 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 10 --warmup 5 --quick
cat $O/summary.log; tail -3 $O/t_multi.log | cut -c1-300; tail -1 $O/bench_n8.log | cut -c1-500; tail -1 $O/bench_n4.log | cut -c1-500
