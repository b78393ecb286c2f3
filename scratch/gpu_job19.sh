#!/bin/bash
O=gpurun_out/r2x; mkdir -p $O
run() { name=$1; shift; ( "$@" ) > $O/$name.log 2>&1; echo "== $name rc=$?" >> $O/summary.log; }
: > $O/summary.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run tl_multi timeout 170 $TR --master-port 29541 scratch/timeline_multi.py
run t_multi timeout 300 python -m pytest tests/test_gpu_multi.py -q -m gpu
run bench_n2 timeout 300 $TR --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 5 --quick
(CUDA_VISIBLE_DEVICES=0 timeout 200 python bench.py --gpus 1 --steps 10 --warmup 5 --quick > $O/solo0.log 2>&1 &
 CUDA_VISIBLE_DEVICES=1 timeout 200 python bench.py --gpus 1 --steps 10 --warmup 5 --quick > $O/solo1.log 2>&1 ; wait)
cat $O/summary.log; grep "^rank" $O/tl_multi.log | cut -c1-250; tail -3 $O/t_multi.log | cut -c1-300; tail -1 $O/bench_n2.log | cut -c1-400; tail -1 $O/solo0.log | cut -c1-300; tail -1 $O/solo1.log | cut -c1-300
