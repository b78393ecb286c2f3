#!/bin/bash
O=gpurun_out/r2v; mkdir -p $O
run() { name=$1; shift; ( "$@" ) > $O/$name.log 2>&1; echo "== $name rc=$?" >> $O/summary.log; }
: > $O/summary.log
nvidia-smi topo -m > $O/topo.log 2>&1
run t_multi timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu
cp gpurun_out/multi_rank_output.log $O/ 2>/dev/null
run bench_n1 timeout 300 python bench.py --gpus 1 --steps 10 --warmup 5 --quick
run bench_n2 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 5 --quick
run bench_n2_old env TONIC_B200_PEER_FUSED=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 10 --warmup 5 --quick
cat $O/summary.log; tail -3 $O/t_multi.log | cut -c1-300; tail -1 $O/bench_n1.log | cut -c1-400; tail -1 $O/bench_n2.log | cut -c1-600; tail -1 $O/bench_n2_old.log | cut -c1-600; tail -30 $O/multi_rank_output.log
