#!/bin/bash
O=gpurun_out/r3b; mkdir -p $O
run() { name=$1; shift; ( "$@" ) > $O/$name.log 2>&1; echo "== $name rc=$?" >> $O/summary.log; }
: > $O/summary.log
run t_step timeout 300 python -m pytest tests/test_gpu_agents.py -q -m gpu -x -k "fused or graph"
run t_kern timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "mlp_forward or mlp_backward or small_batch"
run bench_q timeout 200 python bench.py --steps 20 --warmup 5 --quick
run bench_q0 env TONIC_B200_FUSED_STEP=0 timeout 200 python bench.py --steps 20 --warmup 5 --quick
cat $O/summary.log; tail -4 $O/t_step.log | cut -c1-300; tail -3 $O/t_kern.log | cut -c1-300; tail -1 $O/bench_q.log | cut -c1-400; tail -1 $O/bench_q0.log | cut -c1-400
