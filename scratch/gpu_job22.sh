#!/bin/bash
O=gpurun_out/r3a; mkdir -p $O
run() { name=$1; shift; ( "$@" ) > $O/$name.log 2>&1; echo "== $name rc=$?" >> $O/summary.log; }
: > $O/summary.log
run t_kern timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "mlp_forward or mlp_backward or small_batch"
run sac timeout 200 python bench.py --workload sac --steps 5 --warmup 3
run td3 timeout 200 python bench.py --workload td3 --steps 5 --warmup 3
cat $O/summary.log; tail -5 $O/t_kern.log | cut -c1-300; tail -1 $O/sac.log | cut -c1-1800; echo; tail -1 $O/td3.log | cut -c1-600
