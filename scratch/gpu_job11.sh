#!/bin/bash
O=gpurun_out/r2n; mkdir -p $O
run() { name=$1; shift; ( "$@" ) > $O/$name.log 2>&1; echo "== $name rc=$?" >> $O/summary.log; }
: > $O/summary.log
run t_tc timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_kernels.py -x -q -m gpu
run tl_train timeout 100 python scratch/timeline_train.py 16384
run tl_wgrad timeout 100 python scratch/timeline_wgrad.py 16384 plain
run bench_q timeout 200 python bench.py --steps 20 --warmup 5 --quick
run t_all timeout 900 python -m pytest tests -q -m gpu
cat $O/summary.log; tail -3 $O/t_tc.log; cat $O/tl_train.log $O/tl_wgrad.log; tail -1 $O/bench_q.log | cut -c1-300; tail -4 $O/t_all.log | cut -c1-300
