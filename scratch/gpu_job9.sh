#!/bin/bash
O=gpurun_out/r2l; mkdir -p $O
run() { name=$1; shift; ( "$@" ) > $O/$name.log 2>&1; echo "== $name rc=$?" >> $O/summary.log; }
: > $O/summary.log
run tl_wgrad_p1 timeout 100 python scratch/timeline_wgrad.py 16384 p1
run tl_wgrad_p3 timeout 100 python scratch/timeline_wgrad.py 16384
TONIC_B200_GEMM=tf32 run bench_q_tf32 timeout 200 python bench.py --steps 20 --warmup 5 --quick
TONIC_B200_GEMM=tf32 run tl_train_tf32 timeout 100 python scratch/timeline_train.py 16384
cat $O/summary.log; cat $O/tl_wgrad_p1.log $O/tl_wgrad_p3.log $O/tl_train_tf32.log; tail -1 $O/bench_q_tf32.log | cut -c1-300
