#!/bin/bash
O=gpurun_out/r2i; mkdir -p $O
run() { name=$1; shift; ( "$@" ) > $O/$name.log 2>&1; echo "== $name rc=$?" >> $O/summary.log; }
: > $O/summary.log
run tl_train timeout 100 python scratch/timeline_train.py 16384
run t_bench_shape timeout 300 python -m pytest tests/test_gpu_agents.py -q -m gpu -k benched -s
cat $O/summary.log; cat $O/tl_train.log; tail -3 $O/t_bench_shape.log | cut -c1-300
