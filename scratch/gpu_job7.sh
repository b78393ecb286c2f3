#!/bin/bash
O=gpurun_out/r2j; mkdir -p $O
run() { name=$1; shift; ( "$@" ) > $O/$name.log 2>&1; echo "== $name rc=$?" >> $O/summary.log; }
: > $O/summary.log
run tl_train timeout 100 python scratch/timeline_train.py 16384
run t_train timeout 150 python -m pytest tests/test_gpu_tc.py -x -q -m gpu -k "train_equals"
run bench_q timeout 200 python bench.py --steps 20 --warmup 5 --quick
cat $O/summary.log; cat $O/tl_train.log; tail -3 $O/t_train.log; tail -1 $O/bench_q.log | cut -c1-300
