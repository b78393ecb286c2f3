#!/bin/bash
O=gpurun_out/r2r; mkdir -p $O
run() { name=$1; shift; ( "$@" ) > $O/$name.log 2>&1; echo "== $name rc=$?" >> $O/summary.log; }
: > $O/summary.log
run t_wgrad timeout 120 python -m pytest tests/test_gpu_tc.py -x -q -m gpu -k "wgrad_fused"
run tl_wgrad timeout 100 python scratch/timeline_wgrad.py 16384 plain
run bench_q timeout 200 python bench.py --steps 20 --warmup 5 --quick
cat $O/summary.log; tail -2 $O/t_wgrad.log; cat $O/tl_wgrad.log | cut -c1-700; tail -1 $O/bench_q.log | cut -c1-300
