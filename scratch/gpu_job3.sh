#!/bin/bash
O=gpurun_out/r2d; mkdir -p $O
run() { name=$1; shift; ( "$@" ) > $O/$name.log 2>&1; echo "== $name rc=$?" >> $O/summary.log; }
: > $O/summary.log
run t_wgrad timeout 120 python -m pytest tests/test_gpu_tc.py -x -q -m gpu -k "wgrad_fused"
run tl_wgrad timeout 100 python scratch/timeline_wgrad.py 16384
run tl_wgrad_actor timeout 100 python scratch/timeline_wgrad.py 16384 actor
run t_all timeout 900 python -m pytest tests -q -m gpu -s
run bench_q timeout 200 python bench.py --steps 20 --warmup 5 --quick
cat $O/summary.log; tail -3 $O/t_wgrad.log; cat $O/tl_wgrad.log $O/tl_wgrad_actor.log; tail -3 $O/bench_q.log | cut -c1-400
grep -E "passed|failed" $O/t_all.log | tail -3; grep -E "^FAILED|benched shape|^parity" $O/t_all.log | cut -c1-900
