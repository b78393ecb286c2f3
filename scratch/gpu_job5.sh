#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
run() { name=$1; shift; ( "$@" ) > $O/$name.log 2>&1; echo "== $name rc=$?" >> $O/summary.log; }
: > $O/summary.log
run t_wgrad timeout 120 python -m pytest tests/test_gpu_tc.py -x -q -m gpu -k "wgrad_fused"
run tl_wgrad timeout 100 python scratch/timeline_wgrad.py 16384 plain
run tl_wgrad_split timeout 100 python scratch/timeline_wgrad.py 16384
run tl_fwd timeout 100 python scratch/timeline_fwd.py 16384
run t_all timeout 900 python -m pytest tests -q -m gpu -s
run bench_q timeout 200 python bench.py --steps 20 --warmup 5 --quick
TONIC_B200_PLAIN_ACTS=0 run bench_q_split timeout 200 python bench.py --steps 20 --warmup 5 --quick
cat $O/summary.log; tail -3 $O/t_wgrad.log; cat $O/tl_wgrad.log $O/tl_wgrad_split.log; tail -3 $O/bench_q.log | cut -c1-300; tail -1 $O/bench_q_split.log | cut -c1-300
grep -E "passed|failed" $O/t_all.log | tail -3; grep -E "^FAILED|benched shape|^parity" $O/t_all.log | cut -c1-1200
