#!/bin/bash
O=gpurun_out/r2c; mkdir -p $O
run() { name=$1; shift; ( "$@" ) > $O/$name.log 2>&1; echo "== $name rc=$?" >> $O/summary.log; }
: > $O/summary.log
run tl_wgrad timeout 100 python scratch/timeline_wgrad.py 16384
run tl_wgrad_actor timeout 100 python scratch/timeline_wgrad.py 16384 actor
run t_all timeout 900 python -m pytest tests -q -m gpu -s
run bench_q timeout 200 python bench.py --steps 20 --warmup 5 --quick
run bench_sac timeout 400 python bench.py --workload sac --steps 10 --warmup 3 --quick
cat $O/summary.log; cat $O/tl_wgrad.log $O/tl_wgrad_actor.log; tail -3 $O/bench_q.log | cut -c1-400
grep -E "passed|failed" $O/t_all.log | tail -3; grep -E "^FAILED|benched shape|^parity" $O/t_all.log | cut -c1-600
tail -2 $O/bench_sac.log | cut -c1-1500
