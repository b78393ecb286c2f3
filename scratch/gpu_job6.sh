#!/bin/bash
O=gpurun_out/r2h; mkdir -p $O
run() { name=$1; shift; ( "$@" ) > $O/$name.log 2>&1; echo "== $name rc=$?" >> $O/summary.log; }
: > $O/summary.log
run t_train timeout 150 python -m pytest tests/test_gpu_tc.py -x -q -m gpu -k "train_equals"
run t_all timeout 900 python -m pytest tests -q -m gpu -s -x
run bench_q timeout 200 python bench.py --steps 20 --warmup 5 --quick
TONIC_B200_FUSED_TRAIN=0 run bench_q_nofuse timeout 200 python bench.py --steps 20 --warmup 5 --quick
cat $O/summary.log; tail -15 $O/t_train.log | cut -c1-300; tail -2 $O/bench_q.log | cut -c1-300; tail -1 $O/bench_q_nofuse.log | cut -c1-300
grep -E "passed|failed" $O/t_all.log | tail -3; grep -E "^FAILED|benched shape|^parity" $O/t_all.log | cut -c1-1500
