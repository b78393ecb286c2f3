#!/bin/bash
O=gpurun_out/r2k; mkdir -p $O
run() { name=$1; shift; ( "$@" ) > $O/$name.log 2>&1; echo "== $name rc=$?" >> $O/summary.log; }
: > $O/summary.log
TONIC_B200_PLAIN_ACTS=1 run tl_train_plain timeout 100 python scratch/timeline_train.py 16384
TONIC_B200_PLAIN_ACTS=1 run t_train_plain timeout 150 python -m pytest tests/test_gpu_tc.py -x -q -m gpu -k "train_equals"
TONIC_B200_PLAIN_ACTS=1 run bench_q_plain timeout 200 python bench.py --steps 20 --warmup 5 --quick
TONIC_B200_PLAIN_ACTS=1 run t_agents_plain timeout 400 python -m pytest tests/test_gpu_agents.py -x -q -m gpu
cat $O/summary.log; cat $O/tl_train_plain.log; tail -3 $O/t_train_plain.log; tail -1 $O/bench_q_plain.log | cut -c1-300; tail -3 $O/t_agents_plain.log
