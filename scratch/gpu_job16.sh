#!/bin/bash
O=gpurun_out/r2t; mkdir -p $O
run() { name=$1; shift; ( "$@" ) > $O/$name.log 2>&1; echo "== $name rc=$?" >> $O/summary.log; }
: > $O/summary.log
run t_all timeout 900 python -m pytest tests -q -m gpu
run bench_full timeout 900 python bench.py --steps 20 --warmup 5
run bench_ref timeout 600 python bench.py --impl reference --steps 5 --warmup 1
cat $O/summary.log; tail -3 $O/t_all.log | cut -c1-200; tail -1 $O/bench_ref.log | cut -c1-600
