#!/bin/bash
O=gpurun_out/r3d; mkdir -p $O
run() { name=$1; shift; ( "$@" ) > $O/$name.log 2>&1; echo "== $name rc=$?" >> $O/summary.log; }
: > $O/summary.log
run t_all timeout 500 python -m pytest tests -q -m gpu -x
run bench_full timeout 700 python bench.py --steps 20 --warmup 5
run bench_ref timeout 400 python bench.py --impl reference --steps 5 --warmup 1
run smoke timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
(TB_BENCH_CUDA_PROFILER=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 6000 --csv --log-file $O/bench_launches.csv python bench.py --steps 1 --warmup 4 --quick) > $O/ncu_bench.log 2>&1; echo "== ncu rc=$?" >> $O/summary.log
cat $O/summary.log; tail -3 $O/t_all.log | cut -c1-300; tail -1 $O/bench_ref.log | cut -c1-500; tail -2 $O/smoke.log | cut -c1-200; wc -l $O/bench_launches.csv; tail -1 $O/bench_full.log | cut -c1-900
