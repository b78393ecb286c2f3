#!/bin/bash
O=gpurun_out/r3f; mkdir -p $O
run() { name=$1; shift; ( "$@" ) > $O/$name.log 2>&1; echo "== $name rc=$?" >> $O/summary.log; }
: > $O/summary.log
run t_all timeout 500 python -m pytest tests -q -m gpu -x
run bench_full timeout 700 python bench.py --steps 20 --warmup 5
cat $O/summary.log; tail -3 $O/t_all.log | cut -c1-300; tail -1 $O/bench_full.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['update_ms'], {k:v for k,v in d['e2e'].items() if k!='note'})"
