#!/bin/bash
O=gpurun_out/r2w; mkdir -p $O
run() { name=$1; shift; ( "$@" ) > $O/$name.log 2>&1; echo "== $name rc=$?" >> $O/summary.log; }
: > $O/summary.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run tl_multi timeout 170 $TR --master-port 29541 scratch/timeline_multi.py
run scen env TB_MULTI_WATCHDOG=150 timeout 200 $TR --master-port 29542 tests/multi_rank_scenario.py ppo_small
cat $O/summary.log; grep -v "^\[W\|^W0\|^\*\*\*" $O/tl_multi.log | tail -24 | cut -c1-250; tail -40 $O/scen.log | cut -c1-250
