#!/bin/bash
O=gpurun_out/r2q; mkdir -p $O
run() { name=$1; shift; ( "$@" ) > $O/$name.log 2>&1; echo "== $name rc=$?" >> $O/summary.log; }
: > $O/summary.log

run tl_wgrad timeout 100 python scratch/timeline_wgrad.py 16384 plain

cat $O/summary.log; cat $O/tl_wgrad.log
run tl_wgrad_split timeout 100 python scratch/timeline_wgrad.py 16384; cat $O/tl_wgrad_split.log
