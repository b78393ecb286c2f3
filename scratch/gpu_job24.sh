#!/bin/bash
O=gpurun_out/r3c; mkdir -p $O
for v in 0 1; do (CUDA_LAUNCH_BLOCKING=1 timeout 80 python scratch/dbg_step.py $v PPO) > $O/dbg_$v.log 2>&1; echo "== dbg_$v rc=$?" >> $O/summary.log; done
(CUDA_LAUNCH_BLOCKING=1 timeout 80 python scratch/dbg_step.py 1 A2C) > $O/dbg_a2c.log 2>&1; echo "== dbg_a2c rc=$?" >> $O/summary.log
cat $O/summary.log; tail -25 $O/dbg_0.log | cut -c1-200; echo ----; tail -25 $O/dbg_1.log | cut -c1-200; echo ----; tail -12 $O/dbg_a2c.log | cut -c1-200
