#!/bin/bash
O=gpurun_out/r2s; mkdir -p $O
for d in 0 1 2 3; do (TB_WGRAD_DBG=$d timeout 100 python scratch/timeline_wgrad.py 16384 plain) > $O/dbg$d.log 2>&1; echo "--- dbg=$d"; head -2 $O/dbg$d.log | cut -c1-250; done
