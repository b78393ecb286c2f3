#!/bin/bash
O=gpurun_out/r3g; mkdir -p $O
(timeout 200 python scratch/profile_e2e.py) > $O/prof.log 2>&1; echo "rc=$?"
grep -v "^$" $O/prof.log | cut -c1-170 | head -90
