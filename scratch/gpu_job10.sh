#!/bin/bash
O=gpurun_out/r2m; mkdir -p $O
(timeout 100 python scratch/probe_tf32_truncation.py) > $O/probe.log 2>&1; cat $O/probe.log
