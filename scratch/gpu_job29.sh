#!/bin/bash
O=gpurun_out/r3h; mkdir -p $O
(timeout 600 python bench.py --steps 20 --warmup 5) > $O/bench_full.log 2> $O/bench_full.err; echo "rc=$?"
tail -1 $O/bench_full.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['update_ms'], d['minibatch_updates'])
print({k:v for k,v in d['e2e'].items() if k!='note'})
r=d['roofline']; print(r['kernel'], r['achieved'], r['frac'], r['executed_launches'], r['avg_launch_us'])"
tail -3 $O/bench_full.err
