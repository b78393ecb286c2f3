#!/bin/bash
# GPU validation job (run on the GPU box through gpurun): results under gpurun_out/r2b
O=gpurun_out/r2b; mkdir -p $O
run() { name=$1; shift; ( "$@" ) > $O/$name.log 2>&1; echo "== $name rc=$?" >> $O/summary.log; }
: > $O/summary.log
run t_wgrad timeout 120 python -m pytest tests/test_gpu_tc.py -x -q -m gpu -k "wgrad_fused"
run t_tc timeout 200 python -m pytest tests/test_gpu_tc.py tests/test_gpu_kernels.py -x -q -m gpu
run t_all timeout 600 python -m pytest tests -q -m gpu
if grep -q "failed\|error\|Timeout" $O/t_all.log; then
  TONIC_B200_FUSED_WGRAD=0 run t_all_nofused timeout 600 python -m pytest tests -q -m gpu
fi
run bench_q timeout 200 python bench.py --steps 20 --warmup 5 --quick
TONIC_B200_FUSED_ADAM=0 run bench_q_noadam timeout 200 python bench.py --steps 20 --warmup 5 --quick
TONIC_B200_FUSED_WGRAD=0 run bench_q_nowgrad timeout 200 python bench.py --steps 20 --warmup 5 --quick
TONIC_B200_CLUSTER=2 run t_cluster2 timeout 200 python -m pytest tests/test_gpu_tc.py -x -q -m gpu -k "mlp_forward_fused or mlp_backward_fused"
TONIC_B200_CLUSTER=2 run bench_q_c2 timeout 200 python bench.py --steps 20 --warmup 5 --quick
run tl_fwd timeout 100 python scratch/timeline_fwd.py 16384
run tl_bwd timeout 100 python scratch/timeline_bwd.py 16384
TONIC_B200_CLUSTER=2 run tl_fwd_c2 timeout 100 python scratch/timeline_fwd.py 16384
TONIC_B200_CLUSTER=2 run tl_bwd_c2 timeout 100 python scratch/timeline_bwd.py 16384
run bench_full timeout 900 python bench.py --steps 20 --warmup 5
cat $O/summary.log
for f in t_wgrad t_tc t_all bench_q bench_q_noadam bench_q_nowgrad t_cluster2 bench_q_c2; do echo "--- $f"; tail -4 $O/$f.log | cut -c1-600; done
