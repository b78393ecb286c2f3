#!/bin/bash
O=gpurun_out/r2o; mkdir -p $O
# launch list (durations + dram bytes) of one PPO iteration with ONE 16384-row minibatch per network
(timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --cache-control none --clock-control none --profile-from-start off --csv --log-file $O/chain_launches.csv python scratch/profile_chain.py) > $O/chain_launches.log 2>&1
# full capture of the two tensor-core kernels
(timeout 400 ncu --set full --cache-control none --clock-control none --import-source on --profile-from-start off -k regex:"tc_mlp_train_kernel|tc_wgrad_all_kernel" -c 4 -o $O/r2_chain_full python scratch/profile_chain.py) > $O/chain_full.log 2>&1
ls -la $O; tail -3 $O/chain_launches.log $O/chain_full.log
