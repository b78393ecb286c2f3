"""Run-time switches of the device backend.

noise   'host'   : action / target noise is drawn on the host from the same
                   generators the reference uses (torch's global CPU generator,
                   numpy RandomState(seed)) and copied to the device -- the
                   parity mode, bit-compatible streams ("RNG parity").
        'device' : Philox4x32-10 counter-based noise generated inside the kernels
                   (no host->device traffic); same distributions, different stream.
indices 'host'   : minibatch permutations come from the native numpy-compatible
                   MT19937 stream (bit-identical indices; with several ranks the
                   permutation is global and every rank keeps the rows it owns).
        'device' : permutations are generated on the GPU (Feistel bijection); with
                   several ranks each rank permutes ITS OWN rows and contributes
                   batch_size / world rows to every minibatch (static shapes, no host
                   work that grows with the number of GPUs).
gemm    'tf32x3' : the 256-wide hidden-layer GEMMs (forward layer 2, backward dz1, weight
                   gradient dW2) run on the tensor cores (tcgen05.mma kind::tf32, TMA,
                   TMEM) with the 3xTF32 operand split -> fp32-grade results (default
                   for hidden width 256; other widths use the FFMA kernels).
        'tf32'   : single-pass TF32 on the tensor cores (fast mode, ~1e-3 relative).
        'ffma'   : everything on the FP32 FFMA kernels (csrc/mlp.cu).
                   Environment override: TONIC_B200_GEMM.
graphs  True     : sections with static shapes (PPO / A2C rollout of a whole segment and
                   the update, when noise == indices == 'device' in a single process)
                   are captured into CUDA graphs and replayed (tonic_b200/graphs.py).
peer_reduce True : with several ranks the gradient all-reduce is fused into the Adam
                   kernel over NVLink peer memory (symmetric memory); False = NCCL
                   all-reduce of the flat gradient + statistics, then Adam.
fused_rollout 'auto' : with device noise the whole on-policy segment (T vector steps of actor
                   forward + sampling + environment step + segment store + normaliser record)
                   can run as ONE persistent kernel (csrc/mlp.cu rollout_kernel, FP32 FFMA
                   arithmetic, 64 environments resident per CTA = 64 CTAs at 4096 envs).  'auto'
                   uses it unless the per-step chain can be replayed as a CUDA graph on the
                   tensor-core path (then the chain is faster: measured 4.9 vs 6.5 ms per
                   128-step segment of 4096 envs on B200); True = always, False = never.
fused_step True  : with device noise on the synthetic task one vector step of the on-policy
                   rollout chain is actor forward + ONE kernel (sample, log-prob, normaliser record,
                   environment step: csrc/env_step.cu act_env_step_kernel) instead of five launches.
wgrad_splits     : number of row splits of the weight-gradient kernel.
"""

import os

noise = 'host'
gemm = os.environ.get('TONIC_B200_GEMM', 'tf32x3')
indices = 'host'
graphs = os.environ.get('TONIC_B200_GRAPHS', '1') != '0'
peer_reduce = os.environ.get('TONIC_B200_PEER_REDUCE', '1') != '0'   # fused NVLink reduce + Adam
graphs_multi_gpu = os.environ.get('TONIC_B200_GRAPHS_MULTI', '1') != '0'   # capture NCCL too
fused_rollout = {'0': False, '1': True}.get(os.environ.get('TONIC_B200_FUSED_ROLLOUT', 'auto'), 'auto')
fused_step = os.environ.get('TONIC_B200_FUSED_STEP', '1') != '0'
wgrad_splits = 37      # FFMA: 4 heavy tiles x 37 splits = 148 CTAs = one per B200 SM
wgrad_splits_tc = 74   # tensor cores: 2 row tiles x 74 splits = 148 CTAs
