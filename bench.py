"""Benchmark of the hot path: PPO env-steps/sec (BASELINE.json metric).

    python bench.py --gpus 1 --steps 5 --warmup 3          # our arm (CUDA)
    python bench.py --impl reference --steps 2 --warmup 1   # reference arm (CPU oracle port)
    torchrun --nproc-per-node N ... bench.py --gpus N ...   # weak scaling, 4096 envs / GPU

One "step" = one PPO iteration of BASELINE.json configs[1]: a rollout of T=128
vector steps of N=4096 synthetic HalfCheetah-shape envs per GPU (obs 17, act 6)
followed by the full update (value evaluation, lambda-returns, advantage
normalisation, E=10 epochs x 32 minibatches of 16384 of clipped-ratio actor +
value-regression critic updates with Adam, 2x256 tanh MLPs).  Nothing is
skipped in the timed region.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline     dominant kernel entry point, achieved TFLOP/s from CUDA-event
               timing of every launch (an extra instrumented pass of the same K
               steps) against MEASURED_PEAKS.json's sustained bf16 figure
  cpu_baseline the oracle port (CPU restatement of the reference, pinned to the
               reference's golden vectors) timed on this box's host cores
  e2e          same metric through the reference-facing protocol with HOST numpy
               arrays crossing the boundary every vector step
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

OBS, ACT, ENVS_PER_GPU, SEGMENT, EPOCHS, MINIBATCHES, HIDDEN = 17, 6, 4096, 128, 10, 32, 256
# dram__bytes_read.sum + dram__bytes_write.sum per launch (16384 rows) from the committed
# `ncu --set full` captures under profiles/ (cold cache, serialised)
# dram__bytes_read.sum + dram__bytes_write.sum of one 16384-row critic-minibatch launch, from the committed
# `ncu --set full` captures (profiles/r1_ncu_full_fused_raw.csv, profiles/r1_ncu_full_*_raw.csv)
NCU_TRAFFIC = {'tb_tc_gemm256_bwd': 70.8e6, 'tb_tc_gemm256_fwd': None, 'tb_tc_wgrad256': 69.4e6,
               'tb_tc_mlp_forward': 3.45e6, 'tb_tc_mlp_backward': 60.2e6}
MAX_EPISODE_STEPS = 1000


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(tflops=p.get('bf16_tflops_sustained', 1441.0), hbm=p.get('hbm_gbs', 6574.1),
                    source='MEASURED_PEAKS.json (sustained bf16)')
    return dict(tflops=1400.0, hbm=6650.0, source='fallback (B200_PROFILING.md)')


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with nvidia-smi during the timed region."""

    QUERY = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(
                    ['nvidia-smi', f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits',
                     '-i', str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([c.strip() for c in out.strip().split(',')])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        import numpy as np
        good = [s for s in self.samples if len(s) == 6 and s[0].isdigit()]
        if not good:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(s[2 + i] == 'Active' for s in good)]
        return dict(sm_mhz=float(np.median([int(s[0]) for s in good])),
                    sm_max_mhz=float(good[0][1]), reasons=reasons, samples=len(good))


# ----------------------------------------------------------------------------- ours
def build_ours(envs_local, total_envs, seed=0):
    import torch
    import tonic_b200
    import tonic_b200.torch
    m, n = tonic_b200.torch.models, tonic_b200.torch.normalizers
    spec = tonic_b200.environments.SynthControl('HalfCheetah', max_episode_steps=MAX_EPISODE_STEPS)
    env = tonic_b200.environments.distribute(lambda: spec, 1, total_envs)
    assert env.workers == envs_local
    env.initialize(seed=seed)
    model = m.ActorCritic(
        actor=m.Actor(encoder=m.ObservationEncoder(), torso=m.MLP((HIDDEN, HIDDEN), torch.nn.Tanh),
                      head=m.DetachedScaleGaussianPolicyHead()),
        critic=m.Critic(encoder=m.ObservationEncoder(), torso=m.MLP((HIDDEN, HIDDEN), torch.nn.Tanh),
                        head=m.ValueHead()),
        observation_normalizer=n.MeanStd())
    batch = total_envs * SEGMENT // MINIBATCHES     # GLOBAL minibatch (weak scaling)
    replay = tonic_b200.replays.Segment(size=SEGMENT, batch_iterations=EPOCHS, batch_size=batch)
    agent = tonic_b200.torch.agents.PPO(model=model, replay=replay)
    agent.initialize(env.observation_space, env.action_space, seed=seed)
    return agent, env, batch


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    assert world == args.gpus, f'launch with torchrun for --gpus {args.gpus} (WORLD_SIZE={world})'
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.barrier()
    from tonic_b200 import _lib, config, graphs, kernels
    from tonic_b200.utils import logger
    iterations = dict(actor=0, critic=0)          # statistics are still read back every update

    def capture(key, value, stats=False):
        if key.endswith('/iterations'):
            iterations[key.split('/')[0]] += int(value)
    logger.store = capture
    logger.store_aggregate = lambda *a, **k: None

    total_envs = ENVS_PER_GPU * world
    config.noise = 'device'      # Philox action noise inside the kernels (no host traffic)
    config.indices = args.indices
    agent, env, batch = build_ours(ENVS_PER_GPU, total_envs)
    env.start()

    def iteration():
        done = agent.rollout(env, SEGMENT)       # fills the segment, then runs the update
        assert done == SEGMENT

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        iteration()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count() + graphs.replayed_launches
    iterations.update(actor=0, critic=0)
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record()
    for _ in range(args.steps):
        iteration()
    end.record()
    barrier()
    sampler.stop_flag = True
    elapsed_ms = torch.tensor([start.elapsed_time(end)], device='cuda')
    if world > 1:
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
    elapsed_ms = float(elapsed_ms.item())
    launches = _lib.launch_count() + graphs.replayed_launches - launches0
    timed_iterations = dict(iterations)
    env_steps = args.steps * SEGMENT * total_envs
    value = env_steps / (elapsed_ms / 1e3)

    if args.quick:
        if rank == 0:
            print(json.dumps(dict(value=round(value, 1), ms_per_step=round(elapsed_ms / args.steps, 3),
                                  gpu_launches=launches, quick=True)))
        return
    # ---- instrumented pass: CUDA events around every launch (same workload) ----------
    # (graphs off: the per-launch events are recorded by the C entry points)
    config.graphs = False
    kernels.flops.clear()
    kernels.profile_begin()
    t0 = time.time()
    for _ in range(args.steps):
        iteration()
    prof = kernels.profile_end()
    config.graphs = True
    prof_wall_ms = (time.time() - t0) * 1e3
    total_kernel_ms = sum(ms for _, ms in prof.values())
    profiled_iterations = dict(iterations)
    pk = peaks()
    # FLOPs of launches that the device-side KL flag turned into no-ops are not counted:
    # the skipped actor minibatches are known from the logged iteration counters
    skipped = profiled_iterations['critic'] - timed_iterations['critic'] - (
        profiled_iterations['actor'] - timed_iterations['actor'])
    rows_local = batch // world
    gemm_names = ('tb_tc_mlp_forward', 'tb_tc_mlp_backward', 'tb_tc_gemm256_fwd', 'tb_tc_gemm256_bwd',
                  'tb_tc_wgrad256', 'tb_mlp_forward', 'tb_mlp_backward', 'tb_mlp_wgrad')
    # algorithmic FLOPs of ONE skipped launch (they are actor-minibatch launches)
    actor_launch_flops = {
        'tb_tc_mlp_forward': 2.0 * rows_local * (OBS * HIDDEN + HIDDEN * HIDDEN + HIDDEN * ACT),
        'tb_tc_mlp_backward': 2.0 * rows_local * HIDDEN * (ACT + HIDDEN)}

    def executed_flops(name):
        total = kernels.flops.get(name, 0.0)
        if name in gemm_names:
            per_launch = actor_launch_flops.get(name, 2.0 * rows_local * HIDDEN * HIDDEN)
            if name in ('tb_mlp_forward', 'tb_mlp_backward', 'tb_mlp_wgrad'):
                per_launch = total / max(prof.get(name, (1, 0))[0], 1)
            total -= skipped * per_launch
        return max(total, 0.0)
    gemms = [k for k in prof if k in gemm_names]
    top = max(gemms, key=lambda k: prof[k][1])
    top_count, top_ms = prof[top]
    top_flops = executed_flops(top)
    achieved = top_flops / (top_ms / 1e3) / 1e12 if top_ms > 0 else 0.0
    passes = {'tf32x3': 3, 'tf32': 1}.get(config.gemm, 0)
    roofline = dict(
        kernel=top, bound='tensor', achieved=round(achieved, 3), peak=pk['tflops'],
        unit='TFLOP/s', frac=round(achieved / pk['tflops'], 5),
        traffic=NCU_TRAFFIC.get(top), peak_source=pk['source'], launches=top_count,
        executed_launches=top_count - (skipped if top in gemm_names else 0),
        avg_launch_us=round(top_ms / max(top_count - skipped, 1) * 1e3, 2),
        share_of_kernel_time=round(top_ms / total_kernel_ms, 4),
        flops_per_launch=round(top_flops / max(top_count - skipped, 1), 1),
        tensor_pipe_tflops=round(achieved * max(passes, 1), 3),
        note=('dominant GEMM entry point by CUDA-event time in an eager (graphs off) pass of the '
              'same K steps; achieved = algorithmic fp32-equivalent FLOPs of the executed launches '
              '(fused forward: 2*rows*(d_in*256 + 256*256 + 256*n_out); fused backward: '
              '2*rows*256*(n_out + 256); weight gradient: 2*rows*256*256) / event time; ' +
              ('FP32 FFMA kernels' if passes == 0 else
               f'tcgen05 kind::tf32 with {passes} MMA pass(es) per product, so the tensor pipe '
               f'itself runs {passes}x the achieved figure (tensor_pipe_tflops)') +
              '; peak = measured dense bf16 tensor throughput; traffic = dram read+write bytes per '
              'launch from the committed ncu --set full capture (profiles/)'),
        kernels={k: dict(launches=c, ms=round(ms, 3),
                         tflops=round(executed_flops(k) / (ms / 1e3) / 1e12, 3) if ms and k in gemm_names else None)
                 for k, (c, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1])})

    # ---- e2e: the reference-facing protocol with HOST arrays ------------------------
    e2e = None
    if True:
        def host_iteration(observations, steps_count):
            for _ in range(SEGMENT):
                actions = agent.step(observations, steps_count)
                observations, infos = env.step(actions)
                agent.update(**infos, steps=steps_count)
                steps_count += ENVS_PER_GPU
            return observations, steps_count

        def measure(iters, warm):
            """The reference's protocol with numpy arrays crossing the boundary every vector step."""
            agent.replay.index = 0
            observations = env.start(host=True)
            steps_count = 0
            for _ in range(warm):
                observations, steps_count = host_iteration(observations, steps_count)
            torch.cuda.synchronize()
            barrier()
            kernels.transfers['h2d'] = kernels.transfers['d2h'] = 0
            t0 = time.time()
            for _ in range(iters):
                observations, steps_count = host_iteration(observations, steps_count)
            torch.cuda.synchronize()
            dt = torch.tensor([time.time() - t0], device='cuda', dtype=torch.float64)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)      # slowest rank
            dt = float(dt.item())
            return (iters * SEGMENT * total_envs / dt, kernels.transfers['h2d'] // iters,
                    kernels.transfers['d2h'] // iters)
        # product configuration: device Philox noise + device permutations, update replayed as a
        # CUDA graph (3 warm iterations: 2 eager executions size the workspaces, 1 captures)
        e2e_iters = max(1, min(args.steps, 3))
        fast, h2d, d2h = measure(e2e_iters, 3)
        # parity configuration: host torch RNG noise + numpy-compatible MT19937 permutations
        parity = None
        if world == 1:
            config.noise = config.indices = 'host'
            parity, _, _ = measure(1, 1)
            config.noise, config.indices = 'device', args.indices
            parity = round(parity, 1)
        e2e = dict(value=round(fast, 1), unit='env-steps/s', h2d_bytes_per_step=h2d,
                   d2h_bytes_per_step=d2h, steps=e2e_iters, parity_mode_value=parity,
                   note='agent.step / environment.step / agent.update called with numpy arrays '
                        'every vector step (pinned staging, copies inside the timed region; bytes '
                        'are per rank and per PPO iteration of 128 vector steps); value = product configuration '
                        '(device noise and permutations, graph-replayed update), parity_mode_value '
                        '= host torch RNG noise + host MT19937 permutations (bit-compatible streams)')

    if rank != 0:
        return
    cpu = cpu_baseline(sample_steps=1) if world == 1 else None
    line = dict(
        metric='env-steps/sec (PPO, 4096 envs per GPU, 2x256 MLP)', value=round(value, 1),
        unit='env-steps/s', n_gpus=world, steps=args.steps, warmup=args.warmup,
        ms_per_step=round(elapsed_ms / args.steps, 3), higher_is_better=True, scaling='weak',
        vs_baseline=None, dtype='f32', data='synthetic',
        config=dict(workload='PPO, synthetic HalfCheetah-shape env (obs=17, act=6), 4096 envs '
                             'per GPU, T=128 segment, E=10 epochs x 32 minibatches, 2x256 tanh MLP',
                    envs_per_gpu=ENVS_PER_GPU, segment=SEGMENT, epochs=EPOCHS, minibatch=batch,
                    hidden=HIDDEN, parallelism=f'dp{world} (envs sharded, grad all-reduce)',
                    l2='working set per iteration (segment 92 MB + activations 69 MB/minibatch) '
                       'exceeds the 126 MB L2; no explicit flush',
                    gemm=config.gemm, cuda_graphs=bool(config.graphs), noise='device Philox', indices=('device Feistel permutation' if args.indices == 'device'
                                                          else 'host MT19937 (numpy-compatible)')),
        clocks=sampler.summary(), gpu_launches=launches,
        minibatch_updates=dict(
            critic=timed_iterations['critic'], actor=timed_iterations['actor'],
            note='actor updates stop early inside an iteration when KL > 0.015 (reference '
                 'default, ppo.py:45-46): a device flag turns the remaining actor kernels of '
                 'that iteration into no-ops; critic updates always run'),
        e2e=e2e, roofline=roofline,
        cpu_baseline=cpu,
        profiled_pass=dict(kernel_ms_per_step=round(total_kernel_ms / args.steps, 3),
                           wall_ms_per_step=round(prof_wall_ms / args.steps, 3)))
    print(json.dumps(line))


# ------------------------------------------------------------------------- reference
def oracle_iteration_factory(envs, seed=0):
    """The oracle port's PPO on the same workload shape (bounded number of envs)."""
    from oracle import port
    env = port.VectorEnv(OBS, ACT, envs, MAX_EPISODE_STEPS)
    env.initialize(seed)
    batch = envs * SEGMENT // MINIBATCHES
    agent = port.OnPolicyOracle('PPO', (HIDDEN, HIDDEN),
                                dict(size=SEGMENT, batch_iterations=EPOCHS, batch_size=batch))
    agent.initialize(env.observation_space, env.action_space, seed=seed)
    state = dict(obs=env.start(), steps=0)

    def iteration():
        for _ in range(SEGMENT):         # trainer.py:44-50
            actions = agent.step(state['obs'], state['steps'])
            state['obs'], infos = env.step(actions)
            agent.update(**infos, steps=state['steps'])
            state['steps'] += envs
    return iteration


def cpu_baseline(sample_steps=1, envs=256):
    import torch
    iteration = oracle_iteration_factory(envs)
    iteration()                                   # warm-up (allocations, first update)
    t0 = time.time()
    for _ in range(sample_steps):
        iteration()
    dt = time.time() - t0
    return dict(value=round(sample_steps * SEGMENT * envs / dt, 1), unit='env-steps/s',
                cores=torch.get_num_threads(), kind='port',
                sample=f'{sample_steps} PPO iteration(s) of {envs} envs x {SEGMENT} steps + full '
                       f'update (E={EPOCHS}, {MINIBATCHES} minibatches/epoch, 2x256 MLP); the '
                       'reference steps envs in a per-env Python loop, so steps/s is flat in '
                       'the number of envs',
                seconds=round(dt, 2))


def run_reference(args):
    """Reference arm: the reference's own CPU algorithm (oracle port -- the Python
    reference cannot travel to the GPU box) on the host cores."""
    if int(os.environ.get('RANK', 0)) != 0:
        return
    import torch
    envs = 256
    iteration = oracle_iteration_factory(envs)
    for _ in range(args.warmup):
        iteration()
    t0 = time.time()
    for _ in range(args.steps):
        iteration()
    dt = time.time() - t0
    value = round(args.steps * SEGMENT * envs / dt, 1)
    sample = (f'each step = 1 PPO iteration of {envs} envs x {SEGMENT} steps + full update '
              f'(E={EPOCHS}, {MINIBATCHES} minibatches/epoch, 2x256 MLP)')
    print(json.dumps(dict(
        impl='reference', metric='env-steps/sec (PPO, 4096 envs per GPU, 2x256 MLP)', value=value,
        unit='env-steps/s', n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
        ms_per_step=round(dt / args.steps * 1e3, 3), higher_is_better=True, scaling='weak',
        vs_baseline=None, dtype='f32', data='synthetic',
        config=dict(workload='PPO, synthetic HalfCheetah-shape env (obs=17, act=6), bounded '
                             'sample of 256 envs, T=128, E=10 x 32 minibatches, 2x256 tanh MLP'),
        cpu_baseline=dict(value=value, unit='env-steps/s', cores=torch.get_num_threads(),
                          kind='port', sample=sample),
        e2e=dict(value=value, unit='env-steps/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))))


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--gpus', type=int, default=1)
    parser.add_argument('--steps', type=int, default=5)
    parser.add_argument('--warmup', type=int, default=3)
    parser.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    parser.add_argument('--indices', default='device', choices=['host', 'device'],
                        help="minibatch permutations: 'device' (Feistel, on the GPU) or 'host' "
                             "(numpy-compatible MT19937 stream, bit-identical to the reference)")
    parser.add_argument('--quick', action='store_true',
                        help='timed region only (for runs under ncu): no e2e / cpu_baseline / profile pass')
    args = parser.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        assert args.warmup >= 3 or os.environ.get('TB_ALLOW_SHORT_WARMUP'), 'W >= 3 warm-up steps'
        run_ours(args)


if __name__ == '__main__':
    main()
