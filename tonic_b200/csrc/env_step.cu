// K1: device-resident synthetic vector environment.
// Replaces the per-env Python loop of Sequential.step
// (reference tonic/environments/distributed.py:28-58) and the ActionRescaler
// clip (tonic/environments/wrappers.py:18-22) for SynthControl(O, A); the
// dynamics are the ones defined in oracle/synth_env.py and are bit-identical
// (separately rounded float32 ops, integer-quantised reward).
//
// Layout: state / observations are [N, O] row-major (the reference's layout);
// a CTA stages a contiguous tile of `tile_envs` rows through shared memory with
// coalesced loads/stores; inside the tile one warp owns one env at a time
// (lanes over observation coordinates, warp-shuffle reduction for the cost).
#include "env_dynamics.cuh"

namespace tb {

__global__ void __launch_bounds__(256)
env_start_kernel(TbEnv env, float* __restrict__ obs) {
    const int64_t total = (int64_t)env.n_envs * env.obs_dim;
    const int OW = env.obs_dim + (env.time_feature ? 1 : 0);       // observation row width
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(i / env.obs_dim), j = (int)(i % env.obs_dim);
        const uint32_t seed = (uint32_t)(env.seed + env.first_worker + n);
        const float v = reset_coordinate(reset_key(seed, 0u), j);
        env.d_state[i] = v;
        obs[(size_t)n * OW + j] = v;
        if (j == 0 && env.time_feature) obs[(size_t)n * OW + env.obs_dim] = env.time_low;   // wrappers.py:43
        if (j == 0) {
            env.d_length[n] = 0;
            env.d_episode[n] = 1u;
            env.d_score[n] = 0.0;
        }
    }
}

__global__ void __launch_bounds__(256)
env_step_kernel(TbEnv env, const float* __restrict__ actions, float* __restrict__ obs,
                float* __restrict__ next_obs, float* __restrict__ rewards,
                float* __restrict__ resets, float* __restrict__ terminations,
                int tile_envs) {
    extern __shared__ float smem[];
    const int O = env.obs_dim, A = env.act_dim;
    float* sx = smem;                          // [tile, O] state -> transition obs
    float* so = sx + (size_t)tile_envs * O;    // [tile, O] acting obs (post reset)
    float* sa = so + (size_t)tile_envs * O;    // [tile, A] clipped actions
    float* stf = sa + (size_t)tile_envs * A;   // [tile, 2] time feature of next_obs / acting obs

    const int e0 = blockIdx.x * tile_envs;
    const int count = min(tile_envs, env.n_envs - e0);
    const int tid = threadIdx.x;

    const float* gx = env.d_state + (size_t)e0 * O;
    for (int i = tid; i < count * O; i += blockDim.x) sx[i] = gx[i];
    const float* ga = actions + (size_t)e0 * A;
    for (int i = tid; i < count * A; i += blockDim.x)
        sa[i] = fminf(fmaxf(ga[i], -1.0f), 1.0f);      // wrappers.py:22 np.clip
    __syncthreads();

    const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    for (int e = warp; e < count; e += nwarps) {
        const int n = e0 + e;
        float* x = sx + (size_t)e * O;
        const float* a = sa + (size_t)e * A;
        float reward;
        int term;
        env_transition_warp(x, a, O, A, lane, &reward, &term);
        int reset = 0;
        uint32_t episode = 0;
        if (lane == 0) {
            int length = env.d_length[n] + 1;
            // distributed.py:40 -- a time-out resets without terminating
            reset = term || (length == env.max_episode_steps);
            if (env.time_feature) {     // wrappers.py:49-52 in float64, cast like np.array(.., float32)
                const double prop = (double)length / (double)env.max_episode_steps;
                const float v = (float)((double)env.time_low +
                                        ((double)env.time_high - (double)env.time_low) * prop);
                stf[2 * e] = v;
                stf[2 * e + 1] = reset ? env.time_low : v;
            }
            double score = env.d_score[n] + (double)reward;    // trainer.py:52
            episode = env.d_episode[n];
            if (reset) {                                        // trainer.py:64-71
                const unsigned long long slot = atomicAdd(env.d_ep_count, 1ull);
                if (env.log_cap > 0) {
                    env.d_ep_scores[slot % env.log_cap] = score;
                    env.d_ep_lengths[slot % env.log_cap] = length;
                }
                env.d_episode[n] = episode + 1u;
                length = 0;
                score = 0.0;
            }
            env.d_length[n] = length;
            env.d_score[n] = score;
            rewards[n] = reward;
            resets[n] = reset ? 1.0f : 0.0f;
            terminations[n] = term ? 1.0f : 0.0f;
        }
        reset = __shfl_sync(0xffffffffu, reset, 0);
        episode = __shfl_sync(0xffffffffu, episode, 0);
        float* o = so + (size_t)e * O;
        if (reset) {                                             // distributed.py:46-48
            const uint32_t key = reset_key((uint32_t)(env.seed + env.first_worker + n), episode);
            for (int j = lane; j < O; j += 32) o[j] = reset_coordinate(key, j);
        } else {
            for (int j = lane; j < O; j += 32) o[j] = x[j];
        }
    }
    __syncthreads();

    float* gs = env.d_state + (size_t)e0 * O;
    for (int i = tid; i < count * O; i += blockDim.x) gs[i] = so[i];
    const int OW = O + (env.time_feature ? 1 : 0);
    float* go = obs + (size_t)e0 * OW;
    float* gn = next_obs + (size_t)e0 * OW;
    for (int i = tid; i < count * OW; i += blockDim.x) {
        const int e = i / OW, j = i % OW;
        go[i] = j < O ? so[e * O + j] : stf[2 * e + 1];
        gn[i] = j < O ? sx[e * O + j] : stf[2 * e];
    }
}

// One vector step of the on-policy collector after the actor forward pass, in ONE launch
// (trainer.py:44-50): Normal(loc, scale).sample() + summed log-prob (a2c.py:75-85; the Philox
// stream and arithmetic of gauss_sample_kernel, csrc/heads.cu), MeanStd.record of the acting
// observations (mean_stds.py:36-37: the environment state IS the acting observation row), and
// Sequential.step (env_step_kernel above).  Replaces gauss_sample + counter_add + moments_record
// + env_step of the per-step chain: 5 launches per vector step -> 2.
__global__ void __launch_bounds__(256)
act_env_step_kernel(TbEnv env, const float* __restrict__ loc_pre, const float* __restrict__ log_scale,
                    uint64_t seed, uint64_t counter, const uint64_t* __restrict__ d_counter,
                    float* __restrict__ actions, float* __restrict__ log_probs, double* moment_sums,
                    float* __restrict__ obs, float* __restrict__ next_obs, float* __restrict__ rewards,
                    float* __restrict__ resets, float* __restrict__ terminations, int tile_envs) {
    extern __shared__ float smem[];
    const int O = env.obs_dim, A = env.act_dim;
    float* sx = smem;                          // [tile, O] state (= acting obs) -> transition obs
    float* so = sx + (size_t)tile_envs * O;    // [tile, O] acting obs of the next step (post reset)
    float* sa = so + (size_t)tile_envs * O;    // [tile, A] clipped actions
    __shared__ float s_scale[kMaxAct];
    if (d_counter) counter += *d_counter;      // device-resident stream position (CUDA graphs)

    const int e0 = blockIdx.x * tile_envs;
    const int count = min(tile_envs, env.n_envs - e0);
    const int tid = threadIdx.x;
    const float* gx = env.d_state + (size_t)e0 * O;
    for (int i = tid; i < count * O; i += blockDim.x) sx[i] = gx[i];
    if (tid < A) s_scale[tid] = detached_scale(log_scale[tid], nullptr);
    __syncthreads();

    // ---- MeanStd.record(acting observations): float64 column sums of this tile ------------
    if (moment_sums) {
        if (tid < O) {
            double s = 0.0, q = 0.0;
            for (int e = 0; e < count; ++e) {
                const double v = (double)sx[e * O + tid];
                s += v;
                q += v * v;
            }
            atomicAdd(&moment_sums[tid], s);
            atomicAdd(&moment_sums[O + tid], q);
        } else if (tid == 64) {
            atomicAdd(&moment_sums[2 * O], (double)count);
        }
    }
    // ---- sample (thread per environment; identical to gauss_sample_kernel) -------------------
    if (tid < count) {
        const int64_t n = e0 + tid;
        Philox rng(seed);
        float lp = 0.0f;
        float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int a = 0; a < A; ++a) {
            if ((a & 3) == 0) {
                const uint4 r = rng(counter + (uint64_t)n, (uint64_t)(a >> 2));
                const float2 p = box_muller(r.x, r.y), q = box_muller(r.z, r.w);
                z = make_float4(p.x, p.y, q.x, q.y);
            }
            const float e = (a & 3) == 0 ? z.x : (a & 3) == 1 ? z.y : (a & 3) == 2 ? z.z : z.w;
            const float loc = tanhf(loc_pre[n * A + a]);
            const float sc = s_scale[a];
            const float act = __fadd_rn(__fmul_rn(e, sc), loc);       // Normal.sample
            actions[n * A + a] = act;
            sa[tid * A + a] = fminf(fmaxf(act, -1.0f), 1.0f);         // wrappers.py:22 np.clip
            const float d = act - loc;
            lp += -(d * d) / (2.0f * (sc * sc)) - logf(sc) - kLogSqrt2Pi;
        }
        log_probs[n] = lp;
    }
    __syncthreads();

    // ---- Sequential.step (same as env_step_kernel, no time feature) ---------------------------
    const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    for (int e = warp; e < count; e += nwarps) {
        const int n = e0 + e;
        float* x = sx + (size_t)e * O;
        float reward;
        int term;
        env_transition_warp(x, sa + (size_t)e * A, O, A, lane, &reward, &term);
        int reset = 0;
        uint32_t episode = 0;
        if (lane == 0) {
            int length = env.d_length[n] + 1;
            reset = term || (length == env.max_episode_steps);      // distributed.py:40
            double score = env.d_score[n] + (double)reward;         // trainer.py:52
            episode = env.d_episode[n];
            if (reset) {                                            // trainer.py:64-71
                const unsigned long long slot = atomicAdd(env.d_ep_count, 1ull);
                if (env.log_cap > 0) {
                    env.d_ep_scores[slot % env.log_cap] = score;
                    env.d_ep_lengths[slot % env.log_cap] = length;
                }
                env.d_episode[n] = episode + 1u;
                length = 0;
                score = 0.0;
            }
            env.d_length[n] = length;
            env.d_score[n] = score;
            rewards[n] = reward;
            resets[n] = reset ? 1.0f : 0.0f;
            terminations[n] = term ? 1.0f : 0.0f;
        }
        reset = __shfl_sync(0xffffffffu, reset, 0);
        episode = __shfl_sync(0xffffffffu, episode, 0);
        float* o = so + (size_t)e * O;
        if (reset) {                                                // distributed.py:46-48
            const uint32_t key = reset_key((uint32_t)(env.seed + env.first_worker + n), episode);
            for (int j = lane; j < O; j += 32) o[j] = reset_coordinate(key, j);
        } else {
            for (int j = lane; j < O; j += 32) o[j] = x[j];
        }
    }
    __syncthreads();
    float* gs = env.d_state + (size_t)e0 * O;
    float* go = obs + (size_t)e0 * O;
    float* gn = next_obs + (size_t)e0 * O;
    for (int i = tid; i < count * O; i += blockDim.x) {
        gs[i] = so[i];
        go[i] = so[i];
        gn[i] = sx[i];
    }
}

static int pick_tile(const TbEnv* env, size_t* smem_bytes) {
    const size_t per_env = (size_t)(2 * env->obs_dim + env->act_dim + 2) * sizeof(float);
    int tile = 256;
    while (tile > 8 && ((size_t)tile * per_env > 96 * 1024 ||
                        (int64_t)tile * kNumSMs > (int64_t)env->n_envs))
        tile >>= 1;
    *smem_bytes = (size_t)tile * per_env;
    return tile;
}

// closed-form classic-control tasks (csrc/classic_env.cu)
int classic_env_start(const TbEnv* env, float* d_obs, cudaStream_t s);
int classic_env_step(const TbEnv* env, const float* d_actions, float* d_obs, float* d_next_obs,
                     float* d_rewards, float* d_resets, float* d_terminations, cudaStream_t s);

static int check_task(const TbEnv* env, const char* who) {
    if (env->task == TB_TASK_SYNTH) return 0;
    TB_REQUIRE(env->task == TB_TASK_PENDULUM || env->task == TB_TASK_MOUNTAIN_CAR, TB_EINVAL,
               "%s: unknown task %d", who, env->task);
    TB_REQUIRE(env->d_state64 && env->act_dim == 1 &&
               env->obs_dim == (env->task == TB_TASK_PENDULUM ? 3 : 2), TB_EINVAL,
               "%s: classic task %d needs d_state64, act_dim 1 and its observation size", who, env->task);
    return 0;
}

}  // namespace tb

extern "C" int tb_env_start(const TbEnv* env, float* d_obs, void* stream) {
    tb::ProfScope prof_scope("tb_env_start", stream);
    TB_REQUIRE(env && d_obs && env->n_envs > 0 && env->obs_dim > 0 && env->act_dim > 0,
               TB_EINVAL, "tb_env_start: bad arguments");
    if (int rc = tb::check_task(env, "tb_env_start")) return rc;
    if (env->task != TB_TASK_SYNTH) return tb::classic_env_start(env, d_obs, tb::as_stream(stream));
    const int64_t total = (int64_t)env->n_envs * env->obs_dim;
    const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    cudaMemsetAsync(env->d_ep_count, 0, sizeof(unsigned long long), tb::as_stream(stream));
    tb::env_start_kernel<<<blocks, 256, 0, tb::as_stream(stream)>>>(*env, d_obs);
    return tb::check_launch("tb_env_start");
}

extern "C" int tb_act_env_step(const TbEnv* env, const float* d_loc_pre, const float* d_log_scale,
                               uint64_t seed, uint64_t counter, const uint64_t* d_counter,
                               float* d_actions, float* d_log_probs, double* d_moment_sums,
                               float* d_obs, float* d_next_obs, float* d_rewards, float* d_resets,
                               float* d_terminations, void* stream) {
    tb::ProfScope prof_scope("tb_act_env_step", stream);
    TB_REQUIRE(env && d_loc_pre && d_log_scale && d_actions && d_log_probs && d_obs && d_next_obs &&
               d_rewards && d_resets && d_terminations, TB_EINVAL, "tb_act_env_step: null pointer");
    TB_REQUIRE(env->task == TB_TASK_SYNTH && !env->time_feature, TB_ENOTSUP,
               "tb_act_env_step: SynthControl without the time feature only (use tb_gauss_sample + tb_env_step)");
    TB_REQUIRE(env->act_dim >= 1 && env->act_dim <= tb::kMaxAct && env->obs_dim <= 64, TB_ENOTSUP,
               "tb_act_env_step: needs act_dim <= %d and obs_dim <= 64", tb::kMaxAct);
    size_t smem = 0;
    const int tile = tb::pick_tile(env, &smem);
    TB_REQUIRE(smem <= 200 * 1024, TB_ENOTSUP, "tb_act_env_step: obs_dim %d too large", env->obs_dim);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(tb::act_env_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int blocks = (env->n_envs + tile - 1) / tile;
    tb::act_env_step_kernel<<<blocks, 256, smem, tb::as_stream(stream)>>>(
        *env, d_loc_pre, d_log_scale, seed, counter, d_counter, d_actions, d_log_probs, d_moment_sums,
        d_obs, d_next_obs, d_rewards, d_resets, d_terminations, tile);
    return tb::check_launch("tb_act_env_step");
}

extern "C" int tb_env_step(const TbEnv* env, const float* d_actions, float* d_obs,
                           float* d_next_obs, float* d_rewards, float* d_resets,
                           float* d_terminations, void* stream) {
    tb::ProfScope prof_scope("tb_env_step", stream);
    TB_REQUIRE(env && d_actions && d_obs && d_next_obs && d_rewards && d_resets && d_terminations,
               TB_EINVAL, "tb_env_step: null pointer");
    if (int rc = tb::check_task(env, "tb_env_step")) return rc;
    if (env->task != TB_TASK_SYNTH)
        return tb::classic_env_step(env, d_actions, d_obs, d_next_obs, d_rewards, d_resets, d_terminations,
                                    tb::as_stream(stream));
    size_t smem = 0;
    const int tile = tb::pick_tile(env, &smem);
    TB_REQUIRE(smem <= 200 * 1024, TB_ENOTSUP, "tb_env_step: obs_dim %d too large", env->obs_dim);
    if (smem > 48 * 1024) {
        cudaFuncSetAttribute(tb::env_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem);
    }
    const int blocks = (env->n_envs + tile - 1) / tile;
    tb::env_step_kernel<<<blocks, 256, smem, tb::as_stream(stream)>>>(
        *env, d_actions, d_obs, d_next_obs, d_rewards, d_resets, d_terminations, tile);
    return tb::check_launch("tb_env_step");
}
