// Closed-form classic-control tasks as device-resident vector environments:
// Pendulum-v0/v1 and MountainCarContinuous-v0 behind tonic.environments.Gym(name)
// (reference environments/builders.py:12-16,43-78 builds the Gym task, wraps it in
// ActionRescaler -- wrappers.py:7-22 -- and an optional TimeFeature -- wrappers.py:25-54;
// Sequential.step, environments/distributed.py:28-58, adds auto-reset and time-outs).
// The dynamics are the ones of tonic_b200/environments/classic.py, operation by operation:
// float64 state, individually rounded adds / multiplies (no FMA contraction), sin / cos from
// the same portable polynomial (environments/portable_math.py), counter-based reset stream
// -- so the numpy classes and these kernels agree bit for bit.  One thread per environment.
#include "env_dynamics.cuh"

namespace tb {

// ---- portable_math.sincos -------------------------------------------------------------------
__device__ __forceinline__ void portable_sincos(double x, double* s_out, double* c_out) {
    const double k = rint(__dmul_rn(x, 6.36619772367581382433e-01));
    const double r = __dsub_rn(__dsub_rn(x, __dmul_rn(k, 1.57079632673412561417e+00)),
                               __dmul_rn(k, 6.07710050650619224932e-11));
    const double z = __dmul_rn(r, r);
    double ps = 1.58969099521155010221e-10;
    ps = __dadd_rn(-2.50507602534068634195e-08, __dmul_rn(z, ps));
    ps = __dadd_rn(2.75573137070700676789e-06, __dmul_rn(z, ps));
    ps = __dadd_rn(-1.98412698298579493134e-04, __dmul_rn(z, ps));
    ps = __dadd_rn(8.33333333332248946124e-03, __dmul_rn(z, ps));
    ps = __dadd_rn(-1.66666666666666324348e-01, __dmul_rn(z, ps));
    const double s = __dadd_rn(r, __dmul_rn(__dmul_rn(r, z), ps));
    double pc = -1.13596475577881948265e-11;
    pc = __dadd_rn(2.08757232129817482790e-09, __dmul_rn(z, pc));
    pc = __dadd_rn(-2.75573143513906633035e-07, __dmul_rn(z, pc));
    pc = __dadd_rn(2.48015872894767294178e-05, __dmul_rn(z, pc));
    pc = __dadd_rn(-1.38888888888741095749e-03, __dmul_rn(z, pc));
    pc = __dadd_rn(4.16666666666666019037e-02, __dmul_rn(z, pc));
    const double c = __dadd_rn(__dsub_rn(1.0, __dmul_rn(0.5, z)), __dmul_rn(__dmul_rn(z, z), pc));
    const int q = (int)((long long)k & 3ll);
    *s_out = q == 0 ? s : (q == 1 ? c : (q == 2 ? -s : -c));
    *c_out = q == 0 ? c : (q == 1 ? -s : (q == 2 ? -c : s));
}

// portable_math.reset_uniform: 24 hashed bits -> [0, 1)
__device__ __forceinline__ double reset_uniform(uint32_t seed, uint32_t episode, int coordinate) {
    const uint32_t h = fmix32(reset_key(seed, episode) ^ (0x85EBCA6Bu * (uint32_t)(coordinate + 1)));
    return (double)(h >> 8) * 5.9604644775390625e-08;       // 2^-24, exact
}

constexpr double kPi = 3.141592653589793, kTwoPi = 6.283185307179586;

__device__ __forceinline__ void classic_reset(int task, uint32_t seed, uint32_t episode, double* st) {
    if (task == TB_TASK_PENDULUM) {
        st[0] = __dadd_rn(-kPi, __dmul_rn(kTwoPi, reset_uniform(seed, episode, 0)));
        st[1] = __dadd_rn(-1.0, __dmul_rn(2.0, reset_uniform(seed, episode, 1)));
    } else {
        st[0] = __dadd_rn(-0.6, __dmul_rn(0.2, reset_uniform(seed, episode, 0)));
        st[1] = 0.0;
    }
}

// observation of a state (float32 casts of the float64 values, like np.array(.., np.float32))
__device__ __forceinline__ void classic_observe(int task, const double* st, float* o) {
    if (task == TB_TASK_PENDULUM) {
        double s, c;
        portable_sincos(st[0], &s, &c);
        o[0] = (float)c; o[1] = (float)s; o[2] = (float)st[1];
    } else {
        o[0] = (float)st[0]; o[1] = (float)st[1];
    }
}

// one transition: st advanced in place; `a` = policy action in [-1, 1] (ActionRescaler input)
__device__ __forceinline__ void classic_transition(int task, double* st, float a, double* reward, int* term) {
    const float clipped = fminf(fmaxf(a, -1.0f), 1.0f);                      // wrappers.py:22
    if (task == TB_TASK_PENDULUM) {
        const double u0 = (double)__fadd_rn(0.0f, __fmul_rn(2.0f, clipped));   // bias + scale * clip (float32)
        const double u = fmin(fmax(u0, -2.0), 2.0);
        const double theta = st[0], theta_dot = st[1];
        double wrapped = fmod(__dadd_rn(theta, kPi), kTwoPi);                 // python float % (sign of divisor)
        if (wrapped != 0.0 && wrapped < 0.0) wrapped = __dadd_rn(wrapped, kTwoPi);
        wrapped = __dsub_rn(wrapped, kPi);
        const double cost = __dadd_rn(
            __dadd_rn(__dmul_rn(wrapped, wrapped), __dmul_rn(0.1, __dmul_rn(theta_dot, theta_dot))),
            __dmul_rn(0.001, __dmul_rn(u, u)));
        double s, c;
        portable_sincos(theta, &s, &c);
        double nd = __dadd_rn(theta_dot,
                              __dmul_rn(__dadd_rn(__dmul_rn(15.0, s), __dmul_rn(3.0, u)), 0.05));
        nd = fmin(fmax(nd, -8.0), 8.0);
        st[0] = __dadd_rn(theta, __dmul_rn(nd, 0.05));
        st[1] = nd;
        *reward = -cost;
        *term = 0;
    } else {
        const double force = fmin(fmax((double)__fadd_rn(0.0f, __fmul_rn(1.0f, clipped)), -1.0), 1.0);
        double position = st[0], velocity = st[1];
        double s, c;
        portable_sincos(__dmul_rn(3.0, position), &s, &c);
        velocity = __dadd_rn(velocity, __dsub_rn(__dmul_rn(force, 0.0015), __dmul_rn(0.0025, c)));
        velocity = fmin(fmax(velocity, -0.07), 0.07);
        position = __dadd_rn(position, velocity);
        position = fmin(fmax(position, -1.2), 0.6);
        if (position == -1.2 && velocity < 0.0) velocity = 0.0;
        const int done = position >= 0.45 && velocity >= 0.0;
        *reward = __dsub_rn(done ? 100.0 : 0.0, __dmul_rn(__dmul_rn(force, force), 0.1));
        *term = done;
        st[0] = position;
        st[1] = velocity;
    }
}

__global__ void __launch_bounds__(256)
classic_start_kernel(TbEnv env, float* __restrict__ obs) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= env.n_envs) return;
    const int O = env.obs_dim, OW = O + (env.time_feature ? 1 : 0);
    double st[2];
    classic_reset(env.task, (uint32_t)(env.seed + env.first_worker + n), 0u, st);
    env.d_state64[2 * n] = st[0];
    env.d_state64[2 * n + 1] = st[1];
    float o[3];
    classic_observe(env.task, st, o);
    for (int j = 0; j < O; ++j) obs[(size_t)n * OW + j] = o[j];
    if (env.time_feature) obs[(size_t)n * OW + O] = env.time_low;          // wrappers.py:43
    env.d_length[n] = 0;
    env.d_episode[n] = 1u;
    env.d_score[n] = 0.0;
}

__global__ void __launch_bounds__(256)
classic_step_kernel(TbEnv env, const float* __restrict__ actions, float* __restrict__ obs,
                    float* __restrict__ next_obs, float* __restrict__ rewards,
                    float* __restrict__ resets, float* __restrict__ terminations) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= env.n_envs) return;
    const int O = env.obs_dim, OW = O + (env.time_feature ? 1 : 0);
    double st[2] = {env.d_state64[2 * n], env.d_state64[2 * n + 1]};
    double reward64;
    int term;
    classic_transition(env.task, st, actions[(size_t)n * env.act_dim], &reward64, &term);
    const float reward = (float)reward64;                   // np.array(rewards, np.float32)
    int length = env.d_length[n] + 1;
    const int reset = term || (length == env.max_episode_steps);      // distributed.py:39-40
    float o[3];
    classic_observe(env.task, st, o);
    for (int j = 0; j < O; ++j) next_obs[(size_t)n * OW + j] = o[j];
    float tf_acting = 0.0f;
    if (env.time_feature) {              // wrappers.py:49-52 in float64, then the float32 cast
        const double prop = (double)length / (double)env.max_episode_steps;
        const float v = (float)((double)env.time_low + ((double)env.time_high - (double)env.time_low) * prop);
        next_obs[(size_t)n * OW + O] = v;
        tf_acting = reset ? env.time_low : v;
    }
    double score = env.d_score[n] + (double)reward;          // trainer.py:52
    if (reset) {                                              // trainer.py:64-71, distributed.py:46-48
        const uint32_t episode = env.d_episode[n];
        const unsigned long long slot = atomicAdd(env.d_ep_count, 1ull);
        if (env.log_cap > 0) {
            env.d_ep_scores[slot % env.log_cap] = score;
            env.d_ep_lengths[slot % env.log_cap] = length;
        }
        env.d_episode[n] = episode + 1u;
        classic_reset(env.task, (uint32_t)(env.seed + env.first_worker + n), episode, st);
        classic_observe(env.task, st, o);
        length = 0;
        score = 0.0;
    }
    env.d_state64[2 * n] = st[0];
    env.d_state64[2 * n + 1] = st[1];
    env.d_length[n] = length;
    env.d_score[n] = score;
    for (int j = 0; j < O; ++j) obs[(size_t)n * OW + j] = o[j];
    if (env.time_feature) obs[(size_t)n * OW + O] = tf_acting;
    rewards[n] = reward;
    resets[n] = reset ? 1.0f : 0.0f;
    terminations[n] = term ? 1.0f : 0.0f;
}

int classic_env_start(const TbEnv* env, float* d_obs, cudaStream_t s) {
    cudaMemsetAsync(env->d_ep_count, 0, sizeof(unsigned long long), s);
    classic_start_kernel<<<(env->n_envs + 255) / 256, 256, 0, s>>>(*env, d_obs);
    return check_launch("tb_env_start");
}

int classic_env_step(const TbEnv* env, const float* d_actions, float* d_obs, float* d_next_obs,
                     float* d_rewards, float* d_resets, float* d_terminations, cudaStream_t s) {
    classic_step_kernel<<<(env->n_envs + 255) / 256, 256, 0, s>>>(*env, d_actions, d_obs, d_next_obs,
                                                                   d_rewards, d_resets, d_terminations);
    return check_launch("tb_env_step");
}

}  // namespace tb
