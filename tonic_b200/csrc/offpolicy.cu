// Off-policy heads, targets and actor losses (DDPG / TD3 / SAC).
//
// Reference:
//  * DeterministicPolicyHead            tonic/torch/models/actors.py:101-115
//  * NormalActionNoise / NoActionNoise  tonic/explorations/noisy.py:6-50
//  * TargetActionNoise                  tonic/torch/updaters/critics.py:125-134
//  * GaussianPolicyHead (SAC config) + SquashedMultivariateNormalDiag
//                                       tonic/torch/models/actors.py:7-34,69-98
//  * Q targets  DDPG critics.py:71-75, TD3 :159-167, SAC :205-220
//  * actor losses  DPG actors.py:170-189, soft DPG :238-267
#include "common.cuh"

namespace tb {

constexpr float kLogSqrt2PiO = 0.91893853320467274178f;

__device__ __forceinline__ float softplus_o(float x) { return x > 20.0f ? x : log1pf(expf(x)); }

// SAC scale = clamp(softplus(pre), 1e-4, 1) and its derivative w.r.t. pre
__device__ __forceinline__ float sac_scale(float pre, float* dscale) {
    const float sp = softplus_o(pre);
    const float sc = fminf(fmaxf(sp, 1e-4f), 1.0f);
    if (dscale) *dscale = (sp >= 1e-4f && sp <= 1.0f) ? 1.0f / (1.0f + expf(-pre)) : 0.0f;
    return sc;
}

// Device-resident base of the Philox stream positions (tb_set_noise_base): when set, every noise
// kernel below adds *base to its by-value counter, so a captured CUDA graph draws fresh numbers
// at every replay (the caller advances the base with tb_counter_add inside the graph).
static const uint64_t* g_noise_base = nullptr;

__device__ __forceinline__ uint64_t noise_position(uint64_t counter, const uint64_t* base) {
    return counter + (base ? *base : 0ull);
}

__device__ __forceinline__ float philox_normal(const Philox& rng, uint64_t row, int a, uint64_t counter) {
    const uint4 r = rng(counter + row, (uint64_t)(a >> 2));
    const float2 p = (a & 2) ? box_muller(r.z, r.w) : box_muller(r.x, r.y);
    return (a & 1) ? p.y : p.x;
}

// mode 0: tanh(pre); mode 1: clip(tanh(pre) + clip(scale * eps, +-noise_clip), -1, 1);
// mode 2: uniform(-1, 1) (device warm-up actions)
__global__ void __launch_bounds__(256)
tanh_action_kernel(const float* __restrict__ pre, int64_t total, int A, int mode,
                   const float* __restrict__ noise32, const double* __restrict__ noise64,
                   uint64_t seed, uint64_t counter0, const uint64_t* __restrict__ base,
                   float noise_scale, float noise_clip, float* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const uint64_t counter = noise_position(counter0, base);
    const int64_t row = i / A;
    const int a = (int)(i % A);
    Philox rng(seed);
    if (mode == 2) {
        const uint4 r = rng(counter + (uint64_t)row, (uint64_t)(a >> 2) + (1ull << 32));
        const uint32_t u = (a & 3) == 0 ? r.x : (a & 3) == 1 ? r.y : (a & 3) == 2 ? r.z : r.w;
        out[i] = __fsub_rn(__fmul_rn((float)(u >> 8), 1.1920928955078125e-07f), 1.0f);
        return;
    }
    const float act = tanhf(pre[i]);
    if (mode == 0) { out[i] = act; return; }
    float noisy;
    if (noise64) {        // noisy.py:41-42: float32 actions + float64 noise, cast to float32
        noisy = (float)((double)act + (double)noise_scale * noise64[i]);
    } else {
        const float eps = noise32 ? noise32[i] : philox_normal(rng, row, a, counter);
        float nz = __fmul_rn(noise_scale, eps);                  // critics.py:131
        nz = fminf(fmaxf(nz, -noise_clip), noise_clip);          // critics.py:132
        noisy = __fadd_rn(act, nz);
    }
    out[i] = fminf(fmaxf(noisy, -1.0f), 1.0f);
}

// pre [n, 2A] = [loc | scale pre-activation].  raw = loc + eps * scale, action = tanh(raw),
// log_prob = sum_j Normal(loc, scale).log_prob(raw) - log(1 - action^2 + 1e-6)
__global__ void __launch_bounds__(256)
squashed_sample_kernel(const float* __restrict__ pre, const float* __restrict__ eps_in,
                       uint64_t seed, uint64_t counter0, const uint64_t* __restrict__ base, int64_t n,
                       int A, int greedy, float* __restrict__ actions, float* __restrict__ log_probs,
                       float* __restrict__ eps_out) {
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (row >= n) return;
    const uint64_t counter = noise_position(counter0, base);
    Philox rng(seed);
    float lp = 0.0f;
    for (int a = 0; a < A; ++a) {
        const float loc = pre[row * 2 * A + a];
        if (greedy) {                       // sac.py:48-51: tanh(mean)
            actions[row * A + a] = tanhf(loc);
            continue;
        }
        const float sc = sac_scale(pre[row * 2 * A + A + a], nullptr);
        const float eps = eps_in ? eps_in[row * A + a] : philox_normal(rng, row, a, counter);
        const float raw = __fadd_rn(loc, __fmul_rn(eps, sc));    // Normal.rsample / sample
        const float act = tanhf(raw);
        actions[row * A + a] = act;
        if (eps_out) eps_out[row * A + a] = eps;
        const float d = raw - loc;
        lp += -(d * d) / (2.0f * (sc * sc)) - logf(sc) - kLogSqrt2PiO;
        lp -= logf(1.0f - act * act + 1e-6f);                    // actors.py:16
    }
    if (log_probs && !greedy) log_probs[row] = lp;
}

// targets = r + (1 - term) * gamma * (min(q1, q2) - alpha * logp)   (buffers.py:34-36 discounts)
__global__ void __launch_bounds__(256)
q_target_kernel(const float* __restrict__ rewards, const float* __restrict__ terminations,
                const int64_t* __restrict__ idx, float gamma, const float* __restrict__ q1,
                const float* __restrict__ q2, const float* __restrict__ logp, float alpha,
                int64_t n, float* __restrict__ targets) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t r = idx ? idx[i] : i;
    // gamma < 0: `terminations` is the replay's stored (n-step) discounts column
    const float disc = gamma < 0.0f ? terminations[r] : __fmul_rn(__fsub_rn(1.0f, terminations[r]), gamma);
    float v = q1[i];
    if (q2) v = fminf(v, q2[i]);
    if (logp) v = __fsub_rn(v, __fmul_rn(alpha, logp[i]));
    targets[i] = __fadd_rn(rewards[r], __fmul_rn(disc, v));
}

// loss_i = alpha * logp_i - min(q1_i, q2_i);  dout_k = d loss_i / d q_k (torch.min tie -> 1/2 each)
__global__ void __launch_bounds__(256)
q_actor_loss_kernel(const float* __restrict__ q1, const float* __restrict__ q2,
                    const float* __restrict__ logp, float alpha, int64_t n,
                    float* __restrict__ dout1, float* __restrict__ dout2, double* stats) {
    __shared__ double scratch[32];
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    double loss = 0.0, rows = 0.0;
    if (i < n) {
        float v = q1[i], w1 = 1.0f, w2 = 0.0f;
        if (q2) {
            const float b = q2[i];
            if (b < v) { v = b; w1 = 0.0f; w2 = 1.0f; }
            else if (b == v) { w1 = w2 = 0.5f; }
            dout2[i] = -w2;
        }
        dout1[i] = -w1;
        loss = (logp ? (double)(alpha * logp[i]) : 0.0) - (double)v;
        rows = 1.0;
    }
    double r;
    r = block_sum(loss, scratch); if (threadIdx.x == 0) atomicAdd(&stats[TB_STAT_LOSS], r);
    r = block_sum(rows, scratch); if (threadIdx.x == 0) atomicAdd(&stats[TB_STAT_ROWS], r);
}

// deterministic head: dout = dq/da * (1 - a^2)
__global__ void __launch_bounds__(256)
dpg_head_grad_kernel(const float* __restrict__ dqda, const float* __restrict__ actions,
                     int64_t total, float* __restrict__ dout) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float a = actions[i];
    dout[i] = dqda[i] * (1.0f - a * a);
}

// squashed Gaussian head (SAC actor): gradient of  alpha * logp - min(q1, q2)  w.r.t. the
// head pre-activations [loc | scale_pre]; dqda = dqda1 (+ dqda2) already carry the -w_k factors.
__global__ void __launch_bounds__(256)
sac_head_grad_kernel(const float* __restrict__ pre, const float* __restrict__ eps,
                     const float* __restrict__ actions, const float* __restrict__ dqda1,
                     const float* __restrict__ dqda2, float alpha, int64_t total, int A,
                     float* __restrict__ dout) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t row = i / A;
    const int a = (int)(i % A);
    float dsc;
    const float sc = sac_scale(pre[row * 2 * A + A + a], &dsc);
    const float act = actions[i];
    const float one_m = 1.0f - act * act;
    float dq = dqda1[i];
    if (dqda2) dq += dqda2[i];
    // d/draw [ -alpha * log(1 - a^2 + 1e-6) ] = alpha * 2 a (1 - a^2) / (1 - a^2 + 1e-6)
    const float draw = alpha * (2.0f * act * one_m / (one_m + 1e-6f)) + dq * one_m;
    dout[row * 2 * A + a] = draw;                                   // d/dloc
    dout[row * 2 * A + A + a] = (draw * eps[i] - alpha / sc) * dsc; // d/dscale * dscale/dpre
}

inline int blocks_for(int64_t n) { return (int)((n + 255) / 256); }

// ---- device-resident ring replay (fast mode) ---------------------------------------------------
// out[i] uniform in [0, *d_total): Philox word x 64-bit multiply-high (replays/buffers.py:86
// draws np_random.randint(size * N, size=batch); here the stream is Philox, the support identical)
__global__ void __launch_bounds__(256)
randint_kernel(uint64_t seed, uint64_t stream_id, const uint64_t* __restrict__ d_counter,
               const int64_t* __restrict__ d_total, int64_t n, int64_t* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Philox rng(seed);
    const uint4 r = rng((d_counter ? *d_counter : 0ull) + (uint64_t)i, stream_id);
    const unsigned long long u = ((unsigned long long)r.x << 32) | r.y;
    out[i] = (int64_t)__umul64hi(u, (unsigned long long)*d_total);
}

struct RingCopy {
    const float* src[8];
    float* dst[8];
    int64_t elems[8];       // floats per row of each key
    int n;
};

// row *d_index of every key <- the staged vector-step rows (replays/buffers.py:47-56 store)
__global__ void __launch_bounds__(256)
ring_store_kernel(RingCopy c, const int64_t* __restrict__ d_index) {
    const int64_t row = *d_index;
    const int k = blockIdx.y;
    if (k >= c.n) return;
    float* dst = c.dst[k] + row * c.elems[k];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < c.elems[k];
         i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = c.src[k][i];
}

// index = (index + 1) % max_size; size = min(size + 1, max_size); total = size * n_workers
__global__ void ring_advance_kernel(int64_t* state, int64_t max_size, int64_t n_workers) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        state[0] = (state[0] + 1) % max_size;
        state[1] = min(state[1] + 1, max_size);
        state[2] = state[1] * n_workers;
    }
}

}  // namespace tb

extern "C" int tb_set_noise_base(const uint64_t* d_base) {
    tb::g_noise_base = d_base;
    return 0;
}

extern "C" int tb_randint(uint64_t seed, uint64_t stream_id, const uint64_t* d_counter,
                          const int64_t* d_total, int64_t n, int64_t* d_out, void* stream) {
    tb::ProfScope prof_scope("tb_randint", stream);
    TB_REQUIRE(d_total && d_out && n > 0, TB_EINVAL, "tb_randint: bad arguments");
    tb::randint_kernel<<<tb::blocks_for(n), 256, 0, tb::as_stream(stream)>>>(seed, stream_id, d_counter,
                                                                            d_total, n, d_out);
    return tb::check_launch("tb_randint");
}

extern "C" int tb_ring_store(const float* const* h_src, float* const* h_dst, const int64_t* h_row_elems,
                             int32_t n_keys, const int64_t* d_ring_state, void* stream) {
    tb::ProfScope prof_scope("tb_ring_store", stream);
    TB_REQUIRE(h_src && h_dst && h_row_elems && d_ring_state && n_keys >= 1 && n_keys <= 8, TB_EINVAL,
               "tb_ring_store: bad arguments (1..8 keys)");
    tb::RingCopy c;
    int64_t largest = 0;
    for (int k = 0; k < 8; ++k) {
        c.src[k] = k < n_keys ? h_src[k] : nullptr;
        c.dst[k] = k < n_keys ? h_dst[k] : nullptr;
        c.elems[k] = k < n_keys ? h_row_elems[k] : 0;
        TB_REQUIRE(k >= n_keys || (c.src[k] && c.dst[k] && c.elems[k] > 0), TB_EINVAL,
                   "tb_ring_store: key %d is incomplete", k);
        if (c.elems[k] > largest) largest = c.elems[k];
    }
    c.n = n_keys;
    int bx = (int)((largest + 255) / 256);
    if (bx > 2 * tb::kNumSMs) bx = 2 * tb::kNumSMs;
    tb::ring_store_kernel<<<dim3(bx, n_keys), 256, 0, tb::as_stream(stream)>>>(c, d_ring_state);
    return tb::check_launch("tb_ring_store");
}

extern "C" int tb_ring_advance(int64_t* d_ring_state, int64_t max_size, int64_t n_workers, void* stream) {
    tb::ProfScope prof_scope("tb_ring_advance", stream);
    TB_REQUIRE(d_ring_state && max_size > 0 && n_workers > 0, TB_EINVAL, "tb_ring_advance: bad arguments");
    tb::ring_advance_kernel<<<1, 32, 0, tb::as_stream(stream)>>>(d_ring_state, max_size, n_workers);
    return tb::check_launch("tb_ring_advance");
}

extern "C" int tb_tanh_action(const float* d_pre, int64_t n_rows, int32_t act_dim, int32_t mode,
                              const float* d_noise32, const double* d_noise64, uint64_t seed,
                              uint64_t counter, float noise_scale, float noise_clip,
                              float* d_out, void* stream) {
    tb::ProfScope prof_scope("tb_tanh_action", stream);
    TB_REQUIRE((d_pre || mode == 2) && d_out && n_rows > 0 && act_dim > 0 && mode >= 0 && mode <= 2,
               TB_EINVAL, "tb_tanh_action: bad arguments");
    const int64_t total = n_rows * act_dim;
    tb::tanh_action_kernel<<<tb::blocks_for(total), 256, 0, tb::as_stream(stream)>>>(
        d_pre, total, act_dim, mode, d_noise32, d_noise64, seed, counter, tb::g_noise_base, noise_scale,
        noise_clip, d_out);
    return tb::check_launch("tb_tanh_action");
}

extern "C" int tb_squashed_sample(const float* d_pre, const float* d_eps, uint64_t seed,
                                  uint64_t counter, int64_t n_rows, int32_t act_dim,
                                  int32_t greedy, float* d_actions, float* d_log_probs,
                                  float* d_eps_out, void* stream) {
    tb::ProfScope prof_scope("tb_squashed_sample", stream);
    TB_REQUIRE(d_pre && d_actions && n_rows > 0 && act_dim > 0, TB_EINVAL,
               "tb_squashed_sample: bad arguments");
    tb::squashed_sample_kernel<<<tb::blocks_for(n_rows), 256, 0, tb::as_stream(stream)>>>(
        d_pre, d_eps, seed, counter, tb::g_noise_base, n_rows, act_dim, greedy, d_actions, d_log_probs,
        d_eps_out);
    return tb::check_launch("tb_squashed_sample");
}

extern "C" int tb_q_target(const float* d_rewards, const float* d_terminations,
                           const int64_t* d_idx, double discount_factor, const float* d_q1,
                           const float* d_q2, const float* d_log_probs, double entropy_coeff,
                           int64_t n_rows, float* d_targets, void* stream) {
    tb::ProfScope prof_scope("tb_q_target", stream);
    TB_REQUIRE(d_rewards && d_terminations && d_q1 && d_targets && n_rows > 0, TB_EINVAL,
               "tb_q_target: bad arguments");
    tb::q_target_kernel<<<tb::blocks_for(n_rows), 256, 0, tb::as_stream(stream)>>>(
        d_rewards, d_terminations, d_idx, (float)discount_factor, d_q1, d_q2, d_log_probs,
        (float)entropy_coeff, n_rows, d_targets);
    return tb::check_launch("tb_q_target");
}

extern "C" int tb_q_actor_loss(const float* d_q1, const float* d_q2, const float* d_log_probs,
                               double entropy_coeff, int64_t n_rows, float* d_dout1,
                               float* d_dout2, double* d_stats, void* stream) {
    tb::ProfScope prof_scope("tb_q_actor_loss", stream);
    TB_REQUIRE(d_q1 && d_dout1 && d_stats && n_rows > 0 && (!d_q2 || d_dout2), TB_EINVAL,
               "tb_q_actor_loss: bad arguments");
    tb::q_actor_loss_kernel<<<tb::blocks_for(n_rows), 256, 0, tb::as_stream(stream)>>>(
        d_q1, d_q2, d_log_probs, (float)entropy_coeff, n_rows, d_dout1, d_dout2, d_stats);
    return tb::check_launch("tb_q_actor_loss");
}

extern "C" int tb_dpg_head_grad(const float* d_dqda, const float* d_actions, int64_t n_rows,
                                int32_t act_dim, float* d_dout, void* stream) {
    tb::ProfScope prof_scope("tb_dpg_head_grad", stream);
    TB_REQUIRE(d_dqda && d_actions && d_dout && n_rows > 0 && act_dim > 0, TB_EINVAL,
               "tb_dpg_head_grad: bad arguments");
    const int64_t total = n_rows * act_dim;
    tb::dpg_head_grad_kernel<<<tb::blocks_for(total), 256, 0, tb::as_stream(stream)>>>(
        d_dqda, d_actions, total, d_dout);
    return tb::check_launch("tb_dpg_head_grad");
}

extern "C" int tb_sac_head_grad(const float* d_pre, const float* d_eps, const float* d_actions,
                                const float* d_dqda1, const float* d_dqda2, double entropy_coeff,
                                int64_t n_rows, int32_t act_dim, float* d_dout, void* stream) {
    tb::ProfScope prof_scope("tb_sac_head_grad", stream);
    TB_REQUIRE(d_pre && d_eps && d_actions && d_dqda1 && d_dout && n_rows > 0 && act_dim > 0,
               TB_EINVAL, "tb_sac_head_grad: bad arguments");
    const int64_t total = n_rows * act_dim;
    tb::sac_head_grad_kernel<<<tb::blocks_for(total), 256, 0, tb::as_stream(stream)>>>(
        d_pre, d_eps, d_actions, d_dqda1, d_dqda2, (float)entropy_coeff, total, act_dim, d_dout);
    return tb::check_launch("tb_sac_head_grad");
}

// ---- n-step returns of the ring replay (replays/buffers.py:58-79) -------------------------
// Called after row `index` was written: for the previous min(size, return_steps - 1) rows, while
// no reset separates them from the new transition, rewards += discounts * new_rewards,
// discounts *= new_discounts, next_observations = new next_observations (float32, every product
// and sum rounded separately like the numpy expressions).  Thread (worker, coordinate).
namespace tb {
__global__ void __launch_bounds__(256)
accumulate_n_steps_kernel(float* __restrict__ rewards, float* __restrict__ discounts,
                          float* __restrict__ next_obs, const float* __restrict__ resets,
                          int index, int count, int max_size, int N, int O) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= (int64_t)N * O) return;
    const int w = (int)(t / O), j = (int)(t % O);
    const float new_reward = rewards[(size_t)index * N + w];
    const float new_discount = discounts[(size_t)index * N + w];
    const float new_obs = next_obs[((size_t)index * N + w) * O + j];
    float mask = 1.0f;
    for (int i = 0; i < count; ++i) {
        const int row = ((index - i - 1) % max_size + max_size) % max_size;
        const size_t e = (size_t)row * N + w;
        mask = __fmul_rn(mask, __fsub_rn(1.0f, resets[e]));
        const float keep = __fsub_rn(1.0f, mask);
        if (j == 0) {
            const float r_old = rewards[e], d_old = discounts[e];
            const float r_new = __fadd_rn(r_old, __fmul_rn(d_old, new_reward));
            rewards[e] = __fadd_rn(__fmul_rn(keep, r_old), __fmul_rn(mask, r_new));
            const float d_new = __fmul_rn(d_old, new_discount);
            discounts[e] = __fadd_rn(__fmul_rn(keep, d_old), __fmul_rn(mask, d_new));
        }
        const size_t eo = e * O + j;
        next_obs[eo] = __fadd_rn(__fmul_rn(keep, next_obs[eo]), __fmul_rn(mask, new_obs));
    }
}
}  // namespace tb

extern "C" int tb_replay_accumulate_n_steps(float* d_rewards, float* d_discounts, float* d_next_obs,
                                            const float* d_resets, int32_t index, int32_t size,
                                            int32_t max_size, int32_t n_workers, int32_t obs_dim,
                                            int32_t return_steps, void* stream) {
    tb::ProfScope prof_scope("tb_replay_accumulate_n_steps", stream);
    TB_REQUIRE(d_rewards && d_discounts && d_next_obs && d_resets && index >= 0 && index < max_size &&
               size >= 0 && n_workers > 0 && obs_dim > 0 && return_steps >= 1, TB_EINVAL,
               "tb_replay_accumulate_n_steps: bad arguments");
    const int count = size < return_steps - 1 ? size : return_steps - 1;
    TB_REQUIRE(count < max_size, TB_EINVAL, "tb_replay_accumulate_n_steps: return_steps exceeds the ring");
    if (count <= 0) return 0;
    const int64_t total = (int64_t)n_workers * obs_dim;
    tb::accumulate_n_steps_kernel<<<tb::blocks_for(total), 256, 0, tb::as_stream(stream)>>>(
        d_rewards, d_discounts, d_next_obs, d_resets, index, count, max_size, n_workers, obs_dim);
    return tb::check_launch("tb_replay_accumulate_n_steps");
}

extern "C" int tb_q_target_discounts(const float* d_rewards, const float* d_discounts, const int64_t* d_idx,
                                     const float* d_q1, const float* d_q2, const float* d_log_probs,
                                     double entropy_coeff, int64_t n_rows, float* d_targets, void* stream) {
    tb::ProfScope prof_scope("tb_q_target", stream);
    TB_REQUIRE(d_rewards && d_discounts && d_q1 && d_targets && n_rows > 0, TB_EINVAL,
               "tb_q_target_discounts: bad arguments");
    tb::q_target_kernel<<<tb::blocks_for(n_rows), 256, 0, tb::as_stream(stream)>>>(
        d_rewards, d_discounts, d_idx, -1.0f, d_q1, d_q2, d_log_probs, (float)entropy_coeff, n_rows,
        d_targets);
    return tb::check_launch("tb_q_target_discounts");
}
