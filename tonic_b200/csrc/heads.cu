// Policy / value heads and losses (small elementwise kernels on [rows, n_out]).
//
// Reference:
//  * DetachedScaleGaussianPolicyHead  tonic/torch/models/actors.py:37-66
//  * sampling + log-prob              tonic/torch/agents/a2c.py:75-85
//    (torch.distributions.Normal: sample = eps * scale + loc, separately rounded;
//     log_prob = -((a - loc)^2) / (2 scale^2) - log(scale) - log(sqrt(2 pi)))
//  * ClippedRatio                     tonic/torch/updaters/actors.py:70-112
//  * StochasticPolicyGradient         tonic/torch/updaters/actors.py:21-50
//  * VRegression / Q losses (MSE)     tonic/torch/updaters/critics.py:18-28,77-86
#include "common.cuh"

namespace tb {

__global__ void __launch_bounds__(256)
gauss_sample_kernel(const float* __restrict__ loc_pre, const float* __restrict__ log_scale,
                    const float* __restrict__ eps, uint64_t seed, uint64_t counter,
                    const uint64_t* __restrict__ d_counter, int64_t n_rows, int A,
                    float* __restrict__ actions, float* __restrict__ log_probs) {
    if (d_counter) counter += *d_counter;      // device-resident stream position (CUDA graphs)
    __shared__ float s_scale[kMaxAct];
    if ((int)threadIdx.x < A) s_scale[threadIdx.x] = detached_scale(log_scale[threadIdx.x], nullptr);
    __syncthreads();
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n >= n_rows) return;
    Philox rng(seed);
    float lp = 0.0f;
    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int a = 0; a < A; ++a) {
        float e;
        if (eps) {
            e = eps[n * A + a];
        } else {
            if ((a & 3) == 0) {
                const uint4 r = rng(counter + (uint64_t)n, (uint64_t)(a >> 2));
                const float2 p = box_muller(r.x, r.y), q = box_muller(r.z, r.w);
                z = make_float4(p.x, p.y, q.x, q.y);
            }
            e = (a & 3) == 0 ? z.x : (a & 3) == 1 ? z.y : (a & 3) == 2 ? z.z : z.w;
        }
        const float loc = tanhf(loc_pre[n * A + a]);
        const float sc = s_scale[a];
        const float act = __fadd_rn(__fmul_rn(e, sc), loc);       // Normal.sample
        actions[n * A + a] = act;
        const float d = act - loc;
        lp += -(d * d) / (2.0f * (sc * sc)) - logf(sc) - kLogSqrt2Pi;
    }
    log_probs[n] = lp;
}

__global__ void __launch_bounds__(256)
gauss_policy_loss_kernel(const float* __restrict__ loc_pre, const float* __restrict__ log_scale,
                         const float* __restrict__ actions, const float* __restrict__ advantages,
                         const float* __restrict__ old_log_probs, const int64_t* __restrict__ idx,
                         int64_t n_rows, int A, float ratio_clip, float entropy_coeff,
                         float* __restrict__ dout, double* stats, const int32_t* d_skip) {
    if (skip_requested(d_skip)) return;
    __shared__ float s_scale[kMaxAct], s_dsc[kMaxAct];
    __shared__ double scratch[32];
    if ((int)threadIdx.x < A)
        s_scale[threadIdx.x] = detached_scale(log_scale[threadIdx.x], &s_dsc[threadIdx.x]);
    __syncthreads();
    float ent_row = 0.0f, std_row = 0.0f;
    for (int a = 0; a < A; ++a) {
        ent_row += kEntropyConst + logf(s_scale[a]);
        std_row += s_scale[a];
    }
    double st_loss = 0, st_kl = 0, st_clip = 0, st_nz = 0, st_rows = 0;
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n_rows) {
        const int64_t r = idx ? idx[i] : i;
        const float adv = advantages[r], old_lp = old_log_probs[r];
        float lp = 0.0f;
        for (int a = 0; a < A; ++a) {
            const float loc = tanhf(loc_pre[i * A + a]);
            const float sc = s_scale[a];
            const float d = actions[r * A + a] - loc;
            lp += -(d * d) / (2.0f * (sc * sc)) - logf(sc) - kLogSqrt2Pi;
        }
        float g_lp, loss;
        if (ratio_clip > 0.0f) {                               // actors.py:84-90
            const float ratio = expf(lp - old_lp);
            const float lo = 1.0f - ratio_clip, hi = 1.0f + ratio_clip;
            const float clipped = fminf(fmaxf(ratio, lo), hi);
            const float s1 = adv * ratio, s2 = adv * clipped;
            loss = -fminf(s1, s2);
            // torch.min backward: grad to the smaller, split evenly on ties;
            // clamp passes the gradient inside [lo, hi]
            const float w1 = s1 < s2 ? 1.0f : (s1 == s2 ? 0.5f : 0.0f);
            const float w2 = (1.0f - w1) * ((ratio >= lo && ratio <= hi) ? 1.0f : 0.0f);
            g_lp = -adv * ratio * (w1 + w2);
            st_clip = (ratio > hi || ratio < lo) ? 1.0 : 0.0;  // actors.py:105
        } else {                                               // actors.py:34
            loss = -adv * lp;
            g_lp = -adv;
        }
        for (int a = 0; a < A; ++a) {
            const float loc = tanhf(loc_pre[i * A + a]);
            const float sc = s_scale[a];
            const float d = actions[r * A + a] - loc;
            const float inv_var = 1.0f / (sc * sc);
            dout[i * 2 * A + a] = g_lp * d * inv_var * (1.0f - loc * loc);
            const float dlp_dsc = d * d * inv_var / sc - 1.0f / sc;
            const float dent_dsc = 1.0f / sc;                  // entropy = const + log(scale)
            dout[i * 2 * A + A + a] =
                (g_lp * dlp_dsc - (entropy_coeff / (float)A) * dent_dsc) * s_dsc[a];
        }
        st_loss = loss;
        st_kl = old_lp - lp;                                    // actors.py:103
        st_nz = adv != 0.0f ? 1.0 : 0.0;
        st_rows = 1.0;
    }
    double v;
    v = block_sum(st_loss, scratch); if (threadIdx.x == 0) atomicAdd(&stats[TB_STAT_LOSS], v);
    v = block_sum(st_kl, scratch);   if (threadIdx.x == 0) atomicAdd(&stats[TB_STAT_KL], v);
    v = block_sum(st_clip, scratch); if (threadIdx.x == 0) atomicAdd(&stats[TB_STAT_CLIPPED], v);
    v = block_sum(st_nz, scratch);   if (threadIdx.x == 0) atomicAdd(&stats[TB_STAT_NONZERO_ADV], v);
    v = block_sum(st_rows, scratch);
    if (threadIdx.x == 0) {
        atomicAdd(&stats[TB_STAT_ROWS], v);
        atomicAdd(&stats[TB_STAT_ENTROPY], v * (double)ent_row);
        atomicAdd(&stats[TB_STAT_STD], v * (double)std_row);
    }
}

__global__ void __launch_bounds__(256)
mse_loss_kernel(const float* __restrict__ values, const float* __restrict__ targets,
                const int64_t* __restrict__ idx, int64_t n_rows, float* __restrict__ dout,
                int ld_dout, double* stats, int stat_slot, int count_rows,
                const int32_t* d_skip) {
    if (skip_requested(d_skip)) return;
    __shared__ double scratch[32];
    double st_loss = 0, st_val = 0, st_rows = 0;
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n_rows) {
        const float v = values[i];
        const float t = targets[idx ? idx[i] : i];
        const float d = v - t;
        dout[i * ld_dout] = 2.0f * d;                           // d/dv of (v - t)^2
        st_loss = (double)d * (double)d;
        st_val = v;
        st_rows = 1.0;
    }
    double r;
    r = block_sum(st_loss, scratch); if (threadIdx.x == 0) atomicAdd(&stats[TB_STAT_LOSS], r);
    r = block_sum(st_val, scratch);  if (threadIdx.x == 0) atomicAdd(&stats[stat_slot], r);
    if (count_rows) {
        r = block_sum(st_rows, scratch);
        if (threadIdx.x == 0) atomicAdd(&stats[TB_STAT_ROWS], r);
    }
}

}  // namespace tb

extern "C" int tb_gauss_sample(const float* d_loc_pre, const float* d_log_scale,
                               const float* d_eps, uint64_t seed, uint64_t counter,
                               const uint64_t* d_counter, int64_t n_rows, int32_t act_dim,
                               float* d_actions, float* d_log_probs, void* stream) {
    tb::ProfScope prof_scope("tb_gauss_sample", stream);
    TB_REQUIRE(d_loc_pre && d_log_scale && d_actions && d_log_probs && n_rows > 0 &&
               act_dim >= 1 && act_dim <= tb::kMaxAct, TB_EINVAL, "tb_gauss_sample: bad arguments");
    const int blocks = (int)((n_rows + 255) / 256);
    tb::gauss_sample_kernel<<<blocks, 256, 0, tb::as_stream(stream)>>>(
        d_loc_pre, d_log_scale, d_eps, seed, counter, d_counter, n_rows, act_dim, d_actions,
        d_log_probs);
    return tb::check_launch("tb_gauss_sample");
}

extern "C" int tb_gauss_policy_loss(const float* d_loc_pre, const float* d_log_scale,
                                    const float* d_actions, const float* d_advantages,
                                    const float* d_old_log_probs, const int64_t* d_idx,
                                    int64_t n_rows, int32_t act_dim, float ratio_clip,
                                    float entropy_coeff, float* d_dout, double* d_stats,
                                    const int32_t* d_skip, void* stream) {
    tb::ProfScope prof_scope("tb_gauss_policy_loss", stream);
    TB_REQUIRE(d_loc_pre && d_log_scale && d_actions && d_advantages && d_old_log_probs &&
               d_dout && d_stats && n_rows > 0 && act_dim >= 1 && act_dim <= tb::kMaxAct,
               TB_EINVAL, "tb_gauss_policy_loss: bad arguments");
    const int blocks = (int)((n_rows + 255) / 256);
    tb::gauss_policy_loss_kernel<<<blocks, 256, 0, tb::as_stream(stream)>>>(
        d_loc_pre, d_log_scale, d_actions, d_advantages, d_old_log_probs, d_idx, n_rows, act_dim,
        ratio_clip, entropy_coeff, d_dout, d_stats, d_skip);
    return tb::check_launch("tb_gauss_policy_loss");
}

extern "C" int tb_mse_loss(const float* d_values, const float* d_targets, const int64_t* d_idx,
                           int64_t n_rows, float* d_dout, int32_t ld_dout, double* d_stats,
                           int32_t stat_slot, int32_t count_rows, const int32_t* d_skip,
                           void* stream) {
    tb::ProfScope prof_scope("tb_mse_loss", stream);
    TB_REQUIRE(d_values && d_targets && d_dout && d_stats && n_rows > 0 && ld_dout >= 1 &&
               stat_slot >= 0 && stat_slot < TB_STAT_COUNT, TB_EINVAL, "tb_mse_loss: bad arguments");
    const int blocks = (int)((n_rows + 255) / 256);
    tb::mse_loss_kernel<<<blocks, 256, 0, tb::as_stream(stream)>>>(
        d_values, d_targets, d_idx, n_rows, d_dout, ld_dout, d_stats, stat_slot, count_rows, d_skip);
    return tb::check_launch("tb_mse_loss");
}

// ---- running statistics of an array (trainer.py:46 `logger.store('train/action',
// actions, stats=True)` without shipping the array to the host) -------------------
namespace tb {
__global__ void __launch_bounds__(256)
array_stats_kernel(const float* __restrict__ x, int64_t n, double* acc) {
    __shared__ double scratch[32];
    double s = 0, q = 0, lo = 1e300, hi = -1e300;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double v = (double)x[i];
        s += v; q += v * v;
        lo = fmin(lo, v); hi = fmax(hi, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    double r;
    r = block_sum(s, scratch); if (threadIdx.x == 0) atomicAdd(&acc[1], r);
    r = block_sum(q, scratch); if (threadIdx.x == 0) atomicAdd(&acc[2], r);
    if ((threadIdx.x & 31) == 0) {
        // atomic min / max on doubles through the ordered-integer trick
        auto enc = [](double d) {
            long long b = __double_as_longlong(d);
            return b >= 0 ? b : b ^ 0x7fffffffffffffffll;
        };
        atomicMin(reinterpret_cast<long long*>(&acc[3]), enc(lo));
        atomicMax(reinterpret_cast<long long*>(&acc[4]), enc(hi));
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&acc[0], (double)n);
}
}  // namespace tb

extern "C" int tb_array_stats(const float* d_x, int64_t n, double* d_acc, void* stream) {
    tb::ProfScope prof_scope("tb_array_stats", stream);
    TB_REQUIRE(d_x && d_acc && n > 0, TB_EINVAL, "tb_array_stats: bad arguments");
    int blocks = (int)((n + 255) / 256);
    if (blocks > 4 * tb::kNumSMs) blocks = 4 * tb::kNumSMs;
    tb::array_stats_kernel<<<blocks, 256, 0, tb::as_stream(stream)>>>(d_x, n, d_acc);
    return tb::check_launch("tb_array_stats");
}

// ---- device-side minibatch permutation (fast mode) --------------------------------
// A pseudo-random bijection of [0, n): 4-round Feistel network on the smallest
// even-width power-of-two domain >= n with cycle walking.  Replaces the host
// `RandomState.shuffle` of tonic/replays/segments.py:62 when bit-compatibility with
// numpy's stream is not required (config.indices = 'device'): no host work and no
// host->device copy, O(1) state.
namespace tb {
__global__ void __launch_bounds__(256)
permutation_kernel(uint64_t seed, uint64_t stream_id, const uint64_t* __restrict__ d_counter,
                   int64_t n, int half_bits, int64_t* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (d_counter) stream_id += *d_counter;
    const uint32_t mask = (1u << half_bits) - 1u;
    uint32_t keys[4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
        keys[r] = fmix32((uint32_t)seed ^ fmix32((uint32_t)(seed >> 32) + 0x9E3779B9u * (uint32_t)(r + 1)) ^
                         fmix32((uint32_t)stream_id * 0x85EBCA6Bu + (uint32_t)(stream_id >> 32) + r));
    uint64_t x = (uint64_t)i;
    do {
        uint32_t L = (uint32_t)(x >> half_bits) & mask, R = (uint32_t)x & mask;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const uint32_t t = L ^ (fmix32(R ^ keys[r]) & mask);
            L = R;
            R = t;
        }
        x = ((uint64_t)L << half_bits) | R;
    } while (x >= (uint64_t)n);
    out[i] = (int64_t)x;
}
}  // namespace tb

extern "C" int tb_permutation(uint64_t seed, uint64_t stream_id, const uint64_t* d_counter,
                              int64_t n, int64_t* d_out, void* stream) {
    tb::ProfScope prof_scope("tb_permutation", stream);
    TB_REQUIRE(d_out && n > 0 && n < (1ll << 40), TB_EINVAL, "tb_permutation: bad arguments");
    int bits = 1;
    while ((1ll << bits) < n) ++bits;
    const int half = (bits + 1) / 2;
    tb::permutation_kernel<<<(int)((n + 255) / 256), 256, 0, tb::as_stream(stream)>>>(
        seed, stream_id, d_counter, n, half < 1 ? 1 : half, d_out);
    return tb::check_launch("tb_permutation");
}

// Device-resident stream counters (Philox offsets, permutation ids): kernels captured in a
// CUDA graph read their position from memory, this kernel advances it after they ran.
namespace tb {
__global__ void counter_add_kernel(uint64_t* counter, uint64_t delta) { *counter += delta; }
}  // namespace tb

extern "C" int tb_counter_add(uint64_t* d_counter, uint64_t delta, void* stream) {
    tb::ProfScope prof_scope("tb_counter_add", stream);
    TB_REQUIRE(d_counter, TB_EINVAL, "tb_counter_add: null pointer");
    tb::counter_add_kernel<<<1, 1, 0, tb::as_stream(stream)>>>(d_counter, delta);
    return tb::check_launch("tb_counter_add");
}
