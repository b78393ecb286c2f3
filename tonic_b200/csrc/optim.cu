// K9 Adam step (+ partial-sum reduction of the split weight gradients, refresh of
// the packed forward transposes, and the device-side PPO control flags) and K15
// target-network soft update.
//
// Reference: torch.optim.Adam (single-tensor path, torch 2.11) as constructed at
// tonic/torch/updaters/actors.py:11-12,58-59,161-162,228-229 and
// tonic/torch/updaters/critics.py:9-10,59-60,143-144,190-191 (betas 0.9/0.999,
// eps 1e-8, no weight decay); early stop tonic/torch/agents/ppo.py:45-46 with
// updaters/actors.py:103,112; soft update models/actor_critics.py:68-72,126-130.
#include "common.cuh"

namespace tb {

__device__ __forceinline__ void pack_one(const TbMlpShape& sh, int i, float p, float* packed) {
    const int H = sh.hidden;
    if (i >= sh.off_w1 && i < sh.off_w1 + H * sh.d_in) {           // W1 [H, d_in] -> W1T [d_in, H]
        const int e = i - sh.off_w1, n = e / sh.d_in, k = e % sh.d_in;
        packed[sh.off_w1t + k * H + n] = p;
    } else if (i >= sh.off_w2 && i < sh.off_w2 + H * H) {          // W2 [H, H] -> W2T
        const int e = i - sh.off_w2, n = e / H, k = e % H;
        packed[sh.off_w2t + k * H + n] = p;
        if (sh.off_w2_hi > 0) {      // tf32 splits for the tensor-core path (csrc/tc_gemm.cu)
            const float hi = __uint_as_float(__float_as_uint(p) & 0xFFFFE000u);
            const float lo = p - hi;
            packed[sh.off_w2_hi + e] = hi;
            packed[sh.off_w2_lo + e] = lo;
            packed[sh.off_w2t_hi + k * H + n] = hi;
            packed[sh.off_w2t_lo + k * H + n] = lo;
        }
    }
}

__global__ void __launch_bounds__(256)
adam_kernel(TbAdam opt, TbMlpShape sh, float* __restrict__ packed,
            const float* __restrict__ gpart, int n_split, float grad_scale,
            const int32_t* d_skip, const double* d_stats, float kl_threshold, int32_t* d_stop) {
    if (skip_requested(d_skip)) return;
    if (d_stats && d_stats[TB_STAT_NONZERO_ADV] == 0.0) return;    // actors.py:22,71: no step
    const int t = opt.d_step[0] + 1;                               // incremented by the last block
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < opt.n_params) {
        float g = 0.0f;
        for (int s = 0; s < n_split; ++s) g += gpart[(size_t)s * opt.n_params + i];
        g *= grad_scale;
        // torch/optim/adam.py::_single_tensor_adam
        const float w1 = (float)(1.0 - opt.beta1), w2 = (float)(1.0 - opt.beta2);
        const float m = opt.d_m[i] + w1 * (g - opt.d_m[i]);                          // lerp_
        const float v = opt.d_v[i] * (float)opt.beta2 + w2 * g * g;                  // mul_.addcmul_
        const double bc1 = 1.0 - pow(opt.beta1, (double)t);
        const double bc2 = 1.0 - pow(opt.beta2, (double)t);
        const float step_size = (float)(opt.lr / bc1);
        const float bc2_sqrt = (float)sqrt(bc2);
        const float denom = sqrtf(v) / bc2_sqrt + (float)opt.eps;
        const float p = opt.d_params[i] - step_size * (m / denom);                   // addcdiv_
        opt.d_m[i] = m;
        opt.d_v[i] = v;
        opt.d_params[i] = p;
        if (packed) pack_one(sh, i, p, packed);
    }
    // last block to finish publishes the new step count and the KL early-stop flag
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int done = atomicAdd(&opt.d_step[1], 1);
        is_last = done == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        opt.d_step[0] = t;
        opt.d_step[1] = 0;
        if (d_stop && d_stats && kl_threshold >= 0.0f) {
            const float kl = (float)(d_stats[TB_STAT_KL] / d_stats[TB_STAT_ROWS]);
            if (kl > kl_threshold) *d_stop = 1;
        }
    }
}

__global__ void __launch_bounds__(256)
pack_kernel(TbMlpShape sh, const float* __restrict__ params, float* __restrict__ packed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < sh.n_params) pack_one(sh, i, params[i], packed);
}

__global__ void __launch_bounds__(256)
soft_update_kernel(float* __restrict__ target, const float* __restrict__ online, int64_t n,
                   float keep, float tau) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        // t.mul_(1 - tau); t.add_(tau * o)   (actor_critics.py:70-72), separately rounded
        target[i] = __fadd_rn(__fmul_rn(target[i], keep), __fmul_rn(tau, online[i]));
}

}  // namespace tb

extern "C" int tb_adam_step(const TbAdam* opt, const TbMlpShape* shape, float* d_packed,
                            const float* d_gpart, int32_t n_split, float grad_scale,
                            const int32_t* d_skip, const double* d_stats, float kl_threshold,
                            int32_t* d_stop, void* stream) {
    tb::ProfScope prof_scope("tb_adam_step", stream);
    TB_REQUIRE(opt && opt->d_params && opt->d_m && opt->d_v && opt->d_step && d_gpart &&
               n_split >= 1 && opt->n_params > 0, TB_EINVAL, "tb_adam_step: bad arguments");
    TB_REQUIRE(!d_packed || (shape && shape->n_params == opt->n_params), TB_EINVAL,
               "tb_adam_step: shape/optimizer size mismatch");
    TbMlpShape sh;
    if (shape) sh = *shape; else memset(&sh, 0, sizeof(sh));
    const int blocks = (opt->n_params + 255) / 256;
    tb::adam_kernel<<<blocks, 256, 0, tb::as_stream(stream)>>>(
        *opt, sh, d_packed, d_gpart, n_split, grad_scale, d_skip, d_stats, kl_threshold, d_stop);
    return tb::check_launch("tb_adam_step");
}

extern "C" int tb_mlp_pack(const TbMlpShape* shape, const float* d_params, float* d_packed,
                           void* stream) {
    tb::ProfScope prof_scope("tb_mlp_pack", stream);
    TB_REQUIRE(shape && d_params && d_packed, TB_EINVAL, "tb_mlp_pack: null pointer");
    const int blocks = (shape->n_params + 255) / 256;
    tb::pack_kernel<<<blocks, 256, 0, tb::as_stream(stream)>>>(*shape, d_params, d_packed);
    return tb::check_launch("tb_mlp_pack");
}

extern "C" int tb_soft_update(float* d_target, const float* d_online, int64_t n, double tau,
                              void* stream) {
    tb::ProfScope prof_scope("tb_soft_update", stream);
    TB_REQUIRE(d_target && d_online && n > 0, TB_EINVAL, "tb_soft_update: bad arguments");
    int blocks = (int)((n + 255) / 256);
    if (blocks > 8 * tb::kNumSMs) blocks = 8 * tb::kNumSMs;
    // python computes (1 - tau) in double precision before the float multiply
    tb::soft_update_kernel<<<blocks, 256, 0, tb::as_stream(stream)>>>(
        d_target, d_online, n, (float)(1.0 - tau), (float)tau);
    return tb::check_launch("tb_soft_update");
}

// Sum of the split partial gradients into one flat buffer (the send buffer of the
// multi-GPU gradient all-reduce).
namespace tb {
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ gpart, int n_split, int n, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float g = 0.0f;
    for (int s = 0; s < n_split; ++s) g += gpart[(size_t)s * n + i];
    out[i] = g;
}
}  // namespace tb

extern "C" int tb_reduce_partials(const float* d_gpart, int32_t n_split, int32_t n_params,
                                  float* d_out, const int32_t* d_skip, void* stream) {
    tb::ProfScope prof_scope("tb_reduce_partials", stream);
    TB_REQUIRE(d_gpart && d_out && n_split >= 1 && n_params > 0, TB_EINVAL,
               "tb_reduce_partials: bad arguments");
    (void)d_skip;
    tb::reduce_partials_kernel<<<(n_params + 255) / 256, 256, 0, tb::as_stream(stream)>>>(
        d_gpart, n_split, n_params, d_out);
    return tb::check_launch("tb_reduce_partials");
}
