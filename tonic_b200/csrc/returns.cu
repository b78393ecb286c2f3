// K4 lambda-returns and K5 advantage normalisation.
// Reference: tonic/replays/utils.py:4-19 (reverse scan, float32, the op order
// written there) and tonic/replays/segments.py:41-46 (whole-array mean / std,
// ddof = 0).  Arrays are [T, N] row-major, so a warp reading one time step for
// 32 consecutive envs is fully coalesced; the scan over T is serial per env.
#include "common.cuh"

namespace tb {

__global__ void __launch_bounds__(128)
lambda_returns_kernel(const float* __restrict__ values, const float* __restrict__ next_values,
                      const float* __restrict__ rewards, const float* __restrict__ resets,
                      const float* __restrict__ terminations, float* __restrict__ returns,
                      int T, int N, float gamma, float lam, float one_minus_lam) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float carry = next_values[(size_t)(T - 1) * N + n];            // utils.py:11
    // software prefetch of the next (earlier) time step hides the dependent-load latency
    size_t i = (size_t)(T - 1) * N + n;
    float nv = next_values[i], r = rewards[i], rs = resets[i], tm = terminations[i];
    for (int t = T - 1; t >= 0; --t) {
        float nv_n = 0.f, r_n = 0.f, rs_n = 0.f, tm_n = 0.f;
        if (t > 0) {
            const size_t p = i - N;
            nv_n = next_values[p]; r_n = rewards[p]; rs_n = resets[p]; tm_n = terminations[p];
        }
        // utils.py:13-18, every operation separately rounded like numpy float32
        float boot = __fadd_rn(__fmul_rn(one_minus_lam, nv), __fmul_rn(lam, carry));
        boot = __fmul_rn(boot, __fsub_rn(1.0f, rs));
        boot = __fadd_rn(boot, __fmul_rn(rs, nv));
        boot = __fmul_rn(boot, __fsub_rn(1.0f, tm));
        carry = __fadd_rn(r, __fmul_rn(gamma, boot));
        returns[i] = carry;
        i -= N;
        nv = nv_n; r = r_n; rs = rs_n; tm = tm_n;
    }
    (void)values;
}

// phase 1: adv = returns - values; ws[0] += sum(adv)
__global__ void __launch_bounds__(256)
adv_diff_kernel(const float* __restrict__ returns, const float* __restrict__ values,
                float* __restrict__ adv, int64_t n, double* ws) {
    __shared__ double scratch[32];
    double acc = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const float a = __fsub_rn(returns[i], values[i]);
        adv[i] = a;
        acc += (double)a;
    }
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) atomicAdd(&ws[0], acc);
}

// phase 2: ws[2] += sum((adv - mean)^2), mean = ws[0] / n_global
__global__ void __launch_bounds__(256)
adv_var_kernel(const float* __restrict__ adv, int64_t n, double* ws, double inv_n) {
    __shared__ double scratch[32];
    const double mean = ws[0] * inv_n;
    double acc = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double d = (double)adv[i] - mean;
        acc += d * d;
    }
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) atomicAdd(&ws[2], acc);
}

// phase 3: adv = (adv - mean) / std unless std == 0   (segments.py:44-45)
__global__ void __launch_bounds__(256)
adv_norm_kernel(float* __restrict__ adv, int64_t n, const double* ws, double inv_n) {
    const float mean = (float)(ws[0] * inv_n);
    const float stdv = (float)sqrt(ws[2] * inv_n);
    if (stdv == 0.0f) return;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        adv[i] = __fdiv_rn(__fsub_rn(adv[i], mean), stdv);
}

}  // namespace tb

extern "C" int tb_lambda_returns(const float* d_values, const float* d_next_values,
                                 const float* d_rewards, const float* d_resets,
                                 const float* d_terminations, float* d_returns,
                                 int32_t T, int32_t N, double discount_factor,
                                 double trace_decay, void* stream) {
    tb::ProfScope prof_scope("tb_lambda_returns", stream);
    TB_REQUIRE(T > 0 && N > 0 && d_next_values && d_rewards && d_resets && d_terminations &&
               d_returns, TB_EINVAL, "tb_lambda_returns: bad arguments");
    const int blocks = (N + 127) / 128;
    tb::lambda_returns_kernel<<<blocks, 128, 0, tb::as_stream(stream)>>>(
        d_values, d_next_values, d_rewards, d_resets, d_terminations, d_returns, T, N,
        (float)discount_factor, (float)trace_decay, (float)(1.0 - trace_decay));
    return tb::check_launch("tb_lambda_returns");
}

extern "C" int tb_advantages(const float* d_returns, const float* d_values,
                             float* d_advantages, int64_t n, double* d_workspace,
                             int64_t n_global, int32_t phase, void* stream) {
    tb::ProfScope prof_scope("tb_advantages", stream);
    TB_REQUIRE(n > 0 && n_global >= n && d_advantages && d_workspace, TB_EINVAL,
               "tb_advantages: bad arguments");
    cudaStream_t s = tb::as_stream(stream);
    int blocks = (int)((n + 255) / 256);
    if (blocks > 4 * tb::kNumSMs) blocks = 4 * tb::kNumSMs;
    const double inv_n = 1.0 / (double)n_global;
    int rc = 0;
    if (phase == 0 || phase == 1) {
        cudaMemsetAsync(d_workspace, 0, 4 * sizeof(double), s);
        tb::adv_diff_kernel<<<blocks, 256, 0, s>>>(d_returns, d_values, d_advantages, n, d_workspace);
        if ((rc = tb::check_launch("tb_advantages/diff"))) return rc;
    }
    if (phase == 0 || phase == 2) {
        tb::adv_var_kernel<<<blocks, 256, 0, s>>>(d_advantages, n, d_workspace, inv_n);
        if ((rc = tb::check_launch("tb_advantages/var"))) return rc;
    }
    if (phase == 0 || phase == 3) {
        tb::adv_norm_kernel<<<blocks, 256, 0, s>>>(d_advantages, n, d_workspace, inv_n);
        if ((rc = tb::check_launch("tb_advantages/norm"))) return rc;
    }
    return 0;
}
