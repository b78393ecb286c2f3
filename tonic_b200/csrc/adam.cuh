// Adam step and packed-operand refresh shared by the optimizer kernels (csrc/optim.cu) and the
// fused weight-gradient kernel (csrc/tc_gemm.cu).  Reference: torch.optim.Adam single-tensor
// path (torch/optim/adam.py::_single_tensor_adam) as constructed at
// tonic/torch/updaters/actors.py:11-12,58-59 and critics.py:9-10.
#pragma once
#include "common.cuh"

namespace tb {

__device__ __forceinline__ void pack_one(const TbMlpShape& sh, int i, float p, float* packed) {
    const int H = sh.hidden;
    if (i >= sh.off_w1 && i < sh.off_w1 + H * sh.d_in) {           // W1 [H, d_in] -> W1T [d_in, H]
        const int e = i - sh.off_w1, n = e / sh.d_in, k = e % sh.d_in;
        packed[sh.off_w1t + k * H + n] = p;
        if (sh.off_w1_img_hi > 0) {
            // layer-1 B operand of the fused forward kernel (csrc/tc_mlp.cu): element (n, k) of
            // the K-major [256 x 32] tile with the 128-byte swizzle (8-row groups of 1024 B, 16-byte
            // unit index XOR-ed with the row within the group); columns k >= d_in stay zero
            const int word = (n >> 3) * 256 + (n & 7) * 32 + (((k >> 2) ^ (n & 7)) << 2) + (k & 3);
            const float hi = __uint_as_float(__float_as_uint(p) & 0xFFFFE000u);
            packed[sh.off_w1_img_hi + word] = hi;
            packed[sh.off_w1_img_lo + word] = p - hi;
        }
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

// bias corrections of step t: step_size = lr / (1 - beta1^t), bc2_sqrt = sqrt(1 - beta2^t)
__device__ __forceinline__ void adam_corrections(const TbAdam& opt, int t, float* step_size, float* bc2_sqrt) {
    const double bc1 = 1.0 - pow(opt.beta1, (double)t);
    const double bc2 = 1.0 - pow(opt.beta2, (double)t);
    *step_size = (float)(opt.lr / bc1);
    *bc2_sqrt = (float)sqrt(bc2);
}

// one parameter: g = mean gradient
__device__ __forceinline__ void adam_apply(const TbAdam& opt, const TbMlpShape& sh, float* packed, int i,
                                           float g, float step_size, float bc2_sqrt) {
    const float w1 = (float)(1.0 - opt.beta1), w2 = (float)(1.0 - opt.beta2);
    const float m = opt.d_m[i] + w1 * (g - opt.d_m[i]);                          // lerp_
    const float v = opt.d_v[i] * (float)opt.beta2 + w2 * g * g;                  // mul_.addcmul_
    const float denom = sqrtf(v) / bc2_sqrt + (float)opt.eps;
    const float p = opt.d_params[i] - step_size * (m / denom);                   // addcdiv_
    opt.d_m[i] = m;
    opt.d_v[i] = v;
    opt.d_params[i] = p;
    if (packed) pack_one(sh, i, p, packed);
}

}  // namespace tb
