// SynthControl(O, A) dynamics shared by the stand-alone environment kernel
// (csrc/env_step.cu) and the fused rollout kernel (csrc/rollout.cu).  Bit-identical to
// oracle/synth_env.py: separately rounded float32 ops, integer-quantised reward.
#pragma once

#include "common.cuh"

namespace tb {

__device__ __forceinline__ float reset_coordinate(uint32_t key, int j) {
    const uint32_t h = fmix32(key ^ (0x85EBCA6Bu * (uint32_t)(j + 1)));
    return __fsub_rn(__fmul_rn((float)(h >> 8), 1.1920928955078125e-07f), 1.0f);
}
__device__ __forceinline__ uint32_t reset_key(uint32_t seed, uint32_t episode) {
    return fmix32(seed + 0x9E3779B9u * (episode + 1u));
}

constexpr float kDecay = 0.9f;
constexpr float kGain = 0.1f;
constexpr float kTermLimit = 1.0f;
constexpr float kCostScale = (float)(0.01 * 9.5367431640625e-07);   // 0.01 * 2^-20


// One environment transition executed by a full warp: x[0:O] (shared memory) is advanced in
// place to the transition observation; returns (on every lane) the reward and the termination
// flag.  a[0:A] are the already clipped actions.
__device__ __forceinline__ void env_transition_warp(float* x, const float* a, int O, int A, int lane,
                                                    float* reward, int* term) {
    long long xcost = 0, acost = 0;
    for (int j = lane; j < O; j += 32) {
        const float drive = __fmul_rn(kGain, a[j % A]);
        const float keep = (j == 0) ? x[0] : __fmul_rn(kDecay, x[j]);
        const float nx = __fadd_rn(keep, drive);
        x[j] = nx;
        const long long q = __float2ll_rn(__fmul_rn(nx, 256.0f));
        xcost += q * q;
    }
    for (int k = lane; k < A; k += 32) {
        const long long q = __float2ll_rn(__fmul_rn(a[k], 1024.0f));
        acost += q * q;
    }
    xcost = warp_sum(xcost);
    acost = warp_sum(acost);
    __syncwarp();
    const long long cost = 100ll * acost + 16ll * xcost;
    *reward = __fsub_rn(1.0f, __fmul_rn(__ll2float_rn(cost), kCostScale));
    *term = fabsf(x[0]) > kTermLimit;
}

}  // namespace tb
