// Shared helpers for the tonic_b200 sm_100a kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/tonic_b200.h"

namespace tb {

// ---- host-side error plumbing ------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch();

// Optional per-kernel timing with CUDA events recorded on the launching stream
// (tb_profile_begin / tb_profile_end): used by bench.py for the live roofline.
bool profiling_enabled();
void profile_mark(const char* name, cudaStream_t stream, bool begin);

struct ProfScope {
    const char* name;
    cudaStream_t stream;
    bool on;
    ProfScope(const char* n, void* s) : name(n), stream(reinterpret_cast<cudaStream_t>(s)),
                                        on(profiling_enabled()) {
        if (on) profile_mark(name, stream, true);
    }
    ~ProfScope() {
        if (on) profile_mark(name, stream, false);
    }
};

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    count_launch();
    return 0;
}

#define TB_REQUIRE(cond, code, ...)              \
    do {                                         \
        if (!(cond)) {                           \
            tb::set_error(__VA_ARGS__);          \
            return (code);                       \
        }                                        \
    } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs

// ---- device helpers ----------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ long long warp_sum(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of a double; result valid in thread 0.  `scratch` >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? scratch[threadIdx.x] : 0.0;
    if (warp == 0) v = warp_sum(v);
    return v;
}

// murmur3 finaliser -- shared with oracle/synth_env.py::fmix32
__host__ __device__ __forceinline__ uint32_t fmix32(uint32_t h) {
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}

// Philox4x32-10 counter-based generator (device noise in fast mode).
struct Philox {
    uint32_t key[2];
    __device__ __forceinline__ Philox(uint64_t seed) {
        key[0] = (uint32_t)seed;
        key[1] = (uint32_t)(seed >> 32);
    }
    __device__ __forceinline__ uint4 operator()(uint64_t ctr_lo, uint64_t ctr_hi) const {
        uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32);
        uint32_t c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
        uint32_t k0 = key[0], k1 = key[1];
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            c0 = hi1 ^ c1 ^ k0;
            c1 = lo1;
            c2 = hi0 ^ c3 ^ k1;
            c3 = lo0;
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};

// two uint32 -> two standard normals (Box-Muller)
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
    const float u1 = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f);   // (0,1)
    const float u2 = ((float)(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float r = sqrtf(-2.0f * logf(u1));
    float s, c;
    sincospif(2.0f * u2, &s, &c);
    return make_float2(r * c, r * s);
}

// tanh through one MUFU.EX2 and one fast division: absolute error ~1e-7 (the libm-grade
// tanhf costs ~45 instructions and dominated the tensor-core epilogues)
__device__ __forceinline__ float tanh_fast(float x) {
    const float e = __expf(-2.0f * fabsf(x));
    return copysignf(__fdividef(1.0f - e, 1.0f + e), x);
}

// ---- detached-scale Gaussian policy head (reference tonic/torch/models/actors.py:37-66) ----
constexpr int kMaxAct = 64;
constexpr float kLogSqrt2Pi = 0.91893853320467274178f;     // log(sqrt(2*pi))
constexpr float kEntropyConst = 1.41893853320467274178f;   // 0.5 + 0.5*log(2*pi)

__device__ __forceinline__ float softplus_f(float x) {     // torch softplus, beta=1, threshold=20
    return x > 20.0f ? x : log1pf(expf(x));
}
// scale = clamp(softplus(log_scale) + 1e-8, 1e-4, 1)   (actors.py:63-64)
__device__ __forceinline__ float detached_scale(float log_scale, float* dscale_dls) {
    const float sp = softplus_f(log_scale) + 1e-8f;
    const float sc = fminf(fmaxf(sp, 1e-4f), 1.0f);
    if (dscale_dls) {
        const float sig = 1.0f / (1.0f + expf(-log_scale));
        *dscale_dls = (sp >= 1e-4f && sp <= 1.0f) ? sig : 0.0f;    // clamp passes grad inside
    }
    return sc;
}


__device__ __forceinline__ bool skip_requested(const int32_t* d_skip) {
    return d_skip != nullptr && *reinterpret_cast<const volatile int32_t*>(d_skip) != 0;
}

}  // namespace tb
