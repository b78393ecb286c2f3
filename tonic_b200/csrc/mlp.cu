// K2/K3/K6-K8/K13/K14 building blocks: two-hidden-layer MLP forward, backward
// and weight-gradient kernels (float32, FFMA "parity" path).
//
// Reference: tonic/torch/models/utils.py:4-23 (MLP = Linear+activation x2),
// tonic/torch/models/encoders.py:4-31 (input = [normalise(obs) | actions]),
// tonic/torch/normalizers/mean_stds.py:34-39, and torch autograd for the
// backward pass.  One CTA owns a tile of 64 rows (transitions); the hidden
// activations of the tile live in shared memory; the weight matrices are
// streamed from L2 through a double-buffered cp.async stage (they are shared by
// all CTAs and stay L2 resident: <= 256 KB per matrix in a 126 MB L2).
//
//   forward : xin -> h1 = act(xin W1^T + b1) -> h2 = act(h1 W2^T + b2) -> out = h2 W3^T + b3
//   backward: dout -> dz2 = (dout W3) * act'(h2) -> dz1 = (dz2 W2) * act'(h1) [-> dx = dz1 W1[:, cols]]
//   wgrad   : dW = dz^T [inputs | 1]   (split over rows, partial sums per split)
#include <algorithm>

#include <cstdlib>

#include "common.cuh"

namespace tb {

constexpr int TM = 64;          // rows per CTA tile
constexpr int KC = 16;          // K rows of the streamed operand per stage
constexpr int NTHREADS = 256;

template <int H>
struct Cfg {
    static constexpr int NC = H / 32;                 // output columns per thread
    static constexpr int VEC = NC >= 4 ? 4 : NC;      // vector width of column groups
    static constexpr int LDH = H + 4;                 // smem row stride of activation tiles
    __device__ static __forceinline__ int col(int tx, int i) {
        return (i / VEC) * (32 * VEC) + tx * VEC + (i % VEC);
    }
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

template <int ACT>
__device__ __forceinline__ float activate(float x) {
    if (ACT == TB_ACT_TANH) return tanhf(x);
    return fmaxf(x, 0.0f);
}
// derivative of the activation expressed through its OUTPUT h
template <int ACT>
__device__ __forceinline__ float activate_grad(float h) {
    if (ACT == TB_ACT_TANH) return 1.0f - h * h;
    return h > 0.0f ? 1.0f : 0.0f;
}

// Stage rows [k0, k0+KC) of B (row-major [K, ldb], `bcols` valid columns) into
// Bs [KC][H]; rows >= K and columns >= bcols are zero-filled.
template <int H>
__device__ __forceinline__ void stage_b(float* Bs, const float* __restrict__ B, int ldb, int K,
                                        int bcols, int k0, bool aligned) {
    if (aligned) {                                    // bcols == H, 16-byte aligned rows
        constexpr int V4 = KC * H / 4;                // float4 slots
        for (int v = threadIdx.x; v < V4; v += NTHREADS) {
            const int r = v / (H / 4), c4 = v % (H / 4);
            float* dst = Bs + r * H + c4 * 4;
            if (k0 + r < K) {
                cp_async16(dst, B + (size_t)(k0 + r) * ldb + c4 * 4);
            } else {
                *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    } else {
        for (int v = threadIdx.x; v < KC * H; v += NTHREADS) {
            const int r = v / H, c = v % H;
            Bs[v] = (k0 + r < K && c < bcols) ? __ldg(B + (size_t)(k0 + r) * ldb + c) : 0.0f;
        }
    }
}

// acc[8][NC] += As[rows ty*8.., 0:K] * B[0:K, cols]   (As row-major smem, zero padded
// to a multiple of 4 columns; Bs2 = staging area of 2*KC*H floats).
// Ends with a __syncthreads(): As / Bs2 may be overwritten right after.
template <int H>
__device__ __forceinline__ void gemm_acc(float (&acc)[8][Cfg<H>::NC], const float* As, int lda,
                                         int K, const float* __restrict__ B, int ldb, int bcols,
                                         bool aligned, float* Bs2) {
    using C = Cfg<H>;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int nchunks = (K + KC - 1) / KC;
    stage_b<H>(Bs2, B, ldb, K, bcols, 0, aligned);
    cp_async_commit();
    for (int c = 0; c < nchunks; ++c) {
        float* cur = Bs2 + (c & 1) * (KC * H);
        if (c + 1 < nchunks) {
            stage_b<H>(Bs2 + ((c + 1) & 1) * (KC * H), B, ldb, K, bcols, (c + 1) * KC, aligned);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const int k0 = c * KC;
        const int klen = min(KC, K - k0);
        const float* arow = As + (size_t)(ty * 8) * lda + k0;
        for (int kk = 0; kk < klen; kk += 4) {
            float4 a[8];
#pragma unroll
            for (int r = 0; r < 8; ++r)
                a[r] = *reinterpret_cast<const float4*>(arow + (size_t)r * lda + kk);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float b[C::NC];
                const float* brow = cur + (kk + q) * H;
#pragma unroll
                for (int g = 0; g < C::NC / C::VEC; ++g) {
                    if constexpr (C::VEC == 4) {
                        const float4 v = *reinterpret_cast<const float4*>(brow + g * 128 + tx * 4);
                        b[g * 4 + 0] = v.x; b[g * 4 + 1] = v.y; b[g * 4 + 2] = v.z; b[g * 4 + 3] = v.w;
                    } else {
                        const float2 v = *reinterpret_cast<const float2*>(brow + tx * 2);
                        b[0] = v.x; b[1] = v.y;
                    }
                }
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const float av = q == 0 ? a[r].x : q == 1 ? a[r].y : q == 2 ? a[r].z : a[r].w;
#pragma unroll
                    for (int j = 0; j < C::NC; ++j) acc[r][j] = fmaf(av, b[j], acc[r][j]);
                }
            }
        }
        __syncthreads();
    }
}

template <int H>
__device__ __forceinline__ void zero_acc(float (&acc)[8][Cfg<H>::NC]) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int j = 0; j < Cfg<H>::NC; ++j) acc[r][j] = 0.0f;
}

// Store a thread's 8 x NC micro-tile (after `f`) to a row-major destination.
template <int H, typename F>
__device__ __forceinline__ void store_tile(const float (&acc)[8][Cfg<H>::NC], float* dst, int ld,
                                           int rows_valid, F f) {
    using C = Cfg<H>;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int row = ty * 8 + r;
        if (row >= rows_valid) continue;
#pragma unroll
        for (int g = 0; g < C::NC / C::VEC; ++g) {
            const int c0 = C::col(tx, g * C::VEC);
            if constexpr (C::VEC == 4) {
                float4 v;
                v.x = f(acc[r][g * 4 + 0], row, c0 + 0);
                v.y = f(acc[r][g * 4 + 1], row, c0 + 1);
                v.z = f(acc[r][g * 4 + 2], row, c0 + 2);
                v.w = f(acc[r][g * 4 + 3], row, c0 + 3);
                *reinterpret_cast<float4*>(dst + (size_t)row * ld + c0) = v;
            } else {
                float2 v;
                v.x = f(acc[r][0], row, c0 + 0);
                v.y = f(acc[r][1], row, c0 + 1);
                *reinterpret_cast<float2*>(dst + (size_t)row * ld + c0) = v;
            }
        }
    }
}


// acc = act(acc + bias[col]) in place; FAST selects the MUFU-based tanh (tensor-core path)
template <int H, int ACT, bool FAST = false>
__device__ __forceinline__ void bias_activate(float (&acc)[8][Cfg<H>::NC],
                                              const float* __restrict__ bias) {
    using C = Cfg<H>;
    const int tx = threadIdx.x & 31;
    float b[C::NC];
#pragma unroll
    for (int j = 0; j < C::NC; ++j) b[j] = __ldg(bias + C::col(tx, j));
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int j = 0; j < C::NC; ++j)
            acc[r][j] = (FAST && ACT == TB_ACT_TANH) ? tanh_fast(acc[r][j] + b[j])
                                                     : activate<ACT>(acc[r][j] + b[j]);
}

// acc *= act'(h) with h read from the saved activations (global, row-major [*, H]);
// rows >= valid are zeroed.
template <int H, int ACT>
__device__ __forceinline__ void mul_activation_grad(float (&acc)[8][Cfg<H>::NC],
                                                    const float* __restrict__ h, int valid) {
    using C = Cfg<H>;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int row = ty * 8 + r;
#pragma unroll
        for (int g = 0; g < C::NC / C::VEC; ++g) {
            const int c0 = C::col(tx, g * C::VEC);
            float hv[4] = {0.f, 0.f, 0.f, 0.f};
            if (row < valid) {
                if constexpr (C::VEC == 4) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(h + (size_t)row * H + c0));
                    hv[0] = v.x; hv[1] = v.y; hv[2] = v.z; hv[3] = v.w;
                } else {
                    const float2 v = __ldg(reinterpret_cast<const float2*>(h + (size_t)row * H + c0));
                    hv[0] = v.x; hv[1] = v.y;
                }
            }
#pragma unroll
            for (int e = 0; e < C::VEC; ++e)
                acc[r][g * C::VEC + e] =
                    row < valid ? acc[r][g * C::VEC + e] * activate_grad<ACT>(hv[e]) : 0.0f;
        }
    }
}

template <int H>
constexpr size_t mlp_smem_bytes() {
    return (size_t)(2 * TM * Cfg<H>::LDH + 2 * KC * H) * sizeof(float) + TM * sizeof(int64_t) +
           TM * 72 * sizeof(float);
}

// --------------------------------------------------------------------------------
// Forward
// --------------------------------------------------------------------------------
template <int H, int ACT>
__global__ void __launch_bounds__(NTHREADS, 1)
mlp_forward_kernel(TbMlpShape sh, const float* __restrict__ params,
                   const float* __restrict__ packed, TbMlpInput in, int64_t n_rows,
                   float* __restrict__ out, float* __restrict__ xin_save,
                   float* __restrict__ h1_save, float* __restrict__ h2_save,
                   const int32_t* d_skip) {
    using C = Cfg<H>;
    if (skip_requested(d_skip)) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* bufA = reinterpret_cast<float*>(smem_raw);          // [TM][LDH]
    float* bufB = bufA + TM * C::LDH;                          // [TM][LDH]
    float* Bs2 = bufB + TM * C::LDH;                           // [2][KC][H]
    int64_t* srow = reinterpret_cast<int64_t*>(Bs2 + 2 * KC * H);   // [TM] source rows

    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int64_t m0 = (int64_t)blockIdx.x * TM;
    const int valid = (int)min((int64_t)TM, n_rows - m0);

    if (tid < TM) srow[tid] = tid < valid ? (in.d_idx ? in.d_idx[m0 + tid] : m0 + tid) : -1;
    __syncthreads();

    const int d_in = sh.d_in;
    const int ldx = (d_in + 1 + 3) & ~3;                       // saved-input row stride
    float acc[8][C::NC];
    zero_acc<H>(acc);

    // ---- layer 1: input assembled in pieces of <= H columns into bufB ------------
    for (int k0 = 0; k0 < d_in; k0 += H) {
        const int klen = min(H, d_in - k0);
        const int kpad = (klen + 3) & ~3;
        for (int v = tid; v < TM * kpad; v += NTHREADS) {
            const int m = v / kpad, cc = v % kpad;
            const int c = k0 + cc;
            float val = 0.0f;
            const int64_t r = srow[m];
            if (r >= 0 && cc < klen) {
                if (c < in.dim1) {
                    val = in.d_x1[r * in.dim1 + c];
                    if (in.d_mean)      // mean_stds.py:36  (val - mean) / std
                        val = __fdiv_rn(__fsub_rn(val, in.d_mean[c]), in.d_std[c]);
                } else {
                    const int64_t r2 = in.gather2 ? r : (m0 + m);
                    val = in.d_x2[r2 * in.dim2 + (c - in.dim1)];
                }
                if (xin_save) xin_save[(m0 + m) * ldx + c] = val;
            }
            bufB[m * C::LDH + cc] = val;
        }
        __syncthreads();
        gemm_acc<H>(acc, bufB, C::LDH, klen, packed + sh.off_w1t + (size_t)k0 * H, H, H, true, Bs2);
    }
    if (xin_save) {     // trailing 1 column (bias gradients) and zero padding
        for (int v = tid; v < valid * (ldx - d_in); v += NTHREADS) {
            const int m = v / (ldx - d_in), c = d_in + v % (ldx - d_in);
            xin_save[(m0 + m) * ldx + c] = (c == d_in) ? 1.0f : 0.0f;
        }
    }
    auto ident = [](float a, int, int) { return a; };
    bias_activate<H, ACT>(acc, params + sh.off_b1);
    store_tile<H>(acc, bufA, C::LDH, TM, ident);
    if (h1_save) store_tile<H>(acc, h1_save + m0 * H, H, valid, ident);
    __syncthreads();

    // ---- layer 2 ----------------------------------------------------------------
    zero_acc<H>(acc);
    gemm_acc<H>(acc, bufA, C::LDH, H, packed + sh.off_w2t, H, H, true, Bs2);
    bias_activate<H, ACT>(acc, params + sh.off_b2);
    store_tile<H>(acc, bufB, C::LDH, TM, ident);
    if (h2_save) store_tile<H>(acc, h2_save + m0 * H, H, valid, ident);
    __syncthreads();

    // ---- head: out[m][o] = b3[o] + h2[m,:] . W3[o,:]  (4 lanes per dot product) --
    const int n_out = sh.n_out;
    const float* W3 = params + sh.off_w3;
    const float* b3 = params + sh.off_b3;
    const int quad = tid >> 2, ql = tid & 3;
    for (int p = quad; p < TM * n_out; p += NTHREADS / 4) {
        const int m = p % TM, o = p / TM;
        const float* hrow = bufB + m * C::LDH;
        const float* wrow = W3 + (size_t)o * H;
        float s = 0.0f;
#pragma unroll 4
        for (int i = ql; i < H / 4; i += 4) {
            const float4 hv = *reinterpret_cast<const float4*>(hrow + i * 4);
            const float4 wv = __ldg(reinterpret_cast<const float4*>(wrow + i * 4));
            s = fmaf(hv.x, wv.x, s); s = fmaf(hv.y, wv.y, s);
            s = fmaf(hv.z, wv.z, s); s = fmaf(hv.w, wv.w, s);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (ql == 0 && m < valid) out[(m0 + m) * n_out + o] = s + b3[o];
    }
    (void)tx; (void)ty;
}

// --------------------------------------------------------------------------------
// Backward (activation gradients)
// --------------------------------------------------------------------------------
template <int H, int ACT>
__global__ void __launch_bounds__(NTHREADS, 1)
mlp_backward_kernel(TbMlpShape sh, const float* __restrict__ params,
                    const float* __restrict__ dout, int ld_dout,
                    const float* __restrict__ h1, const float* __restrict__ h2, int64_t n_rows,
                    float* __restrict__ dz2, float* __restrict__ dz1, float* __restrict__ dx,
                    int dx_col0, int dx_cols, const int32_t* d_skip) {
    using C = Cfg<H>;
    if (skip_requested(d_skip)) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* bufA = reinterpret_cast<float*>(smem_raw);
    float* bufB = bufA + TM * C::LDH;
    float* Bs2 = bufB + TM * C::LDH;
    float* sD = Bs2 + 2 * KC * H + TM * 2;                     // [TM][72] (after the srow area)

    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * TM;
    const int valid = (int)min((int64_t)TM, n_rows - m0);
    const int n_out = sh.n_out;
    const int ldo = (n_out + 3) & ~3;                          // <= 72

    for (int v = tid; v < TM * ldo; v += NTHREADS) {
        const int m = v / ldo, o = v % ldo;
        sD[m * 72 + o] = (m < valid && o < n_out) ? dout[(m0 + m) * ld_dout + o] : 0.0f;
    }
    __syncthreads();

    float acc[8][C::NC];
    zero_acc<H>(acc);
    gemm_acc<H>(acc, sD, 72, n_out, params + sh.off_w3, H, H, true, Bs2);     // dh2 = dout W3
    auto ident = [](float a, int, int) { return a; };
    mul_activation_grad<H, ACT>(acc, h2 + m0 * H, valid);
    store_tile<H>(acc, bufB, C::LDH, TM, ident);
    store_tile<H>(acc, dz2 + m0 * H, H, valid, ident);
    __syncthreads();

    zero_acc<H>(acc);
    gemm_acc<H>(acc, bufB, C::LDH, H, params + sh.off_w2, H, H, true, Bs2);   // dh1 = dz2 W2
    mul_activation_grad<H, ACT>(acc, h1 + m0 * H, valid);
    if (dx) store_tile<H>(acc, bufA, C::LDH, TM, ident);
    store_tile<H>(acc, dz1 + m0 * H, H, valid, ident);
    if (!dx) return;
    __syncthreads();

    // dx[:, j] = sum_n dz1[:, n] W1[n, dx_col0 + j]   (W1 is [H, d_in] row-major)
    zero_acc<H>(acc);
    gemm_acc<H>(acc, bufA, C::LDH, H, params + sh.off_w1 + dx_col0, sh.d_in, dx_cols, false, Bs2);
    {
        const int tx = tid & 31, ty = tid >> 5;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int row = ty * 8 + r;
            if (row >= valid) continue;
#pragma unroll
            for (int j = 0; j < C::NC; ++j) {
                const int c = C::col(tx, j);
                if (c < dx_cols) dx[(m0 + row) * dx_cols + c] = acc[r][j];
            }
        }
    }
}

// --------------------------------------------------------------------------------
// Weight gradients:  out[n][k] = sum_m A[m][a_col0+n] * B[m][b_col0+k]
// --------------------------------------------------------------------------------
struct WJob {
    const float* A; const float* B;
    const float* A_lo;          // optional second part added to A on load (tf32 splits)
    int lda, a_col0, a_cols;
    int ldb, b_col0, b_cols;
    int out_off, out_ld;        // gpart offset of out[0][0] and its row stride
    int bias_col, bias_off;     // column k == bias_col goes to bias_off + n (or -1)
    int variant;                // 0: 128x128, 1: 128x32, 2: 16x128
};
constexpr int kMaxJobs = 24;
struct WJobTable { WJob jobs[kMaxJobs]; int n_jobs; };
constexpr int WMC = 16;        // rows of m per smem chunk
constexpr int WST = 4;         // cp.async pipeline depth of the weight-gradient kernel

template <int TN, int TK>
__device__ __forceinline__ void wgrad_tile(const WJob& job, int64_t r0, int64_t r1,
                                           float* __restrict__ gout, float* smem) {
    constexpr int MN = TN / 16, MK = TK / 16;
    float* As = smem;                        // [WST][WMC][TN]
    float* Bsm = As + WST * WMC * TN;        // [WST][WMC][TK]
    float* Alo = Bsm + WST * WMC * TK;       // [WST][WMC][TN]  (only when job.A_lo)
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float acc[MN][MK];
#pragma unroll
    for (int i = 0; i < MN; ++i)
#pragma unroll
        for (int j = 0; j < MK; ++j) acc[i][j] = 0.0f;

    const bool has_lo = job.A_lo != nullptr;
    const bool a_fast = TN == 128 && job.a_cols == 128 && (job.lda & 3) == 0 && (job.a_col0 & 3) == 0;
    const bool b_fast = TK == 128 && job.b_cols == 128 && (job.ldb & 3) == 0 && (job.b_col0 & 3) == 0;

    // every element goes through cp.async (16-byte copies for aligned full-width operands,
    // 4-byte copies for the narrow / unaligned ones), so the next chunk is always in flight
    // while the current one is being multiplied
    auto stage_operand = [&](float* dst, const float* __restrict__ src, int ld, int col0, int cols,
                             bool fast, int64_t mbase, int width) {
        if (fast) {
            for (int v = tid; v < WMC * width / 4; v += NTHREADS) {
                const int r = v / (width / 4), c4 = v % (width / 4);
                float* d = dst + r * width + c4 * 4;
                if (mbase + r < r1) cp_async16(d, src + (mbase + r) * ld + col0 + c4 * 4);
                else *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {
            for (int v = tid; v < WMC * width; v += NTHREADS) {
                const int r = v / width, c = v % width;
                if (mbase + r < r1 && c < cols) cp_async4(dst + v, src + (mbase + r) * ld + col0 + c);
                else dst[v] = 0.0f;
            }
        }
    };
    auto stage = [&](int buf, int64_t mbase) {
        stage_operand(As + buf * WMC * TN, job.A, job.lda, job.a_col0, job.a_cols, a_fast, mbase, TN);
        if (has_lo)
            stage_operand(Alo + buf * WMC * TN, job.A_lo, job.lda, job.a_col0, job.a_cols, a_fast, mbase, TN);
        stage_operand(Bsm + buf * WMC * TK, job.B, job.ldb, job.b_col0, job.b_cols, b_fast, mbase, TK);
    };

    const int nchunks = (int)((r1 - r0 + WMC - 1) / WMC);
    // WST-deep ring: chunks c+1 .. c+WST-1 are in flight while chunk c is multiplied
    for (int c = 0; c < WST - 1; ++c) {
        if (c < nchunks) stage(c, r0 + (int64_t)c * WMC);
        cp_async_commit();
    }
    for (int c = 0; c < nchunks; ++c) {
        cp_async_wait<WST - 2>();
        __syncthreads();             // chunk c landed; everyone finished computing chunk c-1
        if (c + WST - 1 < nchunks) stage((c + WST - 1) % WST, r0 + (int64_t)(c + WST - 1) * WMC);
        cp_async_commit();
        const float* as = As + (c % WST) * WMC * TN;
        const float* al = Alo + (c % WST) * WMC * TN;
        const float* bs = Bsm + (c % WST) * WMC * TK;
#pragma unroll 4
        for (int r = 0; r < WMC; ++r) {
            float a[MN], b[MK];
            if constexpr (MN == 8) {
                float4 v0 = *reinterpret_cast<const float4*>(as + r * TN + ty * 4);
                float4 v1 = *reinterpret_cast<const float4*>(as + r * TN + 64 + ty * 4);
                if (has_lo) {
                    const float4 l0 = *reinterpret_cast<const float4*>(al + r * TN + ty * 4);
                    const float4 l1 = *reinterpret_cast<const float4*>(al + r * TN + 64 + ty * 4);
                    v0.x += l0.x; v0.y += l0.y; v0.z += l0.z; v0.w += l0.w;
                    v1.x += l1.x; v1.y += l1.y; v1.z += l1.z; v1.w += l1.w;
                }
                a[0] = v0.x; a[1] = v0.y; a[2] = v0.z; a[3] = v0.w;
                a[4] = v1.x; a[5] = v1.y; a[6] = v1.z; a[7] = v1.w;
            } else {
                a[0] = as[r * TN + ty] + (has_lo ? al[r * TN + ty] : 0.0f);
            }
            if constexpr (MK == 8) {
                const float4 v0 = *reinterpret_cast<const float4*>(bs + r * TK + tx * 4);
                const float4 v1 = *reinterpret_cast<const float4*>(bs + r * TK + 64 + tx * 4);
                b[0] = v0.x; b[1] = v0.y; b[2] = v0.z; b[3] = v0.w;
                b[4] = v1.x; b[5] = v1.y; b[6] = v1.z; b[7] = v1.w;
            } else {
                const float2 v = *reinterpret_cast<const float2*>(bs + r * TK + tx * 2);
                b[0] = v.x; b[1] = v.y;
            }
#pragma unroll
            for (int i = 0; i < MN; ++i)
#pragma unroll
                for (int j = 0; j < MK; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < MN; ++i) {
        const int n = MN == 8 ? (i / 4) * 64 + ty * 4 + (i % 4) : ty;
        if (n >= job.a_cols) continue;
#pragma unroll
        for (int j = 0; j < MK; ++j) {
            const int k = MK == 8 ? (j / 4) * 64 + tx * 4 + (j % 4) : tx * 2 + j;
            if (k >= job.b_cols) continue;
            if (job.b_col0 + k == job.bias_col) gout[job.bias_off + n] = acc[i][j];
            else gout[job.out_off + (size_t)n * job.out_ld + k] = acc[i][j];
        }
    }
}

__global__ void __launch_bounds__(NTHREADS)
mlp_wgrad_kernel(WJobTable table, int64_t n_rows, int64_t rows_per_split,
                 float* __restrict__ gpart, int n_params, const int32_t* d_skip) {
    if (skip_requested(d_skip)) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* smem = reinterpret_cast<float*>(smem_raw);
    // grid = (splits, jobs): blocks are dispatched job-major, so the heavy 128x128
    // tiles (jobs 0..3) go out first, one per SM when 4 * splits == #SMs
    const WJob& job = table.jobs[blockIdx.y];
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_split;
    const int64_t r1 = min(n_rows, r0 + rows_per_split);
    float* gout = gpart + (size_t)blockIdx.x * n_params;
    if (job.variant == 0) wgrad_tile<128, 128>(job, r0, max(r0, r1), gout, smem);
    else if (job.variant == 1) wgrad_tile<128, 32>(job, r0, max(r0, r1), gout, smem);
    else wgrad_tile<16, 128>(job, r0, max(r0, r1), gout, smem);
}

// --------------------------------------------------------------------------------
// host dispatch
// --------------------------------------------------------------------------------
static int check_shape(const TbMlpShape* sh, const char* who) {
    TB_REQUIRE(sh, TB_EINVAL, "%s: null shape", who);
    TB_REQUIRE(sh->hidden == 64 || sh->hidden == 128 || sh->hidden == 256, TB_ENOTSUP,
               "%s: hidden width %d not supported (64, 128, 256)", who, sh->hidden);
    TB_REQUIRE(sh->act == TB_ACT_TANH || sh->act == TB_ACT_RELU, TB_ENOTSUP,
               "%s: activation %d not supported", who, sh->act);
    TB_REQUIRE(sh->n_out >= 1 && sh->n_out <= 64, TB_ENOTSUP, "%s: n_out %d not in [1,64]", who,
               sh->n_out);
    TB_REQUIRE(sh->d_in >= 1 && sh->d_in <= 4096, TB_ENOTSUP, "%s: d_in %d", who, sh->d_in);
    TB_REQUIRE((sh->off_w1 | sh->off_w2 | sh->off_w3 | sh->off_w1t | sh->off_w2t) % 4 == 0,
               TB_EINVAL, "%s: weight offsets must be multiples of 4 floats", who);
    return 0;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
    return (int)cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

#define TB_DISPATCH_H_ACT(H_, ACT_, CALL)                                   \
    do {                                                                    \
        if (H_ == 64 && ACT_ == TB_ACT_TANH) { CALL(64, TB_ACT_TANH); }     \
        else if (H_ == 64) { CALL(64, TB_ACT_RELU); }                       \
        else if (H_ == 128 && ACT_ == TB_ACT_TANH) { CALL(128, TB_ACT_TANH); } \
        else if (H_ == 128) { CALL(128, TB_ACT_RELU); }                     \
        else if (ACT_ == TB_ACT_TANH) { CALL(256, TB_ACT_TANH); }           \
        else { CALL(256, TB_ACT_RELU); }                                    \
    } while (0)

}  // namespace tb

extern "C" int tb_mlp_forward(const TbMlpShape* shape, const float* d_params,
                              const float* d_packed, const TbMlpInput* in, int64_t n_rows,
                              float* d_out, float* d_xin, float* d_h1, float* d_h2,
                              const int32_t* d_skip, void* stream) {
    tb::ProfScope prof_scope("tb_mlp_forward", stream);
    using namespace tb;
    int rc = check_shape(shape, "tb_mlp_forward");
    if (rc) return rc;
    TB_REQUIRE(d_params && d_packed && in && in->d_x1 && d_out && n_rows > 0, TB_EINVAL,
               "tb_mlp_forward: null pointer");
    TB_REQUIRE(in->dim1 + (in->d_x2 ? in->dim2 : 0) == shape->d_in, TB_EINVAL,
               "tb_mlp_forward: input widths %d+%d != d_in %d", in->dim1,
               in->d_x2 ? in->dim2 : 0, shape->d_in);
    TB_REQUIRE((in->d_mean == nullptr) == (in->d_std == nullptr), TB_EINVAL,
               "tb_mlp_forward: mean/std must be given together");
    const int blocks = (int)((n_rows + TM - 1) / TM);
    cudaStream_t s = as_stream(stream);
#define CALL(H_, A_)                                                                        \
    {                                                                                       \
        const size_t smem = mlp_smem_bytes<H_>();                                           \
        set_smem(mlp_forward_kernel<H_, A_>, smem);                                         \
        mlp_forward_kernel<H_, A_><<<blocks, NTHREADS, smem, s>>>(                          \
            *shape, d_params, d_packed, *in, n_rows, d_out, d_xin, d_h1, d_h2, d_skip);     \
    }
    TB_DISPATCH_H_ACT(shape->hidden, shape->act, CALL);
#undef CALL
    return check_launch("tb_mlp_forward");
}

extern "C" int tb_mlp_backward(const TbMlpShape* shape, const float* d_params,
                               const float* d_dout, int32_t ld_dout, const float* d_h1,
                               const float* d_h2, int64_t n_rows, float* d_dz2, float* d_dz1,
                               float* d_dx, int32_t dx_col0, int32_t dx_cols,
                               const int32_t* d_skip, void* stream) {
    tb::ProfScope prof_scope("tb_mlp_backward", stream);
    using namespace tb;
    int rc = check_shape(shape, "tb_mlp_backward");
    if (rc) return rc;
    TB_REQUIRE(d_params && d_dout && d_h1 && d_h2 && d_dz2 && d_dz1 && n_rows > 0 &&
               ld_dout >= shape->n_out, TB_EINVAL, "tb_mlp_backward: bad arguments");
    TB_REQUIRE(!d_dx || (dx_cols >= 1 && dx_cols <= shape->hidden && dx_col0 >= 0 &&
                         dx_col0 + dx_cols <= shape->d_in), TB_EINVAL,
               "tb_mlp_backward: dx column range [%d,+%d) invalid", dx_col0, dx_cols);
    const int blocks = (int)((n_rows + TM - 1) / TM);
    cudaStream_t s = as_stream(stream);
#define CALL(H_, A_)                                                                        \
    {                                                                                       \
        const size_t smem = mlp_smem_bytes<H_>();                                           \
        set_smem(mlp_backward_kernel<H_, A_>, smem);                                        \
        mlp_backward_kernel<H_, A_><<<blocks, NTHREADS, smem, s>>>(                         \
            *shape, d_params, d_dout, ld_dout, d_h1, d_h2, n_rows, d_dz2, d_dz1, d_dx,      \
            dx_col0, dx_cols, d_skip);                                                      \
    }
    TB_DISPATCH_H_ACT(shape->hidden, shape->act, CALL);
#undef CALL
    return check_launch("tb_mlp_backward");
}

static int launch_wgrad_jobs(const TbMlpShape* shape, const float* d_xin, const float* d_h1,
                             const float* d_h2, const float* d_dz1, const float* d_dz2,
                             const float* d_dz2_lo, const float* d_dout, int32_t ld_dout,
                             int32_t n_extra, int32_t off_extra, int64_t n_rows, float* d_gpart,
                             int32_t n_split, bool include_w2, const int32_t* d_skip, void* stream) {
    // include_w2 == false: W2 and b2 come from the tensor-core kernel (tb_tc_wgrad256)
    using namespace tb;
    const int H = shape->hidden, d_in = shape->d_in, n_out = shape->n_out;
    const int ldx = (d_in + 1 + 3) & ~3;
    WJobTable t;
    t.n_jobs = 0;
    auto add = [&](const float* A, const float* A_lo, int lda, int a0, int an, const float* B, int ldb,
                   int b0, int bn, int out_off, int out_ld, int bias_col, int bias_off, int variant) {
        if (t.n_jobs >= kMaxJobs) return false;
        WJob& j = t.jobs[t.n_jobs++];
        j.A = A; j.A_lo = A_lo; j.lda = lda; j.a_col0 = a0; j.a_cols = an;
        j.B = B; j.ldb = ldb; j.b_col0 = b0; j.b_cols = bn;
        j.out_off = out_off; j.out_ld = out_ld; j.bias_col = bias_col; j.bias_off = bias_off;
        j.variant = variant;
        return true;
    };
    bool ok = true;
    const int ntile = (H + 127) / 128;
    // dW2[n][k] = sum_m dz2[m][n] h1[m][k]  (heavy 128x128 tiles first)
    if (include_w2)
        for (int tn = 0; tn < ntile; ++tn)
            for (int tk = 0; tk < ntile; ++tk) {
                const int an = std::min(128, H - tn * 128), bn = std::min(128, H - tk * 128);
                ok &= add(d_dz2, nullptr, H, tn * 128, an, d_h1, H, tk * 128, bn,
                          shape->off_w2 + tn * 128 * H + tk * 128, H, -1, 0, 0);
            }
    // dW1[n][k] (+ db1 through the trailing ones column of xin)
    for (int tn = 0; tn < ntile; ++tn) {
        const int an = std::min(128, H - tn * 128);
        const int tkw = d_in + 1 <= 32 ? 32 : 128;
        for (int k0 = 0; k0 < d_in + 1; k0 += tkw) {
            const int bn = std::min(tkw, d_in + 1 - k0);
            ok &= add(d_dz1, nullptr, H, tn * 128, an, d_xin, ldx, k0, bn,
                      shape->off_w1 + tn * 128 * d_in + k0, d_in, d_in,
                      shape->off_b1 + tn * 128, tkw == 32 ? 1 : 0);
        }
        // db2[n] = sum_m dz2[m][n]  (ones column of xin)
        if (include_w2)
            ok &= add(d_dz2, d_dz2_lo, H, tn * 128, an, d_xin, ldx, d_in, 1, 0, 1, d_in,
                      shape->off_b2 + tn * 128, 1);
    }
    // dW3[o][k] = sum_m dout[m][o] h2[m][k]
    for (int o0 = 0; o0 < n_out; o0 += 16)
        for (int tk = 0; tk < ntile; ++tk) {
            const int bn = std::min(128, H - tk * 128);
            ok &= add(d_dout, nullptr, ld_dout, o0, std::min(16, n_out - o0), d_h2, H, tk * 128, bn,
                      shape->off_w3 + o0 * H + tk * 128, H, -1, 0, 2);
        }
    // db3 and the extra per-row columns (e.g. log_scale gradients)
    ok &= add(d_dout, nullptr, ld_dout, 0, n_out, d_xin, ldx, d_in, 1, 0, 1, d_in, shape->off_b3, 1);
    if (n_extra > 0)
        ok &= add(d_dout, nullptr, ld_dout, n_out, n_extra, d_xin, ldx, d_in, 1, 0, 1, d_in, off_extra, 1);
    TB_REQUIRE(ok, TB_ENOTSUP, "tb_mlp_wgrad: too many tiles (d_in=%d)", d_in);

    int64_t rows_per_split = (n_rows + n_split - 1) / n_split;
    rows_per_split = (rows_per_split + WMC - 1) / WMC * WMC;
    const size_t smem = (size_t)WST * WMC * (128 + 128 + 128) * sizeof(float);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(mlp_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    dim3 grid(n_split, t.n_jobs);
    mlp_wgrad_kernel<<<grid, NTHREADS, smem, as_stream(stream)>>>(
        t, n_rows, rows_per_split, d_gpart, shape->n_params, d_skip);
    return check_launch("tb_mlp_wgrad");
}

extern "C" int tb_mlp_wgrad(const TbMlpShape* shape, const float* d_xin, const float* d_h1,
                            const float* d_h2, const float* d_dz1, const float* d_dz2,
                            const float* d_dout, int32_t ld_dout, int32_t n_extra,
                            int32_t off_extra, int64_t n_rows, float* d_gpart,
                            int32_t n_split, const int32_t* d_skip, void* stream) {
    tb::ProfScope prof_scope("tb_mlp_wgrad", stream);
    using namespace tb;
    int rc = check_shape(shape, "tb_mlp_wgrad");
    if (rc) return rc;
    TB_REQUIRE(d_xin && d_h1 && d_h2 && d_dz1 && d_dz2 && d_dout && d_gpart && n_rows > 0 &&
               n_split >= 1 && ld_dout >= shape->n_out + n_extra, TB_EINVAL,
               "tb_mlp_wgrad: bad arguments");
    return launch_wgrad_jobs(shape, d_xin, d_h1, d_h2, d_dz1, d_dz2, nullptr, d_dout, ld_dout, n_extra,
                             off_extra, n_rows, d_gpart, n_split, true, d_skip, stream);
}

// =====================================================================================
// Tensor-core variant: the FFMA pieces around csrc/tc_gemm.cu
// =====================================================================================
namespace tb {

__device__ __forceinline__ float tf32_hi(float x) {
    return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
}

// h1 = act(xin W1^T + b1), written as a tf32 split (A operand of the layer-2 GEMM)
template <int H, int ACT>
__global__ void __launch_bounds__(NTHREADS, 2)
mlp_layer1_kernel(TbMlpShape sh, const float* __restrict__ params, const float* __restrict__ packed,
                  TbMlpInput in, int64_t n_rows, float* __restrict__ xin_save,
                  float* __restrict__ h1_hi, float* __restrict__ h1_lo, const int32_t* d_skip) {
    using C = Cfg<H>;
    if (skip_requested(d_skip)) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* bufB = reinterpret_cast<float*>(smem_raw);          // [TM][LDH] input staging
    float* Bs2 = bufB + TM * C::LDH;                           // [2][KC][H]
    int64_t* srow = reinterpret_cast<int64_t*>(Bs2 + 2 * KC * H);
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * TM;
    const int valid = (int)min((int64_t)TM, n_rows - m0);
    if (tid < TM) srow[tid] = tid < valid ? (in.d_idx ? in.d_idx[m0 + tid] : m0 + tid) : -1;
    __syncthreads();
    const int d_in = sh.d_in;
    const int ldx = (d_in + 1 + 3) & ~3;
    float acc[8][C::NC];
    zero_acc<H>(acc);
    for (int k0 = 0; k0 < d_in; k0 += H) {
        const int klen = min(H, d_in - k0);
        const int kpad = (klen + 3) & ~3;
        for (int v = tid; v < TM * kpad; v += NTHREADS) {
            const int m = v / kpad, cc = v % kpad;
            const int c = k0 + cc;
            float val = 0.0f;
            const int64_t r = srow[m];
            if (r >= 0 && cc < klen) {
                if (c < in.dim1) {
                    val = in.d_x1[r * in.dim1 + c];
                    if (in.d_mean) val = __fdiv_rn(__fsub_rn(val, in.d_mean[c]), in.d_std[c]);
                } else {
                    const int64_t r2 = in.gather2 ? r : (m0 + m);
                    val = in.d_x2[r2 * in.dim2 + (c - in.dim1)];
                }
                if (xin_save) xin_save[(m0 + m) * ldx + c] = val;
            }
            bufB[m * C::LDH + cc] = val;
        }
        __syncthreads();
        gemm_acc<H>(acc, bufB, C::LDH, klen, packed + sh.off_w1t + (size_t)k0 * H, H, H, true, Bs2);
    }
    if (xin_save) {
        for (int v = tid; v < valid * (ldx - d_in); v += NTHREADS) {
            const int m = v / (ldx - d_in), c = d_in + v % (ldx - d_in);
            xin_save[(m0 + m) * ldx + c] = (c == d_in) ? 1.0f : 0.0f;
        }
    }
    bias_activate<H, ACT, true>(acc, params + sh.off_b1);
    store_tile<H>(acc, h1_hi + m0 * H, H, valid, [](float a, int, int) { return tf32_hi(a); });
    store_tile<H>(acc, h1_lo + m0 * H, H, valid, [](float a, int, int) { return a - tf32_hi(a); });
}

// ---- small batches (off-policy updates: 100 rows; SURVEY.md 8, replays/buffers.py:8-12) --------
// The 64-row tile kernels above put a 100-row minibatch on 2 of the 148 SMs.  These variants cut
// the same work into many small CTAs (16 rows x 32 output columns) with the weights of the column
// block staged once in shared memory.  Every output is still one sequential fmaf chain over k in
// increasing order, i.e. bit-identical to the tile kernels: results do not depend on which variant
// a batch size selects.
constexpr int SK_ROWS = 16;          // rows per CTA (two per thread)
constexpr int SK_COLS = 32;          // output columns per CTA (one per lane)
constexpr int SK_MAX_ROWS = 4096;    // above this the tile kernels stream fewer weight bytes through L2

__host__ __device__ inline size_t skinny_layer1_smem(int d_in) {
    return (size_t)(SK_ROWS * ((d_in + 3) & ~3) + d_in * SK_COLS) * sizeof(float) + SK_ROWS * sizeof(int64_t);
}

template <int ACT>
__global__ void __launch_bounds__(NTHREADS)
mlp_layer1_skinny_kernel(TbMlpShape sh, const float* __restrict__ params, const float* __restrict__ packed,
                         TbMlpInput in, int64_t n_rows, float* __restrict__ xin_save,
                         float* __restrict__ h1_hi, float* __restrict__ h1_lo, const int32_t* d_skip) {
    constexpr int H = 256;
    if (skip_requested(d_skip)) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int d_in = sh.d_in, kp = (d_in + 3) & ~3, ldx = (d_in + 1 + 3) & ~3;
    float* xs = reinterpret_cast<float*>(smem_raw);            // [SK_ROWS][kp]
    float* ws = xs + SK_ROWS * kp;                             // [d_in][SK_COLS]
    int64_t* srow = reinterpret_cast<int64_t*>(ws + d_in * SK_COLS);
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * SK_ROWS;
    const int c0 = blockIdx.y * SK_COLS;
    const int valid = (int)min((int64_t)SK_ROWS, n_rows - m0);
    if (tid < SK_ROWS) srow[tid] = tid < valid ? (in.d_idx ? in.d_idx[m0 + tid] : m0 + tid) : -1;
    const float* W1T = packed + sh.off_w1t;                    // [d_in][H]
    for (int v = tid; v < d_in * (SK_COLS / 4); v += NTHREADS) {
        const int k = v / (SK_COLS / 4), q = v % (SK_COLS / 4);
        *reinterpret_cast<float4*>(ws + k * SK_COLS + q * 4) =
            __ldg(reinterpret_cast<const float4*>(W1T + (size_t)k * H + c0) + q);
    }
    __syncthreads();
    const bool save = xin_save && blockIdx.y == 0;
    for (int v = tid; v < SK_ROWS * kp; v += NTHREADS) {
        const int m = v / kp, c = v % kp;
        float val = 0.0f;
        const int64_t r = srow[m];
        if (r >= 0 && c < d_in) {
            if (c < in.dim1) {
                val = in.d_x1[r * in.dim1 + c];
                if (in.d_mean) val = __fdiv_rn(__fsub_rn(val, in.d_mean[c]), in.d_std[c]);
            } else {
                const int64_t r2 = in.gather2 ? r : (m0 + m);
                val = in.d_x2[r2 * in.dim2 + (c - in.dim1)];
            }
            if (save) xin_save[(m0 + m) * ldx + c] = val;
        }
        xs[m * kp + c] = val;
    }
    if (save) {
        for (int v = tid; v < valid * (ldx - d_in); v += NTHREADS) {
            const int m = v / (ldx - d_in), c = d_in + v % (ldx - d_in);
            xin_save[(m0 + m) * ldx + c] = (c == d_in) ? 1.0f : 0.0f;
        }
    }
    __syncthreads();
    const int n = tid & 31, m = tid >> 5;                      // rows m and m + 8
    const float* x0 = xs + m * kp;
    const float* x1 = xs + (m + 8) * kp;
    float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll 4
    for (int k = 0; k < d_in; ++k) {
        const float w = ws[k * SK_COLS + n];
        acc0 = fmaf(x0[k], w, acc0);
        acc1 = fmaf(x1[k], w, acc1);
    }
    const float bias = __ldg(params + sh.off_b1 + c0 + n);
    acc0 = ACT == TB_ACT_TANH ? tanh_fast(acc0 + bias) : activate<ACT>(acc0 + bias);
    acc1 = ACT == TB_ACT_TANH ? tanh_fast(acc1 + bias) : activate<ACT>(acc1 + bias);
    if (m < valid) {
        const size_t e = (size_t)(m0 + m) * H + c0 + n;
        h1_hi[e] = tf32_hi(acc0);
        h1_lo[e] = acc0 - tf32_hi(acc0);
    }
    if (m + 8 < valid) {
        const size_t e = (size_t)(m0 + m + 8) * H + c0 + n;
        h1_hi[e] = tf32_hi(acc1);
        h1_lo[e] = acc1 - tf32_hi(acc1);
    }
}

// dx[:, j] = sum_n dz1[:, n] W1[n, col0 + j] for a few columns (the action block of a Q network)
// and a small batch: 16 rows per CTA, thread (row, column lane), W1 slice staged in shared memory.
constexpr int SKD_MAX_COLS = 64;
__host__ __device__ inline size_t skinny_dx_smem(int dx_cols) {
    return (size_t)(SK_ROWS * (256 + 4) + 256 * dx_cols) * sizeof(float);
}
__global__ void __launch_bounds__(NTHREADS)
mlp_dx_skinny_kernel(TbMlpShape sh, const float* __restrict__ params, const float* __restrict__ dz1,
                     int64_t n_rows, float* __restrict__ dx, int dx_col0, int dx_cols, const int32_t* d_skip) {
    constexpr int H = 256, LD = H + 4;
    if (skip_requested(d_skip)) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* as = reinterpret_cast<float*>(smem_raw);            // [SK_ROWS][LD]
    float* ws = as + SK_ROWS * LD;                             // [H][dx_cols]
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * SK_ROWS;
    const int valid = (int)min((int64_t)SK_ROWS, n_rows - m0);
    for (int v = tid; v < SK_ROWS * (H / 4); v += NTHREADS) {
        const int m = v / (H / 4), c4 = v % (H / 4);
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < valid) x = __ldg(reinterpret_cast<const float4*>(dz1 + (m0 + m) * H) + c4);
        *reinterpret_cast<float4*>(as + m * LD + c4 * 4) = x;
    }
    const float* W1 = params + sh.off_w1 + dx_col0;            // [H][d_in], columns col0..
    for (int v = tid; v < H * dx_cols; v += NTHREADS) {
        const int nn = v / dx_cols, j = v % dx_cols;
        ws[v] = __ldg(W1 + (size_t)nn * sh.d_in + j);
    }
    __syncthreads();
    const int m = tid >> 4, jl = tid & 15;
    if (m >= valid) return;
    const float* arow = as + m * LD;
    for (int j = jl; j < dx_cols; j += 16) {
        float acc = 0.0f;
#pragma unroll 8
        for (int nn = 0; nn < H; ++nn) acc = fmaf(arow[nn], ws[nn * dx_cols + j], acc);
        dx[(m0 + m) * dx_cols + j] = acc;
    }
}

// out[m][o] = b3[o] + h2[m, :] . W3[o, :]   (ROWS rows per CTA: 64, or 8 for small batches)
template <int H, int ROWS = TM>
__global__ void __launch_bounds__(NTHREADS, 1)
mlp_head_kernel(TbMlpShape sh, const float* __restrict__ params, const float* __restrict__ h2,
                int64_t n_rows, float* __restrict__ out, const int32_t* d_skip) {
    using C = Cfg<H>;
    constexpr int TM = ROWS;
    if (skip_requested(d_skip)) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* bufB = reinterpret_cast<float*>(smem_raw);          // [TM][LDH]
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * TM;
    const int valid = (int)min((int64_t)TM, n_rows - m0);
    for (int v = tid; v < TM * (H / 4); v += NTHREADS) {
        const int m = v / (H / 4), c4 = v % (H / 4);
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < valid) x = __ldg(reinterpret_cast<const float4*>(h2 + (m0 + m) * H) + c4);
        *reinterpret_cast<float4*>(bufB + m * C::LDH + c4 * 4) = x;
    }
    __syncthreads();
    const int n_out = sh.n_out;
    const float* W3 = params + sh.off_w3;
    const float* b3 = params + sh.off_b3;
    const int quad = tid >> 2, ql = tid & 3;
    for (int p = quad; p < TM * n_out; p += NTHREADS / 4) {
        const int m = p % TM, o = p / TM;
        const float* hrow = bufB + m * C::LDH;
        const float* wrow = W3 + (size_t)o * H;
        float s = 0.0f;
#pragma unroll 4
        for (int i = ql; i < H / 4; i += 4) {
            const float4 hv = *reinterpret_cast<const float4*>(hrow + i * 4);
            const float4 wv = __ldg(reinterpret_cast<const float4*>(wrow + i * 4));
            s = fmaf(hv.x, wv.x, s); s = fmaf(hv.y, wv.y, s);
            s = fmaf(hv.z, wv.z, s); s = fmaf(hv.w, wv.w, s);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (ql == 0 && m < valid) out[(m0 + m) * n_out + o] = s + b3[o];
    }
}

// dz2 = (dout W3) * act'(h2), written as a tf32 split (A operand of the backward GEMM and
// of the weight-gradient GEMM)
template <int H, int ACT>
__global__ void __launch_bounds__(NTHREADS, 2)
mlp_head_backward_kernel(TbMlpShape sh, const float* __restrict__ params,
                         const float* __restrict__ dout, int ld_dout, const float* __restrict__ h2,
                         int64_t n_rows, float* __restrict__ dz2_hi, float* __restrict__ dz2_lo,
                         const int32_t* d_skip) {
    using C = Cfg<H>;
    if (skip_requested(d_skip)) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* Bs2 = reinterpret_cast<float*>(smem_raw);           // [2][KC][H]
    float* sD = Bs2 + 2 * KC * H;                              // [TM][72]
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * TM;
    const int valid = (int)min((int64_t)TM, n_rows - m0);
    const int n_out = sh.n_out;
    const int ldo = (n_out + 3) & ~3;
    for (int v = tid; v < TM * ldo; v += NTHREADS) {
        const int m = v / ldo, o = v % ldo;
        sD[m * 72 + o] = (m < valid && o < n_out) ? dout[(m0 + m) * ld_dout + o] : 0.0f;
    }
    __syncthreads();
    float acc[8][C::NC];
    zero_acc<H>(acc);
    gemm_acc<H>(acc, sD, 72, n_out, params + sh.off_w3, H, H, true, Bs2);
    mul_activation_grad<H, ACT>(acc, h2 + m0 * H, valid);
    store_tile<H>(acc, dz2_hi + m0 * H, H, valid, [](float a, int, int) { return tf32_hi(a); });
    store_tile<H>(acc, dz2_lo + m0 * H, H, valid, [](float a, int, int) { return a - tf32_hi(a); });
}

// dx[:, j] = sum_n dz1[:, n] W1[n, col0 + j]
template <int H>
__global__ void __launch_bounds__(NTHREADS, 1)
mlp_dx_kernel(TbMlpShape sh, const float* __restrict__ params, const float* __restrict__ dz1,
              int64_t n_rows, float* __restrict__ dx, int dx_col0, int dx_cols, const int32_t* d_skip) {
    using C = Cfg<H>;
    if (skip_requested(d_skip)) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* bufA = reinterpret_cast<float*>(smem_raw);          // [TM][LDH]
    float* Bs2 = bufA + TM * C::LDH;
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * TM;
    const int valid = (int)min((int64_t)TM, n_rows - m0);
    for (int v = tid; v < TM * (H / 4); v += NTHREADS) {
        const int m = v / (H / 4), c4 = v % (H / 4);
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < valid) x = __ldg(reinterpret_cast<const float4*>(dz1 + (m0 + m) * H) + c4);
        *reinterpret_cast<float4*>(bufA + m * C::LDH + c4 * 4) = x;
    }
    __syncthreads();
    float acc[8][C::NC];
    zero_acc<H>(acc);
    gemm_acc<H>(acc, bufA, C::LDH, H, params + sh.off_w1 + dx_col0, sh.d_in, dx_cols, false, Bs2);
    const int tx = tid & 31, ty = tid >> 5;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int row = ty * 8 + r;
        if (row >= valid) continue;
#pragma unroll
        for (int j = 0; j < C::NC; ++j) {
            const int c = C::col(tx, j);
            if (c < dx_cols) dx[(m0 + row) * dx_cols + c] = acc[r][j];
        }
    }
}


// =====================================================================================
// Narrow weight gradients in one streaming pass (tensor-core path companion):
//   dW1[n][j] = sum_m dz1[m][n] xin[m][j]  (j <= d_in: the ones column gives db1)
//   dW3[o][n] = sum_m dout[m][o] h2[m][n],   db3[o] / extras = column sums of dout
// (db2 = column sums of dz2 comes out of the tensor-core kernel, csrc/tc_gemm.cu)
// One CTA per row split; thread (g, n): column n of the 256-wide activations, rows of
// parity g.  Every activation element is read exactly once, coalesced; xin / dout rows are
// staged in shared memory and broadcast.  Replaces the generic tile jobs when
// d_in + 1 <= KIN and n_out + n_extra <= NO (they spent most of their time on padding).
// =====================================================================================
constexpr int NW_ROWS = 32;          // rows staged per block
constexpr int NW_COLS = 128;         // activation columns per CTA (2 CTAs per row split)
constexpr int NW_GROUPS = 2;         // row-interleaved thread groups per CTA (2 CTAs per SM)

template <int KIN, int NO>
__global__ void __launch_bounds__(NW_COLS * NW_GROUPS, 2)
narrow_wgrad_kernel(TbMlpShape sh, const float* __restrict__ xin, const float* __restrict__ h2,
                    const float* __restrict__ dz1, const float* __restrict__ dout, int ld_dout,
                    int n_extra, int off_extra, int64_t n_rows, int64_t rows_per_split,
                    float* __restrict__ gpart, const int32_t* d_skip) {
    if (skip_requested(d_skip)) return;
    constexpr int H = 256;
    constexpr int NT = NW_COLS * NW_GROUPS;
    __shared__ float xs[NW_ROWS][KIN];
    __shared__ float ds[NW_ROWS][NO];
    extern __shared__ __align__(16) float comb[];         // [groups-1][128][KIN + NO + 1]
    const int half = blockIdx.y;                          // columns [128 half, 128 half + 128)
    const int n = half * NW_COLS + (threadIdx.x & (NW_COLS - 1)), g = threadIdx.x / NW_COLS;
    const int d_in = sh.d_in, n_out = sh.n_out, nd = n_out + n_extra;
    const int ldx = (d_in + 1 + 3) & ~3;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_split;
    const int64_t r1 = min(n_rows, r0 + rows_per_split);
    float w1[KIN], w3[NO], dsum = 0.0f;
#pragma unroll
    for (int j = 0; j < KIN; ++j) w1[j] = 0.0f;
#pragma unroll
    for (int o = 0; o < NO; ++o) w3[o] = 0.0f;

    for (int64_t base = r0; base < r1; base += NW_ROWS) {
        const int rows = (int)min((int64_t)NW_ROWS, r1 - base);
        __syncthreads();
        for (int v = threadIdx.x; v < NW_ROWS * KIN; v += NT) {
            const int r = v / KIN, j = v % KIN;
            xs[r][j] = (r < rows && j <= d_in) ? xin[(base + r) * ldx + j] : 0.0f;
        }
        for (int v = threadIdx.x; v < NW_ROWS * NO; v += NT) {
            const int r = v / NO, o = v % NO;
            ds[r][o] = (r < rows && o < nd) ? dout[(base + r) * ld_dout + o] : 0.0f;
        }
        __syncthreads();
        if (half == 0 && threadIdx.x < nd)
            for (int r = 0; r < rows; ++r) dsum += ds[r][threadIdx.x];
        // 8 rows per group and block: all 32 loads are issued before the first use
        float a1[NW_ROWS / NW_GROUPS], hv[NW_ROWS / NW_GROUPS];
#pragma unroll
        for (int q = 0; q < NW_ROWS / NW_GROUPS; ++q) {
            const int r = g + q * NW_GROUPS;
            const int64_t e = (base + min(r, rows - 1)) * H + n;
            a1[q] = __ldg(dz1 + e);
            hv[q] = __ldg(h2 + e);
        }
#pragma unroll
        for (int q = 0; q < NW_ROWS / NW_GROUPS; ++q) {
            const int r = g + q * NW_GROUPS;
            if (r < rows) {
#pragma unroll
                for (int j = 0; j < KIN; ++j) w1[j] = fmaf(a1[q], xs[r][j], w1[j]);
#pragma unroll
                for (int o = 0; o < NO; ++o) w3[o] = fmaf(ds[r][o], hv[q], w3[o]);
            }
        }
    }
    // combine the row groups (fixed order), then write this split's partial sums
    constexpr int LDC = KIN + NO;
    const int c = threadIdx.x & (NW_COLS - 1);
    __syncthreads();
    if (g > 0) {
        float* dst = comb + ((size_t)(g - 1) * NW_COLS + c) * LDC;
#pragma unroll
        for (int j = 0; j < KIN; ++j) dst[j] = w1[j];
#pragma unroll
        for (int o = 0; o < NO; ++o) dst[KIN + o] = w3[o];
    }
    __syncthreads();
    if (g == 0) {
        for (int gg = 1; gg < NW_GROUPS; ++gg) {
            const float* src = comb + ((size_t)(gg - 1) * NW_COLS + c) * LDC;
#pragma unroll
            for (int j = 0; j < KIN; ++j) w1[j] += src[j];
#pragma unroll
            for (int o = 0; o < NO; ++o) w3[o] += src[KIN + o];
        }
        float* out = gpart + (size_t)blockIdx.x * sh.n_params;
#pragma unroll
        for (int j = 0; j < KIN; ++j) {
            if (j < d_in) out[sh.off_w1 + n * d_in + j] = w1[j];
            else if (j == d_in) out[sh.off_b1 + n] = w1[j];
        }
#pragma unroll
        for (int o = 0; o < NO; ++o)
            if (o < n_out) out[sh.off_w3 + o * H + n] = w3[o];
        if (half == 0) {
            if ((int)threadIdx.x < n_out) out[sh.off_b3 + threadIdx.x] = dsum;
            else if ((int)threadIdx.x < nd) out[off_extra + (threadIdx.x - n_out)] = dsum;
        }
    }
}

// shared memory actually used by the split kernels (2 or more CTAs per SM instead of 1)
template <int H>
constexpr size_t layer1_smem_bytes() {       // input staging + weight ring + source rows
    return (size_t)(TM * Cfg<H>::LDH + 2 * KC * H) * sizeof(float) + TM * sizeof(int64_t);
}
template <int H>
constexpr size_t head_smem_bytes() { return (size_t)TM * Cfg<H>::LDH * sizeof(float); }
template <int H>
constexpr size_t head_backward_smem_bytes() { return (size_t)(2 * KC * H + TM * 72) * sizeof(float); }
template <int H>
constexpr size_t dx_smem_bytes() { return (size_t)(TM * Cfg<H>::LDH + 2 * KC * H) * sizeof(float); }

// small-batch variants on / off (TONIC_B200_SKINNY=0, or tb_debug_skinny for the equality tests)
static int g_skinny = -1;
static bool skinny_enabled() {
    if (g_skinny < 0) { const char* v = getenv("TONIC_B200_SKINNY"); g_skinny = (v && v[0] == '0') ? 0 : 1; }
    return g_skinny != 0;
}

static int check_tc_shape(const TbMlpShape* sh, const char* who) {
    int rc = check_shape(sh, who);
    if (rc) return rc;
    TB_REQUIRE(sh->hidden == 256 && sh->off_w2_hi > 0, TB_ENOTSUP,
               "%s: the tensor-core path needs hidden == 256 and the tf32 weight splits", who);
    return 0;
}

}  // namespace tb

extern "C" int tb_debug_skinny(int32_t on) {
    tb::g_skinny = on ? 1 : 0;
    return 0;
}

extern "C" int tb_mlp_forward_tc(const TbMlpShape* shape, const float* d_params,
                                 const float* d_packed, const TbMlpInput* in, int64_t n_rows,
                                 float* d_out, float* d_xin, float* d_h1_hi, float* d_h1_lo,
                                 float* d_h2, int32_t passes, const int32_t* d_skip, void* stream) {
    using namespace tb;
    int rc = check_tc_shape(shape, "tb_mlp_forward_tc");
    if (rc) return rc;
    TB_REQUIRE(d_params && d_packed && in && in->d_x1 && d_out && n_rows > 0, TB_EINVAL,
               "tb_mlp_forward_tc: null pointer");
    TB_REQUIRE(in->dim1 + (in->d_x2 ? in->dim2 : 0) == shape->d_in, TB_EINVAL,
               "tb_mlp_forward_tc: input widths do not add up to d_in");
    {
        // narrow inputs and heads: the whole forward pass is one tensor-core kernel
        // (csrc/tc_mlp.cu); TONIC_B200_FUSED_FWD=0 selects the three-kernel chain below
        static const bool fused = [] {
            const char* v = getenv("TONIC_B200_FUSED_FWD");
            return !(v && v[0] == '0');
        }();
        if (fused && shape->off_w1_img_hi > 0 && shape->d_in <= 32 && shape->n_out <= 8)
            return tb_tc_mlp_forward(shape, d_params, d_packed, in, n_rows, d_out, d_xin, d_h1_hi,
                                     d_h1_lo, d_h2, passes, d_skip, stream);
    }
    TB_REQUIRE(d_h1_hi && d_h1_lo && d_h2, TB_EINVAL,
               "tb_mlp_forward_tc: the unfused chain needs the h1 / h2 workspaces");
    const int blocks = (int)((n_rows + TM - 1) / TM);
    cudaStream_t s = as_stream(stream);
    const bool small_batch = skinny_enabled() && n_rows <= SK_MAX_ROWS;
    if (small_batch && skinny_layer1_smem(shape->d_in) <= 200 * 1024) {
        const size_t smem = skinny_layer1_smem(shape->d_in);
        const dim3 grid((unsigned)((n_rows + SK_ROWS - 1) / SK_ROWS), 256 / SK_COLS);
        ProfScope prof_scope("tb_mlp_layer1", stream);
        if (shape->act == TB_ACT_TANH) {
            set_smem(mlp_layer1_skinny_kernel<TB_ACT_TANH>, 200 * 1024);
            mlp_layer1_skinny_kernel<TB_ACT_TANH><<<grid, NTHREADS, smem, s>>>(
                *shape, d_params, d_packed, *in, n_rows, d_xin, d_h1_hi, d_h1_lo, d_skip);
        } else {
            set_smem(mlp_layer1_skinny_kernel<TB_ACT_RELU>, 200 * 1024);
            mlp_layer1_skinny_kernel<TB_ACT_RELU><<<grid, NTHREADS, smem, s>>>(
                *shape, d_params, d_packed, *in, n_rows, d_xin, d_h1_hi, d_h1_lo, d_skip);
        }
        if ((rc = check_launch("tb_mlp_forward_tc/layer1"))) return rc;
    } else {
        const size_t smem = layer1_smem_bytes<256>();
        ProfScope prof_scope("tb_mlp_layer1", stream);
        if (shape->act == TB_ACT_TANH) {
            set_smem(mlp_layer1_kernel<256, TB_ACT_TANH>, smem);
            mlp_layer1_kernel<256, TB_ACT_TANH><<<blocks, NTHREADS, smem, s>>>(
                *shape, d_params, d_packed, *in, n_rows, d_xin, d_h1_hi, d_h1_lo, d_skip);
        } else {
            set_smem(mlp_layer1_kernel<256, TB_ACT_RELU>, smem);
            mlp_layer1_kernel<256, TB_ACT_RELU><<<blocks, NTHREADS, smem, s>>>(
                *shape, d_params, d_packed, *in, n_rows, d_xin, d_h1_hi, d_h1_lo, d_skip);
        }
        if ((rc = check_launch("tb_mlp_forward_tc/layer1"))) return rc;
    }
    const bool fused_head = shape->n_out <= 8;
    rc = tb_tc_gemm256(d_h1_hi, d_h1_lo, d_packed + shape->off_w2_hi, d_packed + shape->off_w2_lo, n_rows,
                       passes, 0, shape->act, d_params + shape->off_b2, nullptr, nullptr, d_h2, nullptr,
                       fused_head ? d_params + shape->off_w3 : nullptr,
                       fused_head ? d_params + shape->off_b3 : nullptr, fused_head ? d_out : nullptr,
                       fused_head ? shape->n_out : 0, d_skip, stream);
    if (rc || fused_head) return rc;
    {
        const size_t smem = head_smem_bytes<256>();
        ProfScope prof_scope("tb_mlp_head", stream);
        if (small_batch) {       // 8-row tiles: same arithmetic per output, 8x the CTAs
            set_smem(mlp_head_kernel<256, 8>, smem);
            mlp_head_kernel<256, 8><<<(unsigned)((n_rows + 7) / 8), NTHREADS, smem, s>>>(
                *shape, d_params, d_h2, n_rows, d_out, d_skip);
        } else {
            set_smem(mlp_head_kernel<256>, smem);
            mlp_head_kernel<256><<<blocks, NTHREADS, smem, s>>>(*shape, d_params, d_h2, n_rows, d_out, d_skip);
        }
        rc = check_launch("tb_mlp_forward_tc/head");
    }
    return rc;
}

extern "C" int tb_mlp_backward_tc(const TbMlpShape* shape, const float* d_params,
                                  const float* d_packed, const float* d_dout, int32_t ld_dout,
                                  const float* d_h1_hi, const float* d_h1_lo, const float* d_h2,
                                  int64_t n_rows, float* d_dz2_hi, float* d_dz2_lo, float* d_dz1,
                                  float* d_dx, int32_t dx_col0, int32_t dx_cols, int32_t passes,
                                  const int32_t* d_skip, void* stream) {
    using namespace tb;
    int rc = check_tc_shape(shape, "tb_mlp_backward_tc");
    if (rc) return rc;
    TB_REQUIRE(d_params && d_packed && d_dout && d_h1_hi && d_h2 && d_dz2_hi &&
               d_dz1 && n_rows > 0 && ld_dout >= shape->n_out, TB_EINVAL,
               "tb_mlp_backward_tc: bad arguments");
    // d_h1_lo == NULL / d_dz2_lo == NULL: plain float32 activations (fused kernels only)
    TB_REQUIRE((d_h1_lo && d_dz2_lo) || shape->n_out <= 8, TB_EINVAL,
               "tb_mlp_backward_tc: the unfused chain needs the tf32 splits (lo arrays)");
    TB_REQUIRE(!d_dx || (dx_cols >= 1 && dx_cols <= 256 && dx_col0 >= 0 &&
                         dx_col0 + dx_cols <= shape->d_in), TB_EINVAL,
               "tb_mlp_backward_tc: dx column range invalid");
    const int blocks = (int)((n_rows + TM - 1) / TM);
    cudaStream_t s = as_stream(stream);
    static const bool fused_bwd = [] {
        const char* v = getenv("TONIC_B200_FUSED_BWD");
        return !(v && v[0] == '0');
    }();
    TB_REQUIRE((d_h1_lo && d_dz2_lo) || (fused_bwd && shape->n_out <= 8), TB_EINVAL,
               "tb_mlp_backward_tc: plain activations need the fused backward kernel");
    if (fused_bwd && shape->n_out <= 8) {
        // head gradient + hidden-layer GEMM + activation gradient in one kernel (csrc/tc_mlp.cu)
        rc = tb_tc_mlp_backward(shape, d_params, d_packed, d_dout, ld_dout, d_h1_hi, d_h1_lo, d_h2, n_rows,
                                d_dz2_hi, d_dz2_lo, d_dz1, passes, d_skip, stream);
    } else {
    {
        const size_t smem = head_backward_smem_bytes<256>();
        ProfScope prof_scope("tb_mlp_head_backward", stream);
        if (shape->act == TB_ACT_TANH) {
            set_smem(mlp_head_backward_kernel<256, TB_ACT_TANH>, smem);
            mlp_head_backward_kernel<256, TB_ACT_TANH><<<blocks, NTHREADS, smem, s>>>(
                *shape, d_params, d_dout, ld_dout, d_h2, n_rows, d_dz2_hi, d_dz2_lo, d_skip);
        } else {
            set_smem(mlp_head_backward_kernel<256, TB_ACT_RELU>, smem);
            mlp_head_backward_kernel<256, TB_ACT_RELU><<<blocks, NTHREADS, smem, s>>>(
                *shape, d_params, d_dout, ld_dout, d_h2, n_rows, d_dz2_hi, d_dz2_lo, d_skip);
        }
        if ((rc = check_launch("tb_mlp_backward_tc/head"))) return rc;
    }
    rc = tb_tc_gemm256(d_dz2_hi, d_dz2_lo, d_packed + shape->off_w2t_hi, d_packed + shape->off_w2t_lo,
                       n_rows, passes, 1, shape->act, nullptr, d_h1_hi, d_h1_lo, d_dz1, nullptr, nullptr,
                       nullptr, nullptr, 0, d_skip, stream);
    }
    if (rc || !d_dx) return rc;
    {
        ProfScope prof_scope("tb_mlp_dx", stream);
        if (skinny_enabled() && n_rows <= SK_MAX_ROWS && dx_cols <= SKD_MAX_COLS) {
            const size_t smem = skinny_dx_smem(dx_cols);
            set_smem(mlp_dx_skinny_kernel, skinny_dx_smem(SKD_MAX_COLS));
            mlp_dx_skinny_kernel<<<(unsigned)((n_rows + SK_ROWS - 1) / SK_ROWS), NTHREADS, smem, s>>>(
                *shape, d_params, d_dz1, n_rows, d_dx, dx_col0, dx_cols, d_skip);
        } else {
            const size_t smem = dx_smem_bytes<256>();
            set_smem(mlp_dx_kernel<256>, smem);
            mlp_dx_kernel<256><<<blocks, NTHREADS, smem, s>>>(*shape, d_params, d_dz1, n_rows, d_dx, dx_col0,
                                                           dx_cols, d_skip);
        }
        rc = check_launch("tb_mlp_backward_tc/dx");
    }
    return rc;
}

extern "C" int tb_mlp_wgrad_tc(const TbMlpShape* shape, const float* d_xin, const float* d_h1_hi,
                               const float* d_h1_lo, const float* d_h2, const float* d_dz1,
                               const float* d_dz2_hi, const float* d_dz2_lo, const float* d_dout,
                               int32_t ld_dout, int32_t n_extra, int32_t off_extra, int64_t n_rows,
                               float* d_gpart, int32_t n_split, int32_t n_split_w2, int32_t passes,
                               const int32_t* d_skip, void* stream) {
    using namespace tb;
    int rc = check_tc_shape(shape, "tb_mlp_wgrad_tc");
    if (rc) return rc;
    TB_REQUIRE(d_xin && d_h1_hi && d_h1_lo && d_h2 && d_dz1 && d_dz2_hi && d_dz2_lo && d_dout &&
               d_gpart && n_rows > 0 && n_split >= 1 && n_split_w2 >= 1 && n_split_w2 <= n_split &&
               ld_dout >= shape->n_out + n_extra, TB_EINVAL, "tb_mlp_wgrad_tc: bad arguments");
    // The tensor-core kernel (dW2, db2) and the narrow-gradient kernel (dW1, db1, dW3, db3) are
    // independent: the second one runs on a side stream between two events, so that inside a
    // captured graph (and in eager mode) both are resident on the SMs at the same time.
    static cudaStream_t side = nullptr;
    static cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
    static const bool overlap = [] {
        const char* v = getenv("TONIC_B200_WGRAD_OVERLAP");
        return !(v && v[0] == '0');
    }();
    const bool narrow = shape->d_in + 1 <= 32 && shape->n_out + n_extra <= 16;
    void* narrow_stream = stream;
    if (overlap && narrow) {
        if (!side) {
            TB_REQUIRE(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking) == cudaSuccess &&
                           cudaEventCreateWithFlags(&fork_ev, cudaEventDisableTiming) == cudaSuccess &&
                           cudaEventCreateWithFlags(&join_ev, cudaEventDisableTiming) == cudaSuccess,
                       TB_ENOTSUP, "tb_mlp_wgrad_tc: cannot create the side stream");
        }
        cudaEventRecord(fork_ev, as_stream(stream));
        cudaStreamWaitEvent(side, fork_ev, 0);
        narrow_stream = side;
    }
    rc = tb_tc_wgrad256(d_dz2_hi, d_dz2_lo, d_h1_hi, d_h1_lo, n_rows, passes, d_gpart, n_split_w2,
                        shape->n_params, shape->off_w2, shape->off_b2, d_skip, stream);
    if (rc) return rc;
    if (narrow) {
        // streaming single-pass kernel for the narrow gradients
        ProfScope prof_scope("tb_mlp_wgrad_narrow", narrow_stream);
        int64_t rows_per_split = ((n_rows + n_split - 1) / n_split + NW_ROWS - 1) / NW_ROWS * NW_ROWS;
        const bool small_in = shape->d_in + 1 <= 20, small_out = shape->n_out + n_extra <= 8;
#define TB_NARROW(KIN_, NO_)                                                                       \
    {                                                                                             \
        const size_t smem = (size_t)(NW_GROUPS - 1) * NW_COLS * (KIN_ + NO_) * sizeof(float); \
        set_smem(narrow_wgrad_kernel<KIN_, NO_>, smem);                                           \
        narrow_wgrad_kernel<KIN_, NO_><<<dim3(n_split, 256 / NW_COLS), NW_COLS * NW_GROUPS, smem, \
                                         as_stream(narrow_stream)>>>(                             \
            *shape, d_xin, d_h2, d_dz1, d_dout, ld_dout, n_extra, off_extra,                      \
            n_rows, rows_per_split, d_gpart, d_skip);                                             \
    }
        if (small_in && small_out) TB_NARROW(20, 8)
        else if (small_in) TB_NARROW(20, 16)
        else if (small_out) TB_NARROW(32, 8)
        else TB_NARROW(32, 16)
#undef TB_NARROW
        rc = check_launch("tb_mlp_wgrad_tc/narrow");
        if (narrow_stream != stream) {
            cudaEventRecord(join_ev, side);
            cudaStreamWaitEvent(as_stream(stream), join_ev, 0);
        }
        return rc;
    }
    ProfScope prof_scope("tb_mlp_wgrad_small", stream);
    return launch_wgrad_jobs(shape, d_xin, d_h1_hi, d_h2, d_dz1, d_dz2_hi, d_dz2_lo, d_dout, ld_dout,
                             n_extra, off_extra, n_rows, d_gpart, n_split, false, d_skip, stream);
}

// =====================================================================================
// Fused rollout: the whole segment (T vector steps) of the on-policy collector in ONE launch.
//
// Reference loop being replaced (tonic/utils/trainer.py:44-50 per vector step):
//   agent.step      actor forward + Normal sample + log-prob      torch/agents/a2c.py:41-52,75-85
//   environment.step  dynamics, auto-reset, time-outs               environments/distributed.py:28-58
//   agent.update    Segment.store + MeanStd.record                  a2c.py:58-69, replays/segments.py:27-36
// The policy is constant during a segment and environments are independent, so a CTA keeps
// its 64 environments resident (state tile in shared memory) for all T steps: per step it
// runs the 2x256 MLP on the tile (same FFMA tile code and summation order as
// mlp_forward_kernel), samples actions with the Philox stream of gauss_sample_kernel, advances
// the environments (env_dynamics.cuh) and writes observation / action / log-prob /
// next-observation / reward / reset / termination rows straight into the [T, N, ...] segment.
// =====================================================================================
#include "env_dynamics.cuh"

namespace tb {

struct RolloutArgs {
    TbEnv env;
    TbMlpShape sh;
    const float* params;
    const float* packed;
    const float* log_scale;
    const float* norm_mean;       // actor normaliser (NULL: reference behaviour)
    const float* norm_std;
    int T;
    float* seg_obs; float* seg_actions; float* seg_next_obs; float* seg_rewards;
    float* seg_resets; float* seg_terms; float* seg_logp;
    float* env_obs;               // acting observations after the last step
    double* moment_sums;          // [2 O + 1] running-normaliser sums or NULL
    uint64_t seed, counter, counter_stride;
    const uint64_t* d_counter;
};

template <int H, int ACT>
__global__ void __launch_bounds__(NTHREADS, 1)
rollout_kernel(const RolloutArgs p) {
    using C = Cfg<H>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* bufA = reinterpret_cast<float*>(smem_raw);
    float* bufB = bufA + TM * C::LDH;
    float* Bs2 = bufB + TM * C::LDH;
    float* sx = Bs2 + 2 * KC * H;                       // [TM][O] environment state tile
    const int O = p.env.obs_dim, A = p.env.act_dim, N = p.env.n_envs;
    float* s_pre = sx + TM * O;                         // [TM][A] head pre-activations
    float* s_act = s_pre + TM * A;                      // [TM][A] clipped actions
    float* s_scale = s_act + TM * A;                    // [A]
    int* s_len = reinterpret_cast<int*>(s_scale + kMaxAct);       // [TM]
    uint32_t* s_epi = reinterpret_cast<uint32_t*>(s_len + TM);    // [TM]
    double* s_score = reinterpret_cast<double*>(s_epi + TM);      // [TM] (8-byte aligned by layout)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n0 = blockIdx.x * TM;
    const int valid = min(TM, N - n0);
    const int kpad = (O + 3) & ~3;
    uint64_t counter = p.counter + (p.d_counter ? *p.d_counter : 0ull);

    for (int i = tid; i < valid * O; i += NTHREADS) sx[i] = p.env.d_state[(size_t)n0 * O + i];
    if (tid < A) s_scale[tid] = detached_scale(p.log_scale[tid], nullptr);
    if (tid < valid) {
        s_len[tid] = p.env.d_length[n0 + tid];
        s_epi[tid] = p.env.d_episode[n0 + tid];
        s_score[tid] = p.env.d_score[n0 + tid];
    }
    double msum = 0.0, msq = 0.0;                       // thread c < O owns observation column c
    __syncthreads();

    for (int t = 0; t < p.T; ++t) {
        const size_t row0 = (size_t)t * N + n0;         // first transition of this tile at step t
        // ---- acting observations: segment row, normaliser statistics, layer-1 input ---------
        for (int i = tid; i < valid * O; i += NTHREADS) p.seg_obs[row0 * O + i] = sx[i];
        if (p.moment_sums && tid < O) {
            for (int m = 0; m < valid; ++m) {
                const double v = (double)sx[m * O + tid];
                msum += v;
                msq += v * v;
            }
        }
        for (int v = tid; v < TM * kpad; v += NTHREADS) {
            const int m = v / kpad, c = v % kpad;
            float val = 0.0f;
            if (m < valid && c < O) {
                val = sx[m * O + c];
                if (p.norm_mean) val = __fdiv_rn(__fsub_rn(val, p.norm_mean[c]), p.norm_std[c]);
            }
            bufB[m * C::LDH + c] = val;
        }
        __syncthreads();
        // ---- actor forward (identical tile arithmetic to mlp_forward_kernel) ----------------
        float acc[8][C::NC];
        auto ident = [](float a, int, int) { return a; };
        zero_acc<H>(acc);
        gemm_acc<H>(acc, bufB, C::LDH, O, p.packed + p.sh.off_w1t, H, H, true, Bs2);
        bias_activate<H, ACT>(acc, p.params + p.sh.off_b1);
        store_tile<H>(acc, bufA, C::LDH, TM, ident);
        __syncthreads();
        zero_acc<H>(acc);
        gemm_acc<H>(acc, bufA, C::LDH, H, p.packed + p.sh.off_w2t, H, H, true, Bs2);
        bias_activate<H, ACT>(acc, p.params + p.sh.off_b2);
        store_tile<H>(acc, bufB, C::LDH, TM, ident);
        __syncthreads();
        {
            const float* W3 = p.params + p.sh.off_w3;
            const float* b3 = p.params + p.sh.off_b3;
            const int quad = tid >> 2, ql = tid & 3;
            for (int q = quad; q < TM * A; q += NTHREADS / 4) {
                const int m = q % TM, o = q / TM;
                const float* hrow = bufB + m * C::LDH;
                const float* wrow = W3 + (size_t)o * H;
                float s = 0.0f;
#pragma unroll 4
                for (int i = ql; i < H / 4; i += 4) {
                    const float4 hv = *reinterpret_cast<const float4*>(hrow + i * 4);
                    const float4 wv = __ldg(reinterpret_cast<const float4*>(wrow + i * 4));
                    s = fmaf(hv.x, wv.x, s); s = fmaf(hv.y, wv.y, s);
                    s = fmaf(hv.z, wv.z, s); s = fmaf(hv.w, wv.w, s);
                }
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                if (ql == 0) s_pre[m * A + o] = s + b3[o];
            }
        }
        __syncthreads();
        // ---- Normal(loc, scale).sample() + summed log-prob (same stream as gauss_sample_kernel)
        if (tid < valid) {
            Philox rng(p.seed);
            const uint64_t ctr = counter + (uint64_t)t * p.counter_stride + (uint64_t)(n0 + tid);
            float lp = 0.0f;
            float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int a = 0; a < A; ++a) {
                if ((a & 3) == 0) {
                    const uint4 r = rng(ctr, (uint64_t)(a >> 2));
                    const float2 u = box_muller(r.x, r.y), w = box_muller(r.z, r.w);
                    z = make_float4(u.x, u.y, w.x, w.y);
                }
                const float e = (a & 3) == 0 ? z.x : (a & 3) == 1 ? z.y : (a & 3) == 2 ? z.z : z.w;
                const float loc = tanhf(s_pre[tid * A + a]);
                const float sc = s_scale[a];
                const float act = __fadd_rn(__fmul_rn(e, sc), loc);
                p.seg_actions[(row0 + tid) * A + a] = act;
                s_act[tid * A + a] = fminf(fmaxf(act, -1.0f), 1.0f);     // wrappers.py:22
                const float d = act - loc;
                lp += -(d * d) / (2.0f * (sc * sc)) - logf(sc) - kLogSqrt2Pi;
            }
            p.seg_logp[row0 + tid] = lp;
        }
        __syncthreads();
        // ---- environment transitions (one warp per environment at a time) ---------------------
        for (int e = warp; e < valid; e += NTHREADS / 32) {
            const int n = n0 + e;
            float* x = sx + e * O;
            float reward;
            int term;
            env_transition_warp(x, s_act + e * A, O, A, lane, &reward, &term);
            for (int j = lane; j < O; j += 32) p.seg_next_obs[(row0 + e) * O + j] = x[j];
            int reset = 0;
            uint32_t episode = 0;
            if (lane == 0) {
                int length = s_len[e] + 1;
                reset = term || (length == p.env.max_episode_steps);     // distributed.py:40
                double score = s_score[e] + (double)reward;
                episode = s_epi[e];
                if (reset) {
                    const unsigned long long slot = atomicAdd(p.env.d_ep_count, 1ull);
                    if (p.env.log_cap > 0) {
                        p.env.d_ep_scores[slot % p.env.log_cap] = score;
                        p.env.d_ep_lengths[slot % p.env.log_cap] = length;
                    }
                    s_epi[e] = episode + 1u;
                    length = 0;
                    score = 0.0;
                }
                s_len[e] = length;
                s_score[e] = score;
                p.seg_rewards[row0 + e] = reward;
                p.seg_resets[row0 + e] = reset ? 1.0f : 0.0f;
                p.seg_terms[row0 + e] = term ? 1.0f : 0.0f;
            }
            reset = __shfl_sync(0xffffffffu, reset, 0);
            episode = __shfl_sync(0xffffffffu, episode, 0);
            if (reset) {                                                 // distributed.py:46-48
                const uint32_t key = reset_key((uint32_t)(p.env.seed + p.env.first_worker + n), episode);
                for (int j = lane; j < O; j += 32) x[j] = reset_coordinate(key, j);
            }
        }
        __syncthreads();
    }
    // ---- write the resident state back ----------------------------------------------------------
    for (int i = tid; i < valid * O; i += NTHREADS) {
        p.env.d_state[(size_t)n0 * O + i] = sx[i];
        p.env_obs[(size_t)n0 * O + i] = sx[i];
    }
    if (tid < valid) {
        p.env.d_length[n0 + tid] = s_len[tid];
        p.env.d_episode[n0 + tid] = s_epi[tid];
        p.env.d_score[n0 + tid] = s_score[tid];
    }
    if (p.moment_sums && tid < O) {
        atomicAdd(&p.moment_sums[tid], msum);
        atomicAdd(&p.moment_sums[O + tid], msq);
    }
    if (p.moment_sums && tid == 0) atomicAdd(&p.moment_sums[2 * O], (double)valid * (double)p.T);
}

}  // namespace tb

extern "C" int tb_rollout_fused(const TbEnv* env, const TbMlpShape* shape, const float* d_params,
                                const float* d_packed, const float* d_log_scale,
                                const float* d_norm_mean, const float* d_norm_std, int32_t T,
                                float* d_seg_obs, float* d_seg_actions, float* d_seg_next_obs,
                                float* d_seg_rewards, float* d_seg_resets, float* d_seg_terms,
                                float* d_seg_logp, float* d_env_obs, double* d_moment_sums,
                                uint64_t seed, uint64_t counter, uint64_t counter_stride,
                                const uint64_t* d_counter, void* stream) {
    using namespace tb;
    ProfScope prof_scope("tb_rollout_fused", stream);
    int rc = check_shape(shape, "tb_rollout_fused");
    if (rc) return rc;
    TB_REQUIRE(env && d_params && d_packed && d_log_scale && d_seg_obs && d_seg_actions &&
               d_seg_next_obs && d_seg_rewards && d_seg_resets && d_seg_terms && d_seg_logp &&
               d_env_obs && T > 0, TB_EINVAL, "tb_rollout_fused: null pointer");
    TB_REQUIRE(!env->time_feature, TB_ENOTSUP, "tb_rollout_fused: the time feature is not supported");
    TB_REQUIRE(shape->d_in == env->obs_dim && shape->n_out == env->act_dim, TB_EINVAL,
               "tb_rollout_fused: network / environment shapes differ");
    TB_REQUIRE(env->obs_dim <= 64 && env->obs_dim <= shape->hidden && env->act_dim <= 16, TB_ENOTSUP,
               "tb_rollout_fused: needs obs_dim <= 64 and act_dim <= 16 (got %d, %d)",
               env->obs_dim, env->act_dim);
    RolloutArgs a;
    a.env = *env; a.sh = *shape; a.params = d_params; a.packed = d_packed; a.log_scale = d_log_scale;
    a.norm_mean = d_norm_mean; a.norm_std = d_norm_std; a.T = T;
    a.seg_obs = d_seg_obs; a.seg_actions = d_seg_actions; a.seg_next_obs = d_seg_next_obs;
    a.seg_rewards = d_seg_rewards; a.seg_resets = d_seg_resets; a.seg_terms = d_seg_terms;
    a.seg_logp = d_seg_logp; a.env_obs = d_env_obs; a.moment_sums = d_moment_sums;
    a.seed = seed; a.counter = counter; a.counter_stride = counter_stride; a.d_counter = d_counter;
    const int blocks = (env->n_envs + TM - 1) / TM;
    const int H = shape->hidden;
    const size_t smem = (size_t)(2 * TM * (H + 4) + 2 * KC * H + TM * env->obs_dim +
                                 2 * TM * env->act_dim + kMaxAct) * sizeof(float) +
                        TM * (sizeof(int) + sizeof(uint32_t) + sizeof(double)) + 16;
    cudaStream_t s = as_stream(stream);
#define CALL(H_, A_)                                                              \
    {                                                                             \
        set_smem(rollout_kernel<H_, A_>, smem);                                   \
        rollout_kernel<H_, A_><<<blocks, NTHREADS, smem, s>>>(a);                 \
    }
    TB_DISPATCH_H_ACT(H, shape->act, CALL);
#undef CALL
    return check_launch("tb_rollout_fused");
}
