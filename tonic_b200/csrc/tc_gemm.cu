// Tensor-core path for the 256-wide hidden-layer GEMMs (sm_100a only):
//
//     out[M, 256] = epilogue( A[M, 256] . B[256, 256]^T )        (both operands K-major)
//
// used for the second hidden layer of the forward pass (A = h1, B = W2, epilogue =
// bias + activation) and for the hidden-layer gradient of the backward pass
// (A = dz2, B = W2^T, epilogue = * act'(h1)); reference arithmetic:
// tonic/torch/models/utils.py:15-23 (Linear + activation) and its autograd.
//
// Design (one CTA per SM, persistent over 128-row tiles):
//   warp 0   TMA producer: cp.async.bulk.tensor 2-D tiles (128B swizzle) of the
//            A row-block and the whole B matrix, K in chunks of 32 floats, into a
//            multi-stage shared-memory ring guarded by full/empty mbarriers;
//   warp 1   MMA issuer: one elected thread issues tcgen05.mma.kind::tf32
//            (M=128, N=256, K=8) with the accumulator in TMEM; two accumulator
//            stages (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of
//            tile i+1; tcgen05.commit releases smem stages / publishes accumulators;
//   warp 2   TMEM allocation / deallocation;
//   warps 4-7 epilogue: tcgen05.ld (32 lanes x 32 columns per warp) -> registers ->
//            padded smem transpose -> bias/activation (or activation gradient) ->
//            coalesced 128-byte global stores.
//
// Precision: float32 operands are pre-split into tf32-exact parts a = a_hi + a_lo
// (a_hi = a with the low 13 mantissa bits cleared, a_lo = a - a_hi, exact).  With
// PASSES = 3 the kernel accumulates a_hi.b_hi + a_lo.b_hi + a_hi.b_lo in the fp32
// TMEM accumulator ("3xTF32"): the dropped terms are O(2^-22) relative, i.e. fp32
// grade, which keeps the 1e-4 loss parity with the reference's fp32 CPU path.
// PASSES = 1 is plain TF32 (fast mode).
#include "tc_common.cuh"
#include "adam.cuh"
#include "peers.cuh"

namespace tb {

enum { TC_EPI_BIAS_ACT = 0, TC_EPI_ACT_GRAD = 1, TC_EPI_NONE = 2 };

struct TcParams {
    int64_t n_rows;
    float* out;                 // [n_rows, 256]
    const float* bias;          // [256]            (EPI_BIAS_ACT)
    const float* aux_hi;        // [n_rows, 256]    (EPI_ACT_GRAD: saved activations, split)
    const float* aux_lo;
    float* out_lo;              // optional: also write the tf32 split of the result (out = hi)
    int act;
    const int32_t* skip;
    // optional fused linear head (EPI_BIAS_ACT): head_out[row, o] = head_b[o] + out[row, :] . head_w[o, :]
    const float* head_w;        // [n_head, 256]
    const float* head_b;        // [n_head]
    float* head_out;            // [n_rows, n_head]
    int n_head;                 // 0 = no fused head, <= TC_MAX_HEAD
};

template <int PASSES, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
               const TcParams p) {
    using Cfg = TcCfg<PASSES>;
    if (skip_requested(p.skip)) return;
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment by pointer arithmetic on the shared symbol (an integer round trip
    // would turn every staging access into a generic LD / ST)
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* stage_base = smem;
    float* epi = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
    float* s_head_w = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::EPI_BYTES);
    float* s_bias = s_head_w + 8 * TC_BN;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::EPI_BYTES +
                                                 Cfg::HEAD_BYTES);
    uint64_t* full_bar = bars;                       // [STAGES]
    uint64_t* empty_bar = bars + Cfg::STAGES;        // [STAGES]
    uint64_t* tmem_full = bars + 2 * Cfg::STAGES;    // [2]
    uint64_t* tmem_empty = tmem_full + 2;            // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (int)((p.n_rows + TC_BM - 1) / TC_BM);

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {        // TMEM: 512 columns = two 128 x 256 fp32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (EPI == TC_EPI_BIAS_ACT) {       // epilogue constants -> shared memory (broadcast reads)
        for (int i = threadIdx.x; i < TC_BN; i += TC_THREADS) s_bias[i] = p.bias[i];
        for (int i = threadIdx.x; i < p.n_head * TC_BN; i += TC_THREADS) s_head_w[i] = p.head_w[i];
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int c = 0; c < TC_K / TC_BK; ++c) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    unsigned char* st = stage_base + stage * Cfg::STAGE_BYTES;
                    mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                    tma_load_2d(st, &map_a_hi, &full_bar[stage], c * TC_BK, tile * TC_BM);
                    tma_load_2d(st + Cfg::PARTS * TC_A_BYTES, &map_b_hi, &full_bar[stage], c * TC_BK, 0);
                    if (PASSES == 3) {
                        tma_load_2d(st + TC_A_BYTES, &map_a_lo, &full_bar[stage], c * TC_BK, tile * TC_BM);
                        tma_load_2d(st + 2 * TC_A_BYTES + TC_B_BYTES, &map_b_lo, &full_bar[stage],
                                    c * TC_BK, 0);
                    }
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TC_BN);
                for (int c = 0; c < TC_K / TC_BK; ++c) {
                    mbar_wait(&full_bar[stage], phase);
                    tcgen05_fence_after();
                    unsigned char* st = stage_base + stage * Cfg::STAGE_BYTES;
                    const uint64_t a_hi = umma_desc_kmajor_sw128(st);
                    const uint64_t a_lo = umma_desc_kmajor_sw128(st + TC_A_BYTES);
                    const uint64_t b_hi = umma_desc_kmajor_sw128(st + Cfg::PARTS * TC_A_BYTES);
                    const uint64_t b_lo = umma_desc_kmajor_sw128(st + 2 * TC_A_BYTES + TC_B_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k) {
                        const uint64_t koff = (uint64_t)(k * 32 >> 4);       // 8 tf32 = 32 bytes
                        if (PASSES == 3) {
                            // small cross terms first, the dominant term last
                            tcgen05_mma_tf32(d_tmem, a_lo + koff, b_hi + koff, kIdescTf32, (c | k) != 0);
                            tcgen05_mma_tf32(d_tmem, a_hi + koff, b_lo + koff, kIdescTf32, 1);
                            tcgen05_mma_tf32(d_tmem, a_hi + koff, b_hi + koff, kIdescTf32, 1);
                        } else {
                            tcgen05_mma_tf32(d_tmem, a_hi + koff, b_hi + koff, kIdescTf32, (c | k) != 0);
                        }
                    }
                    tcgen05_commit(&empty_bar[stage]);          // frees the smem stage
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
                tcgen05_commit(&tmem_full[acc]);                // accumulator ready
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int w = warp - 4;                                 // == warp % 4: TMEM lanes 32w..32w+31
        float* stg = epi + w * 32 * TC_STAGE_ROWSTRIDE;
        float* __restrict__ g_out = p.out;
        float* __restrict__ g_out_lo = p.out_lo;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            mbar_wait(&tmem_full[acc], (it >> 1) & 1);
            tcgen05_fence_after();
            const int64_t row0 = (int64_t)tile * TC_BM + w * 32;
            float hacc[TC_MAX_HEAD];
#pragma unroll
            for (int o = 0; o < TC_MAX_HEAD; ++o) hacc[o] = 0.0f;
#pragma unroll 1
            for (int c = 0; c < TC_BN / 32; ++c) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(w * 32) << 16) + (uint32_t)(acc * TC_BN + c * 32);
                tcgen05_ld_32x32(taddr, v);
                if (EPI == TC_EPI_BIAS_ACT) {
                    // this thread holds columns c*32 .. c*32+31 of ITS row: bias + activation
                    // in registers, then the row's contribution to the fused linear head
                    const float4* b4 = reinterpret_cast<const float4*>(s_bias + c * 32);
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b = b4[j / 4];
                        float x0 = __uint_as_float(v[j]) + b.x, x1 = __uint_as_float(v[j + 1]) + b.y;
                        float x2 = __uint_as_float(v[j + 2]) + b.z, x3 = __uint_as_float(v[j + 3]) + b.w;
                        if (p.act == TB_ACT_TANH) {
                            x0 = tanh_fast(x0); x1 = tanh_fast(x1); x2 = tanh_fast(x2); x3 = tanh_fast(x3);
                        } else {
                            x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); x2 = fmaxf(x2, 0.f); x3 = fmaxf(x3, 0.f);
                        }
                        v[j] = __float_as_uint(x0); v[j + 1] = __float_as_uint(x1);
                        v[j + 2] = __float_as_uint(x2); v[j + 3] = __float_as_uint(x3);
                    }
#pragma unroll
                    for (int o = 0; o < TC_MAX_HEAD; ++o) {
                        if (o < p.n_head) {
                            const float4* w4 = reinterpret_cast<const float4*>(s_head_w + o * TC_BN + c * 32);
                            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;     // independent chains
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 wv = w4[j / 4];
                                s0 = fmaf(__uint_as_float(v[j]), wv.x, s0);
                                s1 = fmaf(__uint_as_float(v[j + 1]), wv.y, s1);
                                s2 = fmaf(__uint_as_float(v[j + 2]), wv.z, s2);
                                s3 = fmaf(__uint_as_float(v[j + 3]), wv.w, s3);
                            }
                            hacc[o] += (s0 + s1) + (s2 + s3);
                        }
                    }
                }
                float* mine = stg + lane * TC_STAGE_ROWSTRIDE;
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(mine + j) =
                        make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                    __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                __syncwarp();
                const int col = c * 32 + lane;
                if (EPI == TC_EPI_ACT_GRAD) {
                    // all 64 loads of the chunk are issued before the first use (the row-by-row
                    // form left ~4 loads in flight per warp and was latency bound)
                    float gh[32], gl[32];
#pragma unroll
                    for (int r = 0; r < 32; ++r) {       // 64 independent loads, no use in between
                        const int64_t row = min(row0 + r, p.n_rows - 1);
                        gh[r] = ldg_nc_volatile(p.aux_hi + row * TC_BN + col);
                        gl[r] = ldg_nc_volatile(p.aux_lo + row * TC_BN + col);
                    }
#pragma unroll
                    for (int r = 0; r < 32; ++r) gh[r] += gl[r];
#pragma unroll
                    for (int r = 0; r < 32; ++r) {
                        const int64_t row = row0 + r;
                        if (row < p.n_rows) {
                            const float h = gh[r];
                            const float g = p.act == TB_ACT_TANH ? (1.0f - h * h) : (h > 0.0f ? 1.0f : 0.0f);
                            g_out[row * TC_BN + col] = stg[r * TC_STAGE_ROWSTRIDE + lane] * g;
                        }
                    }
                } else {
#pragma unroll 8
                    for (int r = 0; r < 32; ++r) {
                        const int64_t row = row0 + r;
                        if (row >= p.n_rows) break;
                        const float x = stg[r * TC_STAGE_ROWSTRIDE + lane];
                        if (g_out_lo) {
                            const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
                            g_out[row * TC_BN + col] = hi;
                            g_out_lo[row * TC_BN + col] = x - hi;
                        } else {
                            g_out[row * TC_BN + col] = x;
                        }
                    }
                }
                __syncwarp();
            }
            if (EPI == TC_EPI_BIAS_ACT && p.n_head > 0 && row0 + lane < p.n_rows) {
#pragma unroll
                for (int o = 0; o < TC_MAX_HEAD; ++o)
                    if (o < p.n_head) p.head_out[(row0 + lane) * p.n_head + o] = hacc[o] + __ldg(p.head_b + o);
            }
            tcgen05_fence_before();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512)
                     : "memory");
    }
}


// =====================================================================================
// Weight gradient of the hidden layer on the tensor cores:
//     dW2[n, k] = sum_m dz2[m, n] * h1[m, k]          (reduction over the rows m)
// Both operands are "MN-major" for the MMA (the reduction index m is the slow axis of the
// row-major activation arrays), which tcgen05 supports for tf32 through the a_major /
// b_major bits of the instruction descriptor.  CTA (tile, split): output rows
// n in [128 tile, 128 tile + 128), all 256 columns k, rows m of split `split`; the fp32
// partial sum goes to gpart[split] (reduced by the Adam kernel in a fixed order).
// For MN-major tf32 operands the only shared-memory layout the tensor core accepts is the
// 128-byte swizzle with a 32-byte atom (cute::UMMA::LayoutType::SWIZZLE_128B_BASE32B,
// Swizzle<2,5,2>; bring-up on B200: with the plain 128B swizzle and a_major / b_major set the
// MMA silently returns zeros).  TMA writes exactly that layout with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B for a box of 32 floats x 32 rows: [32 m-rows][128 B],
// 32-byte chunks XOR-ed with (row % 4).  Swizzle atoms are 4 rows (512 B) apart along K (SBO),
// one MMA (K = 8) consumes two of them (1024 B), 32-column groups along M / N are LBO = 4096 B
// apart.
// =====================================================================================
constexpr int TCW_ROWS = 32;                                   // m rows per pipeline chunk
constexpr int TCW_A_BYTES = TC_BM * TCW_ROWS * 4;               // 16 KB (4 boxes of 4 KB)
constexpr int TCW_B_BYTES = TC_BN * TCW_ROWS * 4;               // 32 KB (8 boxes of 4 KB)

__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128(const void* smem, int lbo = 4096,
                                                            int sbo = 512, int layout = 1) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(smem) >> 4) & 0x3FFF);
    d |= (uint64_t)(lbo >> 4) << 16;        // LBO: next 32-element group along M / N
    d |= (uint64_t)(sbo >> 4) << 32;        // SBO: next 4-row swizzle atom along K
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;            // 1 = SWIZZLE_128B_BASE32B
    return d;
}
constexpr uint32_t kIdescTf32MN = kIdescTf32 | (1u << 15) | (1u << 16);   // A and B MN-major

struct TcWgradParams {
    int64_t n_rows;
    int64_t rows_per_split;     // multiple of TCW_ROWS
    float* gpart;               // [n_split, n_params]
    int n_params;
    int off_w2;                 // offset of W2 [256, 256] in the flat layout
    int off_b2;                 // offset of b2 [256] (column sums of dz2), or -1
    const int32_t* skip;
    int dbg_lbo, dbg_sbo, dbg_kstep, dbg_idesc_xor;   // bring-up knobs (0 = defaults)
};

template <int PASSES>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_wgrad_kernel(const __grid_constant__ CUtensorMap map_dz_hi, const __grid_constant__ CUtensorMap map_dz_lo,
                const __grid_constant__ CUtensorMap map_h_hi, const __grid_constant__ CUtensorMap map_h_lo,
                const TcWgradParams p) {
    using Cfg = TcCfg<PASSES>;
    if (skip_requested(p.skip)) return;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* epi = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
    // [8 x 32] block of ones: B operand of the bias-gradient MMAs (any layout of ones is ones)
    float* ones = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::EPI_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::EPI_BYTES +
                                                 Cfg::HEAD_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + Cfg::STAGES;
    uint64_t* tmem_full = bars + 2 * Cfg::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x;                                 // 0 / 1: output rows 128*tile..
    const int split = blockIdx.y;
    const int64_t m_begin = (int64_t)split * p.rows_per_split;
    const int64_t m_end = min(p.n_rows, m_begin + p.rows_per_split);
    const int n_chunks = m_end > m_begin ? (int)((m_end - m_begin + TCW_ROWS - 1) / TCW_ROWS) : 0;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {       // 256 columns for dW2 + 16 for the column sums (power of two: 512)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (p.off_b2 >= 0) {
        for (int i = threadIdx.x; i < 1024; i += TC_THREADS) ones[i] = 1.0f;
        fence_proxy_async_smem();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int c = 0; c < n_chunks; ++c) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                unsigned char* st = smem + stage * Cfg::STAGE_BYTES;
                mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                const int m0 = (int)(m_begin + (int64_t)c * TCW_ROWS);
                // rows >= n_rows are zero-filled by TMA; rows in [m_end, chunk end) belong to
                // the next split, so the chunk grid must align with the split boundaries
                for (int b = 0; b < TC_BM / 32; ++b)
                    tma_load_2d(st + b * 4096, &map_dz_hi, &full_bar[stage], tile * TC_BM + b * 32, m0);
                unsigned char* bh = st + Cfg::PARTS * TCW_A_BYTES;
                for (int b = 0; b < TC_BN / 32; ++b)
                    tma_load_2d(bh + b * 4096, &map_h_hi, &full_bar[stage], b * 32, m0);
                if (PASSES == 3) {
                    for (int b = 0; b < TC_BM / 32; ++b)
                        tma_load_2d(st + TCW_A_BYTES + b * 4096, &map_dz_lo, &full_bar[stage],
                                    tile * TC_BM + b * 32, m0);
                    unsigned char* bl = st + 2 * TCW_A_BYTES + TCW_B_BYTES;
                    for (int b = 0; b < TC_BN / 32; ++b)
                        tma_load_2d(bl + b * 4096, &map_h_lo, &full_bar[stage], b * 32, m0);
                }
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int c = 0; c < n_chunks; ++c) {
                mbar_wait(&full_bar[stage], phase);
                tcgen05_fence_after();
                unsigned char* st = smem + stage * Cfg::STAGE_BYTES;
                const int lbo = p.dbg_lbo ? p.dbg_lbo : 4096, sbo = p.dbg_sbo ? p.dbg_sbo : 512;
                const int kstep = p.dbg_kstep ? p.dbg_kstep : 1024;
                const uint32_t kIdescTf32MN = tb::kIdescTf32MN ^ (uint32_t)p.dbg_idesc_xor;
                const uint64_t a_hi = umma_desc_mnmajor_sw128(st, lbo, sbo);
                const uint64_t a_lo = umma_desc_mnmajor_sw128(st + TCW_A_BYTES, lbo, sbo);
                const uint64_t b_hi = umma_desc_mnmajor_sw128(st + Cfg::PARTS * TCW_A_BYTES, lbo, sbo);
                const uint64_t b_lo = umma_desc_mnmajor_sw128(st + 2 * TCW_A_BYTES + TCW_B_BYTES, lbo, sbo);
#pragma unroll
                for (int k = 0; k < TCW_ROWS / 8; ++k) {
                    const uint64_t koff = (uint64_t)(k * kstep >> 4);     // next 8-row group
                    if (PASSES == 3) {
                        tcgen05_mma_tf32(tmem_base, a_lo + koff, b_hi + koff, kIdescTf32MN, (c | k) != 0);
                        tcgen05_mma_tf32(tmem_base, a_hi + koff, b_lo + koff, kIdescTf32MN, 1);
                        tcgen05_mma_tf32(tmem_base, a_hi + koff, b_hi + koff, kIdescTf32MN, 1);
                    } else {
                        tcgen05_mma_tf32(tmem_base, a_hi + koff, b_hi + koff, kIdescTf32MN, (c | k) != 0);
                    }
                    if (p.off_b2 >= 0) {
                        // db2[n] = sum_m dz2[m, n]: the same A operand against a block of ones,
                        // N = 16 accumulator columns 256..271 (all 16 columns hold the sum)
                        const uint32_t idesc16 = (kIdescTf32MN & ~(0x3Fu << 17)) | ((16u >> 3) << 17);
                        const uint64_t b_ones = umma_desc_mnmajor_sw128(ones, lbo, sbo);
                        tcgen05_mma_tf32(tmem_base + TC_BN, a_hi + koff, b_ones, idesc16, (c | k) != 0);
                        if (PASSES == 3) tcgen05_mma_tf32(tmem_base + TC_BN, a_lo + koff, b_ones, idesc16, 1);
                    }
                }
                tcgen05_commit(&empty_bar[stage]);
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
            tcgen05_commit(tmem_full);
        }
    } else if (warp >= 4) {
        const int w = warp - 4;
        float* stg = epi + w * 32 * TC_STAGE_ROWSTRIDE;
        float* out = p.gpart + (size_t)split * p.n_params + p.off_w2 + (size_t)(tile * TC_BM + w * 32) * TC_BN;
        if (n_chunks > 0) {
            mbar_wait(tmem_full, 0);
            tcgen05_fence_after();
        }
#pragma unroll 1
        for (int c = 0; c < TC_BN / 32; ++c) {
            if (n_chunks > 0) {
                uint32_t v[32];
                tcgen05_ld_32x32(tmem_base + ((uint32_t)(w * 32) << 16) + (uint32_t)(c * 32), v);
                float* mine = stg + lane * TC_STAGE_ROWSTRIDE;
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(mine + j) =
                        make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                    __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                __syncwarp();
            }
#pragma unroll 4
            for (int r = 0; r < 32; ++r)
                out[(size_t)r * TC_BN + c * 32 + lane] = n_chunks > 0 ? stg[r * TC_STAGE_ROWSTRIDE + lane] : 0.0f;
            __syncwarp();
        }
        if (p.off_b2 >= 0) {
            float sum = 0.0f;
            if (n_chunks > 0) {
                uint32_t v[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
                      "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]),
                      "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(tmem_base + ((uint32_t)(w * 32) << 16) + (uint32_t)TC_BN) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                sum = __uint_as_float(v[0]);
            }
            p.gpart[(size_t)split * p.n_params + p.off_b2 + tile * TC_BM + w * 32 + lane] = sum;
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512)
                     : "memory");
    }
}


// =====================================================================================
// ALL weight gradients of one network-minibatch in ONE launch, reduced inside the kernel:
//
//   warps 0-7   the tensor-core kernel above: dW2 (3xTF32 tcgen05, MN-major operands) and
//               db2 (ones-MMA) for output rows [128 tile, +128) over the rows of split s;
//   warps 8-15  at the same time, on the FFMA pipe the tensor-core mainloop leaves idle:
//               the narrow gradients of the same (tile, split) block
//                   dW1[n][j] = sum_m dz1[m][n] xin[m][j]   (ones column of xin: db1)
//                   dW3[o][n] = sum_m dout[m][o] h2[m][n],  db3 / extras = column sums of dout
//               (autograd of models/utils.py:15-23 and the heads, as narrow_wgrad_kernel in
//               csrc/mlp.cu, which this replaces together with its side stream);
//   all warps   after a grid-wide barrier (all 2 x n_split <= 148 CTAs are resident: one per
//               SM) CTA b sums the n_split partial slots of its slice of the flat parameter
//               vector in a FIXED order ((s%4==0) + (s%4==1)) + ((s%4==2) + (s%4==3)) --
//               the order the Adam kernel used -- and writes the flat gradient, so that the
//               optimizer step (or the multi-GPU exchange) reads ONE vector instead of
//               74-148 partial slots (19 MB -> 0.3 MB per minibatch).
// =====================================================================================
constexpr int TCA_THREADS = 512;
constexpr int TCA_NARROW_WARP0 = 8;          // warps 8..15
constexpr int TCA_NO = 8;                    // head outputs with a dW3 row (n_out <= 8)
constexpr int TCA_ND = 12;                   // dout columns with a column sum (n_out + extras <= 12)
constexpr int TCA_ROWS = 16;                 // rows per pipeline chunk (two K = 8 MMA steps)
constexpr int TCA_STAGES = 3;
constexpr int TCA_XRING = 4;                 // chunk buffers of the xin / dout rows (cp.async, 3 chunks ahead)

// 4-byte asynchronous copy global -> shared; ok == false writes a zero (src-size 0)
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gmem_src, bool ok) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src),
                 "r"(ok ? 4 : 0) : "memory");
}
// One pipeline stage: MMA operands of a 16-row chunk in the MN-major 128B / 32B-atom swizzle
// (boxes of 32 columns x 16 rows = 2 KB) and the narrow warps' dz1 / h2 column halves
// ([16][128] float32, unswizzled).
template <int KIN>
struct TcaLayout {
    static constexpr int BOX = 32 * TCA_ROWS * 4;                 // 2048
    static constexpr int A_BYTES = (TC_BM / 32) * BOX;            // 8 KB  dz2 column half
    static constexpr int B_BYTES = (TC_BN / 32) * BOX;            // 16 KB h1
    static constexpr int N_BYTES = TCA_ROWS * 128 * 4;            // 8 KB
    static constexpr int A_HI = 0, A_LO = A_BYTES, B_HI = 2 * A_BYTES, B_LO = 2 * A_BYTES + B_BYTES;
    static constexpr int N_DZ1 = 2 * (A_BYTES + B_BYTES), N_H2 = N_DZ1 + N_BYTES;
    static constexpr int STAGE_BYTES = N_H2 + N_BYTES;            // 64 KB
    static_assert(STAGE_BYTES % 1024 == 0, "stages keep the 1024-byte swizzle alignment");
};
// dynamic shared memory: stages | epilogue staging | ones block (4 KB) | xin / dout rows of two
// chunks | barriers | alignment slack
template <int PASSES, int KIN>
constexpr int tca_smem_bytes() {
    return TCA_STAGES * TcaLayout<KIN>::STAGE_BYTES + TcCfg<PASSES>::EPI_BYTES + 4096 +
           TCA_XRING * TCA_ROWS * (KIN + TCA_ND) * 4 + 256 + 1024;
}

__device__ __forceinline__ void tcw_stamp(unsigned long long* timeline, int slot) {
    if (timeline && blockIdx.x == 0 && blockIdx.y == 0) timeline[slot] = (unsigned long long)clock64();
}

struct TcWgradAllParams {
    TcWgradParams w;            // tensor-core part (gpart, n_params, off_w2, off_b2, rows_per_split)
    TbMlpShape sh;
    const float* xin;           // [n_rows, ldx] with the ones column
    const float* h2;            // [n_rows, 256]
    const float* dz1;           // [n_rows, 256]
    const float* dout;          // [n_rows, ld_dout]
    int ld_dout, n_extra, off_extra;
    int n_split;
    float* flat;                // [n_params] reduced gradient (sum over rows)
    unsigned long long* sync;   // grid-barrier word: generation << 32 | arrivals
    unsigned long long* timeline;   // profiling aid: clock64() stamps of CTA (0, 0), or NULL
    int tma3d;                  // operand tiles by one 3-D box each (else one 2-D box per column group)
    int dbg;                    // timing experiments only (TB_WGRAD_DBG): 1 = no bias-gradient MMAs, 2 = no FFMA block
    // optional fused optimizer step (single process, no gradient clipping): the reduction phase
    // applies Adam to its slice right away -- same arithmetic and device-side controls as adam_kernel
    int fuse_adam;
    TbAdam opt;
    float* packed;
    float grad_scale;
    const double* stats;        // minibatch statistics (PPO controls) or NULL
    float kl_threshold;
    int32_t* stop;
    // several ranks (peers.world > 1, needs fuse_adam): the reduction phase exchanges the slices
    // over NVLink peer memory (peers.cuh) before Adam; reduce_stats = statistics block summed over
    // the ranks in place (also the PPO controls when `stats` is set)
    TbPeers peers;
    unsigned long long* epoch;
    double* reduce_stats;
    int two_phase;              // force the reduce-scatter + all-gather exchange (TONIC_B200_PEER_TWO_PHASE=1)
};

// Grid-wide barrier of a launch whose CTAs are all resident.  *counter = generation << 32 |
// arrivals: the last arrival resets the arrivals and bumps the generation in ONE atomic add, the
// others spin until the generation moves.  The word is back to "0 arrivals" after every launch, so
// consecutive launches on the same counter may use different grid sizes (ragged last minibatch).
__device__ __forceinline__ void grid_barrier(unsigned long long* counter, unsigned int n_ctas) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long old = atomicAdd(counter, 1ULL);
        const unsigned int generation = (unsigned int)(old >> 32);
        if ((unsigned int)old == n_ctas - 1u) {
            atomicAdd(counter, (1ULL << 32) - (unsigned long long)n_ctas);
        } else {
            unsigned long long seen;
            for (;;) {
                asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(counter) : "memory");
                if ((unsigned int)(seen >> 32) != generation) break;
                __nanosleep(40);
            }
        }
        __threadfence();
    }
    __syncthreads();
}

// PLAIN: dz2 / h1 are single float32 arrays; the TMA lands them in the "hi" operand slots and
// the FFMA warps split every tile in shared memory (hi = tf32 truncation in place, lo = x - hi
// in the "lo" slot: bit-identical to the splits the forward / backward kernels used to store)
// before the MMA thread is released -- half the activation bytes through L2.
template <int PASSES, int KIN, bool PLAIN>
__global__ void __launch_bounds__(TCA_THREADS, 1)
tc_wgrad_all_kernel(const __grid_constant__ CUtensorMap map_dz_hi, const __grid_constant__ CUtensorMap map_dz_lo,
                    const __grid_constant__ CUtensorMap map_h_hi, const __grid_constant__ CUtensorMap map_h_lo,
                    const __grid_constant__ CUtensorMap map_dz1, const __grid_constant__ CUtensorMap map_h2,
                    const __grid_constant__ CUtensorMap map3_dz_hi, const __grid_constant__ CUtensorMap map3_dz_lo,
                    const __grid_constant__ CUtensorMap map3_h_hi, const __grid_constant__ CUtensorMap map3_h_lo,
                    const TcWgradAllParams q) {
    using L = TcaLayout<KIN>;
    const TcWgradParams& p = q.w;
    if (skip_requested(p.skip)) return;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* epi = reinterpret_cast<float*>(smem + TCA_STAGES * L::STAGE_BYTES);
    // [8 x 32] block of ones (B operand of the bias-gradient MMAs), then the narrow warps' xin / dout rows
    float* ones = reinterpret_cast<float*>(smem + TCA_STAGES * L::STAGE_BYTES + TcCfg<PASSES>::EPI_BYTES);
    float* xs = ones + 1024;                             // [XRING][16][KIN]
    float* dsm = xs + TCA_XRING * TCA_ROWS * KIN;        // [XRING][16][TCA_ND]
    uint64_t* bars = reinterpret_cast<uint64_t*>(dsm + TCA_XRING * TCA_ROWS * TCA_ND);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + TCA_STAGES;
    uint64_t* tmem_full = bars + 2 * TCA_STAGES;
    uint64_t* split_bar = tmem_full + 1;                 // [STAGES] operands split (PLAIN)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(split_bar + TCA_STAGES);
    static_assert(KIN % 4 == 0 && TCA_ND % 4 == 0 && TCA_NO % 4 == 0, "16-byte shared-memory rows");
    static_assert(!PLAIN || PASSES == 3, "plain activations are split for the 3-pass mode");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x;                                 // 0 / 1: columns 128*tile..
    const int split = blockIdx.y;
    const int64_t m_begin = (int64_t)split * p.rows_per_split;
    const int64_t m_end = min(p.n_rows, m_begin + p.rows_per_split);
    const int n_chunks = m_end > m_begin ? (int)((m_end - m_begin + TCA_ROWS - 1) / TCA_ROWS) : 0;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < TCA_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 2);          // released by the MMA commit AND by the narrow warps
            mbar_init(&split_bar[s], 1);
        }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 1024; i += TCA_THREADS) ones[i] = 1.0f;
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) { tcw_stamp(q.timeline, 0); }                 // setup done

    if (warp == 0) {
        // ===== TMA producer: per 16-row chunk the MMA operands (dz2 half, h1, hi / lo) and the
        // narrow warps' dz1 / h2 column halves, 3 stages deep =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int c = 0; c < n_chunks; ++c) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (c < 14) tcw_stamp(q.timeline, 16 + c);           // stage free -> TMA issue of chunk c
                unsigned char* st = smem + stage * L::STAGE_BYTES;
                mbar_expect_tx(&full_bar[stage],
                               ((PASSES == 3 && !PLAIN) ? 2 : 1) * (L::A_BYTES + L::B_BYTES) + 2 * L::N_BYTES);
                const int m0 = (int)(m_begin + (int64_t)c * TCA_ROWS);
                if (q.tma3d) {
                    // one 3-D box per operand tile (4 / 8 column groups x 16 rows x 128 B)
                    tma_load_3d(st + L::A_HI, &map3_dz_hi, &full_bar[stage], 0, m0, tile * (TC_BM / 32));
                    tma_load_3d(st + L::B_HI, &map3_h_hi, &full_bar[stage], 0, m0, 0);
                    if (PASSES == 3 && !PLAIN) {
                        tma_load_3d(st + L::A_LO, &map3_dz_lo, &full_bar[stage], 0, m0, tile * (TC_BM / 32));
                        tma_load_3d(st + L::B_LO, &map3_h_lo, &full_bar[stage], 0, m0, 0);
                    }
                    tma_load_2d(st + L::N_DZ1, &map_dz1, &full_bar[stage], tile * TC_BM, m0);
                    tma_load_2d(st + L::N_H2, &map_h2, &full_bar[stage], tile * TC_BM, m0);
                    if (++stage == TCA_STAGES) { stage = 0; phase ^= 1; }
                    continue;
                }
                for (int b = 0; b < TC_BM / 32; ++b)
                    tma_load_2d(st + L::A_HI + b * L::BOX, &map_dz_hi, &full_bar[stage], tile * TC_BM + b * 32, m0);
                for (int b = 0; b < TC_BN / 32; ++b)
                    tma_load_2d(st + L::B_HI + b * L::BOX, &map_h_hi, &full_bar[stage], b * 32, m0);
                if (PASSES == 3 && !PLAIN) {
                    for (int b = 0; b < TC_BM / 32; ++b)
                        tma_load_2d(st + L::A_LO + b * L::BOX, &map_dz_lo, &full_bar[stage], tile * TC_BM + b * 32, m0);
                    for (int b = 0; b < TC_BN / 32; ++b)
                        tma_load_2d(st + L::B_LO + b * L::BOX, &map_h_lo, &full_bar[stage], b * 32, m0);
                }
                tma_load_2d(st + L::N_DZ1, &map_dz1, &full_bar[stage], tile * TC_BM, m0);
                tma_load_2d(st + L::N_H2, &map_h2, &full_bar[stage], tile * TC_BM, m0);
                if (++stage == TCA_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t idesc16 = (kIdescTf32MN & ~(0x3Fu << 17)) | ((16u >> 3) << 17);
            const uint64_t b_ones = umma_desc_mnmajor_sw128(ones, L::BOX);
            for (int c = 0; c < n_chunks; ++c) {
                mbar_wait(&full_bar[stage], phase);
                if (c < 14) tcw_stamp(q.timeline, 32 + c);           // chunk c landed
                if (PLAIN) mbar_wait(&split_bar[stage], phase);
                if (c < 14) tcw_stamp(q.timeline, 48 + c);           // (split done ->) MMAs of chunk c issued next
                tcgen05_fence_after();
                unsigned char* st = smem + stage * L::STAGE_BYTES;
                // 16-row boxes: 32-column groups are BOX = 2048 bytes apart (LBO)
                const uint64_t a_hi = umma_desc_mnmajor_sw128(st + L::A_HI, L::BOX);
                const uint64_t a_lo = umma_desc_mnmajor_sw128(st + L::A_LO, L::BOX);
                const uint64_t b_hi = umma_desc_mnmajor_sw128(st + L::B_HI, L::BOX);
                const uint64_t b_lo = umma_desc_mnmajor_sw128(st + L::B_LO, L::BOX);
#pragma unroll
                for (int k = 0; k < TCA_ROWS / 8; ++k) {
                    const uint64_t koff = (uint64_t)(k * 1024 >> 4);     // next 8-row group
                    if (PASSES == 3) {
                        tcgen05_mma_tf32(tmem_base, a_lo + koff, b_hi + koff, kIdescTf32MN, (c | k) != 0);
                        tcgen05_mma_tf32(tmem_base, a_hi + koff, b_lo + koff, kIdescTf32MN, 1);
                        tcgen05_mma_tf32(tmem_base, a_hi + koff, b_hi + koff, kIdescTf32MN, 1);
                    } else {
                        tcgen05_mma_tf32(tmem_base, a_hi + koff, b_hi + koff, kIdescTf32MN, (c | k) != 0);
                    }
                    // db2[n] = sum_m dz2[m, n]: the A operand against a block of ones (N = 16)
                    if (!(q.dbg & 1)) {
                        tcgen05_mma_tf32(tmem_base + TC_BN, a_hi + koff, b_ones, idesc16, (c | k) != 0);
                        if (PASSES == 3) tcgen05_mma_tf32(tmem_base + TC_BN, a_lo + koff, b_ones, idesc16, 1);
                    }
                }
                tcgen05_commit(&empty_bar[stage]);
                if (++stage == TCA_STAGES) { stage = 0; phase ^= 1; }
            }
            tcgen05_commit(tmem_full);
            tcw_stamp(q.timeline, 1);                                 // all MMAs issued
        }
    } else if (warp >= 4 && warp < TCA_NARROW_WARP0) {
        // ---- epilogue: accumulator -> this split's partial slot (dW2 rows, db2) ----------------
        const int w = warp - 4;
        float* stg = epi + w * 32 * TC_STAGE_ROWSTRIDE;
        float* out = p.gpart + (size_t)split * p.n_params + p.off_w2 + (size_t)(tile * TC_BM + w * 32) * TC_BN;
        if (PLAIN) {
            // during the mainloop these four warps are the SPLITTERS: as soon as a chunk has landed
            // they derive the lo tiles (x - tf32 truncation) of dz2 (8 KB) and h1 (16 KB) -- element
            // positions are the same in the hi and lo tiles, so the swizzle does not matter -- and
            // release the MMA thread; they run up to 3 stages ahead
            const int ts = threadIdx.x - 128;                       // 0..127
            int stage = 0;
            uint32_t phase = 0;
            for (int ch = 0; ch < n_chunks; ++ch) {
                mbar_wait(&full_bar[stage], phase);
                unsigned char* st = smem + stage * L::STAGE_BYTES;
                float4* a_hi = reinterpret_cast<float4*>(st + L::A_HI);
                float4* a_lo = reinterpret_cast<float4*>(st + L::A_LO);
                float4* b_hi = reinterpret_cast<float4*>(st + L::B_HI);
                float4* b_lo = reinterpret_cast<float4*>(st + L::B_LO);
                // the hi slot keeps the plain value (the tensor core ignores its 13 low mantissa
                // bits: scratch/probe_tf32_truncation.py); only lo = x - trunc(x) is written
                auto split4 = [](const float4* hi, float4* lo, int i) {
                    const float4 v = hi[i];
                    float4 l;
                    l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
                    l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
                    l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
                    l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
                    lo[i] = l;
                };
#pragma unroll
                for (int k = 0; k < L::A_BYTES / 16 / 128; ++k) split4(a_hi, a_lo, ts + 128 * k);
#pragma unroll
                for (int k = 0; k < L::B_BYTES / 16 / 128; ++k) split4(b_hi, b_lo, ts + 128 * k);
                fence_proxy_async_smem();                       // generic-proxy writes -> tensor core
                asm volatile("bar.sync 2, 128;" ::: "memory");
                if (ts == 0) mbar_arrive(&split_bar[stage]);
                if (++stage == TCA_STAGES) { stage = 0; phase ^= 1; }
            }
        }
        if (n_chunks > 0) {
            mbar_wait(tmem_full, 0);
            tcgen05_fence_after();
        }
        if (w == 0 && lane == 0) tcw_stamp(q.timeline, 2);             // accumulator complete
#pragma unroll 1
        for (int c = 0; c < TC_BN / 32; ++c) {
            if (n_chunks > 0) {
                uint32_t v[32];
                tcgen05_ld_32x32(tmem_base + ((uint32_t)(w * 32) << 16) + (uint32_t)(c * 32), v);
                float* mine = stg + lane * TC_STAGE_ROWSTRIDE;
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(mine + j) =
                        make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                    __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                __syncwarp();
            }
#pragma unroll 4
            for (int r = 0; r < 32; ++r)
                out[(size_t)r * TC_BN + c * 32 + lane] = n_chunks > 0 ? stg[r * TC_STAGE_ROWSTRIDE + lane] : 0.0f;
            __syncwarp();
        }
        float sum = 0.0f;
        if (n_chunks > 0) {
            uint32_t v[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
                  "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]),
                  "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(tmem_base + ((uint32_t)(w * 32) << 16) + (uint32_t)TC_BN) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            sum = __uint_as_float(v[0]);
        }
        p.gpart[(size_t)split * p.n_params + p.off_b2 + tile * TC_BM + w * 32 + lane] = sum;
        if (w == 0 && lane == 0) tcw_stamp(q.timeline, 4);             // partial slot written
    } else if (warp >= TCA_NARROW_WARP0) {
        // ---- narrow gradients on the FFMA pipe.  Thread (g, c): column n = 128 tile + c, rows of
        // parity g.  dz1 / h2 of the chunk arrive in the stage by TMA (3 chunks ahead of the
        // consumer: plain loads issued here starved behind the operand stream); the few xin / dout
        // rows are prefetched one chunk ahead through registers -----------------------------------
        const int t = threadIdx.x - TCA_NARROW_WARP0 * 32;       // 0..255
        const int c = t & 127, g = t >> 7;
        const int n = tile * 128 + c;
        const TbMlpShape& sh = q.sh;
        const int d_in = sh.d_in, n_out = sh.n_out, nd = n_out + q.n_extra;
        const int ldx = (d_in + 1 + 3) & ~3;
        float w1[KIN], w3[TCA_NO], dsum = 0.0f;
#pragma unroll
        for (int j = 0; j < KIN; ++j) w1[j] = 0.0f;
#pragma unroll
        for (int o = 0; o < TCA_NO; ++o) w3[o] = 0.0f;
        // xin / dout rows of the chunks: cp.async (4-byte, zero-filled outside the split / the used
        // columns) into a ring of TCA_XRING chunk buffers, TCA_XRING - 1 chunks ahead -- the global
        // latency of these small rows paced the whole pipeline when they were staged one chunk ahead
        constexpr int NX = (TCA_ROWS * KIN + 255) / 256, NDS = (TCA_ROWS * TCA_ND + 255) / 256;
        // (row, column) of this thread's elements inside a chunk never change: precomputed
        int x_row[NX], x_off[NX], d_row[NDS], d_off[NDS];
#pragma unroll
        for (int k = 0; k < NX; ++k) {
            const int v = t + 256 * k, r = v / KIN, j = v % KIN;
            x_row[k] = (v < TCA_ROWS * KIN && j <= d_in) ? r : -1;         // -1: zero fill (padding column)
            x_off[k] = r * ldx + j;
        }
#pragma unroll
        for (int k = 0; k < NDS; ++k) {
            const int v = t + 256 * k, r = v / TCA_ND, o = v % TCA_ND;
            d_row[k] = (v < TCA_ROWS * TCA_ND && o < nd) ? r : -1;
            d_off[k] = r * q.ld_dout + o;
        }
        auto stage_async = [&](int chunk) {
            if (chunk < n_chunks) {
                const int64_t base = m_begin + (int64_t)chunk * TCA_ROWS;
                const int valid = (int)min((int64_t)TCA_ROWS, m_end - base);
                float* x = xs + (chunk % TCA_XRING) * TCA_ROWS * KIN;
                float* d = dsm + (chunk % TCA_XRING) * TCA_ROWS * TCA_ND;
                const float* gx = q.xin + base * ldx;
                const float* gd = q.dout + base * q.ld_dout;
#pragma unroll
                for (int k = 0; k < NX; ++k) {
                    if (t + 256 * k < TCA_ROWS * KIN) {
                        const bool ok = x_row[k] >= 0 && x_row[k] < valid;
                        cp_async4(x + t + 256 * k, ok ? gx + x_off[k] : q.xin, ok);
                    }
                }
#pragma unroll
                for (int k = 0; k < NDS; ++k) {
                    if (t + 256 * k < TCA_ROWS * TCA_ND) {
                        const bool ok = d_row[k] >= 0 && d_row[k] < valid;
                        cp_async4(d + t + 256 * k, ok ? gd + d_off[k] : q.dout, ok);
                    }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");      // one group per chunk (may be empty)
        };
#pragma unroll
        for (int k = 0; k < TCA_XRING - 1; ++k) stage_async(k);
        int stage = 0;
        uint32_t phase = 0;
        for (int ch = 0; ch < n_chunks; ++ch) {
            // ONE barrier per chunk: (a) every thread's copies of chunk ch have landed (the
            // TCA_XRING - 2 newer groups may be in flight) and are published, (b) everybody is done
            // with chunk ch - 1: its stage goes back to the producer and its ring buffer is reused
            asm volatile("cp.async.wait_group %0;" ::"n"(TCA_XRING - 2) : "memory");
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (ch > 0 && t == 0) mbar_arrive(&empty_bar[(stage + TCA_STAGES - 1) % TCA_STAGES]);
            stage_async(ch + TCA_XRING - 1);
            mbar_wait(&full_bar[stage], phase);                 // operands and dz1 / h2 of this chunk landed
            const unsigned char* st = smem + stage * L::STAGE_BYTES;
            const float* s_dz1 = reinterpret_cast<const float*>(st + L::N_DZ1);
            const float* s_h2 = reinterpret_cast<const float*>(st + L::N_H2);
            const int rows = (int)min((int64_t)TCA_ROWS, m_end - (m_begin + (int64_t)ch * TCA_ROWS));
            const float* x = xs + (ch % TCA_XRING) * TCA_ROWS * KIN;
            const float* d = dsm + (ch % TCA_XRING) * TCA_ROWS * TCA_ND;
            if (tile == 0 && t < nd) {
                // column sum of dout (rows beyond the split are zero-filled): 16 independent loads,
                // fixed pairwise order
                float dvv[TCA_ROWS];
#pragma unroll
                for (int r = 0; r < TCA_ROWS; ++r) dvv[r] = d[r * TCA_ND + t];
#pragma unroll
                for (int w = TCA_ROWS / 2; w > 0; w >>= 1)
#pragma unroll
                    for (int r = 0; r < w; ++r) dvv[r] += dvv[r + w];
                dsum += dvv[0];
            }
#pragma unroll
            for (int u = 0; u < TCA_ROWS / 2; ++u) {
                const int r = g + 2 * u;
                if (r < rows && !(q.dbg & 2)) {
                    const float a1 = s_dz1[r * 128 + c], hv = s_h2[r * 128 + c];
                    const float4* x4 = reinterpret_cast<const float4*>(x + r * KIN);
#pragma unroll
                    for (int j = 0; j < KIN / 4; ++j) {
                        const float4 xv = x4[j];
                        w1[4 * j] = fmaf(a1, xv.x, w1[4 * j]);
                        w1[4 * j + 1] = fmaf(a1, xv.y, w1[4 * j + 1]);
                        w1[4 * j + 2] = fmaf(a1, xv.z, w1[4 * j + 2]);
                        w1[4 * j + 3] = fmaf(a1, xv.w, w1[4 * j + 3]);
                    }
                    const float4* d4 = reinterpret_cast<const float4*>(d + r * TCA_ND);
#pragma unroll
                    for (int o = 0; o < TCA_NO / 4; ++o) {
                        const float4 dv = d4[o];
                        w3[4 * o] = fmaf(dv.x, hv, w3[4 * o]);
                        w3[4 * o + 1] = fmaf(dv.y, hv, w3[4 * o + 1]);
                        w3[4 * o + 2] = fmaf(dv.z, hv, w3[4 * o + 2]);
                        w3[4 * o + 3] = fmaf(dv.w, hv, w3[4 * o + 3]);
                    }
                }
            }
            if (++stage == TCA_STAGES) { stage = 0; phase ^= 1; }
        }
        // the last chunk's stage (nobody refills it, but the producer's bookkeeping stays balanced)
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (n_chunks > 0 && t == 0) mbar_arrive(&empty_bar[(stage + TCA_STAGES - 1) % TCA_STAGES]);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (t == 0) tcw_stamp(q.timeline, 3);                     // narrow gradients accumulated
        // combine the two row groups (fixed order: group 0 + group 1) in two rounds through the
        // dz1 / h2 block of stage 0 (the producer is done and only these warps read that block),
        // then write this split's partial slot
        float (*comb)[KIN] = reinterpret_cast<float (*)[KIN]>(smem + L::N_DZ1);       // [128][KIN] <= 16 KB
        static_assert(128 * KIN * 4 <= 2 * L::N_BYTES, "exchange block too large");
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (g == 1) {
#pragma unroll
            for (int j = 0; j < KIN; ++j) comb[c][j] = w1[j];
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (g == 0) {
#pragma unroll
            for (int j = 0; j < KIN; ++j) w1[j] += comb[c][j];
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        float (*comb3)[TCA_NO] = reinterpret_cast<float (*)[TCA_NO]>(smem + L::N_DZ1);
        if (g == 1) {
#pragma unroll
            for (int o = 0; o < TCA_NO; ++o) comb3[c][o] = w3[o];
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (g == 0) {
#pragma unroll
            for (int o = 0; o < TCA_NO; ++o) w3[o] += comb3[c][o];
            float* out = p.gpart + (size_t)split * p.n_params;
#pragma unroll
            for (int j = 0; j < KIN; ++j) {
                if (j < d_in) out[sh.off_w1 + n * d_in + j] = w1[j];
                else if (j == d_in) out[sh.off_b1 + n] = w1[j];
            }
#pragma unroll
            for (int o = 0; o < TCA_NO; ++o)
                if (o < n_out) out[sh.off_w3 + o * 256 + n] = w3[o];
            if (tile == 0) {
                if (t < n_out) out[sh.off_b3 + t] = dsum;
                else if (t < nd) out[q.off_extra + (t - n_out)] = dsum;
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512)
                     : "memory");
    }
    // ---- grid-wide barrier, then the fixed-order reduction of this CTA's parameter slice -------
    const unsigned int n_ctas = gridDim.x * gridDim.y;
    if (threadIdx.x == 0) tcw_stamp(q.timeline, 5);                   // arriving at the grid barrier
    grid_barrier(q.sync, n_ctas);
    const int b = blockIdx.y * gridDim.x + blockIdx.x;
    const int per = (((p.n_params + (int)n_ctas - 1) / (int)n_ctas) + 31) & ~31;
    const int lo = b * per, hi = min(p.n_params, lo + per);
    const bool multi = q.peers.world > 1;
    const unsigned long long epoch = multi ? *q.epoch : 0ULL;
    const int slot = (int)(epoch & 1);
    const unsigned int tag = (unsigned int)(epoch + 1);
    // more than two ranks: reduce-scatter + all-gather of flagged words (2 hops, 2 / world of the
    // all-to-all bytes); two ranks: the direct exchange (1 hop)
    const bool two_phase = multi && (q.peers.world > 2 || q.two_phase);
    double* s_stats = reinterpret_cast<double*>(smem + 64);       // global statistics (several ranks)
    // fused Adam (updaters/actors.py:22,71: no step when every advantage of the minibatch is zero)
    bool do_step = q.fuse_adam && !(q.stats && q.stats[TB_STAT_NONZERO_ADV] == 0.0);
    const int t_step = q.fuse_adam ? q.opt.d_step[0] + 1 : 0;
    float* s_corr = reinterpret_cast<float*>(smem);       // step_size, bc2_sqrt (the stages are idle)
    if (q.fuse_adam && threadIdx.x == 0) adam_corrections(q.opt, t_step, &s_corr[0], &s_corr[1]);
    __syncthreads();
    if (threadIdx.x == 0) tcw_stamp(q.timeline, 6);               // grid barrier passed
    for (int i = lo + (int)threadIdx.x; i < hi; i += TCA_THREADS) {
        // same summation order as adam_kernel: g_k = sum over s = k (mod 4) in increasing s, then
        // (g0 + g1) + (g2 + g3); 32 independent L2 loads in flight (__ldcg: written by other SMs)
        const float* src = p.gpart + i;
        float g0 = 0.0f, g1 = 0.0f, g2 = 0.0f, g3 = 0.0f;
        int sidx = 0;
        for (; sidx + 32 <= q.n_split; sidx += 32) {
            float v[32];
#pragma unroll
            for (int u = 0; u < 32; ++u) v[u] = __ldcg(src + (size_t)(sidx + u) * p.n_params);
#pragma unroll
            for (int u = 0; u < 32; u += 4) { g0 += v[u]; g1 += v[u + 1]; g2 += v[u + 2]; g3 += v[u + 3]; }
        }
        for (; sidx + 4 <= q.n_split; sidx += 4) {
            const float v0 = __ldcg(src + (size_t)(sidx + 0) * p.n_params);
            const float v1 = __ldcg(src + (size_t)(sidx + 1) * p.n_params);
            const float v2 = __ldcg(src + (size_t)(sidx + 2) * p.n_params);
            const float v3 = __ldcg(src + (size_t)(sidx + 3) * p.n_params);
            g0 += v0; g1 += v1; g2 += v2; g3 += v3;
        }
        for (int k = 0; sidx < q.n_split; ++sidx, ++k) {
            const float v = __ldcg(src + (size_t)sidx * p.n_params);
            if (k == 0) g0 += v; else if (k == 1) g1 += v; else g2 += v;
        }
        const float g = (g0 + g1) + (g2 + g3);
        if (multi) {
            if (two_phase) {
                // reduce-scatter leg: only to the rank that owns this warp-sized group of elements
                peer_put(peer_fused_grad(q.peers.base[peer_owner(i, q.peers.world)], slot, q.peers.rank, p.n_params) + i,
                         __float_as_uint(g), tag);
            } else {
                // this rank's lane of the slot in every rank's region (NVLink stores, value + tag)
                for (int r = 0; r < q.peers.world; ++r)
                    peer_put(peer_fused_grad(q.peers.base[r], slot, q.peers.rank, p.n_params) + i,
                             __float_as_uint(g), tag);
            }
            continue;
        }
        q.flat[i] = g;
        if (do_step) adam_apply(q.opt, q.sh, q.packed, i, g * q.grad_scale, s_corr[0], s_corr[1]);
    }
    if (multi) {
        void* mine = q.peers.base[q.peers.rank];
        const bool solo = q.dbg == 4;          // timing experiment: do not wait for the other ranks
        if (b == 0 && threadIdx.x < TB_STAT_COUNT) {
            const unsigned long long v = __double_as_longlong(q.reduce_stats ? q.reduce_stats[threadIdx.x] : 0.0);
            for (int r = 0; r < q.peers.world; ++r) {
                uint2* dst = peer_fused_stats(q.peers.base[r], slot, q.peers.rank, p.n_params) + 2 * threadIdx.x;
                peer_put(dst, (unsigned int)v, tag);
                peer_put(dst + 1, (unsigned int)(v >> 32), tag);
            }
        }
        if (threadIdx.x == 0) tcw_stamp(q.timeline, 8);           // slice pushed
        // global statistics: CTA 0 of every rank sent its block to THIS rank's region
        if (threadIdx.x < TB_STAT_COUNT) {
            double sum = 0.0;
            for (int r = 0; r < q.peers.world; ++r) {
                const uint2* src = peer_fused_stats(mine, slot, solo ? q.peers.rank : r, p.n_params) + 2 * threadIdx.x;
                const unsigned long long lo32 = peer_get(src, tag), hi32 = peer_get(src + 1, tag);
                sum += __longlong_as_double((long long)(lo32 | (hi32 << 32)));
            }
            s_stats[threadIdx.x] = sum;
            if (b == 0 && q.reduce_stats) q.reduce_stats[threadIdx.x] = sum;
        }
        __syncthreads();
        if (threadIdx.x == 0) tcw_stamp(q.timeline, 9);           // statistics of every rank seen
        do_step = q.fuse_adam && !(q.stats && s_stats[TB_STAT_NONZERO_ADV] == 0.0);
        if (two_phase) {
            // owners: sum the lanes in rank order, all-gather leg of the sum to every rank
            for (int i = lo + (int)threadIdx.x; i < hi; i += TCA_THREADS) {
                if (peer_owner(i, q.peers.world) != q.peers.rank) continue;
                float g = 0.0f;
                for (int r = 0; r < q.peers.world; ++r)
                    g += __uint_as_float(peer_get(peer_fused_grad(mine, slot, r, p.n_params) + i, tag));
                for (int r = 0; r < q.peers.world; ++r)
                    peer_put(peer_fused_sum(q.peers.base[r], slot, p.n_params) + i, __float_as_uint(g), tag);
            }
        }
        for (int i = lo + (int)threadIdx.x; i < hi; i += TCA_THREADS) {
            float g = 0.0f;
            if (two_phase) {
                g = __uint_as_float(peer_get(peer_fused_sum(mine, slot, p.n_params) + i, tag));
            } else {
                for (int r = 0; r < q.peers.world; ++r)
                    g += __uint_as_float(peer_get(peer_fused_grad(mine, slot, solo ? q.peers.rank : r, p.n_params) + i, tag));
            }
            q.flat[i] = g;
            if (do_step) adam_apply(q.opt, q.sh, q.packed, i, g * q.grad_scale, s_corr[0], s_corr[1]);
        }
    }
    if (threadIdx.x == 0) tcw_stamp(q.timeline, 7);               // reduction (+ Adam) done
    if (q.fuse_adam) {
        // last CTA to finish publishes the new step count and the KL early-stop flag
        // (ppo.py:45-46 with updaters/actors.py:103,112), exactly like adam_kernel
        __shared__ bool is_last;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) is_last = atomicAdd(&q.opt.d_step[1], 1) == (int)n_ctas - 1;
        __syncthreads();
        if (is_last && threadIdx.x == 0) {
            q.opt.d_step[1] = 0;
            if (multi) *q.epoch = epoch + 1;
            if (do_step) {
                q.opt.d_step[0] = t_step;
                if (q.stop && q.stats && q.kl_threshold >= 0.0f) {
                    const float kl = multi ? (float)(s_stats[TB_STAT_KL] / s_stats[TB_STAT_ROWS])
                                           : (float)(q.stats[TB_STAT_KL] / q.stats[TB_STAT_ROWS]);
                    if (kl > q.kl_threshold) *q.stop = 1;
                }
            }
        }
    }
}

// ---- split helper: hi = x with the low 13 mantissa bits cleared, lo = x - hi -------------
__global__ void __launch_bounds__(256)
split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const float v = x[i];
        const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        hi[i] = h;
        lo[i] = v - h;
    }
}

template <int PASSES, int EPI>
static int launch_tc(const CUtensorMap* maps, const TcParams& p, cudaStream_t s) {
    using Cfg = TcCfg<PASSES>;
    auto kernel = tc_gemm_kernel<PASSES, EPI>;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        configured = true;
    }
    const int n_tiles = (int)((p.n_rows + TC_BM - 1) / TC_BM);
    const int grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
    kernel<<<grid, TC_THREADS, Cfg::SMEM_BYTES, s>>>(maps[0], maps[1], maps[2], maps[3], p);
    return 0;
}

}  // namespace tb

extern "C" int tb_split_tf32(const float* d_x, float* d_hi, float* d_lo, int64_t n, void* stream) {
    tb::ProfScope prof_scope("tb_split_tf32", stream);
    TB_REQUIRE(d_x && d_hi && d_lo && n > 0, TB_EINVAL, "tb_split_tf32: bad arguments");
    int blocks = (int)((n + 255) / 256);
    if (blocks > 8 * tb::kNumSMs) blocks = 8 * tb::kNumSMs;
    tb::split_tf32_kernel<<<blocks, 256, 0, tb::as_stream(stream)>>>(d_x, d_hi, d_lo, n);
    return tb::check_launch("tb_split_tf32");
}

extern "C" int tb_tc_gemm256(const float* d_a_hi, const float* d_a_lo, const float* d_b_hi,
                             const float* d_b_lo, int64_t n_rows, int32_t passes, int32_t epilogue,
                             int32_t act, const float* d_bias, const float* d_aux_hi,
                             const float* d_aux_lo, float* d_out, float* d_out_lo,
                             const float* d_head_w, const float* d_head_b, float* d_head_out,
                             int32_t n_head, const int32_t* d_skip, void* stream) {
    using namespace tb;
    TB_REQUIRE(n_head == 0 || (epilogue == TC_EPI_BIAS_ACT && n_head <= TC_MAX_HEAD && d_head_w &&
                               d_head_b && d_head_out), TB_EINVAL,
               "tb_tc_gemm256: the fused head needs epilogue 0, n_head <= 8 and its pointers");
    TB_REQUIRE(d_a_hi && d_b_hi && d_out && n_rows > 0, TB_EINVAL, "tb_tc_gemm256: null pointer");
    TB_REQUIRE(passes == 1 || passes == 3, TB_EINVAL, "tb_tc_gemm256: passes must be 1 or 3");
    TB_REQUIRE(passes == 1 || (d_a_lo && d_b_lo), TB_EINVAL, "tb_tc_gemm256: 3 passes need the lo parts");
    TB_REQUIRE(epilogue >= 0 && epilogue <= 2, TB_EINVAL, "tb_tc_gemm256: bad epilogue");
    TB_REQUIRE(epilogue != TC_EPI_BIAS_ACT || d_bias, TB_EINVAL, "tb_tc_gemm256: bias missing");
    TB_REQUIRE(epilogue != TC_EPI_ACT_GRAD || (d_aux_hi && d_aux_lo), TB_EINVAL, "tb_tc_gemm256: aux missing");
    CUtensorMap maps[4];
    int rc;
    if ((rc = make_map(&maps[0], d_a_hi, n_rows, TC_BM))) return rc;
    if ((rc = make_map(&maps[1], d_a_lo ? d_a_lo : d_a_hi, n_rows, TC_BM))) return rc;
    if ((rc = make_map(&maps[2], d_b_hi, TC_BN, TC_BN))) return rc;
    if ((rc = make_map(&maps[3], d_b_lo ? d_b_lo : d_b_hi, TC_BN, TC_BN))) return rc;
    TcParams p;
    p.n_rows = n_rows; p.out = d_out; p.bias = d_bias; p.aux_hi = d_aux_hi; p.aux_lo = d_aux_lo;
    p.out_lo = d_out_lo; p.act = act; p.skip = d_skip;
    p.head_w = d_head_w; p.head_b = d_head_b; p.head_out = d_head_out; p.n_head = n_head;
    cudaStream_t s = as_stream(stream);
    ProfScope prof_scope(epilogue == 0 ? "tb_tc_gemm256_fwd" : epilogue == 1 ? "tb_tc_gemm256_bwd"
                                                            : "tb_tc_gemm256", stream);
#define TB_TC(P_, E_) launch_tc<P_, E_>(maps, p, s)
    if (passes == 3) {
        if (epilogue == 0) TB_TC(3, 0); else if (epilogue == 1) TB_TC(3, 1); else TB_TC(3, 2);
    } else {
        if (epilogue == 0) TB_TC(1, 0); else if (epilogue == 1) TB_TC(1, 1); else TB_TC(1, 2);
    }
#undef TB_TC
    return check_launch("tb_tc_gemm256");
}

extern "C" int tb_tc_wgrad256(const float* d_dz_hi, const float* d_dz_lo, const float* d_h_hi,
                              const float* d_h_lo, int64_t n_rows, int32_t passes, float* d_gpart,
                              int32_t n_split, int32_t n_params, int32_t off_w2, int32_t off_b2,
                              const int32_t* d_skip, void* stream) {
    using namespace tb;
    TB_REQUIRE(d_dz_hi && d_h_hi && d_gpart && n_rows > 0 && n_split >= 1, TB_EINVAL,
               "tb_tc_wgrad256: bad arguments");
    TB_REQUIRE(passes == 1 || (passes == 3 && d_dz_lo && d_h_lo), TB_EINVAL,
               "tb_tc_wgrad256: passes must be 1, or 3 with the lo parts");
    CUtensorMap maps[4];
    int rc;
    if ((rc = make_map(&maps[0], d_dz_hi, n_rows, TCW_ROWS, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
    if ((rc = make_map(&maps[1], d_dz_lo ? d_dz_lo : d_dz_hi, n_rows, TCW_ROWS, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
    if ((rc = make_map(&maps[2], d_h_hi, n_rows, TCW_ROWS, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
    if ((rc = make_map(&maps[3], d_h_lo ? d_h_lo : d_h_hi, n_rows, TCW_ROWS, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
    TcWgradParams p;
    p.n_rows = n_rows;
    p.rows_per_split = ((n_rows + n_split - 1) / n_split + TCW_ROWS - 1) / TCW_ROWS * TCW_ROWS;
    p.gpart = d_gpart; p.n_params = n_params; p.off_w2 = off_w2; p.off_b2 = off_b2; p.skip = d_skip;
    auto knob = [](const char* name) { const char* v = getenv(name); return v ? atoi(v) : 0; };
    p.dbg_lbo = knob("TB_TCW_LBO"); p.dbg_sbo = knob("TB_TCW_SBO"); p.dbg_kstep = knob("TB_TCW_KSTEP");
    p.dbg_idesc_xor = knob("TB_TCW_IDESC_XOR");
    dim3 grid(TC_BN / TC_BM, n_split);
    cudaStream_t s = as_stream(stream);
    ProfScope prof_scope("tb_tc_wgrad256", stream);
    if (passes == 3) {
        static bool configured = false;
        if (!configured) {
            cudaFuncSetAttribute(tc_wgrad_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 TcCfg<3>::SMEM_BYTES);
            configured = true;
        }
        tc_wgrad_kernel<3><<<grid, TC_THREADS, TcCfg<3>::SMEM_BYTES, s>>>(maps[0], maps[1], maps[2], maps[3], p);
    } else {
        static bool configured = false;
        if (!configured) {
            cudaFuncSetAttribute(tc_wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 TcCfg<1>::SMEM_BYTES);
            configured = true;
        }
        tc_wgrad_kernel<1><<<grid, TC_THREADS, TcCfg<1>::SMEM_BYTES, s>>>(maps[0], maps[1], maps[2], maps[3], p);
    }
    return check_launch("tb_tc_wgrad256");
}

namespace tb { static unsigned long long* g_wgrad_timeline = nullptr; }

// Profiling aid for tb_mlp_wgrad_fused (like tb_tc_timeline): first call allocates 16 clock64()
// slots that CTA (0, 0) of later launches fills; out16 != NULL reads them back (host pointer).
extern "C" int tb_wgrad_timeline(uint64_t* out16) {
    using namespace tb;
    if (!g_wgrad_timeline) {
        TB_REQUIRE(cudaMalloc(&g_wgrad_timeline, 64 * sizeof(unsigned long long)) == cudaSuccess, TB_ENOTSUP,
                   "tb_wgrad_timeline: cudaMalloc failed");
        cudaMemset(g_wgrad_timeline, 0, 64 * sizeof(unsigned long long));
    }
    if (out16)
        TB_REQUIRE(cudaMemcpy(out16, g_wgrad_timeline, 64 * sizeof(unsigned long long),
                              cudaMemcpyDeviceToHost) == cudaSuccess, TB_ENOTSUP, "tb_wgrad_timeline: copy failed");
    return 0;
}

// All weight gradients of one minibatch in one launch, reduced to the flat gradient
// (see tc_wgrad_all_kernel).  d_sync: one uint64 (zero-initialised once), the grid-barrier
// counter of this network.
extern "C" int tb_mlp_wgrad_fused(const TbMlpShape* shape, const float* d_xin, const float* d_h1_hi,
                                  const float* d_h1_lo, const float* d_h2, const float* d_dz1,
                                  const float* d_dz2_hi, const float* d_dz2_lo, const float* d_dout,
                                  int32_t ld_dout, int32_t n_extra, int32_t off_extra, int64_t n_rows,
                                  float* d_gpart, int32_t n_split, float* d_flat, uint64_t* d_sync,
                                  int32_t passes, const TbAdam* opt, float* d_packed, float grad_scale,
                                  const double* d_stats, float kl_threshold, int32_t* d_stop,
                                  const int32_t* d_skip, const TbPeers* peers, uint64_t* d_epoch,
                                  double* d_reduce_stats, void* stream) {
    using namespace tb;
    TB_REQUIRE(!peers || peers->world <= 1 || (opt && d_epoch && peers->world <= kPeerLanes &&
                                               peers->rank >= 0 && peers->rank < peers->world), TB_EINVAL,
               "tb_mlp_wgrad_fused: the peer exchange needs the fused optimizer step, an epoch counter "
               "and at most %d ranks", kPeerLanes);
    TB_REQUIRE(!opt || (opt->d_params && opt->d_m && opt->d_v && opt->d_step && shape &&
                        opt->n_params == shape->n_params), TB_EINVAL,
               "tb_mlp_wgrad_fused: optimizer / shape mismatch");
    TB_REQUIRE(shape && d_xin && d_h1_hi && d_h2 && d_dz1 && d_dz2_hi && d_dout &&
               d_gpart && d_flat && d_sync && n_rows > 0, TB_EINVAL, "tb_mlp_wgrad_fused: null pointer");
    TB_REQUIRE((d_h1_lo == nullptr) == (d_dz2_lo == nullptr), TB_EINVAL,
               "tb_mlp_wgrad_fused: h1 and dz2 must both be tf32 splits or both plain (lo == NULL)");
    TB_REQUIRE(shape->hidden == 256 && shape->off_w2_hi > 0 && shape->d_in + 1 <= 32 &&
               shape->n_out >= 1 && shape->n_out <= TCA_NO && shape->n_out + n_extra <= TCA_ND &&
               ld_dout >= shape->n_out + n_extra, TB_ENOTSUP,
               "tb_mlp_wgrad_fused: needs hidden == 256, d_in <= 31, n_out <= 8, n_out + extras <= 12");
    TB_REQUIRE(n_split >= 1 && 2 * n_split <= kNumSMs, TB_EINVAL,
               "tb_mlp_wgrad_fused: 2 * n_split CTAs must be resident together (<= %d)", kNumSMs);
    TB_REQUIRE(passes == 1 || passes == 3, TB_EINVAL, "tb_mlp_wgrad_fused: passes must be 1 or 3");
    TB_REQUIRE(shape->off_b2 == shape->off_w2 + 256 * 256, TB_EINVAL,
               "tb_mlp_wgrad_fused: b2 must follow W2 in the flat layout");
    CUtensorMap maps[6];
    int rc;
    if ((rc = make_map(&maps[0], d_dz2_hi, n_rows, TCA_ROWS, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
    if ((rc = make_map(&maps[1], d_dz2_lo ? d_dz2_lo : d_dz2_hi, n_rows, TCA_ROWS, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
    if ((rc = make_map(&maps[2], d_h1_hi, n_rows, TCA_ROWS, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
    if ((rc = make_map(&maps[3], d_h1_lo ? d_h1_lo : d_h1_hi, n_rows, TCA_ROWS, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
    if ((rc = make_map_plain(&maps[4], d_dz1, n_rows, 128, TCA_ROWS))) return rc;
    if ((rc = make_map_plain(&maps[5], d_h2, n_rows, 128, TCA_ROWS))) return rc;
    CUtensorMap maps3[4];
    static const bool want3d = [] { const char* v = getenv("TONIC_B200_WGRAD_TMA3D"); return !(v && v[0] == '0'); }();
    bool tma3d = want3d;
    if (tma3d) {
        tma3d = make_map_groups(&maps3[0], d_dz2_hi, n_rows, TCA_ROWS, TC_BM / 32) == 0 &&
                make_map_groups(&maps3[1], d_dz2_lo ? d_dz2_lo : d_dz2_hi, n_rows, TCA_ROWS, TC_BM / 32) == 0 &&
                make_map_groups(&maps3[2], d_h1_hi, n_rows, TCA_ROWS, TC_BN / 32) == 0 &&
                make_map_groups(&maps3[3], d_h1_lo ? d_h1_lo : d_h1_hi, n_rows, TCA_ROWS, TC_BN / 32) == 0;
    }
    if (!tma3d) for (int i = 0; i < 4; ++i) maps3[i] = maps[i];
    TcWgradAllParams q;
    q.tma3d = tma3d ? 1 : 0;
    { const char* v = getenv("TB_WGRAD_DBG"); q.dbg = v ? atoi(v) : 0; }
    q.w.n_rows = n_rows;
    q.w.rows_per_split = ((n_rows + n_split - 1) / n_split + TCA_ROWS - 1) / TCA_ROWS * TCA_ROWS;
    q.w.gpart = d_gpart; q.w.n_params = shape->n_params; q.w.off_w2 = shape->off_w2;
    q.w.off_b2 = shape->off_b2; q.w.skip = d_skip;
    q.w.dbg_lbo = q.w.dbg_sbo = q.w.dbg_kstep = q.w.dbg_idesc_xor = 0;
    q.sh = *shape; q.xin = d_xin; q.h2 = d_h2; q.dz1 = d_dz1; q.dout = d_dout; q.ld_dout = ld_dout;
    q.n_extra = n_extra; q.off_extra = off_extra; q.n_split = n_split; q.flat = d_flat;
    q.sync = reinterpret_cast<unsigned long long*>(d_sync);
    q.timeline = g_wgrad_timeline;
    q.fuse_adam = opt != nullptr;
    if (opt) q.opt = *opt; else memset(&q.opt, 0, sizeof(q.opt));
    q.packed = d_packed; q.grad_scale = grad_scale; q.stats = d_stats; q.kl_threshold = kl_threshold;
    q.stop = d_stop;
    if (peers && peers->world > 1) q.peers = *peers; else memset(&q.peers, 0, sizeof(q.peers));
    q.epoch = reinterpret_cast<unsigned long long*>(d_epoch);
    q.reduce_stats = d_reduce_stats;
    { const char* v = getenv("TONIC_B200_PEER_TWO_PHASE"); q.two_phase = (v && v[0] == '1') ? 1 : 0; }
    dim3 grid(TC_BN / TC_BM, n_split);
    cudaStream_t s = as_stream(stream);
    ProfScope prof_scope("tb_mlp_wgrad_fused", stream);
    const bool small_in = shape->d_in + 1 <= 20;
    const bool plain = d_h1_lo == nullptr;
    TB_REQUIRE(!plain || passes == 3, TB_EINVAL, "tb_mlp_wgrad_fused: plain activations need passes == 3");
#define TB_WGRAD_ALL(P_, K_, PL_)                                                                    \
    {                                                                                                \
        static bool configured = false;                                                              \
        if (!configured) {                                                                           \
            cudaFuncSetAttribute(tc_wgrad_all_kernel<P_, K_, PL_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                 tca_smem_bytes<P_, K_>());                                          \
            configured = true;                                                                       \
        }                                                                                            \
        tc_wgrad_all_kernel<P_, K_, PL_><<<grid, TCA_THREADS, tca_smem_bytes<P_, K_>(), s>>>(        \
            maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], maps3[0], maps3[1], maps3[2], maps3[3], q); \
    }
    if (plain) { if (small_in) TB_WGRAD_ALL(3, 20, true) else TB_WGRAD_ALL(3, 32, true) }
    else if (passes == 3) { if (small_in) TB_WGRAD_ALL(3, 20, false) else TB_WGRAD_ALL(3, 32, false) }
    else { if (small_in) TB_WGRAD_ALL(1, 20, false) else TB_WGRAD_ALL(1, 32, false) }
#undef TB_WGRAD_ALL
    return check_launch("tb_mlp_wgrad_fused");
}
