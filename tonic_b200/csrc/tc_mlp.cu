// Fused forward pass of the 2 x 256 MLP on the tensor cores (sm_100a): one kernel computes
//
//     x   = gather / normalise / concatenate the input rows        (TbMlpInput)
//     h1  = act(x  W1^T + b1)                                      [rows, 256]
//     h2  = act(h1 W2^T + b2)                                      [rows, 256]
//     out = h2 W3^T + b3                                           [rows, n_out <= 8]
//
// for a 128-row tile per CTA without h1 ever leaving the SM as a matmul operand; reference
// arithmetic: tonic/torch/models/utils.py:15-23 (MLP torso), encoders.py:11-30 (normalise /
// concatenate), critics.py:9-20 / actors.py:37-66 (linear heads).  It replaces the chain
// mlp_layer1_kernel -> tc_gemm_kernel<.., EPI_BIAS_ACT> (csrc/mlp.cu, csrc/tc_gemm.cu); with
// one row tile per CTA those two kernels spent most of their time on launch latency, the
// h1 round trip through global memory and an instruction-bound FFMA first layer.
//
// Roles (384 threads):
//   warp 0      producer (one thread): per tile a 1-D bulk copy of the W1 operand image
//               (written by tb_mlp_pack in the swizzled shared-memory layout) and TMA loads of
//               the W2 chunks (hi / lo tf32 parts, 128B swizzle);
//   warp 1      MMA issuer (one thread): tcgen05.mma kind::tf32, M = 128, N = 256;
//               accumulator 0 (TMEM columns 0..255) = layer 1, accumulator 1 = layer 2;
//   warp 2      TMEM allocation;
//   warps 4-11  "row" warps: thread (q, lane) of column half wg owns row 32 q + lane of the
//               tile and 16 of every 32 columns (two warps share a TMEM lane quarter):
//                 a) layer-1 A operand: gather / normalise the input row, zero-pad to K = 32,
//                    split into tf32 hi / lo parts, store in the K-major 128B-swizzle layout
//                    the MMA descriptor expects (and the input copy the weight gradient needs);
//                 b) mid epilogue, per 32-column chunk c of accumulator 0: tcgen05.ld -> + b1 ->
//                    activation -> hi / lo split -> A operand of layer-2 chunk c in shared
//                    memory (and h1 to global memory when the backward pass needs it);
//                 c) final epilogue on accumulator 1: + b2 -> activation -> fused linear head
//                    (-> h2 to global memory when the backward pass needs it).
// Shared-memory stages hold [A hi | A lo | B hi | B lo] = 96 KB (3xTF32) exactly as in
// tc_gemm_kernel; per tile they are used 9 times: once for layer 1 and 8 times for the K
// chunks of layer 2 (A by the row warps, B by the producer).  Measured on B200 with the
// clock64() timeline below (16384 rows, one tile per CTA): the row warps are the critical
// path; the activation stores of the training forward pass (h1 hi / lo, h2: 48 MB) run at
// the chip's L2 write bandwidth.
#include "tc_common.cuh"

namespace tb {

struct TcMlpParams {
    int64_t n_rows;
    TbMlpInput in;
    int d_in;                   // <= 32
    int act;
    const float* w1_img_hi;     // W1 as the layer-1 B operand: byte image of the swizzled
    const float* w1_img_lo;     // [256 x 32] K-major tile (hi / lo), written by tb_mlp_pack
    const float* b1;
    const float* b2;
    float* xin_save;            // [n_rows, ldx] with the ones column, or NULL
    float* h1_hi;               // [n_rows, 256] tf32 split of h1, or NULL (no backward pass)
    float* h1_lo;
    float* h2;                  // [n_rows, 256] or NULL
    const float* head_w;        // [n_head, 256]
    const float* head_b;
    float* head_out;            // [n_rows, n_head]
    int n_head;                 // 0..8
    const int32_t* skip;
    unsigned long long* timeline;   // profiling aid: 64 clock64() stamps of CTA 0, or NULL
    // optional fused squared-error loss on the single head output (value regression,
    // torch/updaters/critics.py:18-28): dout = 2 (v - target), sums into the statistics block
    const float* loss_targets;      // [.] indexed by loss_idx[row] (or row), or NULL
    const int64_t* loss_idx;
    float* loss_dout;               // [n_rows, loss_ld]
    int loss_ld;
    double* loss_stats;             // TB_STAT_* block
    int loss_stat_slot;             // slot receiving the sum of values
    int loss_count_rows;
};

constexpr int TCM_ROW_WARPS = 8;                 // two per TMEM lane quarter (column halves)
constexpr int TCM_THREADS = (4 + TCM_ROW_WARPS) * 32;
constexpr int TCM_STG_STRIDE = 20;               // floats per staged row (16 + pad)

template <int PASSES>
struct TcMlpCfg {
    using G = TcCfg<PASSES>;
    static constexpr int STAGES = 2;
    static constexpr int STAGE_BYTES = G::STAGE_BYTES;
    static constexpr int A_LO = TC_A_BYTES;                              // offset of A lo (3 passes)
    static constexpr int B_HI = G::PARTS * TC_A_BYTES;
    static constexpr int B_LO = 2 * TC_A_BYTES + TC_B_BYTES;
    static constexpr int EPI_BYTES = TCM_ROW_WARPS * 32 * TCM_STG_STRIDE * 4;
    static constexpr int CONST_BYTES = (TC_MAX_HEAD + 2) * TC_BN * 4;    // head weights, b1, b2
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + CONST_BYTES + 1024 + 256;
};

__device__ __forceinline__ float tc_tf32_hi(float x) {
    return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
}
// tcgen05 kind::tf32 ignores the 13 low mantissa bits of its operands (scratch/
// probe_tf32_truncation.py: bit-identical results with masked and unmasked "hi" tiles on B200).
// `plain` = the hi tile keeps the PLAIN float32 value (lo stays x - trunc(x)): the tile is then
// both the MMA operand and, stored by TMA, the saved activation itself -- no separate lo array.
__device__ int g_plain_hi = 0;        // tb_debug_plain_hi: force plain hi tiles everywhere (hardware check)

template <int ACT>
__device__ __forceinline__ float tc_act(float x) {
    return ACT == TB_ACT_TANH ? tanh_fast(x) : fmaxf(x, 0.0f);
}

__device__ __forceinline__ void tc_stamp(unsigned long long* timeline, int slot) {
    if (timeline && blockIdx.x == 0) timeline[slot] = (unsigned long long)clock64();
}

// 16 consecutive 32-bit TMEM columns of this warp's 32 lanes -> 16 registers per thread
__device__ __forceinline__ void tcgen05_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 1-D bulk copy global -> shared memory, completing on an mbarrier (no tensor map)
__device__ __forceinline__ void bulk_load(void* smem, const void* gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(smem)), "l"(gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// TMA store of a [128 x 32]-float swizzled shared-memory tile (bulk async-group completion)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(smem)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the N most recent store groups of this thread have finished READING shared memory
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// barrier among the 4 warps of one row-warp group
__device__ __forceinline__ void group_sync(int wg) {
    asm volatile("bar.sync %0, 128;" ::"r"(2 + wg) : "memory");
}

// 32 consecutive float32 values of one row into a swizzled [128 x 32] staging tile (TMA store source)
__device__ __forceinline__ void store_plain_row(unsigned char* tile, int row, const float (&v)[32]) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
        *reinterpret_cast<float4*>(tile + sw128_offset(row, u)) =
            make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
}

// barrier among the 8 row warps only
__device__ __forceinline__ void row_warps_sync() {
    asm volatile("bar.sync 1, 256;" ::: "memory");
}

// write 32 consecutive K values of one operand row (as hi / lo parts) into swizzled tiles
template <int PASSES>
__device__ __forceinline__ void store_operand_row(unsigned char* hi_tile, unsigned char* lo_tile, int row,
                                                  const float (&v)[32], bool plain = false) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const uint32_t off = sw128_offset(row, u);
        float4 h;
        h.x = tc_tf32_hi(v[4 * u]); h.y = tc_tf32_hi(v[4 * u + 1]);
        h.z = tc_tf32_hi(v[4 * u + 2]); h.w = tc_tf32_hi(v[4 * u + 3]);
        *reinterpret_cast<float4*>(hi_tile + off) =
            (plain || g_plain_hi) ? make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]) : h;
        if (PASSES == 3) {
            float4 l;
            l.x = v[4 * u] - h.x; l.y = v[4 * u + 1] - h.y;
            l.z = v[4 * u + 2] - h.z; l.w = v[4 * u + 3] - h.w;
            *reinterpret_cast<float4*>(lo_tile + off) = l;
        }
    }
}

// write 16 consecutive K values (units 4 half .. 4 half + 3) of one operand row
template <int PASSES>
__device__ __forceinline__ void store_operand_half(unsigned char* hi_tile, unsigned char* lo_tile, int row,
                                                   int half, const float (&v)[16], bool plain = false) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const uint32_t off = sw128_offset(row, 4 * half + u);
        float4 h;
        h.x = tc_tf32_hi(v[4 * u]); h.y = tc_tf32_hi(v[4 * u + 1]);
        h.z = tc_tf32_hi(v[4 * u + 2]); h.w = tc_tf32_hi(v[4 * u + 3]);
        *reinterpret_cast<float4*>(hi_tile + off) =
            (plain || g_plain_hi) ? make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]) : h;
        if (PASSES == 3) {
            float4 l;
            l.x = v[4 * u] - h.x; l.y = v[4 * u + 1] - h.y;
            l.z = v[4 * u + 2] - h.z; l.w = v[4 * u + 3] - h.w;
            *reinterpret_cast<float4*>(lo_tile + off) = l;
        }
    }
}

// CLUSTER = 2 (experimental, TONIC_B200_CLUSTER=2): the two CTAs of a thread-block cluster each
// fetch half of every weight-operand chunk and multicast it to both, which halves the L2 -> SM
// traffic of the operand that bounds this kernel; the shared-memory stages are then released
// by both MMA threads (multicast commit).
template <int PASSES, int ACT, int CLUSTER>
__global__ void __launch_bounds__(TCM_THREADS, 1)
tc_mlp_forward_kernel(const __grid_constant__ CUtensorMap map_b_hi,
                      const __grid_constant__ CUtensorMap map_b_lo,
                      const __grid_constant__ CUtensorMap map_h1_hi,
                      const __grid_constant__ CUtensorMap map_h1_lo,
                      const __grid_constant__ CUtensorMap map_h2, const TcMlpParams p) {
    using Cfg = TcMlpCfg<PASSES>;
    if (skip_requested(p.skip)) return;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* epi = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
    float* s_head_w = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::EPI_BYTES);
    float* s_b1 = s_head_w + TC_MAX_HEAD * TC_BN;
    float* s_b2 = s_b1 + TC_BN;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::EPI_BYTES +
                                                 Cfg::CONST_BYTES);
    uint64_t* full_bar = bars;                        // [STAGES] B operand landed (TMA / bulk copy)
    uint64_t* a_ready = bars + Cfg::STAGES;           // [STAGES] A operand written by the row warps
    uint64_t* empty_bar = bars + 2 * Cfg::STAGES;     // [STAGES] MMAs reading the stage retired
    uint64_t* acc_full = bars + 3 * Cfg::STAGES;      // [2] accumulator complete
    uint64_t* x_ready = acc_full + 2;                 // layer-1 A operand written (all row warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_ready + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (int)((p.n_rows + TC_BM - 1) / TC_BM);
    if (threadIdx.x == 0) tc_stamp(p.timeline, 0);
    constexpr int USES = 1 + TC_K / TC_BK;            // stage uses per tile: layer 1 + 8 chunks

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&a_ready[s], TCM_ROW_WARPS / 2);
            mbar_init(&empty_bar[s], CLUSTER);      // released by the MMA thread of every CTA that multicasts into it
        }
        mbar_init(&acc_full[0], 1);
        mbar_init(&acc_full[1], 1);
        mbar_init(x_ready, TCM_ROW_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < TC_BN; i += TCM_THREADS) {
        s_b1[i] = p.b1[i];
        s_b2[i] = p.b2[i];
    }
    for (int i = threadIdx.x; i < p.n_head * TC_BN; i += TCM_THREADS) s_head_w[i] = p.head_w[i];
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    if (CLUSTER == 2) cluster_sync_all();            // the peer's barriers exist before anything targets them
    const int crank = CLUSTER == 2 ? (int)(blockIdx.x & 1) : 0;
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) tc_stamp(p.timeline, 1);
    const int k_steps1 = (p.d_in + 7) >> 3;           // layer-1 MMAs (K = 8 each) per pass

    if (warp == 0) {
        // ===================== producer: W1 operand image (bulk copy), W2 chunks (TMA) ==========
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; (CLUSTER == 2 ? (tile & ~1) : tile) < n_tiles; tile += gridDim.x) {
                for (int u = 0; u < USES; ++u) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    unsigned char* st = smem + stage * Cfg::STAGE_BYTES;
                    mbar_expect_tx(&full_bar[stage], (PASSES == 3 ? 2 : 1) * TC_B_BYTES);
                    if (CLUSTER == 2) {
                        // this CTA fetches rows [128 crank, 128 crank + 128) of the operand (half of
                        // the bytes) and multicasts them; the other half arrives from the peer
                        const int half = crank * (TC_B_BYTES / 2);
                        if (u == 0) {
                            bulk_load_multicast(st + Cfg::B_HI + half, reinterpret_cast<const unsigned char*>(p.w1_img_hi) + half,
                                                TC_B_BYTES / 2, &full_bar[stage], 3);
                            if (PASSES == 3)
                                bulk_load_multicast(st + Cfg::B_LO + half, reinterpret_cast<const unsigned char*>(p.w1_img_lo) + half,
                                                    TC_B_BYTES / 2, &full_bar[stage], 3);
                        } else {
                            const int c = u - 1;
                            tma_load_2d_multicast(st + Cfg::B_HI + half, &map_b_hi, &full_bar[stage], c * TC_BK,
                                                  crank * (TC_BN / 2), 3);
                            if (PASSES == 3)
                                tma_load_2d_multicast(st + Cfg::B_LO + half, &map_b_lo, &full_bar[stage], c * TC_BK,
                                                      crank * (TC_BN / 2), 3);
                        }
                    } else if (u == 0) {
                        bulk_load(st + Cfg::B_HI, p.w1_img_hi, TC_B_BYTES, &full_bar[stage]);
                        if (PASSES == 3) bulk_load(st + Cfg::B_LO, p.w1_img_lo, TC_B_BYTES, &full_bar[stage]);
                    } else {
                        const int c = u - 1;
                        tma_load_2d(st + Cfg::B_HI, &map_b_hi, &full_bar[stage], c * TC_BK, 0);
                        if (PASSES == 3)
                            tma_load_2d(st + Cfg::B_LO, &map_b_lo, &full_bar[stage], c * TC_BK, 0);
                    }
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int stage = 0, it = 0;
            uint32_t phase = 0;
            uint32_t a_phase = 0;                     // bit s: parity of a_ready[s] (layer-2 uses only)
            for (int tile = blockIdx.x; (CLUSTER == 2 ? (tile & ~1) : tile) < n_tiles; tile += gridDim.x, ++it) {
                for (int u = 0; u < USES; ++u) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_stamp(p.timeline, 40 + u);             // B operand landed
                    if (u == 0) mbar_wait(x_ready, it & 1);
                    else mbar_wait(&a_ready[stage], (a_phase >> stage) & 1u), a_phase ^= 1u << stage;
                    tc_stamp(p.timeline, 50 + u);             // A operand ready -> MMAs issued
                    tcgen05_fence_after();
                    unsigned char* st = smem + stage * Cfg::STAGE_BYTES;
                    const uint64_t a_hi = umma_desc_kmajor_sw128(st);
                    const uint64_t a_lo = umma_desc_kmajor_sw128(st + Cfg::A_LO);
                    const uint64_t b_hi = umma_desc_kmajor_sw128(st + Cfg::B_HI);
                    const uint64_t b_lo = umma_desc_kmajor_sw128(st + Cfg::B_LO);
                    const uint32_t d_tmem = tmem_base + (u == 0 ? 0u : (uint32_t)TC_BN);
                    const int ks = u == 0 ? k_steps1 : TC_BK / 8;
                    const bool fresh = u <= 1;                      // first chunk of an accumulator
                    for (int k = 0; k < ks; ++k) {
                        const uint64_t koff = (uint64_t)(k * 32 >> 4);
                        const uint32_t accumulate = (fresh && k == 0) ? 0u : 1u;
                        if (PASSES == 3) {
                            tcgen05_mma_tf32(d_tmem, a_lo + koff, b_hi + koff, kIdescTf32, accumulate);
                            tcgen05_mma_tf32(d_tmem, a_hi + koff, b_lo + koff, kIdescTf32, 1);
                            tcgen05_mma_tf32(d_tmem, a_hi + koff, b_hi + koff, kIdescTf32, 1);
                        } else {
                            tcgen05_mma_tf32(d_tmem, a_hi + koff, b_hi + koff, kIdescTf32, accumulate);
                        }
                    }
                    if (CLUSTER == 2) tcgen05_commit_multicast(&empty_bar[stage], 3);
                    else tcgen05_commit(&empty_bar[stage]);
                    if (u == 0) tcgen05_commit(&acc_full[0]);
                    if (u == USES - 1) tcgen05_commit(&acc_full[1]);
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== row warps =====================
        // warp 4 + rw: TMEM lane quarter q = rw % 4 (== warp % 4), group wg = rw / 4.  Thread
        // (q, lane) owns row 32 q + lane of the tile.  The input row is split by columns between
        // the two groups (16 each); the 32-column chunks of both epilogues alternate between
        // them (group 0 even chunks, group 1 odd chunks), so that on every SM sub-partition one
        // warp computes while the other waits on TMEM / shared-memory / barrier latency.
        const int rw = warp - 4, q = rw & 3, wg = rw >> 2;
        const int trow = q * 32 + lane;
        const bool stamper = rw == 0 && lane == 0;
        const bool issuer = q == 0 && lane == 0;            // this group's TMA-store thread
        float* stg = epi + rw * 32 * TCM_STG_STRIDE;
        const int d_in = p.d_in;
        const int ldx = (d_in + 1 + 3) & ~3;
        // h1 for the backward pass: as the tf32 split (two arrays, TMA stores of the operand tiles)
        // or, with h1_lo == NULL, as ONE plain float32 array written straight from registers (the
        // fused weight-gradient kernel splits it on the fly: half the activation traffic)
        const bool save_h1 = p.h1_hi != nullptr, save_h2 = p.h2 != nullptr;
        const bool plain_h1 = p.h1_hi != nullptr && p.h1_lo == nullptr;
        const uint32_t t_lane = (uint32_t)(q * 32) << 16;
        int it = 0;
        if (stamper) tc_stamp(p.timeline, 31);
        for (int tile = blockIdx.x; (CLUSTER == 2 ? (tile & ~1) : tile) < n_tiles; tile += gridDim.x, ++it) {
            const int64_t row0 = (int64_t)tile * TC_BM + q * 32;
            const int64_t row = (int64_t)tile * TC_BM + trow;
            const int g0 = it * USES;                 // stage use index of this tile's layer 1
            // ---- a) layer-1 A operand: this thread's 16 input columns ---------------------------
            {
                float xv[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) xv[j] = 0.0f;
                if (row < p.n_rows) {
                    const int64_t r = p.in.d_idx ? p.in.d_idx[row] : row;
                    const int64_t r2 = p.in.gather2 ? r : row;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int c = 16 * wg + j;
                        if (c < d_in) {
                            float val;
                            if (c < p.in.dim1) {
                                val = p.in.d_x1[r * p.in.dim1 + c];
                                if (p.in.d_mean)      // mean_stds.py:36  (val - mean) / std
                                    val = __fdiv_rn(__fsub_rn(val, p.in.d_mean[c]), p.in.d_std[c]);
                            } else {
                                val = p.in.d_x2[r2 * p.in.dim2 + (c - p.in.dim1)];
                            }
                            xv[j] = val;
                        }
                    }
                    if (p.xin_save) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int c = 16 * wg + j;
                            if (c < ldx) p.xin_save[row * ldx + c] = c < d_in ? xv[j] : (c == d_in ? 1.0f : 0.0f);
                        }
                        if (wg == 1)
                            for (int c = 32; c < ldx; ++c) p.xin_save[row * ldx + c] = c == d_in ? 1.0f : 0.0f;
                    }
                }
                const int stage = g0 & 1;
                mbar_wait(&empty_bar[stage], ((g0 >> 1) & 1) ^ 1);
                unsigned char* st = smem + stage * Cfg::STAGE_BYTES;
                store_operand_half<PASSES>(st, st + Cfg::A_LO, trow, wg, xv);
                if (stamper) tc_stamp(p.timeline, 32);        // input row loaded and stored
                fence_proxy_async_smem();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(x_ready);
                if (stamper) tc_stamp(p.timeline, 2);         // layer-1 A operand published
            }
            // ---- b) mid epilogue: h1 chunks -> layer-2 A operand -------------------------------
            mbar_wait(&acc_full[0], it & 1);
            tcgen05_fence_after();
            if (stamper) tc_stamp(p.timeline, 3);             // layer-1 accumulator complete
#pragma unroll 1
            for (int c = wg; c < TC_K / TC_BK; c += 2) {
                uint32_t v[32];
                tcgen05_ld_32x32(tmem_base + t_lane + (uint32_t)(c * 32), v);
                float hv[32];
                const float4* b4 = reinterpret_cast<const float4*>(s_b1 + c * 32);
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b = b4[j / 4];
                    hv[j] = tc_act<ACT>(__uint_as_float(v[j]) + b.x);
                    hv[j + 1] = tc_act<ACT>(__uint_as_float(v[j + 1]) + b.y);
                    hv[j + 2] = tc_act<ACT>(__uint_as_float(v[j + 2]) + b.z);
                    hv[j + 3] = tc_act<ACT>(__uint_as_float(v[j + 3]) + b.w);
                }
                if (stamper) tc_stamp(p.timeline, 4 + c);     // chunk computed
                const int g = g0 + 1 + c, stage = g & 1;
                mbar_wait(&empty_bar[stage], ((g >> 1) & 1) ^ 1);
                if (stamper) tc_stamp(p.timeline, 12 + c);    // stage free
                unsigned char* st = smem + stage * Cfg::STAGE_BYTES;
                if (save_h1) {        // the TMA stores of this group's previous chunk have left the stage
                    if (issuer) bulk_wait_read<0>();
                    group_sync(wg);
                }
                store_operand_row<PASSES>(st, st + Cfg::A_LO, trow, hv, plain_h1);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_ready[stage]);
                if (stamper) tc_stamp(p.timeline, 20 + c);    // operand published
                if (save_h1) {
                    // saved activations for backward: the operand tile IS the tf32 split of h1 in
                    // the layout TMA expects -> one thread stores it asynchronously (the MMA and
                    // the store only read the tile; this group rewrites it two chunks later)
                    group_sync(wg);
                    if (issuer) {
                        tma_store_2d(&map_h1_hi, st, c * TC_BK, tile * TC_BM);
                        if (PASSES == 3 && !plain_h1) tma_store_2d(&map_h1_lo, st + Cfg::A_LO, c * TC_BK, tile * TC_BM);
                        bulk_commit();
                    }
                }
            }
            // ---- c) final epilogue: h2, fused head ----------------------------------------------
            if (stamper) tc_stamp(p.timeline, 28);            // mid epilogue (+ h1 stores) done
            mbar_wait(&acc_full[1], it & 1);
            tcgen05_fence_after();
            if (stamper) tc_stamp(p.timeline, 29);            // layer-2 accumulator complete
            float hacc[TC_MAX_HEAD];
#pragma unroll
            for (int o = 0; o < TC_MAX_HEAD; ++o) hacc[o] = 0.0f;
            unsigned char* own_stage = smem + ((g0 + 1 + wg) & 1) * Cfg::STAGE_BYTES;
            int n_staged = 0;
            if (save_h2 && save_h1) {      // the h1 stores have left this group's operand buffers
                if (issuer) bulk_wait_read<0>();
                group_sync(wg);
            }
#pragma unroll 1
            for (int c = wg; c < TC_BN / 32; c += 2) {
                uint32_t v[32];
                tcgen05_ld_32x32(tmem_base + t_lane + (uint32_t)(TC_BN + c * 32), v);
                float hv[32];
                const float4* b4 = reinterpret_cast<const float4*>(s_b2 + c * 32);
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b = b4[j / 4];
                    hv[j] = tc_act<ACT>(__uint_as_float(v[j]) + b.x);
                    hv[j + 1] = tc_act<ACT>(__uint_as_float(v[j + 1]) + b.y);
                    hv[j + 2] = tc_act<ACT>(__uint_as_float(v[j + 2]) + b.z);
                    hv[j + 3] = tc_act<ACT>(__uint_as_float(v[j + 3]) + b.w);
                }
#pragma unroll
                for (int o = 0; o < TC_MAX_HEAD; ++o) {
                    if (o < p.n_head) {
                        const float4* w4 = reinterpret_cast<const float4*>(s_head_w + o * TC_BN + c * 32);
                        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;     // independent chains
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 wv = w4[j / 4];
                            s0 = fmaf(hv[j], wv.x, s0);
                            s1 = fmaf(hv[j + 1], wv.y, s1);
                            s2 = fmaf(hv[j + 2], wv.z, s2);
                            s3 = fmaf(hv[j + 3], wv.w, s3);
                        }
                        hacc[o] += (s0 + s1) + (s2 + s3);
                    }
                }
                if (save_h2) {
                    // h2 for the backward pass: staged in this group's (now idle) operand buffers
                    // in the TMA layout, stored asynchronously
                    constexpr int NBUF = PASSES == 3 ? 2 : 1;
                    unsigned char* buf = own_stage + (n_staged % NBUF) * TC_A_BYTES;
                    if (n_staged >= NBUF) {
                        if (issuer) bulk_wait_read<NBUF - 1>();
                        group_sync(wg);
                    }
                    store_plain_row(buf, trow, hv);
                    fence_proxy_async_smem();
                    group_sync(wg);
                    if (issuer) {
                        tma_store_2d(&map_h2, buf, c * TC_BK, tile * TC_BM);
                        bulk_commit();
                    }
                    ++n_staged;
                }
            }
            // the two groups' partial head sums of a row meet in shared memory: group 1 publishes
            // its sums in its staging block, group 0 adds them (fixed order) and writes the output
            if (p.n_head > 0) {
                if (wg == 1) {
#pragma unroll
                    for (int o = 0; o < TC_MAX_HEAD; ++o) stg[lane * TCM_STG_STRIDE + o] = hacc[o];
                }
                row_warps_sync();
                double st_loss = 0.0, st_val = 0.0, st_rows = 0.0;
                if (wg == 0 && row < p.n_rows) {
                    const float* other = epi + (rw + 4) * 32 * TCM_STG_STRIDE + lane * TCM_STG_STRIDE;
#pragma unroll
                    for (int o = 0; o < TC_MAX_HEAD; ++o)
                        if (o < p.n_head)
                            p.head_out[row * p.n_head + o] = (hacc[o] + other[o]) + __ldg(p.head_b + o);
                    if (p.loss_targets) {
                        // same arithmetic as mse_loss_kernel (csrc/heads.cu)
                        const float v = (hacc[0] + other[0]) + __ldg(p.head_b);
                        const float t = p.loss_targets[p.loss_idx ? p.loss_idx[row] : row];
                        const float d = v - t;
                        p.loss_dout[row * p.loss_ld] = 2.0f * d;
                        st_loss = (double)d * (double)d;
                        st_val = v;
                        st_rows = 1.0;
                    }
                }
                if (p.loss_targets && wg == 0) {
                    st_loss = warp_sum(st_loss);
                    st_val = warp_sum(st_val);
                    st_rows = warp_sum(st_rows);
                    if (lane == 0) {       // group 0's staging blocks are idle here
                        double* mine = reinterpret_cast<double*>(stg);
                        mine[0] = st_loss; mine[1] = st_val; mine[2] = st_rows;
                    }
                }
                row_warps_sync();
                if (p.loss_targets && rw == 0 && lane == 0) {
                    double tl = 0.0, tv = 0.0, tr = 0.0;
                    for (int k = 0; k < 4; ++k) {          // fixed order over the 4 warps of group 0
                        const double* part = reinterpret_cast<const double*>(epi + k * 32 * TCM_STG_STRIDE);
                        tl += part[0]; tv += part[1]; tr += part[2];
                    }
                    atomicAdd(&p.loss_stats[TB_STAT_LOSS], tl);
                    atomicAdd(&p.loss_stats[p.loss_stat_slot], tv);
                    if (p.loss_count_rows) atomicAdd(&p.loss_stats[TB_STAT_ROWS], tr);
                }
            }
            if (stamper) tc_stamp(p.timeline, 30);            // final epilogue done
            // TMEM reads of this tile are complete before the next tile's MMAs are released
            // (they wait for x_ready of the next tile, arrived after this fence); the barrier
            // keeps the next tile's input rows out of a stage the other group still reads back
            tcgen05_fence_before();
            if ((save_h1 || save_h2) && tile + (int)gridDim.x < n_tiles) {
                // the next tile's input rows go into operand buffers the stores may still read
                if (issuer) bulk_wait_read<0>();
                row_warps_sync();
            }
        }
        if (issuer) bulk_wait_all();
    }
    tcgen05_fence_before();
    __syncthreads();
    if (CLUSTER == 2) cluster_sync_all();            // no CTA leaves while the peer may still write into it
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512)
                     : "memory");
    }
}

// =====================================================================================
// Fused backward pass (activation gradients) on the tensor cores:
//
//     dz2 = (dout W3) * act'(h2)            [rows, 256]   (row warps, from global h2)
//     dz1 = (dz2 W2)  * act'(h1)            [rows, 256]   (tcgen05 GEMM + epilogue)
//
// i.e. the autograd of the head and of the second hidden layer (reference: loss.backward()
// in torch/updaters/actors.py:33,104,186 / critics.py:24,84 through models/utils.py:15-23).
// It replaces mlp_head_backward_kernel -> tc_gemm_kernel<.., EPI_ACT_GRAD>; dz2 (as a tf32
// split) and dz1 still go to global memory because the weight-gradient kernels consume them.
// Same structure as the forward kernel: the row warps build the A operand chunk by chunk
// (even chunks group 0, odd chunks group 1), the producer streams W2^T chunks by TMA, one
// thread issues the MMAs, and the row warps run the epilogue.
// =====================================================================================
struct TcMlpBwdParams {
    int64_t n_rows;
    const float* dout;          // [n_rows, ld_dout], columns 0..n_head-1
    int ld_dout;
    int n_head;                 // <= 8
    const float* head_w;        // W3 [n_head, 256]
    const float* h2;            // [n_rows, 256]
    const float* h1_hi;         // [n_rows, 256] tf32 split of h1
    const float* h1_lo;
    float* dz2_hi;              // [n_rows, 256] tf32 split of dz2
    float* dz2_lo;
    float* dz1;                 // [n_rows, 256]
    const int32_t* skip;
    unsigned long long* timeline;   // profiling aid (see TcMlpParams)
};

template <int ACT>
__device__ __forceinline__ float tc_act_grad(float h) {     // act'(z) written in terms of h = act(z)
    return ACT == TB_ACT_TANH ? (1.0f - h * h) : (h > 0.0f ? 1.0f : 0.0f);
}

template <int PASSES, int ACT, int CLUSTER>      // CLUSTER: see tc_mlp_forward_kernel
__global__ void __launch_bounds__(TCM_THREADS, 1)
tc_mlp_backward_kernel(const __grid_constant__ CUtensorMap map_b_hi,
                       const __grid_constant__ CUtensorMap map_b_lo,
                       const __grid_constant__ CUtensorMap map_dz2_hi,
                       const __grid_constant__ CUtensorMap map_dz2_lo,
                       const __grid_constant__ CUtensorMap map_dz1, const TcMlpBwdParams p) {
    using Cfg = TcMlpCfg<PASSES>;
    if (skip_requested(p.skip)) return;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* epi = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
    float* s_head_w = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::EPI_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::EPI_BYTES +
                                                 Cfg::CONST_BYTES);
    uint64_t* full_bar = bars;                        // [STAGES] W2^T chunk landed
    uint64_t* a_ready = bars + Cfg::STAGES;           // [STAGES] dz2 chunk written by a row-warp group
    uint64_t* empty_bar = bars + 2 * Cfg::STAGES;     // [STAGES] MMAs reading the stage retired
    uint64_t* acc_full = bars + 3 * Cfg::STAGES;      // accumulator complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (int)((p.n_rows + TC_BM - 1) / TC_BM);
    if (threadIdx.x == 0) tc_stamp(p.timeline, 0);
    constexpr int CHUNKS = TC_K / TC_BK;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&a_ready[s], TCM_ROW_WARPS / 2);
            mbar_init(&empty_bar[s], CLUSTER);
        }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(tmem_slot)), "n"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < TC_MAX_HEAD * TC_BN; i += TCM_THREADS)
        s_head_w[i] = i < p.n_head * TC_BN ? p.head_w[i] : 0.0f;
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    if (CLUSTER == 2) cluster_sync_all();
    const int crank = CLUSTER == 2 ? (int)(blockIdx.x & 1) : 0;
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) tc_stamp(p.timeline, 1);

    if (warp == 0) {
        // ===================== producer: W2^T chunks (TMA) =====================
        if (lane == 0) {
            int g = 0;
            for (int tile = blockIdx.x; (CLUSTER == 2 ? (tile & ~1) : tile) < n_tiles; tile += gridDim.x) {
                for (int c = 0; c < CHUNKS; ++c, ++g) {
                    const int stage = g & 1;
                    mbar_wait(&empty_bar[stage], ((g >> 1) & 1) ^ 1);
                    unsigned char* st = smem + stage * Cfg::STAGE_BYTES;
                    mbar_expect_tx(&full_bar[stage], (PASSES == 3 ? 2 : 1) * TC_B_BYTES);
                    if (CLUSTER == 2) {      // half of the rows each, multicast to both CTAs
                        const int half = crank * (TC_B_BYTES / 2);
                        tma_load_2d_multicast(st + Cfg::B_HI + half, &map_b_hi, &full_bar[stage], c * TC_BK,
                                              crank * (TC_BN / 2), 3);
                        if (PASSES == 3)
                            tma_load_2d_multicast(st + Cfg::B_LO + half, &map_b_lo, &full_bar[stage], c * TC_BK,
                                                  crank * (TC_BN / 2), 3);
                    } else {
                        tma_load_2d(st + Cfg::B_HI, &map_b_hi, &full_bar[stage], c * TC_BK, 0);
                        if (PASSES == 3) tma_load_2d(st + Cfg::B_LO, &map_b_lo, &full_bar[stage], c * TC_BK, 0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int g = 0;
            for (int tile = blockIdx.x; (CLUSTER == 2 ? (tile & ~1) : tile) < n_tiles; tile += gridDim.x) {
                for (int c = 0; c < CHUNKS; ++c, ++g) {
                    const int stage = g & 1;
                    const uint32_t parity = (g >> 1) & 1;
                    mbar_wait(&full_bar[stage], parity);
                    tc_stamp(p.timeline, 40 + c);
                    mbar_wait(&a_ready[stage], parity);
                    tc_stamp(p.timeline, 50 + c);
                    tcgen05_fence_after();
                    unsigned char* st = smem + stage * Cfg::STAGE_BYTES;
                    const uint64_t a_hi = umma_desc_kmajor_sw128(st);
                    const uint64_t a_lo = umma_desc_kmajor_sw128(st + Cfg::A_LO);
                    const uint64_t b_hi = umma_desc_kmajor_sw128(st + Cfg::B_HI);
                    const uint64_t b_lo = umma_desc_kmajor_sw128(st + Cfg::B_LO);
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k) {
                        const uint64_t koff = (uint64_t)(k * 32 >> 4);
                        const uint32_t accumulate = (c | k) != 0;
                        if (PASSES == 3) {
                            tcgen05_mma_tf32(tmem_base, a_lo + koff, b_hi + koff, kIdescTf32, accumulate);
                            tcgen05_mma_tf32(tmem_base, a_hi + koff, b_lo + koff, kIdescTf32, 1);
                            tcgen05_mma_tf32(tmem_base, a_hi + koff, b_hi + koff, kIdescTf32, 1);
                        } else {
                            tcgen05_mma_tf32(tmem_base, a_hi + koff, b_hi + koff, kIdescTf32, accumulate);
                        }
                    }
                    if (CLUSTER == 2) tcgen05_commit_multicast(&empty_bar[stage], 3);
                    else tcgen05_commit(&empty_bar[stage]);
                    if (c == CHUNKS - 1) tcgen05_commit(acc_full);
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== row warps =====================
        const int rw = warp - 4, q = rw & 3, wg = rw >> 2;
        const int trow = q * 32 + lane;
        float* stg = epi + rw * 32 * TCM_STG_STRIDE;
        const bool issuer = q == 0 && lane == 0;            // this group's TMA-store thread
        const uint32_t t_lane = (uint32_t)(q * 32) << 16;
        const bool stamper = (rw == 0 || rw == 4) && lane == 0;
        const bool plain_dz2 = p.dz2_lo == nullptr;
        int it = 0;
        for (int tile = blockIdx.x; (CLUSTER == 2 ? (tile & ~1) : tile) < n_tiles; tile += gridDim.x, ++it) {
            const int64_t row0 = (int64_t)tile * TC_BM + q * 32;
            const int64_t row = (int64_t)tile * TC_BM + trow;
            const bool live = row < p.n_rows;
            float dv[TC_MAX_HEAD];
#pragma unroll
            for (int o = 0; o < TC_MAX_HEAD; ++o)
                dv[o] = (live && o < p.n_head) ? __ldg(p.dout + row * p.ld_dout + o) : 0.0f;
            // ---- a) dz2 chunks: A operand of the GEMM + tf32 split to global memory -------------
            // h2 is read with coalesced 64-byte row segments (lane -> 2 rows x 16 columns; row-wise
            // 16-byte loads would pull every sector through the small L1 twice) and transposed to
            // the row-per-thread layout through the warp's staging block; the loads of the group's
            // next chunk are in flight while the current one is computed and published
            float z[32], t[32];
            auto load_h2 = [&](int c) {
#pragma unroll
                for (int half = 0; half < 2; ++half)
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int64_t r = row0 + 2 * i + (lane >> 4);
                        t[16 * half + i] = r < p.n_rows
                            ? __ldg(p.h2 + r * TC_BN + c * 32 + half * 16 + (lane & 15)) : 0.0f;
                    }
            };
            load_h2(wg);
#pragma unroll 1
            for (int c = wg; c < CHUNKS; c += 2) {
#pragma unroll
                for (int half = 0; half < 2; ++half) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        stg[(2 * i + (lane >> 4)) * TCM_STG_STRIDE + (lane & 15)] = t[16 * half + i];
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 h = *reinterpret_cast<const float4*>(stg + lane * TCM_STG_STRIDE + j);
                        z[16 * half + j] = h.x; z[16 * half + j + 1] = h.y;
                        z[16 * half + j + 2] = h.z; z[16 * half + j + 3] = h.w;
                    }
                    __syncwarp();
                }
                if (c + 2 < CHUNKS) load_h2(c + 2);
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                    for (int o = 0; o < TC_MAX_HEAD; ++o) {
                        if (o < p.n_head) {
                            const float4 wv = *reinterpret_cast<const float4*>(s_head_w + o * TC_BN + c * 32 + j);
                            s0 = fmaf(dv[o], wv.x, s0);
                            s1 = fmaf(dv[o], wv.y, s1);
                            s2 = fmaf(dv[o], wv.z, s2);
                            s3 = fmaf(dv[o], wv.w, s3);
                        }
                    }
                    z[j] = s0 * tc_act_grad<ACT>(z[j]);
                    z[j + 1] = s1 * tc_act_grad<ACT>(z[j + 1]);
                    z[j + 2] = s2 * tc_act_grad<ACT>(z[j + 2]);
                    z[j + 3] = s3 * tc_act_grad<ACT>(z[j + 3]);
                }
                if (stamper) tc_stamp(p.timeline, 4 + c);      // chunk computed
                const int g = it * CHUNKS + c, stage = g & 1;     // == wg: a group owns one stage
                mbar_wait(&empty_bar[stage], ((g >> 1) & 1) ^ 1);
                if (stamper) tc_stamp(p.timeline, 12 + c);     // stage free
                unsigned char* st = smem + stage * Cfg::STAGE_BYTES;
                if (issuer) bulk_wait_read<0>();    // stores of the previous chunk have left the stage
                group_sync(wg);
                store_operand_row<PASSES>(st, st + Cfg::A_LO, trow, z, plain_dz2);
                fence_proxy_async_smem();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_ready[stage]);
                if (stamper) tc_stamp(p.timeline, 20 + c);     // published
                // dz2 for the weight-gradient kernels: the operand tile is already in the TMA layout
                // (tf32 split: hi and lo tiles; plain: the hi tile holds the float32 values) ->
                // asynchronous store by one thread
                group_sync(wg);
                if (issuer) {
                    tma_store_2d(&map_dz2_hi, st, c * TC_BK, tile * TC_BM);
                    if (PASSES == 3 && !plain_dz2) tma_store_2d(&map_dz2_lo, st + Cfg::A_LO, c * TC_BK, tile * TC_BM);
                    bulk_commit();
                }
            }
            // ---- b) epilogue: dz1 = acc * act'(h1).  h1 (hi + lo) is read with coalesced 64-byte
            // row segments, the accumulator half chunk goes through the staging block into the same
            // lane -> (2 rows x 16 columns) layout, the products are written into the group's idle
            // operand buffers in the TMA layout and stored by TMA; the h1 loads of the next half
            // chunk are in flight while the current one is processed ------------------------------
            float ua[16], ub[16];
            auto load_h1 = [&](int c, int half) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int64_t r = row0 + 2 * i + (lane >> 4);
                    const int64_t e = r * TC_BN + c * 32 + half * 16 + (lane & 15);
                    ua[i] = r < p.n_rows ? __ldg(p.h1_hi + e) : 0.0f;
                    ub[i] = (PASSES == 3 && p.h1_lo && r < p.n_rows) ? __ldg(p.h1_lo + e) : 0.0f;
                }
            };
            load_h1(wg, 0);
            if (stamper) tc_stamp(p.timeline, 28 + wg * 4);    // dz2 phase done
            mbar_wait(acc_full, it & 1);
            tcgen05_fence_after();
            if (stamper) tc_stamp(p.timeline, 29 + wg * 4);    // accumulator complete
            if (issuer) bulk_wait_read<0>();        // dz2 stores have left this group's buffers
            group_sync(wg);
            unsigned char* own_stage = smem + wg * Cfg::STAGE_BYTES;
            int n_staged = 0;
#pragma unroll 1
            for (int c = wg; c < CHUNKS; c += 2) {
                constexpr int NBUF = PASSES == 3 ? 2 : 1;
                unsigned char* buf = own_stage + (n_staged % NBUF) * TC_A_BYTES;
                if (n_staged >= NBUF) {
                    if (issuer) bulk_wait_read<NBUF - 1>();
                    group_sync(wg);
                }
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    uint32_t v[16];
                    tcgen05_ld_32x16(tmem_base + t_lane + (uint32_t)(c * 32 + half * 16), v);
                    float* mine = stg + lane * TCM_STG_STRIDE;
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        *reinterpret_cast<float4*>(mine + j) = make_float4(
                            __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                            __uint_as_float(v[j + 3]));
                    __syncwarp();
                    const int cc = half * 16 + (lane & 15);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int r = 2 * i + (lane >> 4);
                        const float gval = tc_act_grad<ACT>(ua[i] + ub[i]) * stg[r * TCM_STG_STRIDE + (lane & 15)];
                        *reinterpret_cast<float*>(buf + sw128_offset(q * 32 + r, cc >> 2) + ((cc & 3) << 2)) = gval;
                    }
                    __syncwarp();
                    if (half == 0) load_h1(c, 1);
                    else if (c + 2 < CHUNKS) load_h1(c + 2, 0);
                }
                fence_proxy_async_smem();
                group_sync(wg);
                if (issuer) {
                    tma_store_2d(&map_dz1, buf, c * TC_BK, tile * TC_BM);
                    bulk_commit();
                }
                ++n_staged;
            }
            if (stamper) tc_stamp(p.timeline, 30 + wg * 4);    // epilogue done
            // this tile's TMEM reads (both groups) are complete before the next tile's first MMA
            // overwrites the accumulator: that MMA waits for a_ready of chunk 0, which group 0
            // arrives only after this fence and barrier
            tcgen05_fence_before();
            if (issuer) bulk_wait_read<0>();        // dz1 stores have left the operand buffers
            row_warps_sync();
        }
        if (issuer) bulk_wait_all();
    }
    tcgen05_fence_before();
    __syncthreads();
    if (CLUSTER == 2) cluster_sync_all();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256)
                     : "memory");
    }
}


// =====================================================================================
// One kernel per network-minibatch for the activation passes: forward -> loss -> backward.
//
//     x -> h1 = act(x W1^T + b1) -> h2 = act(h1 W2^T + b2) -> out = h2 W3^T + b3
//     loss / dout:  LOSS_VALUE   squared error against targets[idx]   (updaters/critics.py:18-28)
//                   LOSS_POLICY  clipped-ratio / policy-gradient loss of the detached-scale Gaussian
//                                head (updaters/actors.py:21-50,70-112, models/actors.py:37-66)
//     dz2 = (dout W3) * act'(h2) -> dz1 = (dz2 W2) * act'(h1)
//
// i.e. tc_mlp_forward_kernel (training mode), the loss kernel and tc_mlp_backward_kernel as ONE
// launch per 128-row tile: z2 never leaves TMEM (h2 is recomputed from the accumulator for the
// backward pass instead of being re-read from global memory), the head output and dout stay in
// registers / shared memory, and one launch + pipeline fill / drain disappears from the chain.
// Accumulator 0 (TMEM columns 0..255) holds z1, then -- once the mid epilogue has consumed it --
// the dz1 GEMM; accumulator 1 holds z2.  Per tile the two shared-memory stages are used 17 times:
// layer 1, 8 K chunks of layer 2 (B = W2), 8 K chunks of the backward GEMM (B = W2^T).
// What still goes to global memory is what the weight-gradient kernel consumes: xin, h1
// (tf32 split or plain), h2, dz2 (split or plain), dz1, dout.
// =====================================================================================
enum { TC_LOSS_VALUE = 0, TC_LOSS_POLICY = 1 };

struct TcTrainParams {
    TcMlpParams f;              // forward part (input assembly, W1 image, biases, head, saves)
    // loss
    const int64_t* idx;         // minibatch indices of the rows (NULL = identity)
    const float* targets;       // LOSS_VALUE: returns, indexed by idx[row]
    const float* log_scale;     // LOSS_POLICY: [A]
    const float* actions;       // [., A] indexed by idx[row]
    const float* advantages;    // [.]
    const float* old_log_probs; // [.]
    float ratio_clip, entropy_coeff;
    double* stats;              // TB_STAT_* block (zeroed by the caller)
    float* dout;                // [n_rows, ld_dout]: head gradient (+ log_scale columns) for the weight gradients
    int ld_dout;
    // backward part
    float* dz2_hi; float* dz2_lo;   // tf32 split, or plain float32 when dz2_lo == NULL
    float* dz1;
};

template <int PASSES, int ACT, int LOSS>
__global__ void __launch_bounds__(TCM_THREADS, 1)
tc_mlp_train_kernel(const __grid_constant__ CUtensorMap map_w2_hi, const __grid_constant__ CUtensorMap map_w2_lo,
                    const __grid_constant__ CUtensorMap map_w2t_hi, const __grid_constant__ CUtensorMap map_w2t_lo,
                    const __grid_constant__ CUtensorMap map_h1_hi, const __grid_constant__ CUtensorMap map_h1_lo,
                    const __grid_constant__ CUtensorMap map_h2, const __grid_constant__ CUtensorMap map_dz2_hi,
                    const __grid_constant__ CUtensorMap map_dz2_lo, const __grid_constant__ CUtensorMap map_dz1,
                    const TcTrainParams q) {
    using Cfg = TcMlpCfg<PASSES>;
    const TcMlpParams& p = q.f;
    if (skip_requested(p.skip)) return;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* epi = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
    float* s_head_w = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::EPI_BYTES);
    float* s_b1 = s_head_w + TC_MAX_HEAD * TC_BN;
    float* s_b2 = s_b1 + TC_BN;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::EPI_BYTES +
                                                 Cfg::CONST_BYTES);
    uint64_t* full_bar = bars;                        // [STAGES] B operand landed
    uint64_t* a_ready = bars + Cfg::STAGES;           // [STAGES] A operand written by a row-warp group
    uint64_t* empty_bar = bars + 2 * Cfg::STAGES;     // [STAGES] MMAs reading the stage retired
    uint64_t* acc_full = bars + 3 * Cfg::STAGES;      // [2] accumulator complete
    uint64_t* x_ready = acc_full + 2;                 // layer-1 A operand written (all row warps)
    uint64_t* h1_full = x_ready + 1;                  // [2 groups][2 slots] h1 tile re-loaded by TMA
    uint64_t* tile_free = h1_full + 4;                // all row warps are done with the stages (next tile may load)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tile_free + 1);
    __shared__ float s_scale[TC_MAX_HEAD], s_dsc[TC_MAX_HEAD];     // policy head: scale, dscale/dlog_scale

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (int)((p.n_rows + TC_BM - 1) / TC_BM);
    if (threadIdx.x == 0) tc_stamp(p.timeline, 0);
    constexpr int CH = TC_K / TC_BK;                  // 8 chunks
    constexpr int USES = 1 + 2 * CH;                  // stage uses per tile: layer 1 + 8 + 8

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&a_ready[s], TCM_ROW_WARPS / 2);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&acc_full[0], 1);
        mbar_init(&acc_full[1], 1);
        mbar_init(x_ready, TCM_ROW_WARPS);
        for (int k = 0; k < 4; ++k) mbar_init(&h1_full[k], 1);
        mbar_init(tile_free, TCM_ROW_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < TC_BN; i += TCM_THREADS) {
        s_b1[i] = p.b1[i];
        s_b2[i] = p.b2[i];
    }
    for (int i = threadIdx.x; i < TC_MAX_HEAD * TC_BN; i += TCM_THREADS)
        s_head_w[i] = i < p.n_head * TC_BN ? p.head_w[i] : 0.0f;
    if (LOSS == TC_LOSS_POLICY && (int)threadIdx.x < p.n_head)
        s_scale[threadIdx.x] = detached_scale(q.log_scale[threadIdx.x], &s_dsc[threadIdx.x]);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int k_steps1 = (p.d_in + 7) >> 3;           // layer-1 MMAs (K = 8 each) per pass
    if (threadIdx.x == 0) tc_stamp(p.timeline, 1);    // setup done

    if (warp == 0) {
        // ===================== producer: W1 image, W2 chunks, W2^T chunks =====================
        if (lane == 0) {
            int g = 0, itp = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++itp) {
                // the row warps re-use the B regions of both stages in their last phase (h1 tiles)
                if (itp > 0) mbar_wait(tile_free, (itp - 1) & 1);
                for (int u = 0; u < USES; ++u, ++g) {
                    const int stage = g & 1;
                    mbar_wait(&empty_bar[stage], ((g >> 1) & 1) ^ 1);
                    unsigned char* st = smem + stage * Cfg::STAGE_BYTES;
                    mbar_expect_tx(&full_bar[stage], (PASSES == 3 ? 2 : 1) * TC_B_BYTES);
                    if (u == 0) {
                        bulk_load(st + Cfg::B_HI, p.w1_img_hi, TC_B_BYTES, &full_bar[stage]);
                        if (PASSES == 3) bulk_load(st + Cfg::B_LO, p.w1_img_lo, TC_B_BYTES, &full_bar[stage]);
                    } else if (u <= CH) {
                        tma_load_2d(st + Cfg::B_HI, &map_w2_hi, &full_bar[stage], (u - 1) * TC_BK, 0);
                        if (PASSES == 3) tma_load_2d(st + Cfg::B_LO, &map_w2_lo, &full_bar[stage], (u - 1) * TC_BK, 0);
                    } else {
                        tma_load_2d(st + Cfg::B_HI, &map_w2t_hi, &full_bar[stage], (u - 1 - CH) * TC_BK, 0);
                        if (PASSES == 3)
                            tma_load_2d(st + Cfg::B_LO, &map_w2t_lo, &full_bar[stage], (u - 1 - CH) * TC_BK, 0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int g = 0, it = 0;
            uint32_t a_phase = 0;                     // bit s: parity of the next completion of a_ready[s]
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                for (int u = 0; u < USES; ++u, ++g) {
                    const int stage = g & 1;
                    const uint32_t parity = (g >> 1) & 1;
                    mbar_wait(&full_bar[stage], parity);
                    if (u == 0) mbar_wait(x_ready, it & 1);
                    else mbar_wait(&a_ready[stage], (a_phase >> stage) & 1u), a_phase ^= 1u << stage;
                    tcgen05_fence_after();
                    unsigned char* st = smem + stage * Cfg::STAGE_BYTES;
                    const uint64_t a_hi = umma_desc_kmajor_sw128(st);
                    const uint64_t a_lo = umma_desc_kmajor_sw128(st + Cfg::A_LO);
                    const uint64_t b_hi = umma_desc_kmajor_sw128(st + Cfg::B_HI);
                    const uint64_t b_lo = umma_desc_kmajor_sw128(st + Cfg::B_LO);
                    // accumulator 0: layer 1 (u = 0) and the backward GEMM (u > 8); accumulator 1: layer 2
                    const uint32_t d_tmem = tmem_base + ((u >= 1 && u <= CH) ? (uint32_t)TC_BN : 0u);
                    const int ks = u == 0 ? k_steps1 : TC_BK / 8;
                    const bool fresh = u == 0 || u == 1 || u == CH + 1;
                    for (int k = 0; k < ks; ++k) {
                        const uint64_t koff = (uint64_t)(k * 32 >> 4);
                        const uint32_t accumulate = (fresh && k == 0) ? 0u : 1u;
                        if (PASSES == 3) {
                            tcgen05_mma_tf32(d_tmem, a_lo + koff, b_hi + koff, kIdescTf32, accumulate);
                            tcgen05_mma_tf32(d_tmem, a_hi + koff, b_lo + koff, kIdescTf32, 1);
                            tcgen05_mma_tf32(d_tmem, a_hi + koff, b_hi + koff, kIdescTf32, 1);
                        } else {
                            tcgen05_mma_tf32(d_tmem, a_hi + koff, b_hi + koff, kIdescTf32, accumulate);
                        }
                    }
                    tcgen05_commit(&empty_bar[stage]);
                    if (u == 0 || u == USES - 1) tcgen05_commit(&acc_full[0]);
                    if (u == CH) tcgen05_commit(&acc_full[1]);
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== row warps =====================
        const int rw = warp - 4, qd = rw & 3, wg = rw >> 2;
        const int trow = qd * 32 + lane;
        const bool issuer = qd == 0 && lane == 0;            // this group's TMA-store thread
        float* stg = epi + rw * 32 * TCM_STG_STRIDE;
        const int d_in = p.d_in;
        const int ldx = (d_in + 1 + 3) & ~3;
        const bool split_h1 = p.h1_lo != nullptr, split_dz2 = q.dz2_lo != nullptr;
        const uint32_t t_lane = (uint32_t)(qd * 32) << 16;
        const int A = p.n_head;
        const bool stamper = rw == 0 && lane == 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int64_t row0 = (int64_t)tile * TC_BM + qd * 32;
            const int64_t row = (int64_t)tile * TC_BM + trow;
            const bool live = row < p.n_rows;
            const int g0 = it * USES;                 // stage use index of this tile's layer 1
            const int64_t src = live ? (q.idx ? q.idx[row] : row) : 0;      // loss inputs of this row
            // ---- a) layer-1 A operand: this thread's 16 input columns ---------------------------
            {
                float xv[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) xv[j] = 0.0f;
                if (live) {
                    const int64_t r = p.in.d_idx ? p.in.d_idx[row] : row;
                    const int64_t r2 = p.in.gather2 ? r : row;
                    // all 16 gathered loads in flight before the first use (two dependent DRAM
                    // round trips -- index, then row -- are this phase's critical path)
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int c = 16 * wg + j;
                        if (c < d_in)
                            xv[j] = c < p.in.dim1 ? ldg_nc_volatile(p.in.d_x1 + r * p.in.dim1 + c)
                                                  : ldg_nc_volatile(p.in.d_x2 + r2 * p.in.dim2 + (c - p.in.dim1));
                    }
                    if (p.in.d_mean) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int c = 16 * wg + j;
                            if (c < p.in.dim1 && c < d_in)      // mean_stds.py:36  (val - mean) / std
                                xv[j] = __fdiv_rn(__fsub_rn(xv[j], p.in.d_mean[c]), p.in.d_std[c]);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int c = 16 * wg + j;
                        if (c < ldx) p.xin_save[row * ldx + c] = c < d_in ? xv[j] : (c == d_in ? 1.0f : 0.0f);
                    }
                    if (wg == 1)
                        for (int c = 32; c < ldx; ++c) p.xin_save[row * ldx + c] = c == d_in ? 1.0f : 0.0f;
                }
                const int stage = g0 & 1;
                mbar_wait(&empty_bar[stage], ((g0 >> 1) & 1) ^ 1);
                unsigned char* st = smem + stage * Cfg::STAGE_BYTES;
                store_operand_half<PASSES>(st, st + Cfg::A_LO, trow, wg, xv);
                fence_proxy_async_smem();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(x_ready);
                if (stamper) tc_stamp(p.timeline, 2);         // layer-1 A operand published
            }
            // ---- b) mid epilogue: h1 chunks -> layer-2 A operand (+ saved for the weight gradient) --
            mbar_wait(&acc_full[0], 0);               // first completion of the tile (two per tile)
            tcgen05_fence_after();
            if (stamper) tc_stamp(p.timeline, 3);             // z1 complete
#pragma unroll 1
            for (int c = wg; c < CH; c += 2) {
                uint32_t v[32];
                tcgen05_ld_32x32(tmem_base + t_lane + (uint32_t)(c * 32), v);
                float hv[32];
                const float4* b4 = reinterpret_cast<const float4*>(s_b1 + c * 32);
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b = b4[j / 4];
                    hv[j] = tc_act<ACT>(__uint_as_float(v[j]) + b.x);
                    hv[j + 1] = tc_act<ACT>(__uint_as_float(v[j + 1]) + b.y);
                    hv[j + 2] = tc_act<ACT>(__uint_as_float(v[j + 2]) + b.z);
                    hv[j + 3] = tc_act<ACT>(__uint_as_float(v[j + 3]) + b.w);
                }
                const int g = g0 + 1 + c, stage = g & 1;
                mbar_wait(&empty_bar[stage], ((g >> 1) & 1) ^ 1);
                unsigned char* st = smem + stage * Cfg::STAGE_BYTES;
                // the TMA stores of this group's previous chunk have left the stage
                if (issuer) bulk_wait_read<0>();
                group_sync(wg);
                // plain mode: the hi tile keeps the float32 value (the tensor core ignores the low
                // mantissa bits) and IS the saved h1; split mode: tf32-exact hi + lo, both saved
                store_operand_row<PASSES>(st, st + Cfg::A_LO, trow, hv, !split_h1);
                fence_proxy_async_smem();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_ready[stage]);
                group_sync(wg);
                if (issuer) {
                    tma_store_2d(&map_h1_hi, st, c * TC_BK, tile * TC_BM);
                    if (PASSES == 3 && split_h1) tma_store_2d(&map_h1_lo, st + Cfg::A_LO, c * TC_BK, tile * TC_BM);
                    bulk_commit();
                }
            }
            // ---- c) pass 1 over z2: h2 (saved), head output ------------------------------------
            if (stamper) tc_stamp(p.timeline, 4);             // mid epilogue done (group 0)
            mbar_wait(&acc_full[1], it & 1);
            tcgen05_fence_after();
            if (stamper) tc_stamp(p.timeline, 5);             // z2 complete
            float hacc[TC_MAX_HEAD];
#pragma unroll
            for (int o = 0; o < TC_MAX_HEAD; ++o) hacc[o] = 0.0f;
            // both stages are idle now (all layer-2 MMAs retired): group wg stages h2 in the A
            // buffers of stage wg; the h1 stores that may still read them must have left
            unsigned char* own_stage = smem + wg * Cfg::STAGE_BYTES;
            if (issuer) bulk_wait_read<0>();
            row_warps_sync();                         // (the other group's stores read both stages over time)
            int n_staged = 0;
#pragma unroll 1
            for (int c = wg; c < CH; c += 2) {
                uint32_t v[32];
                tcgen05_ld_32x32(tmem_base + t_lane + (uint32_t)(TC_BN + c * 32), v);
                float hv[32];
                const float4* b4 = reinterpret_cast<const float4*>(s_b2 + c * 32);
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b = b4[j / 4];
                    hv[j] = tc_act<ACT>(__uint_as_float(v[j]) + b.x);
                    hv[j + 1] = tc_act<ACT>(__uint_as_float(v[j + 1]) + b.y);
                    hv[j + 2] = tc_act<ACT>(__uint_as_float(v[j + 2]) + b.z);
                    hv[j + 3] = tc_act<ACT>(__uint_as_float(v[j + 3]) + b.w);
                }
#pragma unroll
                for (int o = 0; o < TC_MAX_HEAD; ++o) {
                    if (o < A) {
                        const float4* w4 = reinterpret_cast<const float4*>(s_head_w + o * TC_BN + c * 32);
                        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 wv = w4[j / 4];
                            s0 = fmaf(hv[j], wv.x, s0);
                            s1 = fmaf(hv[j + 1], wv.y, s1);
                            s2 = fmaf(hv[j + 2], wv.z, s2);
                            s3 = fmaf(hv[j + 3], wv.w, s3);
                        }
                        hacc[o] += (s0 + s1) + (s2 + s3);
                    }
                }
                // h2 for the weight-gradient kernel: staged in the TMA layout, stored asynchronously
                constexpr int NBUF = PASSES == 3 ? 2 : 1;
                unsigned char* buf = own_stage + (n_staged % NBUF) * TC_A_BYTES;
                if (n_staged >= NBUF) {
                    if (issuer) bulk_wait_read<NBUF - 1>();
                    group_sync(wg);
                }
                store_plain_row(buf, trow, hv);
                fence_proxy_async_smem();
                group_sync(wg);
                if (issuer) {
                    tma_store_2d(&map_h2, buf, c * TC_BK, tile * TC_BM);
                    bulk_commit();
                }
                ++n_staged;
            }
            // ---- head output, loss, dout: group 1 hands its partial sums to group 0, group 0
            // computes the row's loss terms and hands the head gradient back -----------------------
            if (stamper) tc_stamp(p.timeline, 6);             // pass 1 (h2, head sums) done
            if (wg == 1) {
#pragma unroll
                for (int o = 0; o < TC_MAX_HEAD; ++o) stg[lane * TCM_STG_STRIDE + o] = hacc[o];
            }
            row_warps_sync();
            float dv[TC_MAX_HEAD];
#pragma unroll
            for (int o = 0; o < TC_MAX_HEAD; ++o) dv[o] = 0.0f;
            double st_a = 0.0, st_b = 0.0, st_c = 0.0, st_d = 0.0, st_rows = 0.0;
            if (wg == 0) {
                const float* other = epi + (rw + 4) * 32 * TCM_STG_STRIDE + lane * TCM_STG_STRIDE;
                float outv[TC_MAX_HEAD];
#pragma unroll
                for (int o = 0; o < TC_MAX_HEAD; ++o)
                    outv[o] = o < A ? (hacc[o] + other[o]) + __ldg(p.head_b + o) : 0.0f;
                if (live) {
                    if (p.head_out) {
#pragma unroll
                        for (int o = 0; o < TC_MAX_HEAD; ++o)
                            if (o < A) p.head_out[row * A + o] = outv[o];
                    }
                    if (LOSS == TC_LOSS_VALUE) {
                        // same arithmetic as mse_loss_kernel (csrc/heads.cu)
                        const float d = outv[0] - q.targets[src];
                        dv[0] = 2.0f * d;
                        q.dout[row * q.ld_dout] = dv[0];
                        st_a = (double)d * (double)d;
                        st_b = outv[0];
                        st_rows = 1.0;
                    } else {
                        // same arithmetic as gauss_policy_loss_kernel (csrc/heads.cu)
                        const float adv = q.advantages[src], old_lp = q.old_log_probs[src];
                        float lp = 0.0f, dd[TC_MAX_HEAD], loc[TC_MAX_HEAD];
#pragma unroll
                        for (int a = 0; a < TC_MAX_HEAD; ++a) {
                            if (a < A) {
                                loc[a] = tanhf(outv[a]);
                                const float sc = s_scale[a];
                                dd[a] = q.actions[src * A + a] - loc[a];
                                lp += -(dd[a] * dd[a]) / (2.0f * (sc * sc)) - logf(sc) - kLogSqrt2Pi;
                            }
                        }
                        float g_lp, loss;
                        if (q.ratio_clip > 0.0f) {                               // actors.py:84-90
                            const float ratio = expf(lp - old_lp);
                            const float lo = 1.0f - q.ratio_clip, hi = 1.0f + q.ratio_clip;
                            const float clipped = fminf(fmaxf(ratio, lo), hi);
                            const float s1 = adv * ratio, s2 = adv * clipped;
                            loss = -fminf(s1, s2);
                            const float w1 = s1 < s2 ? 1.0f : (s1 == s2 ? 0.5f : 0.0f);
                            const float w2 = (1.0f - w1) * ((ratio >= lo && ratio <= hi) ? 1.0f : 0.0f);
                            g_lp = -adv * ratio * (w1 + w2);
                            st_c = (ratio > hi || ratio < lo) ? 1.0 : 0.0;        // actors.py:105
                        } else {                                                 // actors.py:34
                            loss = -adv * lp;
                            g_lp = -adv;
                        }
#pragma unroll
                        for (int a = 0; a < TC_MAX_HEAD; ++a) {
                            if (a < A) {
                                const float sc = s_scale[a];
                                const float inv_var = 1.0f / (sc * sc);
                                dv[a] = g_lp * dd[a] * inv_var * (1.0f - loc[a] * loc[a]);
                                q.dout[row * q.ld_dout + a] = dv[a];
                                const float dlp_dsc = dd[a] * dd[a] * inv_var / sc - 1.0f / sc;
                                const float dent_dsc = 1.0f / sc;
                                q.dout[row * q.ld_dout + A + a] =
                                    (g_lp * dlp_dsc - (q.entropy_coeff / (float)A) * dent_dsc) * s_dsc[a];
                            }
                        }
                        st_a = loss;
                        st_b = old_lp - lp;                                       // actors.py:103
                        st_d = adv != 0.0f ? 1.0 : 0.0;
                        st_rows = 1.0;
                    }
                }
                // head gradient of this row for the other group
#pragma unroll
                for (int o = 0; o < TC_MAX_HEAD; ++o) stg[lane * TCM_STG_STRIDE + 8 + o] = dv[o];
                st_a = warp_sum(st_a); st_b = warp_sum(st_b); st_rows = warp_sum(st_rows);
                if (LOSS == TC_LOSS_POLICY) { st_c = warp_sum(st_c); st_d = warp_sum(st_d); }
            }
            row_warps_sync();
            if (wg == 1) {
                const float* other = epi + (rw - 4) * 32 * TCM_STG_STRIDE + lane * TCM_STG_STRIDE + 8;
#pragma unroll
                for (int o = 0; o < TC_MAX_HEAD; ++o) dv[o] = other[o];
            } else if (lane == 0) {
                // one atomic per warp and statistic (4 warps per tile)
                if (LOSS == TC_LOSS_VALUE) {
                    atomicAdd(&q.stats[TB_STAT_LOSS], st_a);
                    atomicAdd(&q.stats[TB_STAT_VALUE], st_b);
                    atomicAdd(&q.stats[TB_STAT_ROWS], st_rows);
                } else {
                    float ent_row = 0.0f, std_row = 0.0f;
                    for (int a = 0; a < A; ++a) {
                        ent_row += kEntropyConst + logf(s_scale[a]);
                        std_row += s_scale[a];
                    }
                    atomicAdd(&q.stats[TB_STAT_LOSS], st_a);
                    atomicAdd(&q.stats[TB_STAT_KL], st_b);
                    atomicAdd(&q.stats[TB_STAT_CLIPPED], st_c);
                    atomicAdd(&q.stats[TB_STAT_NONZERO_ADV], st_d);
                    atomicAdd(&q.stats[TB_STAT_ROWS], st_rows);
                    atomicAdd(&q.stats[TB_STAT_ENTROPY], st_rows * (double)ent_row);
                    atomicAdd(&q.stats[TB_STAT_STD], st_rows * (double)std_row);
                }
            }
            // ---- d) pass 2 over z2: dz2 = (dout W3) * act'(h2) -> A operand of the backward GEMM ----
            // the h2 stores must have left the A buffers before they are rewritten
            if (issuer) bulk_wait_read<0>();
            row_warps_sync();
            if (stamper) tc_stamp(p.timeline, 7);             // loss / head gradient exchanged, h2 stores read
#pragma unroll 1
            for (int c = wg; c < CH; c += 2) {
                uint32_t v[32];
                tcgen05_ld_32x32(tmem_base + t_lane + (uint32_t)(TC_BN + c * 32), v);
                float z[32];
                const float4* b4 = reinterpret_cast<const float4*>(s_b2 + c * 32);
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b = b4[j / 4];
                    z[j] = tc_act<ACT>(__uint_as_float(v[j]) + b.x);
                    z[j + 1] = tc_act<ACT>(__uint_as_float(v[j + 1]) + b.y);
                    z[j + 2] = tc_act<ACT>(__uint_as_float(v[j + 2]) + b.z);
                    z[j + 3] = tc_act<ACT>(__uint_as_float(v[j + 3]) + b.w);
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                    for (int o = 0; o < TC_MAX_HEAD; ++o) {
                        if (o < A) {
                            const float4 wv = *reinterpret_cast<const float4*>(s_head_w + o * TC_BN + c * 32 + j);
                            s0 = fmaf(dv[o], wv.x, s0);
                            s1 = fmaf(dv[o], wv.y, s1);
                            s2 = fmaf(dv[o], wv.z, s2);
                            s3 = fmaf(dv[o], wv.w, s3);
                        }
                    }
                    z[j] = s0 * tc_act_grad<ACT>(z[j]);
                    z[j + 1] = s1 * tc_act_grad<ACT>(z[j + 1]);
                    z[j + 2] = s2 * tc_act_grad<ACT>(z[j + 2]);
                    z[j + 3] = s3 * tc_act_grad<ACT>(z[j + 3]);
                }
                const int g = g0 + 1 + CH + c, stage = g & 1;
                mbar_wait(&empty_bar[stage], ((g >> 1) & 1) ^ 1);
                unsigned char* st = smem + stage * Cfg::STAGE_BYTES;
                if (issuer) bulk_wait_read<0>();    // stores of the previous chunk have left the stage
                group_sync(wg);
                store_operand_row<PASSES>(st, st + Cfg::A_LO, trow, z, !split_dz2);
                fence_proxy_async_smem();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_ready[stage]);
                group_sync(wg);
                if (issuer) {
                    tma_store_2d(&map_dz2_hi, st, c * TC_BK, tile * TC_BM);
                    if (PASSES == 3 && split_dz2) tma_store_2d(&map_dz2_lo, st + Cfg::A_LO, c * TC_BK, tile * TC_BM);
                    bulk_commit();
                }
            }
            // ---- e) epilogue of the backward GEMM: dz1 = acc0 * act'(h1).  h1 comes back from L2 by
            // TMA into the (now idle) B regions of this group's stage -- the rows this CTA stored a few
            // microseconds ago, in the operand-tile layout, so every thread reads its own row with
            // 16-byte shared-memory loads: no global-load latency chain, no transposes ----------------
            if (stamper) tc_stamp(p.timeline, 8);             // pass 2 (dz2 operands) done
            if (issuer) bulk_wait_all();              // this group's h1 TMA stores are complete
            mbar_wait(&acc_full[0], 1);               // second completion of the tile: all MMAs retired
            tcgen05_fence_after();
            if (stamper) tc_stamp(p.timeline, 9);             // backward GEMM complete
            if (issuer) bulk_wait_read<0>();          // dz2 stores have left the A buffers
            row_warps_sync();
            own_stage = smem + wg * Cfg::STAGE_BYTES;
            constexpr int SLOT_BYTES = (PASSES == 3 ? 2 : 1) * TC_A_BYTES;     // hi (+ lo) tile of one chunk
            unsigned char* slots = own_stage + Cfg::B_HI;
            auto load_h1_tile = [&](int k) {          // issuer only: chunk wg + 2 k -> slot k & 1
                const int c = wg + 2 * k, slot = k & 1;
                uint64_t* bar = &h1_full[wg * 2 + slot];
                fence_proxy_async_smem();
                mbar_expect_tx(bar, (PASSES == 3 && split_h1 ? 2 : 1) * TC_A_BYTES);
                tma_load_2d(slots + slot * SLOT_BYTES, &map_h1_hi, bar, c * TC_BK, tile * TC_BM);
                if (PASSES == 3 && split_h1)
                    tma_load_2d(slots + slot * SLOT_BYTES + TC_A_BYTES, &map_h1_lo, bar, c * TC_BK, tile * TC_BM);
            };
            if (issuer) { load_h1_tile(0); load_h1_tile(1); }
#pragma unroll 1
            for (int k = 0; k < CH / 2; ++k) {
                const int c = wg + 2 * k, slot = k & 1;
                constexpr int NBUF = PASSES == 3 ? 2 : 1;
                unsigned char* buf = own_stage + (k % NBUF) * TC_A_BYTES;
                mbar_wait(&h1_full[wg * 2 + slot], (uint32_t)((it * (CH / 4) + (k >> 1)) & 1));
                uint32_t v[32];
                tcgen05_ld_32x32(tmem_base + t_lane + (uint32_t)(c * 32), v);
                float gv[32];
                const unsigned char* hi_tile = slots + slot * SLOT_BYTES;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    float4 h = *reinterpret_cast<const float4*>(hi_tile + sw128_offset(trow, u));
                    if (PASSES == 3 && split_h1) {
                        const float4 l = *reinterpret_cast<const float4*>(hi_tile + TC_A_BYTES + sw128_offset(trow, u));
                        h.x += l.x; h.y += l.y; h.z += l.z; h.w += l.w;
                    }
                    gv[4 * u] = tc_act_grad<ACT>(h.x) * __uint_as_float(v[4 * u]);
                    gv[4 * u + 1] = tc_act_grad<ACT>(h.y) * __uint_as_float(v[4 * u + 1]);
                    gv[4 * u + 2] = tc_act_grad<ACT>(h.z) * __uint_as_float(v[4 * u + 2]);
                    gv[4 * u + 3] = tc_act_grad<ACT>(h.w) * __uint_as_float(v[4 * u + 3]);
                }
                if (k >= NBUF) {                      // the store that read this staging buffer has left it
                    if (issuer) bulk_wait_read<NBUF - 1>();
                    group_sync(wg);
                }
                store_plain_row(buf, trow, gv);
                fence_proxy_async_smem();
                group_sync(wg);                       // staging tile complete; slot fully read by the group
                if (issuer) {
                    tma_store_2d(&map_dz1, buf, c * TC_BK, tile * TC_BM);
                    bulk_commit();
                    if (k + 2 < CH / 2) load_h1_tile(k + 2);
                }
            }
            // this tile's TMEM reads are complete before the next tile's MMAs are released, and
            // the stores have left the operand buffers before the next tile rewrites them
            tcgen05_fence_before();
            if (issuer) bulk_wait_read<0>();
            row_warps_sync();
            if (lane == 0) mbar_arrive(tile_free);
            if (stamper) tc_stamp(p.timeline, 10);            // dz1 epilogue done
        }
        if (issuer) bulk_wait_all();
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512)
                     : "memory");
    }
}

template <int PASSES, int ACT, int CLUSTER>
static int launch_tc_mlp_bwd(const CUtensorMap* maps, const TcMlpBwdParams& p, cudaStream_t s) {
    using Cfg = TcMlpCfg<PASSES>;
    auto kernel = tc_mlp_backward_kernel<PASSES, ACT, CLUSTER>;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        configured = true;
    }
    const int n_tiles = (int)((p.n_rows + TC_BM - 1) / TC_BM);
    if (CLUSTER == 1) {
        const int grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
        kernel<<<grid, TCM_THREADS, Cfg::SMEM_BYTES, s>>>(maps[0], maps[1], maps[2], maps[3], maps[4], p);
        return 0;
    }
    int grid = (n_tiles + 1) & ~1;
    if (grid > (kNumSMs & ~1)) grid = kNumSMs & ~1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(TCM_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, maps[0], maps[1], maps[2], maps[3], maps[4], p);
    return 0;
}

template <int PASSES, int ACT, int CLUSTER>
static int launch_tc_mlp(const CUtensorMap* maps, const TcMlpParams& p, cudaStream_t s) {
    using Cfg = TcMlpCfg<PASSES>;
    auto kernel = tc_mlp_forward_kernel<PASSES, ACT, CLUSTER>;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        configured = true;
    }
    const int n_tiles = (int)((p.n_rows + TC_BM - 1) / TC_BM);
    if (CLUSTER == 1) {
        const int grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
        kernel<<<grid, TCM_THREADS, Cfg::SMEM_BYTES, s>>>(maps[0], maps[1], maps[2], maps[3], maps[4], p);
        return 0;
    }
    // clusters of 2 CTAs: an even grid, both CTAs of a cluster run the same number of tiles
    int grid = (n_tiles + 1) & ~1;
    if (grid > (kNumSMs & ~1)) grid = kNumSMs & ~1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(TCM_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, maps[0], maps[1], maps[2], maps[3], maps[4], p);
    return 0;
}

static unsigned long long* g_timeline = nullptr;     // device buffer, 64 stamps

}  // namespace tb

extern "C" int tb_tc_timeline(uint64_t* out64) {
    // enable (first call) / read back the clock64() stamps CTA 0 of the last fused forward wrote
    using namespace tb;
    if (!g_timeline) {
        TB_REQUIRE(cudaMalloc(&g_timeline, 64 * sizeof(unsigned long long)) == cudaSuccess, TB_ENOTSUP,
                   "tb_tc_timeline: cudaMalloc failed");
        cudaMemset(g_timeline, 0, 64 * sizeof(unsigned long long));
    }
    if (out64) {
        TB_REQUIRE(cudaMemcpy(out64, g_timeline, 64 * sizeof(unsigned long long), cudaMemcpyDeviceToHost) ==
                       cudaSuccess, TB_ENOTSUP, "tb_tc_timeline: copy failed");
    }
    return 0;
}

namespace tb {
struct VLoss {
    const float* targets; const int64_t* idx; float* dout; int ld; double* stats; int slot; int count_rows;
};
}  // namespace tb

static int tc_mlp_forward_impl(const TbMlpShape* shape, const float* d_params, const float* d_packed,
                               const TbMlpInput* in, int64_t n_rows, float* d_out, float* d_xin,
                               float* d_h1_hi, float* d_h1_lo, float* d_h2, int32_t passes,
                               const tb::VLoss* loss, const int32_t* d_skip, void* stream) {
    using namespace tb;
    TB_REQUIRE(shape && d_params && d_packed && in && in->d_x1 && d_out && n_rows > 0, TB_EINVAL,
               "tb_tc_mlp_forward: null pointer");
    TB_REQUIRE(shape->hidden == 256 && shape->off_w2_hi > 0 && shape->off_w1_img_hi > 0 && shape->d_in >= 1 && shape->d_in <= 32 &&
               shape->n_out >= 1 && shape->n_out <= TC_MAX_HEAD, TB_ENOTSUP,
               "tb_tc_mlp_forward: needs hidden == 256, d_in <= 32, n_out <= 8 (got %d, %d, %d)",
               shape->hidden, shape->d_in, shape->n_out);
    TB_REQUIRE(in->dim1 + (in->d_x2 ? in->dim2 : 0) == shape->d_in, TB_EINVAL,
               "tb_tc_mlp_forward: input widths do not add up to d_in");
    TB_REQUIRE(passes == 1 || passes == 3, TB_EINVAL, "tb_tc_mlp_forward: passes must be 1 or 3");
    TB_REQUIRE(d_h1_hi || !d_h1_lo, TB_EINVAL, "tb_tc_mlp_forward: h1 lo without hi");
    // experimental: 2-CTA clusters with multicast of the weight-operand chunks (3-pass mode only;
    // written and compiled in round 1, not yet validated on hardware -> off unless requested)
    static const bool cluster_env = [] {
        const char* v = getenv("TONIC_B200_CLUSTER");
        return v && v[0] == '2';
    }();
    const bool cluster2 = cluster_env && passes == 3;
    CUtensorMap maps[5];
    int rc;
    const int b_box_rows = cluster2 ? TC_BN / 2 : TC_BN;       // each CTA of a cluster loads half the rows
    if ((rc = make_map(&maps[0], d_packed + shape->off_w2_hi, TC_BN, b_box_rows))) return rc;
    if ((rc = make_map(&maps[1], d_packed + shape->off_w2_lo, TC_BN, b_box_rows))) return rc;
    maps[2] = maps[3] = maps[4] = maps[0];       // placeholders when nothing is saved
    if (d_h1_hi) {      // tf32 split (two arrays) or, with d_h1_lo == NULL, ONE plain float32 array
        if ((rc = make_map(&maps[2], d_h1_hi, n_rows, TC_BM))) return rc;
        if ((rc = make_map(&maps[3], d_h1_lo ? d_h1_lo : d_h1_hi, n_rows, TC_BM))) return rc;
    }
    if (d_h2 && (rc = make_map(&maps[4], d_h2, n_rows, TC_BM))) return rc;
    TcMlpParams p;
    p.n_rows = n_rows; p.in = *in; p.d_in = shape->d_in; p.act = shape->act;
    p.w1_img_hi = d_packed + shape->off_w1_img_hi; p.w1_img_lo = d_packed + shape->off_w1_img_lo;
    p.b1 = d_params + shape->off_b1; p.b2 = d_params + shape->off_b2;
    p.xin_save = d_xin; p.h1_hi = d_h1_hi; p.h1_lo = d_h1_lo; p.h2 = d_h2;
    p.head_w = d_params + shape->off_w3; p.head_b = d_params + shape->off_b3; p.head_out = d_out;
    p.n_head = shape->n_out; p.skip = d_skip; p.timeline = g_timeline;
    p.loss_targets = nullptr; p.loss_idx = nullptr; p.loss_dout = nullptr; p.loss_ld = 1;
    p.loss_stats = nullptr; p.loss_stat_slot = 0; p.loss_count_rows = 0;
    if (loss) {
        TB_REQUIRE(shape->n_out == 1 && loss->targets && loss->dout && loss->stats && loss->ld >= 1 &&
                   loss->slot >= 0 && loss->slot < TB_STAT_COUNT, TB_EINVAL,
                   "tb_tc_mlp_forward_vloss: needs a single-output head, targets, dout and stats");
        p.loss_targets = loss->targets; p.loss_idx = loss->idx; p.loss_dout = loss->dout;
        p.loss_ld = loss->ld; p.loss_stats = loss->stats; p.loss_stat_slot = loss->slot;
        p.loss_count_rows = loss->count_rows;
    }
    ProfScope prof_scope("tb_tc_mlp_forward", stream);
    const bool tanh_act = shape->act == TB_ACT_TANH;
    if (passes == 3 && cluster2) {
        if (tanh_act) launch_tc_mlp<3, TB_ACT_TANH, 2>(maps, p, as_stream(stream));
        else launch_tc_mlp<3, TB_ACT_RELU, 2>(maps, p, as_stream(stream));
    } else if (passes == 3) {
        if (tanh_act) launch_tc_mlp<3, TB_ACT_TANH, 1>(maps, p, as_stream(stream));
        else launch_tc_mlp<3, TB_ACT_RELU, 1>(maps, p, as_stream(stream));
    } else {
        if (tanh_act) launch_tc_mlp<1, TB_ACT_TANH, 1>(maps, p, as_stream(stream));
        else launch_tc_mlp<1, TB_ACT_RELU, 1>(maps, p, as_stream(stream));
    }
    return check_launch("tb_tc_mlp_forward");
}

extern "C" int tb_tc_mlp_forward(const TbMlpShape* shape, const float* d_params, const float* d_packed,
                                 const TbMlpInput* in, int64_t n_rows, float* d_out, float* d_xin,
                                 float* d_h1_hi, float* d_h1_lo, float* d_h2, int32_t passes,
                                 const int32_t* d_skip, void* stream) {
    return tc_mlp_forward_impl(shape, d_params, d_packed, in, n_rows, d_out, d_xin, d_h1_hi, d_h1_lo, d_h2,
                               passes, nullptr, d_skip, stream);
}

extern "C" int tb_tc_mlp_forward_vloss(const TbMlpShape* shape, const float* d_params, const float* d_packed,
                                       const TbMlpInput* in, int64_t n_rows, float* d_out, float* d_xin,
                                       float* d_h1_hi, float* d_h1_lo, float* d_h2, int32_t passes,
                                       const float* d_targets, const int64_t* d_idx, float* d_dout,
                                       int32_t ld_dout, double* d_stats, int32_t stat_slot,
                                       int32_t count_rows, const int32_t* d_skip, void* stream) {
    tb::VLoss loss{d_targets, d_idx, d_dout, ld_dout, d_stats, stat_slot, count_rows};
    return tc_mlp_forward_impl(shape, d_params, d_packed, in, n_rows, d_out, d_xin, d_h1_hi, d_h1_lo, d_h2,
                               passes, &loss, d_skip, stream);
}

extern "C" int tb_tc_mlp_backward(const TbMlpShape* shape, const float* d_params, const float* d_packed,
                                  const float* d_dout, int32_t ld_dout, const float* d_h1_hi,
                                  const float* d_h1_lo, const float* d_h2, int64_t n_rows,
                                  float* d_dz2_hi, float* d_dz2_lo, float* d_dz1, int32_t passes,
                                  const int32_t* d_skip, void* stream) {
    using namespace tb;
    TB_REQUIRE(shape && d_params && d_packed && d_dout && d_h1_hi && d_h2 && d_dz2_hi &&
               d_dz1 && n_rows > 0, TB_EINVAL, "tb_tc_mlp_backward: null pointer");
    TB_REQUIRE(shape->hidden == 256 && shape->off_w2t_hi > 0 && shape->n_out >= 1 &&
               shape->n_out <= TC_MAX_HEAD && ld_dout >= shape->n_out, TB_ENOTSUP,
               "tb_tc_mlp_backward: needs hidden == 256 and n_out <= 8 (got %d, %d)", shape->hidden,
               shape->n_out);
    TB_REQUIRE(passes == 1 || passes == 3, TB_EINVAL, "tb_tc_mlp_backward: passes must be 1 or 3");
    static const bool cluster_env = [] {        // experimental, see tb_tc_mlp_forward
        const char* v = getenv("TONIC_B200_CLUSTER");
        return v && v[0] == '2';
    }();
    const bool cluster2 = cluster_env && passes == 3;
    const int b_box_rows = cluster2 ? TC_BN / 2 : TC_BN;
    CUtensorMap maps[5];
    int rc;
    if ((rc = make_map(&maps[0], d_packed + shape->off_w2t_hi, TC_BN, b_box_rows))) return rc;
    if ((rc = make_map(&maps[1], d_packed + shape->off_w2t_lo, TC_BN, b_box_rows))) return rc;
    if ((rc = make_map(&maps[2], d_dz2_hi, n_rows, TC_BM))) return rc;
    if ((rc = make_map(&maps[3], d_dz2_lo ? d_dz2_lo : d_dz2_hi, n_rows, TC_BM))) return rc;
    if ((rc = make_map(&maps[4], d_dz1, n_rows, TC_BM))) return rc;
    TcMlpBwdParams p;
    p.n_rows = n_rows; p.dout = d_dout; p.ld_dout = ld_dout; p.n_head = shape->n_out;
    p.head_w = d_params + shape->off_w3; p.h2 = d_h2; p.h1_hi = d_h1_hi; p.h1_lo = d_h1_lo;
    p.dz2_hi = d_dz2_hi; p.dz2_lo = d_dz2_lo; p.dz1 = d_dz1; p.skip = d_skip; p.timeline = g_timeline;
    ProfScope prof_scope("tb_tc_mlp_backward", stream);
    const bool tanh_act = shape->act == TB_ACT_TANH;
    if (passes == 3 && cluster2) {
        if (tanh_act) launch_tc_mlp_bwd<3, TB_ACT_TANH, 2>(maps, p, as_stream(stream));
        else launch_tc_mlp_bwd<3, TB_ACT_RELU, 2>(maps, p, as_stream(stream));
    } else if (passes == 3) {
        if (tanh_act) launch_tc_mlp_bwd<3, TB_ACT_TANH, 1>(maps, p, as_stream(stream));
        else launch_tc_mlp_bwd<3, TB_ACT_RELU, 1>(maps, p, as_stream(stream));
    } else {
        if (tanh_act) launch_tc_mlp_bwd<1, TB_ACT_TANH, 1>(maps, p, as_stream(stream));
        else launch_tc_mlp_bwd<1, TB_ACT_RELU, 1>(maps, p, as_stream(stream));
    }
    return check_launch("tb_tc_mlp_backward");
}

namespace tb {

template <int PASSES, int ACT, int LOSS>
static int launch_tc_train(const CUtensorMap* maps, const TcTrainParams& q, cudaStream_t s) {
    using Cfg = TcMlpCfg<PASSES>;
    auto kernel = tc_mlp_train_kernel<PASSES, ACT, LOSS>;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        configured = true;
    }
    const int n_tiles = (int)((q.f.n_rows + TC_BM - 1) / TC_BM);
    const int grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
    kernel<<<grid, TCM_THREADS, Cfg::SMEM_BYTES, s>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5],
                                                    maps[6], maps[7], maps[8], maps[9], q);
    return 0;
}

}  // namespace tb

extern "C" int tb_tc_mlp_train(const TbMlpShape* shape, const float* d_params, const float* d_packed,
                               const TbMlpInput* in, int64_t n_rows, int32_t loss_kind,
                               const int64_t* d_idx, const float* d_targets, const float* d_log_scale,
                               const float* d_actions, const float* d_advantages,
                               const float* d_old_log_probs, float ratio_clip, float entropy_coeff,
                               double* d_stats, float* d_out, float* d_xin, float* d_h1_hi, float* d_h1_lo,
                               float* d_h2, float* d_dout, int32_t ld_dout, float* d_dz2_hi,
                               float* d_dz2_lo, float* d_dz1, int32_t passes, const int32_t* d_skip,
                               void* stream) {
    using namespace tb;
    TB_REQUIRE(shape && d_params && d_packed && in && in->d_x1 && n_rows > 0 && d_stats && d_xin && d_h1_hi &&
               d_h2 && d_dout && d_dz2_hi && d_dz1, TB_EINVAL, "tb_tc_mlp_train: null pointer");
    TB_REQUIRE(shape->hidden == 256 && shape->off_w2_hi > 0 && shape->off_w2t_hi > 0 && shape->off_w1_img_hi > 0 &&
               shape->d_in >= 1 && shape->d_in <= 32 && shape->n_out >= 1 && shape->n_out <= TC_MAX_HEAD,
               TB_ENOTSUP, "tb_tc_mlp_train: needs hidden == 256, d_in <= 32, n_out <= 8 (got %d, %d, %d)",
               shape->hidden, shape->d_in, shape->n_out);
    TB_REQUIRE(in->dim1 + (in->d_x2 ? in->dim2 : 0) == shape->d_in, TB_EINVAL,
               "tb_tc_mlp_train: input widths do not add up to d_in");
    TB_REQUIRE(passes == 1 || passes == 3, TB_EINVAL, "tb_tc_mlp_train: passes must be 1 or 3");
    TB_REQUIRE((d_h1_lo == nullptr) == (d_dz2_lo == nullptr), TB_EINVAL,
               "tb_tc_mlp_train: h1 and dz2 both as tf32 splits or both plain (lo == NULL)");
    if (loss_kind == TC_LOSS_VALUE) {
        TB_REQUIRE(shape->n_out == 1 && d_targets && ld_dout >= 1, TB_EINVAL,
                   "tb_tc_mlp_train: the value loss needs a single output and targets");
    } else {
        TB_REQUIRE(loss_kind == TC_LOSS_POLICY && d_log_scale && d_actions && d_advantages && d_old_log_probs &&
                   ld_dout >= 2 * shape->n_out, TB_EINVAL,
                   "tb_tc_mlp_train: the policy loss needs log_scale, actions, advantages, old log-probs "
                   "and ld_dout >= 2 * act_dim");
    }
    CUtensorMap maps[10];
    int rc;
    if ((rc = make_map(&maps[0], d_packed + shape->off_w2_hi, TC_BN, TC_BN))) return rc;
    if ((rc = make_map(&maps[1], d_packed + shape->off_w2_lo, TC_BN, TC_BN))) return rc;
    if ((rc = make_map(&maps[2], d_packed + shape->off_w2t_hi, TC_BN, TC_BN))) return rc;
    if ((rc = make_map(&maps[3], d_packed + shape->off_w2t_lo, TC_BN, TC_BN))) return rc;
    if ((rc = make_map(&maps[4], d_h1_hi, n_rows, TC_BM))) return rc;
    if ((rc = make_map(&maps[5], d_h1_lo ? d_h1_lo : d_h1_hi, n_rows, TC_BM))) return rc;
    if ((rc = make_map(&maps[6], d_h2, n_rows, TC_BM))) return rc;
    if ((rc = make_map(&maps[7], d_dz2_hi, n_rows, TC_BM))) return rc;
    if ((rc = make_map(&maps[8], d_dz2_lo ? d_dz2_lo : d_dz2_hi, n_rows, TC_BM))) return rc;
    if ((rc = make_map(&maps[9], d_dz1, n_rows, TC_BM))) return rc;
    TcTrainParams q;
    TcMlpParams& p = q.f;
    p.n_rows = n_rows; p.in = *in; p.d_in = shape->d_in; p.act = shape->act;
    p.w1_img_hi = d_packed + shape->off_w1_img_hi; p.w1_img_lo = d_packed + shape->off_w1_img_lo;
    p.b1 = d_params + shape->off_b1; p.b2 = d_params + shape->off_b2;
    p.xin_save = d_xin; p.h1_hi = d_h1_hi; p.h1_lo = d_h1_lo; p.h2 = d_h2;
    p.head_w = d_params + shape->off_w3; p.head_b = d_params + shape->off_b3; p.head_out = d_out;
    p.n_head = shape->n_out; p.skip = d_skip; p.timeline = g_timeline;
    p.loss_targets = nullptr; p.loss_idx = nullptr; p.loss_dout = nullptr; p.loss_ld = 1;
    p.loss_stats = nullptr; p.loss_stat_slot = 0; p.loss_count_rows = 0;
    q.idx = d_idx; q.targets = d_targets; q.log_scale = d_log_scale; q.actions = d_actions;
    q.advantages = d_advantages; q.old_log_probs = d_old_log_probs; q.ratio_clip = ratio_clip;
    q.entropy_coeff = entropy_coeff; q.stats = d_stats; q.dout = d_dout; q.ld_dout = ld_dout;
    q.dz2_hi = d_dz2_hi; q.dz2_lo = d_dz2_lo; q.dz1 = d_dz1;
    ProfScope prof_scope("tb_tc_mlp_train", stream);
    cudaStream_t s = as_stream(stream);
    const bool tanh_act = shape->act == TB_ACT_TANH;
#define TB_TRAIN(P_, L_)                                                                \
    { if (tanh_act) launch_tc_train<P_, TB_ACT_TANH, L_>(maps, q, s);                   \
      else launch_tc_train<P_, TB_ACT_RELU, L_>(maps, q, s); }
    if (passes == 3) { if (loss_kind == TC_LOSS_VALUE) TB_TRAIN(3, TC_LOSS_VALUE) else TB_TRAIN(3, TC_LOSS_POLICY) }
    else { if (loss_kind == TC_LOSS_VALUE) TB_TRAIN(1, TC_LOSS_VALUE) else TB_TRAIN(1, TC_LOSS_POLICY) }
#undef TB_TRAIN
    return check_launch("tb_tc_mlp_train");
}

extern "C" int tb_debug_plain_hi(int32_t on) {
    int v = on;
    TB_REQUIRE(cudaMemcpyToSymbol(tb::g_plain_hi, &v, sizeof(int)) == cudaSuccess, TB_ENOTSUP,
               "tb_debug_plain_hi: cudaMemcpyToSymbol failed");
    return 0;
}
