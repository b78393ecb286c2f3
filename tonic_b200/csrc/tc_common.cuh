// Shared pieces of the tensor-core kernels (csrc/tc_gemm.cu, csrc/tc_mlp.cu): tile
// constants, PTX wrappers for mbarrier / TMA / tcgen05, shared-memory matrix descriptors and
// the host-side tensor-map encoder.  sm_100a only.
#pragma once
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"

namespace tb {

constexpr int TC_BM = 128;      // rows per tile (UMMA M)
constexpr int TC_BN = 256;      // output columns (UMMA N) = hidden width
constexpr int TC_BK = 32;       // K chunk: 32 floats = 128 bytes = one swizzle row
constexpr int TC_K = 256;       // reduction length = hidden width
constexpr int TC_THREADS = 256;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;   // 16 KB
constexpr int TC_B_BYTES = TC_BN * TC_BK * 4;   // 32 KB
constexpr int TC_STAGE_ROWSTRIDE = 36;          // padded floats per staged row (bank-conflict free)

template <int PASSES>
struct TcCfg {
    static constexpr int PARTS = PASSES == 3 ? 2 : 1;
    static constexpr int STAGE_BYTES = PARTS * (TC_A_BYTES + TC_B_BYTES);
    static constexpr int STAGES = PASSES == 3 ? 2 : 4;
    static constexpr int EPI_BYTES = 4 * 32 * TC_STAGE_ROWSTRIDE * 4;
    static constexpr int HEAD_BYTES = (8 + 1) * TC_BN * 4;      // fused head weights [8, 256] + bias [256]
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + HEAD_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

// ---- PTX wrappers ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0),
        "r"(c1) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                                 uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns -> 32 registers per thread
__device__ __forceinline__ void tcgen05_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Read-only global load that the compiler may not reorder or sink (used to put a whole batch
// of independent loads in flight before the first use).
__device__ __forceinline__ float ldg_nc_volatile(const float* p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle, dense [rows x 128 B]
// tile (cute::UMMA::SmemDescriptor: start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46 |
// layout SWIZZLE_128B (2) <<61); SBO = 1024 B between 8-row groups.
__device__ __forceinline__ uint64_t umma_desc_kmajor_sw128(const void* smem) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(smem) >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                 // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset
    d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
    return d;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, K-major both,
// N = 256, M = 128.
constexpr uint32_t kIdescTf32 =
    (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

constexpr int TC_MAX_HEAD = 8;

// ---- thread-block cluster helpers (experimental 2-CTA variant of the fused kernels) ----------
// TMA load with multicast: the box lands at the same shared-memory offset of every CTA in
// `cta_mask` and completes the transaction count of the mbarrier at the same offset in each.
__device__ __forceinline__ void tma_load_2d_multicast(void* smem, const CUtensorMap* map, uint64_t* bar,
                                                      int c0, int c1, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
        "[%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0),
        "r"(c1), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void bulk_load_multicast(void* smem, const void* gmem, uint32_t bytes,
                                                    uint64_t* bar, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
        "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(smem)), "l"(gmem), "r"(bytes), "r"(smem_u32(bar)),
        "h"(cta_mask) : "memory");
}
// commit: arrive on the mbarrier at this offset in every CTA of `cta_mask`
__device__ __forceinline__ void tcgen05_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n"
                 "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// generic-proxy writes to shared memory (st.shared) -> visible to the async proxy
// (tcgen05.mma / TMA reading shared memory)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// Byte offset of element (row, 16-byte unit j) inside a K-major [rows x 128 B] operand tile
// with the 128-byte swizzle (what TMA's CU_TENSOR_MAP_SWIZZLE_128B writes for a box of 32
// floats: 8-row groups of 1024 B, unit index XOR-ed with the row within the group).
__device__ __forceinline__ uint32_t sw128_offset(int row, int unit) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((unit ^ (row & 7)) << 4));
}

// ---- host: tensor maps --------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// 2-D row-major float matrix [rows, 256]; box = [box_rows, 32 floats]; 128-byte swizzle.
static inline int make_map(CUtensorMap* map, const float* base, int64_t rows, int box_rows,
                    CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn fn = encode_fn();
    TB_REQUIRE(fn, TB_ENOTSUP, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)TC_K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)TC_K * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TB_REQUIRE(r == CUDA_SUCCESS, TB_EINVAL, "cuTensorMapEncodeTiled failed with %d", (int)r);
    return 0;
}

// The same [rows, 256] matrix seen as (32 columns, rows, 8 column groups): ONE box of
// (32, box_rows, n_groups) lands in shared memory as n_groups consecutive [box_rows x 128 B]
// blocks -- the MN-major operand layout (SWIZZLE_128B_ATOM_32B) that otherwise takes one 2-D
// box per 32-column group (TMA issue, not bandwidth, bounded the weight-gradient mainloop).
static inline int make_map_groups(CUtensorMap* map, const float* base, int64_t rows, int box_rows,
                                  int n_groups) {
    EncodeTiledFn fn = encode_fn();
    TB_REQUIRE(fn, TB_ENOTSUP, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[3] = {32, (cuuint64_t)rows, (cuuint64_t)(TC_K / 32)};
    cuuint64_t strides[2] = {(cuuint64_t)TC_K * sizeof(float), 32 * sizeof(float)};
    cuuint32_t box[3] = {32, (cuuint32_t)box_rows, (cuuint32_t)n_groups};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;      // the caller falls back to per-group 2-D boxes
}

__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0),
        "r"(c1), "r"(c2) : "memory");
}

// 2-D row-major float matrix [rows, 256]; box = [box_cols floats, box_rows]; no swizzle (plain
// row-major tile in shared memory).
static inline int make_map_plain(CUtensorMap* map, const float* base, int64_t rows, int box_cols,
                                 int box_rows) {
    EncodeTiledFn fn = encode_fn();
    TB_REQUIRE(fn, TB_ENOTSUP, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)TC_K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)TC_K * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TB_REQUIRE(r == CUDA_SUCCESS, TB_EINVAL, "cuTensorMapEncodeTiled (plain) failed with %d", (int)r);
    return 0;
}

}  // namespace tb
