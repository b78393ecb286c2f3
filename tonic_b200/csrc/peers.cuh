// Symmetric-region layout of the gradient exchange that runs INSIDE the fused weight-gradient
// kernel (csrc/tc_gemm.cu, tc_wgrad_all_kernel) when several ranks train one replica each
// (SURVEY.md section 8e).  Push model with flagged words (the "LL" idea of NCCL's low-latency
// protocol): after the grid barrier CTA b of rank r owns parameter slice b; every thread stores
// its reduced element TOGETHER WITH the epoch tag as one 8-byte word into lane r of the slot of
// EVERY rank's region (NVLink stores, fire and forget -- no fence, no separate flag, no
// round trip).  The consumer then reads only LOCAL memory: it polls the words of its elements in
// lanes 0..world-1 until the tag matches and sums them in rank order -- every rank computes the
// same bits, replicas stay identical.
//
//     slot s (= epoch & 1): {float bits, tag} grad[8 lanes][n_pad] | {float bits, tag} sum[n_pad] |
//                           {half of a double, tag} stats[8 lanes][2 * TB_STAT_COUNT]
//
// More than two ranks: two hops instead of the all-to-all.  Element i goes only to its OWNER
// (groups of 32 elements, round robin over the ranks); the owner sums the lanes in rank order and
// stores the result into sum[i] of every rank -- 2 / world of the bytes, still a fixed order,
// still one value for all replicas.
//     tag = low 32 bits of epoch + 1 (the region starts zeroed; a tag recurs after 2^32 epochs
//     of the same slot parity, far beyond any run)
//
// A lane of slot s is rewritten two epochs later; the writer has by then consumed the reader's
// words of the epoch in between, which that rank sends only after its previous launch (the one
// that read the lane) has finished.
#pragma once
#include "common.cuh"

namespace tb {

constexpr int kPeerLanes = 8;

__host__ __device__ inline int peer_pad(int n_params) { return (n_params + 31) & ~31; }
__host__ __device__ inline size_t peer_fused_slot_bytes(int n_params) {
    return ((size_t)(kPeerLanes + 1) * peer_pad(n_params) * 8 + (size_t)kPeerLanes * 2 * TB_STAT_COUNT * 8 + 255) / 256 * 256;
}
// owner of element i in the two-phase exchange: warp-sized groups of 32 elements, round robin
__host__ __device__ inline int peer_owner(int i, int world) { return (i >> 5) % world; }
__device__ __forceinline__ uint2* peer_fused_grad(void* base, int slot, int lane, int n_params) {
    return reinterpret_cast<uint2*>(reinterpret_cast<char*>(base) + slot * peer_fused_slot_bytes(n_params)) +
           (size_t)lane * peer_pad(n_params);
}
// the reduced gradient the owners broadcast (two-phase exchange)
__device__ __forceinline__ uint2* peer_fused_sum(void* base, int slot, int n_params) {
    return peer_fused_grad(base, slot, kPeerLanes, n_params);
}
__device__ __forceinline__ uint2* peer_fused_stats(void* base, int slot, int lane, int n_params) {
    return reinterpret_cast<uint2*>(reinterpret_cast<char*>(base) + slot * peer_fused_slot_bytes(n_params) +
                                    (size_t)(kPeerLanes + 1) * peer_pad(n_params) * 8) + (size_t)lane * 2 * TB_STAT_COUNT;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// one 8-byte store: payload and tag become visible together
__device__ __forceinline__ void peer_put(uint2* dst, unsigned int bits, unsigned int tag) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(bits), "r"(tag) : "memory");
}
// Spin until the word carries `tag`, return its payload.  A peer that never arrives (crashed rank)
// would hang every GPU of the job: after 20 s the kernel traps instead, so the error surfaces on
// the host.
__device__ __forceinline__ unsigned int peer_get(const uint2* src, unsigned int tag) {
    unsigned int bits, seen;
    unsigned long long t0 = 0;
    for (unsigned int spins = 0;; ++spins) {
        asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(bits), "=r"(seen) : "l"(src) : "memory");
        if (seen == tag) return bits;
        if ((spins & 1023u) == 1023u) {
            const unsigned long long now = global_timer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 20000000000ull) asm volatile("trap;");
        }
    }
}

}  // namespace tb
