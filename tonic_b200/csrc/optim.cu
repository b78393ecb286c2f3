// K9 Adam step (+ partial-sum reduction of the split weight gradients, refresh of
// the packed forward transposes, and the device-side PPO control flags) and K15
// target-network soft update.
//
// Reference: torch.optim.Adam (single-tensor path, torch 2.11) as constructed at
// tonic/torch/updaters/actors.py:11-12,58-59,161-162,228-229 and
// tonic/torch/updaters/critics.py:9-10,59-60,143-144,190-191 (betas 0.9/0.999,
// eps 1e-8, no weight decay); early stop tonic/torch/agents/ppo.py:45-46 with
// updaters/actors.py:103,112; soft update models/actor_critics.py:68-72,126-130.
#include "adam.cuh"
#include "peers.cuh"

namespace tb {

__global__ void __launch_bounds__(256)
adam_kernel(TbAdam opt, TbMlpShape sh, float* __restrict__ packed,
            const float* __restrict__ gpart, int n_split_all, int n_split_w2, float grad_scale,
            const int32_t* d_skip, const double* d_stats, float kl_threshold, int32_t* d_stop) {
    if (skip_requested(d_skip)) return;
    if (d_stats && d_stats[TB_STAT_NONZERO_ADV] == 0.0) return;    // actors.py:22,71: no step
    const int t = opt.d_step[0] + 1;                               // incremented by the last block
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    // bias corrections once per block (two double-precision pow() per thread were a third of
    // this kernel's instructions)
    __shared__ float s_step_size, s_bc2_sqrt;
    if (threadIdx.x == 0) adam_corrections(opt, t, &s_step_size, &s_bc2_sqrt);
    __syncthreads();
    if (i < opt.n_params) {
        // one parameter per thread keeps ~72k threads in flight; the n_split partial loads
        // of a thread are independent (4-way unrolled sums)
        // the W2 block may have been produced with fewer row splits than the narrow gradients
        // (W2 and, with the tensor-core kernel, b2 right behind it: tb_tc_wgrad256)
        const int w2_end = sh.off_w2 + sh.hidden * sh.hidden + (sh.off_w2_hi > 0 ? sh.hidden : 0);
        const int n_split = (n_split_w2 > 0 && i >= sh.off_w2 && i < w2_end) ? n_split_w2 : n_split_all;
        float g0 = 0.0f, g1 = 0.0f, g2 = 0.0f, g3 = 0.0f;
        int s = 0;
        // 16 independent loads in flight per thread (the sum order stays s mod 4 -> g0..g3)
        for (; s + 16 <= n_split; s += 16) {
            float v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) v[u] = __ldg(gpart + (size_t)(s + u) * opt.n_params + i);
#pragma unroll
            for (int u = 0; u < 16; u += 4) { g0 += v[u]; g1 += v[u + 1]; g2 += v[u + 2]; g3 += v[u + 3]; }
        }
        for (; s + 4 <= n_split; s += 4) {
            g0 += gpart[(size_t)(s + 0) * opt.n_params + i];
            g1 += gpart[(size_t)(s + 1) * opt.n_params + i];
            g2 += gpart[(size_t)(s + 2) * opt.n_params + i];
            g3 += gpart[(size_t)(s + 3) * opt.n_params + i];
        }
        for (; s < n_split; ++s) g0 += gpart[(size_t)s * opt.n_params + i];
        const float g = ((g0 + g1) + (g2 + g3)) * grad_scale;
        adam_apply(opt, sh, packed, i, g, s_step_size, s_bc2_sqrt);
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
                            const float* d_gpart, int32_t n_split, int32_t n_split_w2, float grad_scale,
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
        *opt, sh, d_packed, d_gpart, n_split, shape ? n_split_w2 : 0, grad_scale, d_skip, d_stats,
        kl_threshold, d_stop);
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
reduce_partials_kernel(const float* __restrict__ gpart, int n_split_all, int n_split_w2, int w2_lo,
                       int w2_hi, int n, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int n_split = (n_split_w2 > 0 && i >= w2_lo && i < w2_hi) ? n_split_w2 : n_split_all;
    float g = 0.0f;
    for (int s = 0; s < n_split; ++s) g += gpart[(size_t)s * n + i];
    out[i] = g;
}
}  // namespace tb

extern "C" int tb_reduce_partials(const float* d_gpart, int32_t n_split, int32_t n_split_w2,
                                  int32_t w2_begin, int32_t w2_end, int32_t n_params,
                                  float* d_out, const int32_t* d_skip, void* stream) {
    tb::ProfScope prof_scope("tb_reduce_partials", stream);
    TB_REQUIRE(d_gpart && d_out && n_split >= 1 && n_params > 0, TB_EINVAL,
               "tb_reduce_partials: bad arguments");
    (void)d_skip;
    tb::reduce_partials_kernel<<<(n_params + 255) / 256, 256, 0, tb::as_stream(stream)>>>(
        d_gpart, n_split, n_split_w2, w2_begin, w2_end, n_params, d_out);
    return tb::check_launch("tb_reduce_partials");
}

// =====================================================================================
// Global-norm gradient clipping (torch.nn.utils.clip_grad_norm_ at
// tonic/torch/updaters/actors.py:37-38,96-98,176-177,256-257 and critics.py:24-25,82-83,
// 177-178,230-231): the flat (summed) gradient of every network of an updater adds its
// sum of squares to one double; the coefficient min(1, max_norm / (norm + 1e-6)) then
// scales each flat gradient before the Adam step.  One block, fixed summation order.
// =====================================================================================
namespace tb {
__global__ void __launch_bounds__(1024)
grad_sqnorm_kernel(const float* __restrict__ g, int n, double* __restrict__ sumsq, const int32_t* d_skip) {
    if (skip_requested(d_skip)) return;
    __shared__ double scratch[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += (double)g[i] * (double)g[i];
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) *sumsq += acc;
}

__global__ void __launch_bounds__(256)
grad_clip_kernel(float* __restrict__ g, int n, const double* __restrict__ sumsq, float grad_scale,
                 float max_norm, const int32_t* d_skip) {
    if (skip_requested(d_skip)) return;
    // the norm of the MEAN gradient (the partial sums are scaled by grad_scale in the Adam kernel)
    const float total_norm = (float)(sqrt(*sumsq) * (double)grad_scale);
    const float coef = fminf(max_norm / (total_norm + 1e-6f), 1.0f);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) g[i] *= coef;
}
}  // namespace tb

extern "C" int tb_grad_sqnorm(const float* d_grad, int32_t n, double* d_sumsq, const int32_t* d_skip,
                              void* stream) {
    tb::ProfScope prof_scope("tb_grad_sqnorm", stream);
    TB_REQUIRE(d_grad && d_sumsq && n > 0, TB_EINVAL, "tb_grad_sqnorm: bad arguments");
    tb::grad_sqnorm_kernel<<<1, 1024, 0, tb::as_stream(stream)>>>(d_grad, n, d_sumsq, d_skip);
    return tb::check_launch("tb_grad_sqnorm");
}

extern "C" int tb_grad_clip(float* d_grad, int32_t n, const double* d_sumsq, float grad_scale,
                            float max_norm, const int32_t* d_skip, void* stream) {
    tb::ProfScope prof_scope("tb_grad_clip", stream);
    TB_REQUIRE(d_grad && d_sumsq && n > 0 && max_norm > 0.0f, TB_EINVAL, "tb_grad_clip: bad arguments");
    tb::grad_clip_kernel<<<(n + 255) / 256, 256, 0, tb::as_stream(stream)>>>(d_grad, n, d_sumsq, grad_scale,
                                                                             max_norm, d_skip);
    return tb::check_launch("tb_grad_clip");
}

// =====================================================================================
// Fused gradient all-reduce + Adam over NVLink peer memory (SURVEY.md section 8e, K9+K10).
//
// Every rank owns a "symmetric" region (same layout on every GPU, mapped into every
// process: torch.distributed._symmetric_memory):
//     flags  uint64 [2 slots][8 ranks]            (256 B)
//     slot s: float grad[n_params] | double stats[TB_STAT_COUNT]     (s = epoch & 1)
// tb_peer_publish : flat gradient (sum of the split partials) and the statistics block of
//                   this rank -> its own slot; then a release-store of epoch+1 into
//                   flags[slot][rank] of EVERY peer (st.release.sys over NVLink).
// tb_adam_step_peers: waits until all flags of the slot show epoch+1, sums the slot of
//                   every rank with direct peer loads in RANK ORDER (identical result on all
//                   ranks -> replicas stay bit-identical), runs the Adam update, writes the
//                   global statistics back and advances the epoch.
// A slot is rewritten two epochs later; the flag wait of the epoch in between orders that
// write after every peer's reads.  Replaces 2 NCCL all-reduces (~30 us each at this
// message size) + 1 reduction kernel per network and minibatch.
// =====================================================================================
namespace tb {

constexpr int kMaxPeers = 8;
constexpr int kPeerFlagBytes = 256;

__host__ __device__ inline size_t peer_slot_bytes(int n_params) {
    return ((size_t)n_params * 4 + TB_STAT_COUNT * 8 + 255) / 256 * 256;
}
__device__ __forceinline__ unsigned long long* peer_flags(void* base, int slot) {
    return reinterpret_cast<unsigned long long*>(base) + slot * kMaxPeers;
}
__device__ __forceinline__ float* peer_grad(void* base, int slot, int n_params) {
    return reinterpret_cast<float*>(reinterpret_cast<char*>(base) + kPeerFlagBytes +
                                    slot * peer_slot_bytes(n_params));
}
__device__ __forceinline__ double* peer_stats(void* base, int slot, int n_params) {
    return reinterpret_cast<double*>(reinterpret_cast<char*>(peer_grad(base, slot, n_params)) +
                                     ((size_t)n_params * 4 + 15) / 16 * 16);
}

__global__ void __launch_bounds__(256)
peer_publish_kernel(TbPeers peers, const float* __restrict__ gpart, int n_split_all, int n_split_w2,
                    int w2_lo, int w2_hi, int n_params,
                    const double* __restrict__ stats, const unsigned long long* d_epoch,
                    int* block_counter, const int32_t* d_skip) {
    if (skip_requested(d_skip)) return;
    const unsigned long long epoch = *d_epoch;
    const int slot = (int)(epoch & 1);
    void* mine = peers.base[peers.rank];
    float* flat = peer_grad(mine, slot, n_params);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_params) {
        float g = 0.0f;
        const int n_split = (n_split_w2 > 0 && i >= w2_lo && i < w2_hi) ? n_split_w2 : n_split_all;
        if (gpart)
            for (int s = 0; s < n_split; ++s) g += gpart[(size_t)s * n_params + i];
        flat[i] = g;
    }
    if (blockIdx.x == 0 && threadIdx.x < TB_STAT_COUNT)
        peer_stats(mine, slot, n_params)[threadIdx.x] = stats ? stats[threadIdx.x] : 0.0;
    // last block to finish: make this rank's slot visible system-wide, then raise its flag
    // in every peer's region
    __shared__ bool is_last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(block_counter, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (is_last) {
        if (threadIdx.x == 0) *block_counter = 0;
        __threadfence_system();
        if ((int)threadIdx.x < peers.world) {
            unsigned long long* flag = peer_flags(peers.base[threadIdx.x], slot) + peers.rank;
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(epoch + 1) : "memory");
        }
    }
}

__global__ void __launch_bounds__(256)
adam_peers_kernel(TbAdam opt, TbMlpShape sh, float* __restrict__ packed, TbPeers peers,
                  float grad_scale, unsigned long long* d_epoch, const int32_t* d_skip,
                  double* d_stats, int use_stats, float kl_threshold, int32_t* d_stop) {
    if (skip_requested(d_skip)) return;
    const unsigned long long epoch = *d_epoch;
    const int slot = (int)(epoch & 1);
    const int n_params = opt.n_params;
    __shared__ double s_stats[TB_STAT_COUNT];
    // wait for the flag of every rank (one lane per rank), then gather the global statistics
    if ((int)threadIdx.x < peers.world) {
        const unsigned long long* flag = peer_flags(peers.base[peers.rank], slot) + threadIdx.x;
        unsigned long long seen;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(flag) : "memory");
        } while (seen < epoch + 1);
    }
    __syncthreads();
    if (threadIdx.x < TB_STAT_COUNT) {
        double s = 0.0;
        for (int r = 0; r < peers.world; ++r)
            s += __ldcv(peer_stats(peers.base[r], slot, n_params) + threadIdx.x);
        s_stats[threadIdx.x] = s;
        if (blockIdx.x == 0 && d_stats) d_stats[threadIdx.x] = s;      // global statistics
    }
    __shared__ float s_step_size, s_bc2_sqrt;
    const int t = opt.d_step[0] + 1;
    if (threadIdx.x == 0) adam_corrections(opt, t, &s_step_size, &s_bc2_sqrt);
    __syncthreads();
    const bool do_step = !(use_stats && s_stats[TB_STAT_NONZERO_ADV] == 0.0);   // actors.py:22,71
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (do_step && i < n_params) {
        float g = 0.0f;
        for (int r = 0; r < peers.world; ++r) g += __ldcv(peer_grad(peers.base[r], slot, n_params) + i);
        g *= grad_scale;
        adam_apply(opt, sh, packed, i, g, s_step_size, s_bc2_sqrt);
    }
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(&opt.d_step[1], 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        opt.d_step[1] = 0;
        if (do_step) opt.d_step[0] = t;
        *d_epoch = epoch + 1;
        if (do_step && use_stats && d_stop && kl_threshold >= 0.0f) {
            const float kl = (float)(s_stats[TB_STAT_KL] / s_stats[TB_STAT_ROWS]);
            if (kl > kl_threshold) *d_stop = 1;
        }
    }
}

}  // namespace tb

extern "C" int64_t tb_peer_region_bytes(int32_t n_params) {
    return (int64_t)(tb::kPeerFlagBytes + 2 * tb::peer_slot_bytes(n_params));
}

extern "C" int64_t tb_peer_region_bytes_fused(int32_t n_params) {
    return (int64_t)(2 * tb::peer_fused_slot_bytes(n_params));
}

extern "C" int tb_peer_publish(const TbPeers* peers, const float* d_gpart, int32_t n_split,
                               int32_t n_split_w2, int32_t w2_begin, int32_t w2_end,
                               int32_t n_params, const double* d_stats, const uint64_t* d_epoch,
                               int32_t* d_block_counter, const int32_t* d_skip, void* stream) {
    tb::ProfScope prof_scope("tb_peer_publish", stream);
    TB_REQUIRE(peers && peers->world >= 1 && peers->world <= tb::kMaxPeers && d_epoch &&
               d_block_counter && n_params > 0 && n_split >= 0, TB_EINVAL,
               "tb_peer_publish: bad arguments");
    tb::peer_publish_kernel<<<(n_params + 255) / 256, 256, 0, tb::as_stream(stream)>>>(
        *peers, d_gpart, n_split, n_split_w2, w2_begin, w2_end, n_params, d_stats,
        reinterpret_cast<const unsigned long long*>(d_epoch), d_block_counter, d_skip);
    return tb::check_launch("tb_peer_publish");
}

extern "C" int tb_adam_step_peers(const TbAdam* opt, const TbMlpShape* shape, float* d_packed,
                                  const TbPeers* peers, float grad_scale, uint64_t* d_epoch,
                                  const int32_t* d_skip, double* d_stats, int32_t use_stats,
                                  float kl_threshold, int32_t* d_stop, void* stream) {
    tb::ProfScope prof_scope("tb_adam_step_peers", stream);
    TB_REQUIRE(opt && opt->d_params && opt->d_m && opt->d_v && opt->d_step && peers && d_epoch &&
               peers->world >= 1 && peers->world <= tb::kMaxPeers, TB_EINVAL,
               "tb_adam_step_peers: bad arguments");
    TbMlpShape sh;
    if (shape) sh = *shape; else memset(&sh, 0, sizeof(sh));
    tb::adam_peers_kernel<<<(opt->n_params + 255) / 256, 256, 0, tb::as_stream(stream)>>>(
        *opt, sh, d_packed, *peers, grad_scale, reinterpret_cast<unsigned long long*>(d_epoch),
        d_skip, d_stats, use_stats, kl_threshold, d_stop);
    return tb::check_launch("tb_adam_step_peers");
}
