/*
 * tonic_b200 -- C ABI of the B200-native (sm_100a) backend for the data-parallel
 * hot path of fabiopardo/tonic: the synchronous vector-environment collector and
 * the PPO / A2C / DDPG / TD3 / SAC update loops with their replays.
 *
 * The reference has no FFI of its own (the boundary is a
 * duck-typed Python protocol).  This header is the drop-in boundary one level
 * below that protocol: one entry point per reference operation on the path, each
 * citing the reference code it replaces (paths relative to the reference's
 * `tonic/` package).  The Python classes in `tonic_b200/` that mirror the
 * reference's classes call these through ctypes.
 *
 * Conventions
 *   - every pointer named d_* is a DEVICE pointer, h_* a HOST pointer;
 *   - all floating-point data are float32, row-major; "rows" are transitions
 *     (flat index t*N+n, replays/utils.py:22-25);
 *   - no entry point allocates device memory or synchronises the device; work is
 *     enqueued on `stream` (a cudaStream_t passed as void*), so calls can be
 *     captured in CUDA graphs;
 *   - return value: 0 on success, a cudaError_t (>0) or a negative TB_E* code
 *     otherwise; tb_last_error() gives a readable message.
 */
#ifndef TONIC_B200_H
#define TONIC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TB_VERSION 1

#define TB_EINVAL (-1)       /* bad argument / unsupported shape            */
#define TB_ENOTSUP (-2)      /* configuration outside the implemented subset */

/* activation ids (models/utils.py:15-16 passes torch.nn.Tanh / ReLU)         */
#define TB_ACT_TANH 0
#define TB_ACT_RELU 1

int tb_version(void);
const char* tb_last_error(void);
/* Number of kernels launched by this library since process start (bench.py's
 * `gpu_launches`).                                                            */
int64_t tb_launch_count(void);

/* Per-entry-point device timing with CUDA events recorded on the launching
 * stream (bench.py's live roofline).  tb_profile_end synchronises the device and
 * writes "name launches total_ms" lines into h_buf.                           */
int tb_profile_begin(void);
int tb_profile_end(char* h_buf, int32_t size);

/* ------------------------------------------------------------------------ */
/* Vector environment  -- replaces environments/distributed.py:8-58           */
/* (Sequential.{initialize,start,step}) and the ActionRescaler clip           */
/* (environments/wrappers.py:18-22) for the synthetic SynthControl(O,A) env.  */
/* ------------------------------------------------------------------------ */
typedef struct {
    int32_t n_envs;            /* N workers on this device                    */
    int32_t obs_dim;           /* O                                           */
    int32_t act_dim;           /* A                                           */
    int32_t max_episode_steps; /* time-limit: reset, NOT a termination (:40)  */
    int64_t seed;              /* env i uses seed + first_worker + i (:18-20) */
    int64_t first_worker;      /* global index of this device's first worker  */
    float* d_state;            /* [N,O] current state x                       */
    int32_t* d_length;         /* [N]  steps in the running episode (:25,38)  */
    uint32_t* d_episode;       /* [N]  episodes started so far                */
    double* d_score;           /* [N]  running episode score (trainer.py:37)  */
    /* finished-episode log (trainer.py:64-71): ring of capacity log_cap      */
    double* d_ep_scores;       /* [log_cap]                                   */
    int32_t* d_ep_lengths;     /* [log_cap]                                   */
    unsigned long long* d_ep_count; /* [1] total finished episodes            */
    int32_t log_cap;
    /* TimeFeature wrapper (environments/wrappers.py:25-54): observation rows get one
     * more column, low + (high - low) * episode_steps / max_episode_steps (low after
     * a reset); 0 = off.  Observation arrays are then [N, obs_dim + 1].             */
    int32_t time_feature;
    float time_low, time_high;
    /* Task: TB_TASK_SYNTH = SynthControl(O, A) (d_state, float32 [N, O]); the closed-form
     * classic-control tasks behind Gym(name) (environments/builders.py:12-16) keep a float64
     * state [N, 2] in d_state64: TB_TASK_PENDULUM (theta, theta_dot; obs 3, act 1),
     * TB_TASK_MOUNTAIN_CAR (position, velocity; obs 2, act 1).  Their ActionRescaler bounds
     * (wrappers.py:7-22) are part of the task.                                            */
    int32_t task;
    double* d_state64;
} TbEnv;
enum { TB_TASK_SYNTH = 0, TB_TASK_PENDULUM = 1, TB_TASK_MOUNTAIN_CAR = 2 };

/* Sequential.start (:22-26): reset every env, lengths=0, writes obs [N,O]
 * ([N,O+1] with the time feature).                                           */
int tb_env_start(const TbEnv* env, float* d_obs, void* stream);

/* Sequential.step (:28-58).  d_actions [N,A].  Outputs: d_obs [N,O] = acting
 * observations (post auto-reset), d_next_obs [N,O] = transition observations
 * (pre-reset), d_rewards/d_resets/d_terminations [N] float32 (0/1 flags, the
 * dtype the replays store: replays/segments.py:33).                          */
int tb_env_step(const TbEnv* env, const float* d_actions, float* d_obs,
                float* d_next_obs, float* d_rewards, float* d_resets,
                float* d_terminations, void* stream);

/* One vector step of the on-policy collector after the actor forward pass in ONE launch
 * (tonic/utils/trainer.py:44-50): tb_gauss_sample (a2c.py:75-85; d_eps == NULL: Philox stream at
 * counter + *d_counter + env index) -> d_actions [N,A] (unclipped, what the replay stores) and
 * d_log_probs [N]; tb_moments_record of the acting observations (the environment state rows;
 * d_moment_sums NULL = no normaliser); tb_env_step with the clipped actions.  SynthControl
 * without the time feature only.  The caller advances the stream position (tb_counter_add)
 * once per rollout, not per step.                                                        */
int tb_act_env_step(const TbEnv* env, const float* d_loc_pre, const float* d_log_scale,
                    uint64_t seed, uint64_t counter, const uint64_t* d_counter,
                    float* d_actions, float* d_log_probs, double* d_moment_sums,
                    float* d_obs, float* d_next_obs, float* d_rewards, float* d_resets,
                    float* d_terminations, void* stream);

/* Fused rollout of a whole on-policy segment (T vector steps) in ONE launch: per step
 * the actor forward + Normal sample + log-prob (torch/agents/a2c.py:41-52,75-85), the
 * environment transition with auto-reset (environments/distributed.py:28-58), the segment
 * rows (replays/segments.py:27-36) and the normaliser sums (mean_stds.py:44-48); the
 * caller of the reference loop is utils/trainer.py:44-50.  A CTA keeps 64 environments
 * resident in shared memory for all T steps.  Segment buffers are [T, N, ...] float32;
 * `shape` / d_params / d_packed describe the actor MLP (declared below); d_counter is the
 * device-resident Philox position, advanced by the caller by T * counter_stride.       */
struct TbMlpShape_;
int tb_rollout_fused(const TbEnv* env, const struct TbMlpShape_* shape, const float* d_params,
                     const float* d_packed, const float* d_log_scale, const float* d_norm_mean,
                     const float* d_norm_std, int32_t T, float* d_seg_obs, float* d_seg_actions,
                     float* d_seg_next_obs, float* d_seg_rewards, float* d_seg_resets,
                     float* d_seg_terms, float* d_seg_logp, float* d_env_obs,
                     double* d_moment_sums, uint64_t seed, uint64_t counter,
                     uint64_t counter_stride, const uint64_t* d_counter, void* stream);

/* ------------------------------------------------------------------------ */
/* Observation normaliser -- replaces torch/normalizers/mean_stds.py:44-74    */
/* ------------------------------------------------------------------------ */
/* MeanStd.record (:44-48): d_sums[0:dim] += sum_n x, d_sums[dim:2dim] += sum_n
 * x^2 (float64 accumulators), d_sums[2*dim] += n.                            */
int tb_moments_record(const float* d_x, int64_t n_rows, int32_t dim,
                      double* d_sums, void* stream);
/* MeanStd.update (:50-74): merges the recorded sums into the running
 * mean / mean_sq (float32, d_running = [mean(dim) | mean_sq(dim)]), total count
 * in d_count[0], writes d_mean/d_std (the `_mean`/`_std` parameters), clears
 * d_sums.  eps = 1e-2 (:15,65-70).  No-op when nothing was recorded.          */
int tb_moments_update(double* d_sums, float* d_running, double* d_count,
                      float* d_mean, float* d_std, int32_t dim, float eps,
                      void* stream);

/* ------------------------------------------------------------------------ */
/* Segment replay helpers -- replays/utils.py:4-19, replays/segments.py:38-78 */
/* ------------------------------------------------------------------------ */
/* lambda_returns (utils.py:4-19): all arrays [T,N]; reverse scan over T.     */
int tb_lambda_returns(const float* d_values, const float* d_next_values,
                      const float* d_rewards, const float* d_resets,
                      const float* d_terminations, float* d_returns,
                      int32_t T, int32_t N, double discount_factor,
                      double trace_decay, void* stream);
/* Segment.get_full advantage normalisation (segments.py:41-46):
 * adv = returns - values; if std(adv) != 0: adv = (adv - mean) / std
 * (population std, two-pass).  d_workspace: >= 4 doubles.  When `world` > 1 the
 * caller all-reduces d_workspace[0:2] between phase 1 and 2 and d_workspace[2]
 * between phase 2 and 3; single-GPU callers use phase 0 (= all).  */
int tb_advantages(const float* d_returns, const float* d_values,
                  float* d_advantages, int64_t n, double* d_workspace,
                  int64_t n_global, int32_t phase, void* stream);

/* ------------------------------------------------------------------------ */
/* Two-hidden-layer MLP (torch/models/utils.py:4-23) with a linear head        */
/* ------------------------------------------------------------------------ */
typedef struct TbMlpShape_ {
    int32_t d_in;      /* input width (obs, or obs+act for Q critics)         */
    int32_t hidden;    /* H: both hidden layers (64, 128 or 256)              */
    int32_t n_out;     /* head rows (A, 2A for loc+scale heads, 1 for values) */
    int32_t act;       /* TB_ACT_*                                            */
    /* float offsets into the flat parameter buffer (each multiple of 4)      */
    int32_t off_w1, off_b1, off_w2, off_b2, off_w3, off_b3;
    int32_t n_params;  /* padded length of the flat buffer                    */
    /* float offsets into the packed buffer (transposes for the forward pass) */
    int32_t off_w1t, off_w2t;
    int32_t n_packed;
    /* tensor-core path (hidden == 256 only; 0 = absent): tf32 splits of W2 [H, H]
     * (B operand of the forward GEMM) and of W2^T (B operand of the backward GEMM) */
    int32_t off_w2_hi, off_w2_lo, off_w2t_hi, off_w2t_lo;
    /* fused forward kernel (hidden == 256 and d_in <= 32; 0 = absent): W1 as the layer-1
     * B operand, i.e. the byte image of the [256 x 32] K-major 128B-swizzled shared-memory
     * tile (tf32 hi / lo parts, zero-padded columns), 8192 floats each, 16-byte aligned  */
    int32_t off_w1_img_hi, off_w1_img_lo;
} TbMlpShape;

typedef struct {
    /* source 1: rows of width dim1, optionally gathered through d_idx and
     * normalised as (x - mean) / std (torch/normalizers/mean_stds.py:34-39)  */
    const float* d_x1;
    int32_t dim1;
    const float* d_mean;   /* NULL = no normalisation (actor quirk) */
    const float* d_std;
    /* source 2 (optional): rows of width dim2 concatenated after source 1
     * (models/encoders.py:28-31), gathered through d_idx iff gather2 != 0     */
    const float* d_x2;
    int32_t dim2;
    int32_t gather2;
    const int64_t* d_idx;  /* NULL = rows are 0..n_rows-1                     */
} TbMlpInput;

/* Forward pass for n_rows rows.  d_out [n_rows, n_out] = head pre-activations.
 * Optional saves for the backward pass (NULL to skip): d_xin [n_rows, ldx] the
 * assembled input with a trailing 1 column (ldx = round_up(d_in + 1, 4)),
 * d_h1 / d_h2 [n_rows, H].  d_skip (optional, device int32): when *d_skip != 0
 * the kernel exits immediately (device-side early stop, ppo.py:45-46).        */
int tb_mlp_forward(const TbMlpShape* shape, const float* d_params,
                   const float* d_packed, const TbMlpInput* in, int64_t n_rows,
                   float* d_out, float* d_xin, float* d_h1, float* d_h2,
                   const int32_t* d_skip, void* stream);

/* Backward pass: given d_dout [n_rows, ld_dout >= n_out] (gradient w.r.t. head
 * pre-activations) and the saved h1/h2, writes d_dz2, d_dz1 [n_rows, H]
 * (gradients w.r.t. hidden pre-activations) and, if d_dx != NULL, the gradient
 * w.r.t. input columns [dx_col0, dx_col0 + dx_cols) into d_dx [n_rows, dx_cols]
 * (used for dQ/da, updaters/actors.py:177-181,254-257).                       */
int tb_mlp_backward(const TbMlpShape* shape, const float* d_params,
                    const float* d_dout, int32_t ld_dout, const float* d_h1, const float* d_h2,
                    int64_t n_rows, float* d_dz2, float* d_dz1, float* d_dx,
                    int32_t dx_col0, int32_t dx_cols, const int32_t* d_skip,
                    void* stream);

/* Weight gradients: per-split partial sums over rows, written (not
 * accumulated) to d_gpart [n_split, n_params] in the flat parameter layout.
 * `d_dout` has ld_dout >= n_out columns; columns [n_out, n_out + n_extra) are
 * summed over rows into the flat buffer at off_extra (e.g. the per-sample
 * log_scale gradients of the detached-scale Gaussian head).                   */
int tb_mlp_wgrad(const TbMlpShape* shape, const float* d_xin, const float* d_h1,
                 const float* d_h2, const float* d_dz1, const float* d_dz2,
                 const float* d_dout, int32_t ld_dout, int32_t n_extra,
                 int32_t off_extra, int64_t n_rows, float* d_gpart,
                 int32_t n_split, const int32_t* d_skip, void* stream);

/* ---- tensor-core path (tcgen05.mma kind::tf32, TMA, TMEM) for the 256-wide ---- */
/* hidden-layer GEMMs of the MLP above (models/utils.py:15-23 and its autograd)    */
/* hi = x with the low 13 mantissa bits cleared (tf32-exact), lo = x - hi.        */
int tb_split_tf32(const float* d_x, float* d_hi, float* d_lo, int64_t n, void* stream);

/* out[n_rows, 256] = epilogue(A[n_rows, 256] . B[256, 256]^T), A and B given as
 * tf32 splits (hi [+ lo]); passes = 3: a_hi.b_hi + a_lo.b_hi + a_hi.b_lo (fp32
 * grade), passes = 1: plain TF32.  epilogue 0: act(. + bias) (forward layer 2,
 * B = W2); 1: . * act'(aux_hi + aux_lo) (backward dz1, B = W2^T); 2: none.
 * d_out_lo != NULL: the result is written as a tf32 split (d_out = hi).
 * n_head in [1, 8] (epilogue 0 only): additionally d_head_out[row, o] = d_head_b[o] +
 * out[row, :] . d_head_w[o, :], the linear head fused into the epilogue.          */
int tb_tc_gemm256(const float* d_a_hi, const float* d_a_lo, const float* d_b_hi,
                  const float* d_b_lo, int64_t n_rows, int32_t passes, int32_t epilogue,
                  int32_t act, const float* d_bias, const float* d_aux_hi,
                  const float* d_aux_lo, float* d_out, float* d_out_lo,
                  const float* d_head_w, const float* d_head_b, float* d_head_out,
                  int32_t n_head, const int32_t* d_skip, void* stream);

/* Whole forward pass of a 2 x 256 MLP with d_in <= 32 and n_out <= 8 as ONE tensor-core
 * kernel (csrc/tc_mlp.cu): input gather / normalisation (encoders.py:11-30), both hidden
 * layers (models/utils.py:15-23) and the linear head; d_xin / d_h1_hi / d_h1_lo / d_h2 may be
 * NULL when no backward pass follows (value / action evaluation).  tb_mlp_forward_tc routes
 * here when the shape allows it.                                                          */
int tb_tc_mlp_forward(const TbMlpShape* shape, const float* d_params, const float* d_packed,
                      const TbMlpInput* in, int64_t n_rows, float* d_out, float* d_xin,
                      float* d_h1_hi, float* d_h1_lo, float* d_h2, int32_t passes,
                      const int32_t* d_skip, void* stream);

/* tb_tc_mlp_forward for a single-output value head with the squared-error loss of
 * VRegression (torch/updaters/critics.py:18-28) fused into its epilogue: the arguments of
 * tb_mse_loss (targets gathered by d_idx, dout = 2 (v - target), sums into the TB_STAT_*
 * block) in addition to those of tb_tc_mlp_forward; n_out must be 1.                     */
int tb_tc_mlp_forward_vloss(const TbMlpShape* shape, const float* d_params, const float* d_packed,
                            const TbMlpInput* in, int64_t n_rows, float* d_out, float* d_xin,
                            float* d_h1_hi, float* d_h1_lo, float* d_h2, int32_t passes,
                            const float* d_targets, const int64_t* d_idx, float* d_dout,
                            int32_t ld_dout, double* d_stats, int32_t stat_slot,
                            int32_t count_rows, const int32_t* d_skip, void* stream);

/* Activation gradients of the same network as ONE tensor-core kernel (csrc/tc_mlp.cu):
 * dz2 = (dout W3) * act'(h2) and dz1 = (dz2 W2) * act'(h1), the autograd of the head and of
 * the second hidden layer behind loss.backward() (torch/updaters/actors.py:33,104,186,
 * critics.py:24,84); n_out <= 8.  dz2 is written as a tf32 split, dz1 as float32, both
 * [n_rows, 256], for the weight-gradient kernels.  tb_mlp_backward_tc routes here.          */
int tb_tc_mlp_backward(const TbMlpShape* shape, const float* d_params, const float* d_packed,
                       const float* d_dout, int32_t ld_dout, const float* d_h1_hi,
                       const float* d_h1_lo, const float* d_h2, int64_t n_rows, float* d_dz2_hi,
                       float* d_dz2_lo, float* d_dz1, int32_t passes, const int32_t* d_skip,
                       void* stream);

/* Forward -> loss -> backward of one network-minibatch in ONE launch (tcgen05, one 128-row
 * tile per CTA): tb_tc_mlp_forward (training mode) + the loss kernel + tb_tc_mlp_backward.
 *   loss_kind 0: value regression, dout = 2 (v - d_targets[d_idx[row]])    (updaters/critics.py:18-28)
 *   loss_kind 1: clipped-ratio (ratio_clip > 0) / policy-gradient loss of the detached-scale
 *                Gaussian head, arithmetic of tb_gauss_policy_loss      (updaters/actors.py:21-50,70-112)
 * z2 stays in TMEM (h2 is recomputed for the backward pass), the head output and its gradient stay
 * on the SM.  Written for the weight-gradient kernel: d_xin, h1 (d_h1_hi / d_h1_lo tf32 split, or
 * plain float32 in d_h1_hi when d_h1_lo == NULL), d_h2, dz2 (same convention), d_dz1, d_dout
 * [n_rows, ld_dout] (policy: columns [0, A) loc gradients, [A, 2A) log_scale terms).  d_out
 * (optional): head outputs.  d_stats: TB_STAT_* sums (zeroed by the caller).  Same shape limits
 * as tb_tc_mlp_forward / tb_tc_mlp_backward.                                                   */
int tb_tc_mlp_train(const TbMlpShape* shape, const float* d_params, const float* d_packed,
                    const TbMlpInput* in, int64_t n_rows, int32_t loss_kind, const int64_t* d_idx,
                    const float* d_targets, const float* d_log_scale, const float* d_actions,
                    const float* d_advantages, const float* d_old_log_probs, float ratio_clip,
                    float entropy_coeff, double* d_stats, float* d_out, float* d_xin, float* d_h1_hi,
                    float* d_h1_lo, float* d_h2, float* d_dout, int32_t ld_dout, float* d_dz2_hi,
                    float* d_dz2_lo, float* d_dz1, int32_t passes, const int32_t* d_skip, void* stream);

/* Hardware check, not on the product path: on != 0 makes the fused forward / backward / train
 * kernels store the PLAIN float32 value in the "hi" operand tile (lo stays x - trunc(x)); results
 * bit-identical to on == 0 show that tcgen05 kind::tf32 ignores the 13 low mantissa bits.      */
int tb_debug_plain_hi(int32_t on);

/* Test aid, not on the product path: batches of at most 4096 rows run the first layer of wide
 * inputs (d_in > 32), the input gradient dx and wide heads (n_out > 8) of the unfused tensor-core
 * chain on small-CTA kernels (16 rows x 32 columns per CTA instead of 64 x 256: a 100-row
 * off-policy minibatch covers 56 SMs instead of 2).  Every output is the same sequential fmaf
 * chain, so on == 0 (the 64-row tile kernels only) must give bit-identical results.
 * TONIC_B200_SKINNY=0 has the same effect from the environment.                              */
int tb_debug_skinny(int32_t on);

/* Profiling aid for the fused forward kernel: the first call allocates a device buffer
 * of 64 clock64() stamps that CTA 0 of every later tb_tc_mlp_forward launch fills
 * (slots documented in csrc/tc_mlp.cu); a non-NULL `out64` reads them back (host
 * pointer, 64 values, synchronising copy).  Not used on the product path.               */
int tb_tc_timeline(uint64_t* out64);

/* The MLP entry points above with the two hidden-layer GEMMs (and the W2 weight
 * gradient) on the tensor cores: layer 1, the head, the head gradient and the
 * narrow weight gradients stay FFMA kernels; activations that feed a tensor-core
 * GEMM are kept as tf32 splits (d_h1_hi/lo, d_dz2_hi/lo).  passes = 3 (fp32 grade)
 * or 1 (TF32).  Same semantics as tb_mlp_forward / _backward / _wgrad.            */
int tb_mlp_forward_tc(const TbMlpShape* shape, const float* d_params, const float* d_packed,
                      const TbMlpInput* in, int64_t n_rows, float* d_out, float* d_xin,
                      float* d_h1_hi, float* d_h1_lo, float* d_h2, int32_t passes,
                      const int32_t* d_skip, void* stream);
int tb_mlp_backward_tc(const TbMlpShape* shape, const float* d_params, const float* d_packed,
                       const float* d_dout, int32_t ld_dout, const float* d_h1_hi,
                       const float* d_h1_lo, const float* d_h2, int64_t n_rows,
                       float* d_dz2_hi, float* d_dz2_lo, float* d_dz1, float* d_dx,
                       int32_t dx_col0, int32_t dx_cols, int32_t passes, const int32_t* d_skip,
                       void* stream);
int tb_mlp_wgrad_tc(const TbMlpShape* shape, const float* d_xin, const float* d_h1_hi,
                    const float* d_h1_lo, const float* d_h2, const float* d_dz1,
                    const float* d_dz2_hi, const float* d_dz2_lo, const float* d_dout,
                    int32_t ld_dout, int32_t n_extra, int32_t off_extra, int64_t n_rows,
                    float* d_gpart, int32_t n_split, int32_t n_split_w2, int32_t passes,
                    const int32_t* d_skip, void* stream);

/* Hidden-layer weight gradient on the tensor cores (MN-major tf32 operands):
 * gpart[s, off_w2 + n * 256 + k] = sum over the rows m of split s of
 * dz2[m, n] * h1[m, k]; operands as tf32 splits like tb_tc_gemm256.  With off_b2 >= 0
 * the same pass also writes the bias gradient gpart[s, off_b2 + n] = sum_m dz2[m, n]
 * (extra N = 16 MMAs of the dz2 operand against a block of ones).                   */
int tb_tc_wgrad256(const float* d_dz_hi, const float* d_dz_lo, const float* d_h_hi,
                   const float* d_h_lo, int64_t n_rows, int32_t passes, float* d_gpart,
                   int32_t n_split, int32_t n_params, int32_t off_w2, int32_t off_b2,
                   const int32_t* d_skip, void* stream);

/* ------------------------------------------------------------------------ */
/* Optimiser -- torch.optim.Adam as constructed at updaters/actors.py:11-12,   */
/* 58-59,161-162,228-229 and updaters/critics.py:9-10,59-60,143-144,190-191    */
/* ------------------------------------------------------------------------ */
typedef struct {
    double lr, beta1, beta2, eps;   /* python floats of the reference ctor */
    int32_t n_params;
    float* d_params;       /* flat parameters                                 */
    float* d_m;            /* exp_avg                                         */
    float* d_v;            /* exp_avg_sq                                      */
    int32_t* d_step;       /* [2] steps taken (device resident), block counter */
} TbAdam;

/* grad[i] = grad_scale * sum_s d_gpart[s, i] (n_split_w2 > 0: the W2 block of `shape`
 * only has n_split_w2 partial sums -- the tensor-core weight-gradient kernel uses fewer
 * row splits than the narrow gradients); Adam step; then refreshes the
 * packed transposes (W1^T, W2^T) of `shape` in d_packed.
 * Device-side control (all optional):
 *   d_skip      -- if *d_skip != 0 nothing happens;
 *   d_stats     -- double[TB_STAT_COUNT] statistics of the current minibatch
 *                  produced by the loss kernel; if stats[TB_STAT_NONZERO_ADV]
 *                  == 0 the step is skipped (updaters/actors.py:22,71);
 *   kl_threshold-- if >= 0 and stats kl mean > kl_threshold, *d_stop is set to
 *                  1 AFTER the step (updaters/actors.py:103,112; ppo.py:45-46).*/
int tb_adam_step(const TbAdam* opt, const TbMlpShape* shape, float* d_packed,
                 const float* d_gpart, int32_t n_split, int32_t n_split_w2, float grad_scale,
                 const int32_t* d_skip, const double* d_stats,
                 float kl_threshold, int32_t* d_stop, void* stream);

/* d_out[i] = sum_s d_gpart[s, i]: the flat gradient that is all-reduced over
 * ranks before tb_adam_step(n_split = 1) in multi-GPU runs.      */
int tb_reduce_partials(const float* d_gpart, int32_t n_split, int32_t n_split_w2,
                       int32_t w2_begin, int32_t w2_end, int32_t n_params, float* d_out,
                       const int32_t* d_skip, void* stream);

/* base[r] = address, in THIS process, of rank r's symmetric region of
 * tb_peer_region_bytes(n_params) bytes (zero-initialised; obtained from
 * torch.distributed._symmetric_memory or cudaIpc / fabric handles).            */
typedef struct {
    int32_t world, rank;
    void* base[8];
} TbPeers;
/* ALL weight gradients of one network-minibatch in one launch, reduced to the flat gradient
 * (reference: loss.backward() filling .grad of every variable, torch/updaters/actors.py:33,95,
 * critics.py:23,81).  dW2 / db2 run on the tensor cores (tcgen05, 3xTF32), the narrow
 * gradients dW1 / db1 / dW3 / db3 / extras on the FFMA pipe of the same CTAs at the same time;
 * after a grid-wide barrier (2 * n_split <= 148 CTAs, all resident) every CTA sums the
 * n_split partial slots of its slice of the parameter vector in a fixed order into
 * d_flat[n_params] (sums over rows: the caller scales by 1 / rows in tb_adam_step with
 * n_split = 1).  d_gpart: [n_split, n_params] scratch; d_sync: one zero-initialised uint64
 * per network (grid-barrier counter).  Needs hidden == 256, d_in <= 31, n_out <= 8,
 * n_out + n_extra <= 12.
 * opt != NULL (single process, no gradient clipping): the reduction phase also performs the
 * optimizer step of tb_adam_step on its slice -- g = grad_scale * flat[i], Adam, refresh of
 * d_packed, and the same device-side controls (d_stats / kl_threshold / d_stop) -- so the
 * chain forward -> backward -> weight gradients + Adam is three launches.
 * peers != NULL with world > 1 (one replica per GPU; needs opt): the gradient
 * all-reduce runs inside the same launch.  After the grid barrier every thread stores its
 * reduced element, paired with the epoch tag in one 8-byte word, into its rank's lane of EVERY
 * rank's symmetric region (tb_peer_region_bytes_fused bytes each, zero-initialised) with NVLink
 * stores -- no fence, no flag round trip; it then polls the same element in every lane of its
 * LOCAL region until the tags match, sums the lanes in rank order and applies Adam with
 * grad_scale = 1 / (global rows).  d_reduce_stats (double[TB_STAT_COUNT] or NULL) is
 * summed over the ranks in place; the PPO controls (d_stats != NULL) act on that global block.
 * d_epoch: one zero-initialised uint64 per region, advanced by the launch.  Every rank must
 * launch the same sequence of exchanges (a rank that never arrives traps the waiting kernels
 * after 20 s instead of hanging the node).                                                    */
int tb_mlp_wgrad_fused(const TbMlpShape* shape, const float* d_xin, const float* d_h1_hi,
                       const float* d_h1_lo, const float* d_h2, const float* d_dz1,
                       const float* d_dz2_hi, const float* d_dz2_lo, const float* d_dout,
                       int32_t ld_dout, int32_t n_extra, int32_t off_extra, int64_t n_rows,
                       float* d_gpart, int32_t n_split, float* d_flat, uint64_t* d_sync,
                       int32_t passes, const TbAdam* opt, float* d_packed, float grad_scale,
                       const double* d_stats, float kl_threshold, int32_t* d_stop,
                       const int32_t* d_skip, const TbPeers* peers, uint64_t* d_epoch,
                       double* d_reduce_stats, void* stream);

/* Profiling aid for tb_mlp_wgrad_fused: 64 clock64() stamps of CTA (0, 0) (out16: 64 values; slots: 0 setup done,
 * 1 MMAs issued, 2 accumulator complete, 3 narrow gradients done, 4 partial slot written,
 * 5 at the grid barrier, 6 barrier passed, 7 reduction + Adam done; several ranks: 8 slice pushed,
 * 9 statistics of every rank seen).  Not on the product path. */
int tb_wgrad_timeline(uint64_t* out16);

/* ---- global-norm gradient clipping ---------------------------------------------
 * Reference: torch.nn.utils.clip_grad_norm_(self.variables, self.gradient_clip) between
 * loss.backward() and optimizer.step() (tonic/torch/updaters/actors.py:37-38,96-98,
 * 176-177,256-257; critics.py:24-25,82-83,177-178,230-231).
 * tb_grad_sqnorm: *d_sumsq += sum_i d_grad[i]^2 (float64, fixed order) -- called once
 *                 per network whose variables the updater's optimizer owns.
 * tb_grad_clip:   d_grad[i] *= min(1, max_norm / (sqrt(*d_sumsq) * grad_scale + 1e-6)),
 *                 grad_scale = 1 / (rows of the global minibatch): the flat gradient
 *                 holds SUMS over rows, the reference clips the MEAN gradient.       */
int tb_grad_sqnorm(const float* d_grad, int32_t n, double* d_sumsq, const int32_t* d_skip,
                   void* stream);
int tb_grad_clip(float* d_grad, int32_t n, const double* d_sumsq, float grad_scale,
                 float max_norm, const int32_t* d_skip, void* stream);

/* ---- fused gradient all-reduce + Adam over NVLink peer memory (multi-GPU) ---- */
int64_t tb_peer_region_bytes(int32_t n_params);
/* Region size of the exchange that runs inside tb_mlp_wgrad_fused (push model: every rank
 * stores its reduced parameter slices into its lane of every rank's region; layout in
 * csrc/peers.cuh).                                                               */
int64_t tb_peer_region_bytes_fused(int32_t n_params);
/* This rank's flat gradient (sum of the n_split partial sums; zeros when d_gpart is
 * NULL) and statistics block -> its slot (epoch & 1) of the region, then a
 * system-scope release flag into every peer's region.                          */
int tb_peer_publish(const TbPeers* peers, const float* d_gpart, int32_t n_split,
                    int32_t n_split_w2, int32_t w2_begin, int32_t w2_end,
                    int32_t n_params, const double* d_stats, const uint64_t* d_epoch,
                    int32_t* d_block_counter, const int32_t* d_skip, void* stream);
/* Waits for every rank's flag, sums the slots of all ranks with peer loads in rank
 * order (grad_scale * sum), applies tb_adam_step's update, writes the global
 * statistics to d_stats, advances *d_epoch.  use_stats != 0: the PPO controls of
 * tb_adam_step (skip when all advantages are zero, KL early stop) on the GLOBAL
 * statistics.                                                                   */
int tb_adam_step_peers(const TbAdam* opt, const TbMlpShape* shape, float* d_packed,
                       const TbPeers* peers, float grad_scale, uint64_t* d_epoch,
                       const int32_t* d_skip, double* d_stats, int32_t use_stats,
                       float kl_threshold, int32_t* d_stop, void* stream);

/* (Re)builds the packed transposes from the flat parameters.                 */
int tb_mlp_pack(const TbMlpShape* shape, const float* d_params, float* d_packed,
                void* stream);

/* Target networks (models/actor_critics.py:68-72,126-130):
 * target = (1 - tau) * target + tau * online, elementwise over n floats.      */
int tb_soft_update(float* d_target, const float* d_online, int64_t n, double tau,
                   void* stream);

/* ------------------------------------------------------------------------ */
/* Policy / value heads and losses                                            */
/* ------------------------------------------------------------------------ */
/* layout of the per-minibatch statistics block (doubles, sums over rows)     */
enum {
    TB_STAT_ROWS = 0,        /* rows in the minibatch                          */
    TB_STAT_LOSS = 1,        /* sum of per-row loss terms                      */
    TB_STAT_KL = 2,          /* sum(log_prob_old - log_prob_new)               */
    TB_STAT_ENTROPY = 3,     /* sum over rows and action dims of the entropy   */
    TB_STAT_CLIPPED = 4,     /* number of clipped ratios                       */
    TB_STAT_NONZERO_ADV = 5, /* number of rows with advantage != 0             */
    TB_STAT_STD = 6,         /* sum over rows and dims of the policy std       */
    TB_STAT_VALUE = 7,       /* sum of predicted values (v / q / q1)           */
    TB_STAT_VALUE2 = 8,      /* sum of q2 (twin critics)                       */
    TB_STAT_COUNT = 12
};

/* DetachedScaleGaussianPolicyHead (models/actors.py:37-66) + Normal.sample +
 * log_prob.sum(-1) (agents/a2c.py:75-85).  d_loc_pre [n, A] head
 * pre-activations; loc = tanh(pre); scale = clamp(softplus(log_scale) + 1e-8,
 * 1e-4, 1).  Noise: d_eps [n, A] host-generated standard normals (parity mode:
 * action = loc + eps * scale, separately rounded like torch) or NULL to draw
 * Philox normals from (seed, counter [+ *d_counter, a device-resident stream
 * position advanced with tb_counter_add so the call can live in a CUDA graph]).  */
int tb_gauss_sample(const float* d_loc_pre, const float* d_log_scale,
                    const float* d_eps, uint64_t seed, uint64_t counter,
                    const uint64_t* d_counter, int64_t n_rows, int32_t act_dim,
                    float* d_actions, float* d_log_probs, void* stream);

/* ClippedRatio (updaters/actors.py:70-112) when ratio_clip > 0, else
 * StochasticPolicyGradient (updaters/actors.py:21-50).  Inputs are gathered
 * through d_idx (NULL = identity).  Writes d_dout [n, 2A]: columns [0, A)
 * gradient w.r.t. the loc pre-activation, [A, 2A) per-row gradient w.r.t.
 * log_scale; gradients are for the SUM loss (the 1/rows factor is applied by
 * tb_adam_step's grad_scale).  Accumulates d_stats (must be zeroed).          */
int tb_gauss_policy_loss(const float* d_loc_pre, const float* d_log_scale,
                         const float* d_actions, const float* d_advantages,
                         const float* d_old_log_probs, const int64_t* d_idx,
                         int64_t n_rows, int32_t act_dim, float ratio_clip,
                         float entropy_coeff, float* d_dout, double* d_stats,
                         const int32_t* d_skip, void* stream);

/* MSE value regression (updaters/critics.py:18-28, and the Q losses at
 * :77-86,169-182,222-235): d_dout[i * ld_dout] = 2 (v - target) for the SUM
 * loss; targets gathered through d_idx when given.  Accumulates the squared
 * error into TB_STAT_LOSS, the values into stat_slot (TB_STAT_VALUE/VALUE2) and,
 * if count_rows != 0, the row count into TB_STAT_ROWS.                        */
int tb_mse_loss(const float* d_values, const float* d_targets,
                const int64_t* d_idx, int64_t n_rows, float* d_dout,
                int32_t ld_dout, double* d_stats, int32_t stat_slot,
                int32_t count_rows, const int32_t* d_skip, void* stream);

/* ------------------------------------------------------------------------ */
/* Off-policy heads, targets and actor losses (DDPG / TD3 / SAC)              */
/* ------------------------------------------------------------------------ */
/* DeterministicPolicyHead (models/actors.py:101-115) with optional noise.
 * mode 0: out = tanh(pre)                                   (greedy, ddpg.py:78-81)
 * mode 1: out = clip(tanh(pre) + clip(noise_scale * eps, +-noise_clip), -1, 1)
 *         eps = d_noise64 (host numpy stream, explorations/noisy.py:38-47: float64
 *         noise added to float32 actions then cast), else d_noise32 (host torch
 *         stream, TargetActionNoise updaters/critics.py:125-134), else Philox;
 * mode 2: out = uniform(-1, 1) from Philox (device warm-up actions, noisy.py:44-46) */
int tb_tanh_action(const float* d_pre, int64_t n_rows, int32_t act_dim, int32_t mode,
                   const float* d_noise32, const double* d_noise64, uint64_t seed,
                   uint64_t counter, float noise_scale, float noise_clip, float* d_out,
                   void* stream);

/* GaussianPolicyHead in the SAC configuration + SquashedMultivariateNormalDiag
 * (models/actors.py:7-34,69-98).  d_pre [n, 2A] = [loc | scale pre-activation];
 * raw = loc + eps * clamp(softplus(.), 1e-4, 1); action = tanh(raw);
 * log_prob = sum_j N(raw) - log(1 - action^2 + 1e-6).  greedy != 0: action =
 * tanh(loc) (sac.py:48-51).  d_eps NULL = Philox; d_eps_out receives the noise
 * used (needed by tb_sac_head_grad).                                          */
int tb_squashed_sample(const float* d_pre, const float* d_eps, uint64_t seed,
                       uint64_t counter, int64_t n_rows, int32_t act_dim, int32_t greedy,
                       float* d_actions, float* d_log_probs, float* d_eps_out, void* stream);

/* targets = r + (1 - termination) * gamma * (min(q1, q2) - alpha * log_prob);
 * r / termination gathered through d_idx; d_q2 / d_log_probs may be NULL.
 * DDPG critics.py:71-75, TD3 :159-167, SAC :205-220; discounts buffers.py:34-36. */
int tb_q_target(const float* d_rewards, const float* d_terminations, const int64_t* d_idx,
                double discount_factor, const float* d_q1, const float* d_q2,
                const float* d_log_probs, double entropy_coeff, int64_t n_rows,
                float* d_targets, void* stream);

/* Same with the replay's stored discounts column (n-step returns: the product of the
 * per-step discounts, replays/buffers.py:58-79) instead of (1 - termination) * gamma. */
int tb_q_target_discounts(const float* d_rewards, const float* d_discounts, const int64_t* d_idx,
                          const float* d_q1, const float* d_q2, const float* d_log_probs,
                          double entropy_coeff, int64_t n_rows, float* d_targets, void* stream);

/* Buffer.accumulate_n_steps (replays/buffers.py:58-79), called after row `index` of the ring
 * [max_size, N, ...] was written and before `size` is incremented: back-fills rewards /
 * discounts / next_observations of the previous min(size, return_steps - 1) rows while no
 * reset separates them from the new transition.                                         */
int tb_replay_accumulate_n_steps(float* d_rewards, float* d_discounts, float* d_next_obs,
                                 const float* d_resets, int32_t index, int32_t size,
                                 int32_t max_size, int32_t n_workers, int32_t obs_dim,
                                 int32_t return_steps, void* stream);

/* Actor losses through the critics: loss_i = alpha * log_prob_i - min(q1_i, q2_i)
 * (DPG actors.py:177-179 with q2 = log_prob = NULL; soft DPG :254-257).  Writes
 * d loss_i / d q_k into d_dout1 / d_dout2 and accumulates TB_STAT_LOSS / ROWS.  */
int tb_q_actor_loss(const float* d_q1, const float* d_q2, const float* d_log_probs,
                    double entropy_coeff, int64_t n_rows, float* d_dout1, float* d_dout2,
                    double* d_stats, void* stream);

/* Chain rule through the deterministic head: dout = dq/da * (1 - a^2).        */
int tb_dpg_head_grad(const float* d_dqda, const float* d_actions, int64_t n_rows,
                     int32_t act_dim, float* d_dout, void* stream);

/* Chain rule through the squashed Gaussian head for the SAC actor loss; d_dqda2
 * may be NULL.  Writes d_dout [n, 2A] (loc | scale pre-activation gradients).   */
int tb_sac_head_grad(const float* d_pre, const float* d_eps, const float* d_actions,
                     const float* d_dqda1, const float* d_dqda2, double entropy_coeff,
                     int64_t n_rows, int32_t act_dim, float* d_dout, void* stream);

/* Running statistics of a float32 array (trainer.py:46 logs the actions with
 * stats=True): d_acc[0] += n, [1] += sum, [2] += sum of squares, [3] / [4] =
 * min / max encoded as order-preserving int64 (initialise with
 * tb_array_stats_init values: +inf / -inf encodings).                         */
int tb_array_stats(const float* d_x, int64_t n, double* d_acc, void* stream);

/* Device-side pseudo-random permutation of [0, n) (Feistel bijection with cycle
 * walking), the fast-mode replacement of the host `RandomState.shuffle` of
 * replays/segments.py:62 (not numpy's stream; uniform minibatch coverage).     */
int tb_permutation(uint64_t seed, uint64_t stream_id, const uint64_t* d_counter, int64_t n,
                   int64_t* d_out, void* stream);
/* *d_counter += delta (device-resident RNG stream positions).                  */
int tb_counter_add(uint64_t* d_counter, uint64_t delta, void* stream);

/* ---- device-resident off-policy collection / sampling (fast mode, CUDA-graph safe) ------
 * tb_set_noise_base: base pointer (device uint64, or NULL) that tb_tanh_action and
 *   tb_squashed_sample add to their by-value `counter`: a captured update draws fresh
 *   Philox numbers at every replay (advance the base with tb_counter_add).
 * tb_randint: d_out[i] uniform in [0, *d_total) -- the sample indices of
 *   replays/buffers.py:84-88 (np_random.randint(size * N, size=batch)) for all
 *   batch_iterations at once; Philox stream (seed, stream_id, *d_counter + i).
 * Ring state (device int64[3]): {index, size, size * n_workers}.
 * tb_ring_store: Buffer.store (replays/buffers.py:47-56): for each of n_keys (<= 8) host-listed
 *   keys, row *d_ring_state[0] of h_dst[k] ([max_size, row_elems] float32) <- h_src[k]
 *   (the staged rows of this vector step: h_row_elems[k] floats).
 * tb_ring_advance: index = (index + 1) % max_size, size = min(size + 1, max_size).          */
int tb_set_noise_base(const uint64_t* d_base);
int tb_randint(uint64_t seed, uint64_t stream_id, const uint64_t* d_counter, const int64_t* d_total,
               int64_t n, int64_t* d_out, void* stream);
int tb_ring_store(const float* const* h_src, float* const* h_dst, const int64_t* h_row_elems,
                  int32_t n_keys, const int64_t* d_ring_state, void* stream);
int tb_ring_advance(int64_t* d_ring_state, int64_t max_size, int64_t n_workers, void* stream);

/* ------------------------------------------------------------------------ */
/* Host-side numpy-compatible MT19937 streams (legacy numpy.random.RandomState)*/
/* used for bit-exact minibatch / replay indices and exploration noise:        */
/* replays/segments.py:20,62  replays/buffers.py:22,86  explorations/noisy.py  */
/* ------------------------------------------------------------------------ */
typedef struct TbRandomState TbRandomState;
TbRandomState* tb_rs_create(uint32_t seed);            /* RandomState(seed)    */
void tb_rs_destroy(TbRandomState* rs);
/* RandomState.shuffle(x) for a 1-D int64 array                                */
void tb_rs_shuffle_i64(TbRandomState* rs, int64_t* h_x, int64_t n);
/* RandomState.randint(high, size=n) (dtype int64)                             */
void tb_rs_randint(TbRandomState* rs, int64_t high, int64_t* h_out, int64_t n);
/* RandomState.uniform(low, high, n) / RandomState.normal(size=n) (float64)    */
void tb_rs_uniform(TbRandomState* rs, double low, double high, double* h_out, int64_t n);
void tb_rs_normal(TbRandomState* rs, double* h_out, int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* TONIC_B200_H */
