/*
 * dreamer_b200.h -- C-ABI of libdreamer_b200.so (sm_100a only).
 *
 * The reference (youngers2006/Dreamer) is pure PyTorch and has no FFI / plugin layer: its
 * "ABI" is the Python class surface.  This header is therefore the boundary a
 * maintainer binds to when swapping the hot path: every entry point names the reference
 * function (file:line under /root/reference) whose arithmetic it replaces.  The Python mirror
 * of the reference classes (the modules under dreamer_b200/) is a thin ctypes layer over exactly these calls.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller, row-major contiguous unless a
 *     leading dimension is given; the library never frees or retains caller memory beyond the
 *     call, except inside explicit handles (drm_rssm, drm_rollout, ...) the caller destroys;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued there, no host sync;
 *   - return value: 0 = DRM_OK, negative = error; drm_last_error() gives a thread-local message;
 *   - there is NO CPU fallback: on a device that is not compute capability 10.x every entry
 *     point returns DRM_ERR_ARCH.
 *
 * Sampling contract (bit-exact indices given the same fp32 logits and uniforms):
 *     p   = 0.99 * softmax(logits) + 0.01 / C
 *     cdf = inclusive left-to-right fp32 prefix sum of p
 *     idx = min(C - 1, #{k : cdf[k] <= u})
 *     z   = (onehot(idx) + p) - p              (DynamicsPredictors.py:38-39)
 */
#ifndef DREAMER_B200_H_
#define DREAMER_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRM_OK 0
#define DRM_ERR_SHAPE (-1)
#define DRM_ERR_ALIGN (-2)
#define DRM_ERR_ARCH (-3)
#define DRM_ERR_CUDA (-4)
#define DRM_ERR_ARG (-5)

#define DRM_ABI_VERSION 1

/* ------------------------------------------------------------------------------------------ */
/* library                                                                                      */
/* ------------------------------------------------------------------------------------------ */
int drm_abi_version(void);
const char* drm_last_error(void);
/* 0 if the current device is sm_100-class, else DRM_ERR_ARCH. */
int drm_device_check(void);
/* number of kernels this library has launched in this process (for bench.py's gpu_launches). */
int64_t drm_launch_count(void);
/* Per-stage device timing for bench.py's roofline line: while enabled, every fused tcgen05 stage  */
/* launch is bracketed by CUDA events on its stream.  drm_profile_read synchronises those events    */
/* and returns the summed milliseconds and launch count of `stage` (DRM_STAGE_*), then clears it.   */
#define DRM_STAGE_GRU 0
#define DRM_STAGE_PRIOR_L1 1
#define DRM_STAGE_PRIOR_L2 2
#define DRM_STAGE_PRIOR_CAT 3
#define DRM_STAGE_HEADS_L1 4
#define DRM_STAGE_HEADS_L2 5
#define DRM_STAGE_HEADS_OUT 6
#define DRM_STAGE_OTHER 7
#define DRM_STAGE_ROLLOUT 8 /* the persistent rollout kernel (one launch per drm_rollout_run) */
#define DRM_STAGE_COUNT 9
/* Runtime switches between kernel paths that compute the same thing (tests/test_gpu_errors.py compares them):           */
/*   "persist"    (default 1)  drm_rollout_run as ONE persistent kernel for the whole horizon (rollout_persist.cuh) whenever its */
/*                             static schedule fits the machine (e.g. <= 1024 start states at the reference sizes); 0 = always */
/*                             the launch-per-stage chain (7 launches per imagined step)                                      */
/*   "ln_cluster" (default 1)  launch-per-stage LN-SiLU stages of small grids split over clusters of 4 CTAs (DSMEM statistics exchange) */
/*   "gru_ksplit" (default 1)  single-m-tile grids: the x part and the h part of every GRU tile on two CTAs of a cluster (half the MMA issues */
/*                             per CTA), rows swapped through distributed shared memory for the epilogue                             */
/*   "conv_persist" (default 1) conv layers with <= 64 output channels on the persistent double-buffered GEMM (0: one tile per CTA)   */
/*   "gru_pair"   (default -1) GRU stage on CTA pairs (tcgen05 cta_group::2, M = 256 MMAs, each SM stages half the weight tile):  */
/*                             -1 = automatic (large grids, where it is 11-18 % faster), 0 = never, 1 = whenever >= 2 m-tiles      */
/*   "gru_band"   (default 0)  CTA-pair GRU kernel: m-tiles per band of the tile order (0 = 16)                           */
/*   "small_a"    (default 1)  stages with <= 32 rows load 32-row A boxes by TMA instead of whole 128-row tiles          */
/*   "gru_u"      (default 0)  GRU tile width of the launch-per-stage path: 0 = automatic, 32 or 64 hidden units per tile */
int drm_set_option(const char* name, int32_t value);
/* Debug probe: with on = 1 CTA (0,0) of every fused stage records {globaltimer ns, clock64} at 8 points */
/* (entry, setup done, first TMA issued, first operands landed, last MMA issued, accumulator ready,     */
/* epilogue done, TMEM freed) and every CTA records {entry ns, dependency wait over ns, exit ns, SM id};  */
/* on = 0 copies DRM_STAGE_COUNT * (16 + 1024) u64 to out_host (probes, then 256 CTA records per stage)  */
/* and disables it.                                                                                     */
int drm_debug_timeline(int32_t on, unsigned long long* out_host);
int drm_profile_enable(int32_t on);
int drm_profile_read(int32_t stage, double* total_ms, int64_t* launches);

/* ------------------------------------------------------------------------------------------ */
/* (2) fused 32-class categorical head                                                          */
/*   replaces DynamicsPredictor.predict  DynamicsPredictors.py:33-39                            */
/*            Encoder.encode            VariationalAutoEncoder.py:88-98                         */
/* ------------------------------------------------------------------------------------------ */
/* logits [n_rows, 32] fp32, uniforms [n_rows] fp32.  Outputs (each may be NULL):               */
/*   idx [n_rows] u8, z_st [n_rows, 32] fp32, probs [n_rows, 32] fp32 (unimixed),               */
/*   z_bf16 [n_rows, 32] bf16 one-hot (raw uint16 storage).                                     */
int drm_categorical32_fwd(const float* logits, const float* uniforms, uint8_t* idx, float* z_st, float* probs,
                          uint16_t* z_bf16, int64_t n_rows, void* stream);
/* One-hot expansion of sampled classes: idx [n_rows] u8 (< 32) -> out [n_rows, 32] fp32, 16-byte aligned.  Lets a caller hand a   */
/* latent state across the host link as 32 bytes instead of 4 KB (the start states of Dreamer.dream_episodes, Dreamer.py:143).    */
int drm_onehot32(const uint8_t* idx, float* out, int64_t n_rows, void* stream);
/* Teacher-forced forward (the class idx [n_rows] u8 is given, not drawn): z_st = (onehot(idx) + p) - p and/or  */
/* probs p = 0.99 softmax + 0.01/32.  Used by the gradient tail, which replays the classes the scan sampled.    */
int drm_categorical32_st(const float* logits, const uint8_t* idx, float* z_st, float* probs, int64_t n_rows, void* stream);
/* Backward of the straight-through sample: dlogits = 0.99 * s * (g - sum_j s_j g_j) [+ dl_add], s = softmax(logits),         */
/* g = dz [+ dz2]; the two optional addends let one BPTT step fuse its accumulations.  All [n_rows, 32] fp32.               */
int drm_categorical32_bwd(const float* logits, const float* dz, const float* dz2, const float* dl_add, float* dlogits,
                          int64_t n_rows, void* stream);

/* ---- elementwise pieces of the hand-scheduled BPTT (the GEMMs between them: drm_gemm_tf32 below) ---- */
/* Backward of y = SiLU(LayerNorm(a) * gamma + beta) (the Linear-LN-SiLU blocks of every MLP, e.g.                           */
/* VariationalAutoEncoder.py:48-52): dy, a [rows, n] (n <= 1024) -> da [rows, n]; dln (optional) = dy * silu'(ln) for the     */
/* batched dgamma = sum(dln * xhat), dbeta = sum(dln).  Statistics are recomputed from a (eps as in the forward, 1e-5).       */
int drm_ln_silu_bwd(const float* dy, const float* a, const float* gamma, const float* beta, float* da, float* dln,
                    int64_t rows, int32_t n, float eps, void* stream);
/* ... with a third optional output dlnx = dln * xhat (xhat = the normalised activation): its column sums are d(loss)/d(gamma),   */
/* dln's are d(loss)/d(beta) (drm_colsum).                                                                                   */
int drm_ln_silu_bwd_affine(const float* dy, const float* a, const float* gamma, const float* beta, float* da, float* dln, float* dlnx,
                           int64_t rows, int32_t n, float eps, void* stream);
/* out[c] (=, or += when accumulate) sum over the rows of x [rows, n] (row pitch ld): the bias / LayerNorm-affine gradients of the  */
/* batched backward (autograd's sum-to-size nodes).  Deterministic (fixed reduction order).                                     */
/* scratch: device memory of drm_colsum_scratch_bytes(rows, n) bytes (0 up to 64 rows: may be NULL; taller matrices are summed in    */
/* two stages).                                                                                                              */
int64_t drm_colsum_scratch_bytes(int64_t rows, int32_t n);
int drm_colsum(const float* x, int64_t rows, int32_t n, int64_t ld, float* out, int32_t accumulate, void* scratch, void* stream);
/* The same for a bf16 matrix (n, ld even; x 4-byte aligned): the conv layers' bias gradients from grad_output viewed as              */
/* [N * H * W, C] channels-last rows (autograd's conv-bias reduction).  scratch: drm_colsum_bf16_scratch_bytes(rows, n) bytes.          */
/* The decoder's image layer inside the TRAINING graph: out = tanh(conv_transpose2d(x, weight, bias, stride 2, padding 1)) for a 4 x 4     */
/* kernel and 1 - 3 output channels (VariationalAutoEncoder.py:134-137; what the world-model step's autograd graph evaluates).          */
/* x: bf16 NHWC [N, Hin, Win, C_in] (C_in = 8 / 16 / 32 / 64), weight fp32 [C_in][C_out][4][4] (the module's parameter, rounded to bf16   */
/* in the kernel), bias fp32 [C_out] -> out fp32 NCHW [N, C_out, 2 Hin, 2 Win].                                                  */
int drm_convt_image_fwd(const void* x_nhwc_bf16, const float* weight, const float* bias, float* out_nchw, int32_t N, int32_t Hin,
                        int32_t Win, int32_t C_in, int32_t C_out, void* stream);
int64_t drm_colsum_bf16_scratch_bytes(int64_t rows, int32_t n);
int drm_colsum_bf16(const void* x, int64_t rows, int32_t n, int64_t ld, float* out, int32_t accumulate, void* scratch, void* stream);
/* Backward of one nn.GRUCell step (SequenceModel.py:13,19-24) from its pre-activations gi = x W_ih^T + b_ih,                */
/* gh = h W_hh^T + b_hh [rows, 3D] (gate order r, u, n), h_prev [rows, D] (NULL = zeros) and dh [rows, D]:                    */
/*   dgi = [dr, du, dn], dgh = [dr, du, dn * r] (pre-activation gradients), dh_prev (=, or += when accumulate) dh * u.        */
/* The caller completes dh_prev += dgh W_hh and dx = dgi W_ih with library GEMMs.                                             */
/* Backward of the actor head a = tanh(mu + sigma eps), sigma = softplus(clamp(ls, -5, 2)) + 1e-3 (Agent.py:199-209) for one BPTT */
/* step: d_head [rows, 2A] = [g_mu + du | (g_sigma + du eps) sigmoid(ls) 1(-5 < ls < 2)], du = da (1 - a^2); da may be NULL (= 0).   */
int drm_actor_head_bwd(const float* g_mu, const float* g_sigma, const float* da, const float* a, const float* eps,
                       const float* log_sigma, float* d_head, int64_t rows, int32_t A, void* stream);
/* log-probability of the taken action under tanh(Normal(mu, sigma)) summed over the action dimension (Agent.py:110-115) and its   */
/* gradient with respect to mu, sigma (the action is a constant there):  a, mu, sigma [rows, A] -> logp [rows] and / or                */
/* g_mu, g_sigma [rows, A] = coef[row] * d logp / d(mu, sigma)  (coef may be NULL = 1; each output may be NULL).                      */
int drm_tanh_normal_logp(const float* a, const float* mu, const float* sigma, const float* coef, float* logp, float* g_mu,
                         float* g_sigma, int64_t rows, int32_t A, void* stream);
/* The two percentiles of Agent.update_S (Agent.py:78-88; torch.quantile's linear rule) of x [n] without a sort (radix select, one     */
/* launch): out[0] = percentile p_lo, out[1] = percentile p_hi, out[2] = 1.0 if every value is finite else 0.0 (device floats).       */
int drm_percentile_pair(const float* x, int64_t n, double p_lo, double p_hi, float* out, void* stream);
int drm_gru_bwd(const float* dh, const float* gi, const float* gh, const float* h_prev, float* dgi, float* dgh,
                float* dh_prev, int32_t accumulate, int64_t rows, int32_t D, void* stream);
/* ... with dh = dh + dh_add (dh_add may be NULL): lets the recurrent term dgh W_hh arrive from a GEMM on a side stream */
int drm_gru_bwd_add(const float* dh, const float* dh_add, const float* gi, const float* gh, const float* h_prev, float* dgi,
                    float* dgh, float* dh_prev, int32_t accumulate, int64_t rows, int32_t D, void* stream);
/* KL balance terms of WorldModel.training_step  WorldModel.py:175-181:                         */
/*   kl[g] = sum over the `rows_per_group` categorical rows of group g of KL(Cat(post)||Cat(prior)) */
/* post/prior logits [n_groups * rows_per_group, 32] fp32 -> kl [n_groups] fp32.                 */
int drm_categorical32_kl(const float* post_logits, const float* prior_logits, float* kl, int64_t n_groups,
                         int rows_per_group, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* (4) replay ring gather -- replaces Buffer.sample_sequences' gather  Buffer.py:49-61           */
/* ------------------------------------------------------------------------------------------ */
/* ring_obs [cap, frame_bytes] u8 (frame_bytes % 16 == 0), ring_act [cap, A], ring_rew/ring_con  */
/* [cap, 1] fp32, starts [B] int64 (window = (start + t) % cap).  Outputs: obs_out [B, L,        */
/* frame_bytes] fp32 holding 0..255 (normalise = 0, as the reference returns) or x/255 - 0.5     */
/* (normalise = 1, WorldModel.py:156 fused), act_out [B, L, A], rew_out/con_out [B, L, 1].       */
int drm_replay_gather(const uint8_t* ring_obs, const float* ring_act, const float* ring_rew, const float* ring_con,
                      const int64_t* starts, float* obs_out, float* act_out, float* rew_out, float* con_out,
                      int32_t B, int32_t L, int64_t cap, int32_t frame_bytes, int32_t A, int32_t normalise,
                      void* stream);
/* Buffer.add_to_buffer  Buffer.py:19-30 for `n` consecutive transitions already on the device   */
/* (obs u8 [n, frame_bytes], act [n, A], rew raw [n], con [n]); rewards are stored symlog'd.     */
int drm_replay_insert(uint8_t* ring_obs, float* ring_act, float* ring_rew, float* ring_con, const uint8_t* obs,
                      const float* act, const float* rew, const float* con, int64_t next_idx, int32_t n,
                      int64_t cap, int32_t frame_bytes, int32_t A, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* returns / losses (fp32, HBM-bound)                                                            */
/* ------------------------------------------------------------------------------------------ */
/* Agent.compute_batched_R_lambda_returns' reverse scan  Agent.py:158-171.                       */
/* rew, cont [B, H]; value [B, H + 1] -> out [B, H].                                             */
int drm_lambda_return(const float* rew, const float* cont, const float* value, float* out, int32_t B, int32_t H,
                      float gamma, float lambda_, void* stream);
/* sum(twohot(v) * log_softmax(logits))  DreamerUtils.py:39-50 + WorldModel.py:137-138 /         */
/* Agent.py:129-134, without materialising the two-hot.  logits [N, NB], value [N] (symlog is    */
/* applied first when apply_symlog != 0), buckets [NB] -> ll [N] (log-likelihood, not negated).  */
int drm_twohot_ce(const float* logits, const float* value, const float* buckets, float* ll, int64_t N, int32_t NB,
                  int32_t apply_symlog, void* stream);
/* Backward of drm_twohot_ce (what autograd does behind WorldModel.py:193 / Agent.py:141-146 for the two-hot losses):              */
/*   dlogits[row][c] = scale * (scale_dev ? *scale_dev : 1) * (coef ? coef[row] : 1) * (twohot(value[row])[c] - softmax(logits[row])[c]) */
/* coef [N] (a mask / per-row weight) and scale_dev (a DEVICE scalar, e.g. 1 / global element count) may be NULL.                   */
int drm_twohot_ce_bwd(const float* logits, const float* value, const float* buckets, const float* coef, const float* scale_dev,
                      float scale, float* dlogits, int64_t N, int32_t NB, int32_t apply_symlog, void* stream);
/* symexp(sum(softmax(logits) * buckets))  DynamicsPredictors.py:70-74, Agent.py:237-241.        */
int drm_bucket_value(const float* logits, const float* buckets, float* value, int64_t N, int32_t NB, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* (1)(2)(5) RSSM: packed weights + fused tcgen05 stages                                         */
/* ------------------------------------------------------------------------------------------ */
typedef struct drm_dims {
  int32_t D;          /* hidden_state_dims (GRU width)                                          */
  int32_t R, C;       /* latent rows x classes; C must be 32, R * C a multiple of 256            */
  int32_t A;          /* action_dims (<= 16)                                                    */
  int32_t NB;         /* reward / critic buckets (<= 256)                                       */
  int32_t h_prior[2]; /* dyn_pred_hidden_num_nodes_{1,2} (<= 256)                               */
  int32_t h_head[2];  /* hidden sizes shared by reward / continue / actor / critic MLPs (<= 256) */
} drm_dims;

/* One Linear-LN-SiLU-Linear-LN-SiLU(-Linear) head in reference state_dict layout, fp32:         */
/* w0 [h1, in], b0, g0, be0 [h1]; w1 [h2, h1], b1, g1, be1 [h2]; w2 [out, h2], b2 [out] or NULL. */
typedef struct drm_mlp_w {
  const float *w0, *b0, *g0, *be0, *w1, *b1, *g1, *be1, *w2, *b2;
} drm_mlp_w;

typedef struct drm_rssm_weights {
  /* SequenceModel.GRU (nn.GRUCell)  SequenceModel.py:13-17; gate order [r; z; n]              */
  const float *gru_w_ih, *gru_w_hh, *gru_b_ih, *gru_b_hh; /* [3D, R*C + A], [3D, D], [3D], [3D] */
  drm_mlp_w prior;   /* DynamicsPredictor.logit_net        DynamicsPredictors.py:15-23  in = D  */
  drm_mlp_w reward;  /* RewardPredictor.logit_net          DynamicsPredictors.py:52-60  in=[h,z] */
  drm_mlp_w cont;    /* ContinuePredictor.logit_generator  DynamicsPredictors.py:85-93  in=[h,z] */
  drm_mlp_w actor;   /* Actor.base_net (w2 = b2 = NULL)    Agent.py:178-185             in=[h,z] */
  const float *actor_mu_w, *actor_mu_b, *actor_ls_w, *actor_ls_b; /* Agent.py:186-187  [A, h2]  */
  drm_mlp_w critic;        /* Critic.value_net        Agent.py:219-227 (may be all NULL)        */
  drm_mlp_w target_critic; /* Agent.target_critic     Agent.py:50      (may be all NULL)        */
  const float* buckets_rew;  /* [NB]  DynamicsPredictors.py:61-62                               */
  const float* buckets_crit; /* [NB]  Agent.py:228-229 (NULL if no critic)                      */
} drm_rssm_weights;

typedef struct drm_rssm drm_rssm;       /* packed bf16 weights (a cache: re-pack after every optimiser step) */
typedef struct drm_rollout drm_rollout; /* workspace + TMA descriptors for a fixed (B, H) */

int drm_rssm_create(const drm_dims* dims, drm_rssm** out);
/* The same with an explicit operand precision.  DRM_PRECISION_BF16 (drm_rssm_create): bf16 operands, fp32 accumulation -- the
 * north star's "bf16 <= 1e-2" mode and the only one the persistent kernels, the observe / VAE path and the CTA-pair / cluster
 * stage variants implement.  DRM_PRECISION_TF32: fp32 state, activations and packed weights rounded to TF32, tcgen05.mma
 * kind::tf32 -- the precision class the reference itself runs its imagination path in on a GPU
 * (torch.backends.cuda.matmul.allow_tf32, train_car_racer.py:13); launch-per-stage kernels, rollout / step-level entry points only
 * (drm_vae_create rejects such a handle). */
enum { DRM_PRECISION_BF16 = 0, DRM_PRECISION_TF32 = 1 };
int drm_rssm_create_ex(const drm_dims* dims, int32_t precision, drm_rssm** out);
int drm_rssm_pack(drm_rssm* m, const drm_rssm_weights* w, void* stream);
int drm_rssm_destroy(drm_rssm* m);

int drm_rollout_create(drm_rssm* m, int32_t B, int32_t H, drm_rollout** out);
int drm_rollout_destroy(drm_rollout* r);

/* Dreamer.dream_episodes  Dreamer.py:143-175  (Actor.act Agent.py:202-210 ->                    */
/* WorldModel.imagine_step WorldModel.py:72-77, H times), with host-supplied randomness.        */
/*   z0 [B, R*C] fp32 (a one-hot / straight-through latent), h0 [B, D] fp32,                     */
/*   uniforms [H, B, R], normals [H, B, A]                                                       */
/*   -> latent [B, H+1, R*C], hidden [B, H+1, D], actions/mu/sigma [B, H, A],                    */
/*      rewards/continues [B, H] (all fp32), idx [B, H, R] u8 (may be NULL).                     */
int drm_rollout_run(drm_rollout* r, const float* z0, const float* h0, const float* uniforms, const float* normals,
                    float* latent, float* hidden, float* actions, float* rewards, float* continues, float* mu,
                    float* sigma, uint8_t* idx, void* stream);

/* Which path the next drm_rollout_run takes and, after a failed launch, what the persistent kernel was waiting for:        */
/* out[0] = 1 persistent kernel / 0 launch-per-stage chain, out[1] = GRU tile width, out[2] = sampling tile width,           */
/* out[3] = CTAs, out[4] = number of timed-out waits recorded, then records {code, seen, want, thread, cta} (n >= 5 words). */
int drm_rollout_info(drm_rollout* r, uint32_t* out, int32_t n);

/* Debug: per-tile timestamps of the persistent kernel.  out == NULL: record the tiles of states [j0, j0 + nj) of the rollouts   */
/* that follow (32 tiles per CTA are kept).  out != NULL: copy CTAs * 32 records of 8 u64 {code, start, dependency seen, first  */
/* operands landed, epilogue ready, accumulator ready, epilogue done, published} (globaltimer ns; code = kind << 24 | layer << 16 */
/* | state << 8 | m-tile) into out (n_words >= CTAs * 256) and stop recording.                                               */
int drm_rollout_trace(drm_rollout* r, int32_t j0, int32_t nj, unsigned long long* out, int64_t n_words);

/* Step-level entry points behind the drop-in classes.  They run on the rollout workspace (any   */
/* N <= B of drm_rollout_create).                                                                */
/* SequenceModel.forward  SequenceModel.py:19-24:  z [N, R*C], h [N, D], a [N, A] -> h_out [N, D] */
int drm_gru_step(drm_rollout* r, const float* z, const float* h, const float* a, float* h_out, int32_t N,
                 void* stream);
/* DynamicsPredictor.forward / .predict  DynamicsPredictors.py:25-40: h [N, D] -> logits         */
/* [N, R*C]; when uniforms [N, R] != NULL also z_st [N, R*C] and idx [N, R] (each may be NULL).  */
int drm_prior(drm_rollout* r, const float* h, const float* uniforms, float* logits, float* z_st, uint8_t* idx,
              int32_t N, void* stream);
/* Reward / Continue / Actor / Critic heads on [h, z]  DynamicsPredictors.py:64-74, 95-105;      */
/* Agent.py:191-210, 231-241.  `heads` is a bit mask of DRM_HEAD_*; outputs of unselected heads   */
/* and any NULL output are skipped.  reward/value [N] (symexp'd), *_logits [N, NB], cont_prob /   */
/* cont_logit [N], mu/sigma/action [N, A] (action = tanh(mu + sigma * normals), normals [N, A]).  */
#define DRM_HEAD_REWARD 1
#define DRM_HEAD_CONT 2
#define DRM_HEAD_ACTOR 4
#define DRM_HEAD_CRITIC 8
#define DRM_HEAD_TARGET_CRITIC 16
typedef struct drm_heads_out {
  float *reward, *reward_logits, *cont_prob, *cont_logit, *mu, *sigma, *action;
  float *value, *value_logits, *target_value;
} drm_heads_out;
int drm_heads(drm_rollout* r, const float* h, const float* z, const float* normals, int32_t heads,
              const drm_heads_out* out, int32_t N, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* (3) VAE encoder / decoder + the posterior (observe) scan                                      */
/* ------------------------------------------------------------------------------------------ */
typedef struct drm_vae_dims {
  int32_t H, W;         /* observation_dims (multiples of 16)                                     */
  int32_t e1, e2;       /* encoder_filter_num_1/2: conv channels 3 -> e1 -> e2 -> 2 e2 -> 4 e2    */
  int32_t d1, d2;       /* decoder_filter_num_1/2: convT channels 4 d2 -> 2 d2 -> d2 -> d1 -> 3   */
  int32_t h_enc, h_dec; /* encoder / decoder_hidden_layer_nodes (<= 256)                          */
} drm_vae_dims;

typedef struct drm_vae_weights {
  /* Encoder.feature_extractor.{0,2,4,6}  VariationalAutoEncoder.py:33-42: Conv2d k4 s2 p1, [co, ci, 4, 4], [co] */
  const float* enc_conv_w[4];
  const float* enc_conv_b[4];
  /* Encoder.latent_mapper.{0,1,3}  :50-55: Linear(feat + D -> h_enc) input [features, h], LN, Linear(h_enc -> R*C) */
  const float *enc_l1_w, *enc_l1_b, *enc_ln_g, *enc_ln_b, *enc_l2_w, *enc_l2_b;
  /* Decoder.upscaler.{0,1,3}  :119-125: Linear(R*C + D -> h_dec) input [h, z], LN, Linear(h_dec -> 4 d2 * H/16 * W/16) */
  const float *dec_l1_w, *dec_l1_b, *dec_ln_g, *dec_ln_b, *dec_l2_w, *dec_l2_b;
  /* Decoder.image_builder.{0,2,4,6}  :128-137: ConvTranspose2d k4 s2 p1, [ci, co, 4, 4], [co] */
  const float* dec_conv_w[4];
  const float* dec_conv_b[4];
} drm_vae_weights;

typedef struct drm_vae drm_vae;         /* packed conv / dense weights of Encoder + Decoder */
typedef struct drm_observe drm_observe; /* workspace for B sequences x T steps (time-major bf16 state, conv scratch) */

int drm_vae_create(drm_rssm* m, const drm_vae_dims* dims, drm_vae** out);
int drm_vae_pack(drm_vae* v, const drm_vae_weights* w, void* stream);
int drm_vae_destroy(drm_vae* v);
int drm_observe_create(drm_rssm* m, drm_vae* v, int32_t B, int32_t T, drm_observe** out);
int drm_observe_destroy(drm_observe* o);

/* debug: per-tile %globaltimer records of the persistent posterior-scan kernel (protocol of drm_rollout_trace) */
int drm_observe_trace(drm_observe* o, int32_t j0, int32_t nj, unsigned long long* out, int64_t n_words);

/* The posterior scan.  mode 0 = WorldModel.unroll_model's loop  WorldModel.py:92-107 (step 0 runs a GRU  */
/* step on the all-zero state); mode 1 = Dreamer.warm_start_generator  Dreamer.py:252-261 (frame 0 is       */
/* encoded with h = 0 and no GRU step).  The conv stack runs once over all B*T frames (it does not depend  */
/* on h), only GRU + the h part of latent_mapper.0 + LN + latent_mapper.3 + sampling stay in the loop.      */
/*   obs [B, T, 3, H, W] fp32 already normalised to [-0.5, 0.5]; act [B, T, A]; uniforms [T, B, R]           */
/*   -> latent [B, T, R*C], hidden [B, T, D], post_logits [B, T, R*C] (fp32), idx [B, T, R] u8 (NULL ok)     */
int drm_observe_scan(drm_observe* o, const float* obs, const float* act, const float* uniforms, int32_t mode,
                     float* latent, float* hidden, float* post_logits, uint8_t* idx, void* stream);
/* The batched heads of WorldModel.unroll_model  WorldModel.py:116-119 on the states of the last scan:      */
/*   prior_logits [B, T, R*C], dec_mu [B, T, 3, H, W], reward_logits [B, T-1, NB], cont_logit [B, T-1]       */
/* (reward / continue use steps 1..T-1).  Any output may be NULL.                                           */
int drm_observe_heads(drm_observe* o, float* prior_logits, float* dec_mu, float* reward_logits, float* cont_logit,
                      void* stream);
/* Encoder.forward / .encode  VariationalAutoEncoder.py:57-99 on N <= B*T rows: h [N, D], obs [N, 3, H, W]  */
/* -> logits [N, R*C]; with uniforms [N, R] also z_st [N, R*C] and idx [N, R].                              */
int drm_encoder_fwd(drm_observe* o, const float* h, const float* obs, const float* uniforms, float* logits, float* z_st,
                    uint8_t* idx, int32_t N, void* stream);
/* Decoder.forward  VariationalAutoEncoder.py:139-161: h [N, D], z [N, R*C] -> mu [N, 3, H, W] (tanh).      */
int drm_decoder_fwd(drm_observe* o, const float* h, const float* z, float* mu, int32_t N, void* stream);
/* -sum((a - b)^2) over each row of `len` floats: WorldModel.py:129.  a, b [rows, len] -> out [rows].       */
int drm_neg_sse_rows(const float* a, const float* b, float* out, int64_t rows, int32_t len, void* stream);

/* ---- fused optimiser tail on a FLAT parameter group ------------------------------------------------ */
/* Replaces nn.utils.clip_grad_norm_(params, 100) + torch.optim.AdamW.step (WorldModel.py:195-200; Agent.py:141-151) and,   */
/* with ema_target, Agent.soft_update_target (Agent.py:90-94).  All buffers are device fp32 [n], 16-byte aligned:            */
/*   g' = g * min(1, max_norm / (||g||_2 + 1e-6));  p *= 1 - lr*wd;  m += (1-b1)(g'-m);  v = b2 v + (1-b2) g'^2;            */
/*   p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps);   ema = (1-tau) ema + tau p;   g = 0 when zero_grad                */
/* state: 8 device floats owned by the caller, zero-initialised: [0] step count t (advanced here), [1] gradient norm,        */
/* [2] clip coefficient, [3] 1.0 when the step was SKIPPED because the norm was not finite (p, m, v, ema, t untouched),       */
/* [4], [5] bias-correction terms.  scratch: drm_adamw_scratch_bytes() device bytes.  max_norm <= 0 disables clipping.       */
/* Three launches, no host synchronisation: safe inside CUDA-graph capture.                                                  */
int64_t drm_adamw_scratch_bytes(void);
int drm_adamw_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float* state, void* scratch,
                   float lr, float beta1, float beta2, float eps, float weight_decay, float max_norm, float* ema_target,
                   float tau, int32_t zero_grad, void* stream);
/* ||grad||_2 over the flat buffer -> norm_out[0] (device float).  Deterministic two-stage reduction (fp64 partials).        */
int drm_grad_norm(const float* grad, int64_t n, void* scratch, float* norm_out, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Contractions of the backward passes: what autograd's mm / addmm nodes do behind                     */
/* WorldModel.training_step's loss.backward() (WorldModel.py:193) and Agent.train_step's (Agent.py:141-151), on this          */
/* library's TMA / tcgen05 kind::tf32 kernel (csrc/gemm_tf32.cu) instead of a library GEMM:                                   */
/*   C[M, N] (+)= op(A)[M, K] * op(B)[N, K]^T (+ bias[n]),  fp32 row-major in / out, operands rounded to TF32, fp32 accumulate */
/*   A: [M, K] (lda), or [K, M] with DRM_GEMM_TRANS_A;   B: [N, K] (ldb), or [K, N] with DRM_GEMM_TRANS_B;   bias may be NULL  */
/*   forward  Y = X W^T + b: flags 0;   input gradient dX = dY W: TRANS_B;   weight gradient dW += dY^T X: TRANS_A|TRANS_B|ACCUMULATE */
/* Both operands are first rewritten K-major and ROUNDED to nearest TF32 by a pack kernel (a tensor core truncates a plain fp32     */
/* operand; the bias compounds over the steps of a recurrence) -- except: an operand flagged DRM_GEMM_A_DIRECT / _B_DIRECT (the      */
/* caller has rounded it already, e.g. weights through drm_pack_tf32 once per backward) is read in place by TMA in either            */
/* orientation when its base is 16-byte aligned and its leading dimension a multiple of 4; and the few-row operand A of a skinny      */
/* problem (M <= 64, not transposed) is rounded inside the GEMM kernel.                                                         */
/* workspace: device memory, 256-byte aligned, >= drm_gemm_tf32_workspace_bytes(M, N, K).  Its first 1024 bytes (split-K tile      */
/* tickets) must be ZERO before the first call and are left zero by every call; calls that share a workspace must be             */
/* stream-ordered.  1 - 2 launches, no host synchronisation, deterministic (a K split is summed in split order by the last       */
/* CTA of each tile): safe inside CUDA-graph capture.                                                                         */
#define DRM_GEMM_TRANS_A 1
#define DRM_GEMM_TRANS_B 2
#define DRM_GEMM_ACCUMULATE 4
#define DRM_GEMM_A_DIRECT 8
#define DRM_GEMM_B_DIRECT 16
#define DRM_GEMM_NO_PDL 32      /* launch without programmatic stream serialisation (profiling) */
int64_t drm_gemm_tf32_workspace_bytes(int32_t M, int32_t N, int32_t K);
int drm_gemm_tf32(int32_t M, int32_t N, int32_t K, const float* A, int64_t lda, const float* B, int64_t ldb, float* C,
                  int64_t ldc, const float* bias, int32_t flags, void* workspace, int64_t workspace_bytes, void* stream);
/* out[r][k] = tf32_round(trans ? in[k][r] : in[r][k]) for r < rows, k < K, zeros up to ld_out (a multiple of 4): a K-major,     */
/* pre-rounded operand for drm_gemm_tf32 (weights that stay fixed over the steps of a recurrence).                            */
int drm_pack_tf32(int32_t rows, int32_t K, const float* in, int64_t ld, int32_t trans, float* out, int64_t ld_out, void* stream);

/* Test hook: plain bf16 GEMM  out[M, N] = A[M, K] * W[N, K]^T + bias  through the same TMA /      */
/* tcgen05 main loop the fused stages use (fp32 inputs are rounded to bf16 on the device).        */
int drm_test_gemm(const float* A, const float* W, const float* bias, float* out, int32_t M, int32_t N, int32_t K,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DREAMER_B200_H_ */
