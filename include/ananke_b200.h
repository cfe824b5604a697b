/* ananke_b200 -- C ABI of the B200-native GAT-ODE hot path.
 *
 * Plain C: pointers, sizes, an opaque stream handle.  No torch types cross this boundary; the caller
 * (PyTorch on the host side) owns every buffer, the library never allocates or frees device memory,
 * keeps no global state and is stream-ordered and re-entrant.  Every entry point returns 0 on success
 * or a negative `ab200_status`; nothing throws across the ABI.  Asynchronous CUDA faults surface at the
 * caller's next synchronisation (the Python wrapper checks `ab200_last_cuda_error`).
 *
 * Each entry point cites the reference interface (under /root/reference/src/ananke_abm/models/) it
 * stands in for.  The solver arithmetic itself lives in the reference's un-vendored dependency
 * torchdiffeq==0.2.5 (uv.lock:2896-2897); "tdq:" cites that package's module.
 */
#ifndef ANANKE_B200_H_
#define ANANKE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AB200_ABI_VERSION 1

typedef void* ab200_stream_t; /* a cudaStream_t */

typedef enum ab200_status {
  AB200_OK = 0,
  AB200_ERR_BAD_ARG = -1,       /* null pointer, non-positive size, T < 1 ... */
  AB200_ERR_UNSUPPORTED = -2,   /* drift-net shape not instantiated in this build */
  AB200_ERR_WORKSPACE = -3,     /* workspace too small: call *_workspace_bytes */
  AB200_ERR_CUDA = -4,          /* launch failed: see ab200_last_cuda_error */
  AB200_ERR_NOT_MONOTONE = -5,  /* time grid must be strictly increasing (tdq: odeint.py _check_inputs) */
  AB200_ERR_DT_UNDERFLOW = -6,  /* tdq: rk_common.py "underflow in dt" */
  AB200_ERR_MAX_STEPS = -7      /* tdq: rk_common.py "max_num_steps exceeded" */
} ab200_status;

/* Arithmetic mode of the drift-net GEMMs. */
typedef enum ab200_precision {
  AB200_PREC_F32 = 0,   /* strict: fp32 FFMA, fp32 accumulate -- the 1e-5 parity path           */
  AB200_PREC_BF16 = 1,  /* tcgen05 bf16 x bf16 -> fp32 (TMEM accumulators); state stays fp32     */
  AB200_PREC_BF16X3 = 2 /* tcgen05 3-term bf16 split of both operands (6 MMAs), fp32-grade       */
} ab200_precision;

/* The "second-order residual-MLP drift" both reference models use as ODE right-hand side:
 *   y = [p(P), v(P), h(H)],   dy/dt = [v, net([p, v, h, sin(2 pi t/period), cos(2 pi t/period)]) + corr(p), 0]
 *   net = Linear(2P+H+2, hid) -> ReLU -> n_res x ResidualBlock(hid) -> Linear(hid, P)
 *   ResidualBlock(x) = act(x + Linear(act(Linear(x))))
 * mode_sep:   P=64, H=32, hid=128, n_res=2, act=ReLU, no correction
 *             (mode_sep/architecture/model.py:16-38 ODEFunc/ResidualBlock, :56-73 WrappedSDE.forward)
 * latent_ode: P=16, H=32, hid=128, n_res=2, act=Tanh, potential correction on p[idx_a], p[idx_b]
 *             (latent_ode/architecture/model.py:9-17, :46-51, :56-74, :77-117)
 */
typedef struct ab200_drift_desc {
  int32_t pos_dim;        /* P */
  int32_t ctx_dim;        /* H */
  int32_t hidden;         /* hid */
  int32_t n_res;          /* residual blocks */
  int32_t res_act;        /* 0 = ReLU, 1 = Tanh (activation inside/after residual blocks) */
  int32_t potential;      /* 0 = none; 1 = latent_ode (sigmoid(p[a]) + sigmoid(p[b]) - 1)^2 correction */
  int32_t pot_idx_a;      /* index into p of the "is_moving" logit      (latent: 12) */
  int32_t pot_idx_b;      /* index into p of the "is_stationary" logit  (latent: 8)  */
  float pot_strength;     /* config.correction_strength */
  float time_period;      /* 24.0 */
} ab200_drift_desc;

/* Number of fp32 parameters of the drift net, in `nn.Module.parameters()` order:
 *   w_in[hid, 2P+H+2], b_in[hid], { w_a[hid,hid], b_a[hid], w_b[hid,hid], b_b[hid] } x n_res,
 *   w_out[P, hid], b_out[P]          (row-major [out, in], exactly torch's nn.Linear storage). */
int64_t ab200_drift_param_count(const ab200_drift_desc* d);

int ab200_abi_version(void);
const char* ab200_status_string(int status);
/* Last CUDA runtime error string seen by this thread inside the library ("" if none). */
const char* ab200_last_cuda_error(void);

/* ---- fixed-grid RK4 (3/8 rule), whole trajectory in one launch --------------------------------
 * Replaces  odeint(self.odefunc, y0, times_union, method="rk4", rtol=, atol=)
 *           mode_sep/architecture/model.py:184-191     (tdq: fixed_grid.py RK4, rk_common.py rk4_alt_step_func)
 * y0     [B, D]  fp32, D = 2P+H          t [T] fp32 strictly increasing (host or device copy `t_host` is
 * y_path [T, B, D] fp32, row 0 = y0      used for validation only and may be NULL)
 */
size_t ab200_rk4_workspace_bytes(const ab200_drift_desc* d, int64_t B, int32_t T, int32_t precision);
int ab200_rk4_forward(const ab200_drift_desc* d, const float* w_flat, const float* y0, const float* t_dev,
                      const float* t_host, int64_t B, int32_t T, float* y_path, void* workspace,
                      size_t workspace_bytes, int32_t precision, ab200_stream_t stream);

/* Discrete adjoint of the call above == reverse-mode autograd through every solver op, which is what the
 * live reference training path does (mode_sep/train/train.py:162 `total.backward()`).
 * grad_y_path [T, B, D] : dL/dy_path (may alias nothing).   Outputs: grad_y0 [B, D], grad_w_flat
 * [ab200_drift_param_count] (OVERWRITTEN, not accumulated). */
size_t ab200_rk4_backward_workspace_bytes(const ab200_drift_desc* d, int64_t B, int32_t T, int32_t precision);
int ab200_rk4_backward(const ab200_drift_desc* d, const float* w_flat, const float* t_dev, const float* y_path,
                       const float* grad_y_path, int64_t B, int32_t T, float* grad_y0, float* grad_w_flat,
                       void* workspace, size_t workspace_bytes, int32_t precision, ab200_stream_t stream);

/* ---- one stand-alone drift evaluation f(t, y) (parity probe; also used by the generic path) ----
 * Replaces WrappedSDE.forward / ODEFunc.forward (files above).  y, out: [B, D]. */
size_t ab200_drift_eval_workspace_bytes(const ab200_drift_desc* d, int64_t B, int32_t precision);
int ab200_drift_eval(const ab200_drift_desc* d, const float* w_flat, float t, const float* y, int64_t B,
                     float* out, void* workspace, size_t workspace_bytes, int32_t precision,
                     ab200_stream_t stream);

/* Vector-Jacobian product of ONE drift evaluation f(t, y) = [v, net(p, v, h, t) (+ potential term), 0] in strict fp32:
 *   grad_y [B][D]  = J_y^T grad_out = [J_p^T g_v, g_p + J_v^T g_v, J_h^T g_v]     (OVERWRITTEN)
 *   grad_w_flat    = sum over agents of d(net)/dw ^T g_v, ab200_drift_param_count order     (OVERWRITTEN)
 * This is the backward of one `func(t, y)` call: it makes a solver written in PyTorch ops over the kernel-evaluated
 * drift differentiable (the reference trains by autograd through the solver: mode_sep/train/train.py:162,
 * latent_ode/train/train.py:73 -- dopri5 there), and it is the a^T df/dy, a^T df/dtheta evaluation of the continuous adjoint
 * (torchdiffeq odeint_adjoint, the call at latent_ode/architecture/ode_components.py:50).  Same shapes as ab200_rk4_backward. */
size_t ab200_drift_vjp_workspace_bytes(const ab200_drift_desc* d, int64_t B);
int ab200_drift_vjp(const ab200_drift_desc* d, const float* w_flat, float t, const float* y, const float* grad_out, int64_t B,
                    float* grad_y, float* grad_w_flat, void* workspace, size_t workspace_bytes, ab200_stream_t stream);

/* Embedding-space terms of the mode_sep training loss in one pass over pred_emb / v_t [B][T][E] (E = 64; element (b, t, e) at
 * ptr[b * stride_b + t * stride_t + e], strides in floats, multiples of 4):
 *   sums12[0] = sum over is_gt rows of |pred_emb - class_table[y_union]|^2,      [1] = number of is_gt rows      (mse_at_snaps,
 *   sums12[2], [3] = the same over stay_non_gt rows against y_stay                mode_sep/architecture/losses.py:24-31; train.py:126-135)
 *   sums12[4] = sum over travel rows of (m_travel - (d_prev - d_dest))+,          [5] = number of travel rows    (losses.py:56-73)
 *   sums12[6], [7] = sums of the two monotonicity hinges over pairs (t, t+1) of one travel segment, [8] = pairs  (losses.py:76-115)
 *   sums12[9] = sum over stay_non_gt rows of |v|^2 ;  sums12[10] = sum over gt_interior rows of (v_min - |v|)+^2 + (|v| - v_max)+^2,
 *   sums12[11] = number of gt_interior rows                                                          (mode_sep/train/train.py:137-153)
 * accumulated into 12 device doubles (zeroed by the caller).  The backward takes coef6 (device) = d total / d of the six sums
 * {mse_gt, mse_stay, margin, each monotonicity hinge, stay_vel, move_vel} and OVERWRITES grad_pred_emb / grad_v ([B][T][E]
 * contiguous) and ACCUMULATES into grad_class_table [Z][E].  Masks are bytes (torch.bool), indices int64 (-1 = none). */
int ab200_emb_losses_forward(const float* pred_emb, int64_t emb_stride_b, int64_t emb_stride_t, const float* v_t, int64_t v_stride_b,
                             int64_t v_stride_t, const float* class_table, const int64_t* y_union, const uint8_t* is_gt,
                             const int64_t* y_stay, const uint8_t* stay_non_gt, const uint8_t* travel_mask, const int64_t* prev_idx,
                             const int64_t* dest_idx, const uint8_t* gt_interior, int64_t B, int32_t T, int32_t E, int32_t Z,
                             float m_travel, float epsilon_mono, float v_min_move, float v_max_move, double* sums12, void* stream);
int ab200_emb_losses_backward(const float* pred_emb, int64_t emb_stride_b, int64_t emb_stride_t, const float* v_t, int64_t v_stride_b,
                              int64_t v_stride_t, const float* class_table, const int64_t* y_union, const uint8_t* is_gt,
                              const int64_t* y_stay, const uint8_t* stay_non_gt, const uint8_t* travel_mask, const int64_t* prev_idx,
                              const int64_t* dest_idx, const uint8_t* gt_interior, int64_t B, int32_t T, int32_t E, int32_t Z,
                              float m_travel, float epsilon_mono, float v_min_move, float v_max_move, const float* coef6,
                              float* grad_pred_emb, float* grad_v, float* grad_class_table, void* stream);

/* ---- generic-func path: fused Runge-Kutta stage combine (+ error norm) ------------------------
 * For an arbitrary `func` evaluated by the caller.  out[i] = y[i] + dt * sum_j coef[j] * k_j[i]
 * (tdq: rk_common.py `_runge_kutta_step`: yi = y0 + sum(k[..., :i+1] * (beta_i * dt))).
 * `k` is an array of `n_k` device pointers held in HOST memory; n_k <= 8. */
int ab200_rk_stage_combine(const float* y, const float* const* k, const float* coef_host, int32_t n_k, float dt,
                           float* out, int64_t n, ab200_stream_t stream);
/* Same pass additionally forms the embedded error estimate and accumulates
 *   sum_i ( (dt * sum_j cerr[j] k_j[i]) / (atol + rtol * max(|y0[i]|, |y1[i]|)) )^2   into *sumsq (device,
 * fp32, must be zeroed by the caller) -- tdq: misc.py `_compute_error_ratio` with `_rms_norm`.
 * y1 = y0 + dt * sum_j csol[j] k_j  is written to `y1_out` (skipped when `y1_out` is NULL). */
int ab200_rk_combine_errnorm(const float* y0, const float* const* k, const float* csol_host,
                             const float* cerr_host, int32_t n_k, float dt, float rtol, float atol,
                             float* y1_out, float* sumsq, int64_t n, ab200_stream_t stream);

/* ---- tensor-core STAGE path: one drift evaluation per launch, any explicit Runge-Kutta tableau ----------------
 * The drift is second order (dp/dt = v), so every stage input, step solution, dense-output row and embedded error
 * estimate of torchdiffeq's rk4 (rk_common.py rk4_alt_step_func) and dopri5 (rk_common.py _runge_kutta_step,
 * _adaptive_step; interp.py; misc.py _compute_error_ratio) is LINEAR in the step's base state y0 = [p0, v0, h] and
 * the stage accelerations a_1..a_s.  A stage launch evaluates
 *     p = p0 + in_cpv v0 + sum_j in_cpa[j] a_j ,  v = v0 + sum_j in_cva[j] a_j ,  a_out = net(p, v, h, t)
 * on the tcgen05 tensor cores (bf16 operands, fp32 accumulation in tensor memory) and optionally
 *     y_out = [p0 + out_cpv v0 + sum_j out_cpa[j] a_j ,  v0 + sum_j out_cva[j] a_j , h]     (index n_a = a_out)
 *     *err_sumsq += sum_i ( e_i / (atol + rtol max(|y0_i|, |y_out_i|)) )^2 ,  e.p = sum_j err_pa[j] a_j, e.v likewise.
 * Replaces the Python-level stage loop inside torchdiffeq that the reference's odeint calls run
 * (mode_sep/architecture/model.py:184-191, latent_ode/architecture/model.py:196).  mode_sep drift shape only. */
#define AB200_STAGE_MAX_A 7
typedef struct ab200_stage_desc {
  int32_t n_a;                                   /* accelerations combined into the stage input (0..7) */
  float in_cpv, in_cpa[AB200_STAGE_MAX_A], in_cva[AB200_STAGE_MAX_A];
  float t;                                       /* stage time */
  float out_cpv, out_cpa[AB200_STAGE_MAX_A + 1], out_cva[AB200_STAGE_MAX_A + 1];
  float err_pa[AB200_STAGE_MAX_A + 1], err_va[AB200_STAGE_MAX_A + 1];
  float rtol, atol;
} ab200_stage_desc;

/* LAYOUT of every per-agent fp32 buffer of the stage path (y0, y_out, a[], a_out, g_a, G_y0, G_a[] and the operands
 * of ab200_pv_combine*): "tile-blocked".  Rows are padded to Bp = ceil(B / 128) * 128; inside a 128-agent tile the
 * float4 with features 4 f4 .. 4 f4 + 3 of agent r is float4 number (tile * F/4 + f4) * 128 + r of the buffer, so a warp
 * whose lanes own consecutive agents (= consecutive tensor-memory lanes) touches one contiguous 512-byte segment per
 * access.  Padding rows of INPUT buffers must hold zeros (ab200_rows_block writes them; ab200_stage_forward* write zeros to
 * the padding rows of a_out / y_out, so output buffers may be uninitialised; no other kernel stores to padding rows).
 * ab200_rows_block / ab200_rows_unblock convert from / to the reference's row-major [B][F] tensors
 * (accumulate = 1: blocked += row-major). */
int ab200_rows_block(const float* src_rowmajor, float* dst_blocked, int64_t B, int32_t F, int32_t accumulate,
                     ab200_stream_t stream);
int ab200_rows_unblock(const float* src_blocked, float* dst_rowmajor, int64_t B, int32_t F, ab200_stream_t stream);

/* bf16 UMMA image of the drift weights (+ a status word); build once per parameter update with ab200_stage_pack */
size_t ab200_stage_image_bytes(const ab200_drift_desc* d);
int ab200_stage_pack(const ab200_drift_desc* d, const float* w_flat, void* image, size_t image_bytes, ab200_stream_t stream);
/* `a` : host array of n_a device pointers (blocked [Bp][P] fp32 each).  a_out (blocked [Bp][P]), y_out (blocked
 * [Bp][D]) and err_sumsq (device double, accumulated) may each be NULL.  operand_format: 0 = bf16 operands,
 * 1 = IEEE fp16 operands (same speed, 8x less rounding noise), 2 = fp16 weights with every ACTIVATION entering the
 * tensor core as a two-term fp16 split hi + lo (two MMAs per K-step) and biases / time features added in fp32: the
 * evaluation then carries no activation rounding at all, only the fixed fp16 rounding of the weights, which is what
 * dopri5's embedded error estimate needs at rtol = atol = 1e-5 (formats 1 / 0 make it take 2.4x / 14x the steps of the
 * fp32 reference there).  The image holds all three encodings of the weights.  A kernel whose bounded barrier wait
 * expired adds NaN to *err_sumsq (format 2) and sets the status word (ab200_stage_status_offset). */
int ab200_stage_forward(const ab200_drift_desc* d, const void* image, const float* y0, const float* const* a,
                        const ab200_stage_desc* s, int64_t B, float* a_out, float* y_out, double* err_sumsq,
                        int32_t operand_format, ab200_stream_t stream);
/* Several consecutive stages of ONE step (or step attempt) in a single launch: stage s reads a[0 .. stages[s].n_a) --
 * which may include buffers an earlier stage of the same launch wrote through a_out[] -- and writes a_out[s] (NULL =
 * not stored); only the last stage may produce y_out / err_sumsq (its out_* / err_* fields).  Each tile runs all stages
 * before the next tile, so the accelerations are re-read from L2, not HBM.  `a` must hold AB200_STAGE_MAX_A entries. */
int ab200_stage_forward_fused(const ab200_drift_desc* d, const void* image, const float* y0, const float* const* a,
                              const ab200_stage_desc* stages, int32_t n_stage, float* const* a_out, int64_t B,
                              float* y_out, double* err_sumsq, int32_t operand_format, ab200_stream_t stream);

/* One attempted Dormand-Prince 5(4) step (tdq dopri5.py tableau, rk_common.py `_runge_kutta_step`) = ab200_stage_forward_fused over
 * stages 2..7 with every stage descriptor built HERE from (t0, dt): a[0] holds the acceleration of the step's first stage (k_1,
 * the previous step's FSAL evaluation), a[1..6] receive a_2..a_7; y_out = the 5th-order solution, *err_sumsq += the squared
 * error-ratio sum of the embedded 4th-order estimate.  Exists because an adaptive solver reads the error norm on the host after
 * every attempt: the time between that read and the next launch is GPU idle time, and assembling ~300 tableau coefficients in
 * the caller's interpreter costs more than a forward launch over a 100k-agent shard takes. */
int ab200_dopri5_attempt(const ab200_drift_desc* d, const void* image, const float* y0, float* const* a, double t0, double dt,
                         int64_t B, float* y_out, double* err_sumsq, float rtol, float atol, int32_t operand_format, void* x_blobs,
                         int32_t save_level, ab200_stream_t stream);
/* `x_blobs` (operand_format 2 only; NULL = not wanted): 6 * ab200_stage_xblob_bytes(d, B, save_level) bytes that receive, for each
 * of the six evaluations of the attempt (stage i -> x_blobs + (i - 1) * ab200_stage_xblob_bytes), what the backward pass would
 * otherwise recompute, as the bf16 operand images its kernels consume:
 *   save_level 1: the stage input (352 B per agent-stage);
 *   save_level 2: + the five hidden activations and their ReLU masks (1,712 B per agent-stage).
 * A training step keeps the buffer of an ACCEPTED attempt and hands it to ab200_stage_backward_fused / ab200_wgrad_accumulate
 * with the same save_level: at level 1 the backward kernel loads the stage input instead of rebuilding it from (y0, a_j); at
 * level 2 it recomputes nothing (no stage input, no forward GEMMs, no activation spills) and reads neither y0 nor a. */
size_t ab200_stage_xblob_bytes(const ab200_drift_desc* d, int64_t B, int32_t save_level);
/* ab200_stage_forward_fused in the split-activation format (operand_format 2) that also SAVES, for every stage s of the launch, what
 * its backward pass would recompute into x_outs[s] (ab200_stage_xblob_bytes(d, B, save_level) bytes each; levels as above): the
 * fixed-grid rk4 training step uses it for its four stages (stage.py rk4_forward), ab200_dopri5_attempt is the same launch with the
 * Dormand-Prince descriptors built in C. */
int ab200_stage_forward_fused_save(const ab200_drift_desc* d, const void* image, const float* y0, const float* const* a,
                                   const ab200_stage_desc* stages, int32_t n_stage, float* const* a_out, int64_t B, float* y_out,
                                   double* err_sumsq, void* const* x_outs, int32_t save_level, ab200_stream_t stream);
/* The dense-output rows of an accepted dopri5 step at relative positions x[q] = (t_q - t0) / dt in (0, 1] (tdq interp.py
 * `_interp_fit` / `_interp_evaluate`: the quartic through y0, y1, y_mid, f0, f1, expanded over the seven stage derivatives):
 * out_rowmajor[q] (row-major [B][D]) for q < n_rows, all in one pass over (y0, a[0..6]). */
int ab200_dopri5_dense_rows(const ab200_drift_desc* d, const float* y0, const float* const* a, double dt, int32_t n_rows,
                            const double* x_host, int64_t B, float* const* out_rowmajor, ab200_stream_t stream);

/* Vector-Jacobian product of one stage = what autograd does for the ops of one `func` call inside the solver
 * (mode_sep/train/train.py:162).  The upstream gradient is assembled in the kernel as
 *     dL/da_out = g_base + sum_{l < n_g} dp[l] gx[l].p + dv[l] gx[l].v
 * where g_base (blocked [Bp][P], may be NULL when n_g > 0) is the step-level part (ab200_pv_combine_backward of the
 * step's outputs) and gx[l] (blocked [Bp][D]) are the gx_out of the LATER stages whose input used a_out with
 * coefficients (dp[l], dv[l]) = that stage's (in_cpa, in_cva) entry for a_out.  WRITES gx_out (blocked [Bp][D]) =
 * dL/d(stage input) = [g_p, g_v, g_h] and the layer (activation, gradient) blobs of tile i to blob index blob0 + i
 * of `spill` (sized by ab200_stage_spill_bytes(nblobs)); ab200_wgrad_accumulate turns filled blobs into weight
 * gradients.  `partial` (ab200_wgrad_partial_bytes, ZEROED by the caller before the first stage of a backward pass)
 * holds the per-SM partial weight gradients; ab200_wgrad_finalize writes grad_w_flat (OVERWRITTEN) in
 * ab200_drift_param_count order.  ab200_adjoint_gather folds the stages of one step into dL/dy0:
 *     out = base + sum_l [gx[l].p ; cpv[l] gx[l].p + gx[l].v ; gx[l].h]      (cpv[l] = that stage's in_cpv). */
size_t ab200_stage_spill_bytes(const ab200_drift_desc* d, int32_t nblobs);
size_t ab200_wgrad_partial_bytes(const ab200_drift_desc* d);
int ab200_stage_backward(const ab200_drift_desc* d, const void* image, const float* y0, const float* const* a,
                         const ab200_stage_desc* s, int64_t B, const float* g_base, const float* const* gx, int32_t n_g,
                         const float* dp_host, const float* dv_host, float* gx_out, void* spill, size_t spill_bytes,
                         int32_t blob0, int32_t nblobs, void* partial, ab200_stream_t stream);
int ab200_adjoint_gather(const ab200_drift_desc* d, const float* base, const float* const* gx, int32_t n,
                         const float* cpv_host, int64_t B, float* out, ab200_stream_t stream);
/* The backward stages of ONE step in a single launch, latest stage first: entry s of `stages` / g_base / gx_out / n_g is
 * one stage; its n_g[s] gradient sources are gx_src[s * 7 + l] >= 0 -> gx_out of an EARLIER entry of this call, or
 * < 0 -> the external buffer gx_ext[-1 - index]; coefficients dp/dv[s * 7 + l].  Blobs of entry s go to blob indices
 * blob0 + s * ntiles + tile.  Each tile runs all stages before the next tile, so the gx of later stages and the stage
 * inputs are re-read from L2 (the blobs are written with streaming stores so that they do not evict them).
 * `a` must hold AB200_STAGE_MAX_A entries. */
int ab200_stage_backward_fused(const ab200_drift_desc* d, const void* image, const float* y0, const float* const* a,
                               const ab200_stage_desc* stages, int32_t n_stage, const float* const* g_base,
                               float* const* gx_out, const int32_t* n_g, const int32_t* gx_src, const float* const* gx_ext,
                               const float* dp_host, const float* dv_host, int64_t B, void* spill, size_t spill_bytes,
                               int32_t blob0, int32_t nblobs, void* partial, const void* const* x_blobs, int32_t save_level,
                               float* y0_accum, float* const* upstream_out, ab200_stream_t stream);
/* GATHER entry: when gx_out[n_stage - 1] is NULL the last entry (allowed on top of AB200_STAGE_MAX_A stages) runs no network
 * backward and writes no blobs (it does not count towards blob indices); from the gx of its sources -- entries of this call --
 *   upstream_out[n_stage - 1] (if upstream_out and that entry are non-NULL; fp32 blocked [Bp][P])
 *        = g_base + sum_l dp[l] gx_l.p + dv[l] gx_l.v        (e.g. the gradient handed to the FSAL evaluation of the previous step)
 *   y0_accum (if non-NULL; blocked [Bp][D], initialised by the caller with the step-level dL/dy0), in place
 *       += sum_l [gx_l.p ; in_cpv_l gx_l.p + gx_l.v ; gx_l.h]                   (what ab200_adjoint_gather does in a pass of its own)
 * while those gx tiles are still in L2.  y0_accum / upstream_out are ignored (must be NULL) without a gather entry. */
/* x_blobs: NULL, or a host array of n_stage device pointers; entry s non-NULL = what the forward launch saved for that stage
 * (ab200_dopri5_attempt, same save_level): it is loaded instead of rebuilt / recomputed, and the corresponding parts of that
 * stage's spill blobs are NOT written (pass the same pointers and level to ab200_wgrad_accumulate). */
/* ab200_adjoint_gather and ab200_stage_upstream over the same gx list in ONE pass (each gx is read once):
 *     out = base + sum_l [...]  and  g_a_out = g_base + sum_l dp[l] gx[l].p + dv[l] gx[l].v   (dp/dv may be 0 for some l). */
int ab200_adjoint_gather_upstream(const ab200_drift_desc* d, const float* base, const float* const* gx, int32_t n,
                                  const float* cpv_host, int64_t B, float* out, const float* g_base, const float* dp_host,
                                  const float* dv_host, float* g_a_out, ab200_stream_t stream);
/* ab200_pv_combine_backward for up to 6 linear outputs of the same step in ONE pass (source i: gradient g[i] and its
 * combination cpv[i], cpa[i * 8 + j], cva[i * 8 + j], j < n_a): the accumulators are written once instead of once per output
 * (a dopri5 step's end state plus the dense-output rows that fall inside it).  Source i is blocked [Bp][D], or -- bit i of
 * rowmajor_mask set -- a ROW-MAJOR [B][D] row of the caller's own gradient tensor (dL/dy_path[k]), read in place.
 * add_a (or NULL): blocked [Bp][P] buffer added to G_a[add_index] in the same pass (dopri5: the gradient the following step hands
 * back to this step's FSAL evaluation a_7). */
int ab200_pv_combine_backward_multi(const ab200_drift_desc* d, const float* const* g, int32_t n_src, const float* cpv_host,
                                    const float* cpa_host, const float* cva_host, int32_t n_a, int64_t B, float* G_y0,
                                    float* const* G_a, int32_t accumulate, int32_t rowmajor_mask, const float* add_a, int32_t add_index,
                                    ab200_stream_t stream);
/* The upstream gradient of a stage written out as a buffer instead of being consumed by ab200_stage_backward:
 *     g_a_out (blocked [Bp][P]) = g_base + sum_l dp[l] gx[l].p + dv[l] gx[l].v
 * dopri5's first stage of a step IS the last (FSAL) evaluation of the previous step (tdq rk_common.py _adaptive_step:
 * f1 is carried over as f0), so its gradient is handed to that step's stage 7 rather than differentiated twice. */
int ab200_stage_upstream(const ab200_drift_desc* d, const float* g_base, const float* const* gx, int32_t n_g,
                         const float* dp_host, const float* dv_host, int64_t B, float* g_a_out, ab200_stream_t stream);
int ab200_wgrad_accumulate(const ab200_drift_desc* d, const void* spill, int32_t nblobs, int32_t used, void* partial,
                           const void* const* x_blobs, int32_t n_x_blobs, int32_t ntiles, int32_t save_level,
                           ab200_stream_t stream);
/* x_blobs (n_x_blobs <= 8 host entries, one per run of `ntiles` consecutive blobs; NULL entry / n_x_blobs 0 = read the layer
 * inputs from `spill`): forward-saved buffers of the stages whose blobs occupy [k * ntiles, (k + 1) * ntiles). */
int ab200_wgrad_finalize(const ab200_drift_desc* d, const void* partial, float* grad_w_flat, ab200_stream_t stream);
/* 1 if any tensor-core kernel that used `image` / `partial` hit its bounded mbarrier wait (a bug, never expected) */
int ab200_stage_status_offset(const ab200_drift_desc* d, int64_t* image_status_byte, int64_t* partial_status_byte);

/* Elementwise companions (one HBM pass): out = [p0 + cpv v0 + sum cpa[j] a_j, v0 + sum cva[j] a_j, h]  (dense-output
 * rows: tdq interp.py _interp_evaluate; step solutions) and the adjoint scatter
 *   G_y0.p (+)= g.p ; G_y0.v (+)= cpv g.p + g.v ; G_y0.h (+)= g.h ; G_a[j] (+)= cpa[j] g.p + cva[j] g.v
 * (accumulate = 0 overwrites).  n_a <= 8; `a` / `G_a` are host arrays of device pointers. */
int ab200_pv_combine(const ab200_drift_desc* d, const float* y0, const float* const* a, int32_t n_a, float cpv,
                     const float* cpa_host, const float* cva_host, int64_t B, float* out, ab200_stream_t stream);
/* ab200_pv_combine whose result is the reference's ROW-MAJOR [B][D] trajectory row (blocked inputs, transposed on chip) */
int ab200_pv_combine_rowmajor(const ab200_drift_desc* d, const float* y0, const float* const* a, int32_t n_a, float cpv,
                              const float* cpa_host, const float* cva_host, int64_t B, float* out_rowmajor, ab200_stream_t stream);
/* n_rows row-major outputs of the SAME (y0, a[]) in one pass (the dense-output rows that fall inside one dopri5 step: y0 and
 * the accelerations are read once): out_rowmajor[q] = combination (cpv[q], cpa[q * 8 + j], cva[q * 8 + j], j < n_a). */
int ab200_pv_combine_rowmajor_multi(const ab200_drift_desc* d, const float* y0, const float* const* a, int32_t n_a, int32_t n_rows,
                                    const float* cpv_host, const float* cpa_host, const float* cva_host, int64_t B,
                                    float* const* out_rowmajor, ab200_stream_t stream);
int ab200_pv_combine_backward(const ab200_drift_desc* d, const float* g, int32_t n_a, float cpv, const float* cpa_host,
                              const float* cva_host, int64_t B, float* G_y0, float* const* G_a, int32_t accumulate,
                              ab200_stream_t stream);

/* ---- SDE branch: one Euler-Maruyama step ----------------------------------------------------------------------
 * y_out = y + drift * dt + diffusion * sqrt(dt) * xi,  xi ~ N(0, 1): the step of torchsde's `sdeint(..., method="euler")`
 * for diagonal Ito noise (latent_ode/architecture/model.py:119-130,192-194; mode_sep/architecture/model.py:79-89,158-182).
 * diffusion is [D] (diffusion_per_row = 0) or [B][D].  The noise is counter-based (Philox4x32-10 keyed by `seed`,
 * counter = (element group, step); Box-Muller), specified in csrc/sde_em.cu and restated in oracle/sde_oracle.py: an
 * agent's noise does not depend on the batch it is in.  xi_out (may be NULL) receives the normals.  D % 4 == 0. */
int ab200_sde_euler_step(const float* y, const float* drift, const float* diffusion, int32_t diffusion_per_row, int64_t B,
                         int32_t D, float dt, uint64_t seed, uint64_t step, float* y_out, float* xi_out, void* stream);

/* ---- optimiser step over the flat parameter / gradient buffers ----------------------------------------------------
 * Replaces  torch.nn.utils.clip_grad_norm_(params, max_norm) ; torch.optim.Adam(lr, weight_decay).step()
 * (mode_sep/train/train.py:68,163-164) on the flat fp32 buffer the NCCL gradient all-reduce already uses: two launches,
 * no host synchronisation.  ab200_grad_sumsq writes sum(g^2) (double, device); ab200_adam_step applies the clip
 * coefficient min(1, max_grad_norm / (sqrt(sumsq) + 1e-6)) when max_grad_norm > 0 and grad_sumsq != NULL, then Adam
 * with L2 weight decay (torch.optim.Adam semantics, not AdamW), `step` counted from 1. */
int ab200_grad_sumsq(const float* flat_grad, int64_t n, double* sumsq_out, void* stream);
int ab200_adam_step(float* flat_param, const float* flat_grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                    float beta1, float beta2, float eps, float weight_decay, int32_t step, float max_grad_norm,
                    const double* grad_sumsq, void* stream);

/* ---- fused classification head + label prediction ------------------------------------------------------------
 * Replaces  emb_norm = pred_emb / (|pred_emb| + 1e-8); table_norm = class_table / (|class_table| + 1e-8);
 *           logits = einsum("bte,ze->btz", emb_norm, table_norm) / softmax_tau      mode_sep/architecture/model.py:196-199
 *           labels = logits.argmax(-1)                      mode_sep/inference/inference.py:57,63; train/train.py:168
 * without materialising [B, T, Z] (3.9 TB at 1M agents x 97 x 10k zones).  pred_emb [M][E] fp32 row-major with
 * M = B*T rows, class_table [Z][E] fp32, E = 64.  labels [M] int64 (first index on ties); best_logit [M] fp32 (the
 * winning logit, may be NULL).  tcgen05 nominates the two best zones per row on 2-term bf16 splits, the winner is
 * decided by an fp32 re-score.  Workspace: ab200_head_workspace_bytes(Z, E). */
size_t ab200_head_workspace_bytes(int32_t Z, int32_t E);
int ab200_head_argmax(const float* pred_emb, const float* class_table, int64_t M, int32_t Z, int32_t E, float tau,
                      int64_t* labels, float* best_logit, void* workspace, size_t workspace_bytes, ab200_stream_t stream);

/* Cross-entropy forward of the same head -- F.cross_entropy(logits[mask], y[mask]) of `ce_at_snaps`
 * (mode_sep/architecture/losses.py:14-22) without the [M, Z] logits: per row lse[m] = log sum_z exp(logit[m, z]) (streamed
 * over the zones on the tensor cores, split-bf16 operands) and target_logit[m] = logit[m, target[m]] (fp32 re-score);
 * row loss = lse - target_logit, masking and the mean are the caller's.  target values outside [0, Z) are read as zone 0.
 * labels may be NULL; when given it receives the argmax of ab200_head_argmax from the same pass.
 * dist_mat [Z][Z] + expected_dist [M] (both or neither): the same sweep also yields `expected_distance_at_snaps`
 * (losses.py:34-44) per row, expected_dist[m] = sum_z softmax[m,z] dist_mat[target[m], z]; rows sorted by target share
 * their distance row in L1/L2 (the Python wrapper sorts). */
int ab200_head_ce_forward(const float* pred_emb, const float* class_table, const int64_t* target, int64_t M, int32_t Z,
                          int32_t E, float tau, float* lse, float* target_logit, int64_t* labels, const float* dist_mat,
                          float* expected_dist, void* workspace, size_t workspace_bytes, void* stream);

/* Backward of the cross-entropy head on tensor cores: with dlogit[m,z] = grad_rows[m] (softmax[m,z] - [z = target[m]])
 * recomputed tile by tile from `lse` (never stored), writes the gradients w.r.t. the NORMALISED vectors
 *   grad_emb_normalised[m,:] = (1/tau) sum_z dlogit[m,z] table^[z,:]      grad_table_normalised[z,:] = (1/tau) sum_m dlogit[m,z] emb^[m,:]
 * (autograd through the reference's einsum + F.cross_entropy, model.py:196-199 / losses.py:14-22); the caller applies
 * the Jacobian of x / (|x| + 1e-8).  With grad_dist_rows / expected_dist / dist_mat (all three or none) the expected-distance
 * term is added: dlogit[m,z] += grad_dist_rows[m] softmax[m,z] (dist_mat[target[m], z] - expected_dist[m]).
 * Deterministic (no atomics).  ab200_head_ce_backward_status copies the kernels'
 * barrier-timeout word to the host (synchronises the stream; 0 = clean). */
size_t ab200_head_ce_backward_workspace_bytes(int64_t M, int32_t Z, int32_t E);
int ab200_head_ce_backward(const float* pred_emb, const float* class_table, const int64_t* target, const float* lse,
                           const float* grad_rows, const float* grad_dist_rows, const float* expected_dist, const float* dist_mat,
                           int64_t M, int32_t Z, int32_t E, float tau, float* grad_emb_normalised, float* grad_table_normalised,
                           void* workspace, size_t workspace_bytes, void* stream);
int ab200_head_ce_backward_status(const void* workspace, int64_t M, int32_t Z, int32_t* status_host, void* stream);

/* ---- graph attention over the zone graph (the `gnn_embed` slot) ------------------------------------
 * No reference implementation exists (README.md:5,57,80 promise it; pyproject.toml:25 declares torch-geometric
 * 2.6.1, never imported): semantics are PyG `GATConv` -- x'_i = ||_h sum_{j in N(i)+i}
 * alpha^h_ij W^h x_j + b, alpha = softmax_j LeakyReLU(a_src^h . W^h x_j + a_dst^h . W^h x_i).
 * Graph: CSR sorted by DESTINATION over the symmetrised zone graph with self loops (data_generator/
 * load_data.py:104-110 builds the same adjacency): rowptr[Z+1], col[nnz] = source of each in-edge (int32).
 * The transposed CSR (rowptr_t/col_t by SOURCE) and eid_t[nnz] (index of that edge in the by-destination order)
 * are needed by the backward pass only.  W [heads*F_out, F_in] (nn.Linear storage), att_* [heads*F_out].
 * Saved for backward (caller-allocated): xw [Z, heads*F_out], a_src/a_dst [Z, heads], alpha [nnz, heads].
 * out: [Z, heads*F_out] (concat=1) or [Z, F_out] (concat=0, mean over heads).  F_out: power of two >= 4. */
int ab200_gat_forward(const int32_t* rowptr, const int32_t* col, int32_t Z, int32_t nnz, const float* x, int32_t F_in,
                      const float* W, const float* att_src, const float* att_dst, const float* bias, int32_t heads,
                      int32_t F_out, int32_t concat, float negative_slope, float* out, float* xw, float* a_src,
                      float* a_dst, float* alpha, ab200_stream_t stream);
size_t ab200_gat_backward_workspace_bytes(int32_t Z, int32_t nnz, int32_t heads, int32_t F_out);
/* grad_x and grad_bias may be NULL.  All gradient outputs are OVERWRITTEN. */
int ab200_gat_backward(const int32_t* rowptr, const int32_t* col, const int32_t* rowptr_t, const int32_t* col_t,
                       const int32_t* eid_t, int32_t Z, int32_t nnz, const float* x, int32_t F_in, const float* W,
                       const float* att_src, const float* att_dst, int32_t heads, int32_t F_out, int32_t concat,
                       float negative_slope, const float* xw, const float* a_src, const float* a_dst,
                       const float* alpha, const float* grad_out, float* grad_x, float* grad_W, float* grad_att_src,
                       float* grad_att_dst, float* grad_bias, void* workspace, size_t workspace_bytes,
                       ab200_stream_t stream);

/* ---- self-test of the tcgen05 plumbing (one CTA): D[128,N] = A[128,K] * B[N,K]^T, bf16 operands, fp32 result.
 * a_mode 0/1/2 = A in TMEM / smem un-swizzled / smem 128B-swizzled; b_mode 1/2 = B un-swizzled / 128B-swizzled;
 * mode 3 (either operand) = smem MN-major un-swizzled, mode 4 = the same bytes with LBO/SBO exchanged.
 * `status` (device int) receives 0, or 1 if the MMA completion barrier timed out.  No reference counterpart. */
int ab200_debug_umma_probe(const float* A, const float* B, float* D, int32_t N, int32_t K, int32_t a_mode,
                           int32_t b_mode, int32_t* status, ab200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ANANKE_B200_H_ */
