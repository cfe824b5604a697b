// extern "C" surface of libananke_b200.so -- see include/ananke_b200.h for the contract of each symbol.
#include <string.h>
#include "common.cuh"
#include "wgrad_layout.cuh"

namespace ab200 {
static thread_local char g_cuda_err[256] = "";
void set_cuda_error(cudaError_t e) {
  strncpy(g_cuda_err, cudaGetErrorString(e), sizeof(g_cuda_err) - 1);
  g_cuda_err[sizeof(g_cuda_err) - 1] = 0;
}

// implemented in the kernel translation units
int pack_drift(const ab200_drift_desc* d, const float* w_flat, float* packed, cudaStream_t st);
int rk4_forward_f32(const ab200_drift_desc* d, const float* packed, const float* y0, const float* t_dev, int64_t B, int T,
                    float* y_path, cudaStream_t st);
int drift_eval_f32(const ab200_drift_desc* d, const float* packed, float t, const float* y, int64_t B, float* out,
                   cudaStream_t st);
size_t rk4_backward_f32_workspace(const ab200_drift_desc* d, int64_t B, int T);
int rk4_backward_f32(const ab200_drift_desc* d, const float* w_flat, const float* t_dev, const float* y_path,
                     const float* grad_y_path, int64_t B, int T, float* grad_y0, float* grad_w_flat, void* ws, size_t ws_bytes,
                     cudaStream_t st);
int drift_vjp_f32(const ab200_drift_desc* d, const float* w_flat, float t, const float* y, const float* g, int64_t B, float* grad_y,
                  float* grad_w_flat, void* ws, size_t ws_bytes, cudaStream_t st);
int rk_stage_combine(const float* y, const float* const* k, const float* coef, int n_k, float dt, float* out, int64_t n,
                     cudaStream_t st);
int rk_combine_errnorm(const float* y0, const float* const* k, const float* csol, const float* cerr, int n_k, float dt,
                       float rtol, float atol, float* y1_out, float* sumsq, int64_t n, cudaStream_t st);

int umma_probe(const float* A, const float* B, float* D, int N, int K, int a_mode, int b_mode, int* status, cudaStream_t st);

size_t rk4_tc_workspace(const ab200_drift_desc* d);
int rk4_forward_tc(const ab200_drift_desc* d, const float* w_flat, const float* y0, const float* t_dev, int64_t B, int T,
                   float* y_path, void* ws, size_t ws_bytes, cudaStream_t st);

int gat_forward(const int* rowptr, const int* col, int Z, int nnz, const float* x, int F_in, const float* W, const float* att_src,
                const float* att_dst, const float* bias, int heads, int F_out, int concat, float slope, float* out, float* xw,
                float* a_src, float* a_dst, float* alpha, cudaStream_t st);
size_t gat_backward_workspace(int Z, int nnz, int heads, int F_out);
int gat_backward(const int* rowptr, const int* col, const int* rowptr_t, const int* col_t, const int* eid_t, int Z, int nnz,
                 const float* x, int F_in, const float* W, const float* att_src, const float* att_dst, int heads, int F_out,
                 int concat, float slope, const float* xw, const float* a_src, const float* a_dst, const float* alpha,
                 const float* gout, float* grad_x, float* grad_W, float* grad_att_src, float* grad_att_dst, float* grad_bias,
                 void* ws, size_t ws_bytes, cudaStream_t st);

size_t stage_tc_image_bytes();
int stage_tc_pack(const float* w_flat, uint8_t* image, cudaStream_t st);
int stage_fwd_tc(const ab200_drift_desc* d, const uint8_t* image, const float* y0, const float* const* a_ptrs, const void* desc_v,
                 int64_t B, float* a_out, float* y_out, double* err_sumsq, int half_ops, cudaStream_t st);
int stage_fwd_tc_multi(const ab200_drift_desc* d, const uint8_t* image, const float* y0, const float* const* a_ptrs, const void* descs_v,
                       int n_stage, float* const* a_outs, int64_t B, float* y_out, double* err_sumsq, int half_ops, cudaStream_t st);
int stage_fwd2_tc_multi(const ab200_drift_desc* d, const uint8_t* image, const float* y0, const float* const* a_ptrs, const void* descs_v,
                        int n_stage, float* const* a_outs, int64_t B, float* y_out, double* err_sumsq, void* const* x1_outs, int save_level,
                        cudaStream_t st);
int stage_bwd_tc(const ab200_drift_desc* d, const uint8_t* image, const float* y0, const float* const* a_ptrs, const void* desc_v,
                 int64_t B, const float* g_base, const float* const* gx_ptrs, int n_g, const float* dp, const float* dv, float* gx_out,
                 void* spill, int blob0, int nblobs, float* g_bout, cudaStream_t st);
int pv_combine_bwd_multi(const ab200_drift_desc* d, const float* const* g, int n_src, const float* cpv, const float* cpa, const float* cva,
                         int n_a, int64_t B, float* G_y0, float* const* G_a, int accumulate, int rowmajor_mask, const float* add_a, int add_idx,
                         cudaStream_t st);
int ga_assemble(const ab200_drift_desc* d, const float* base, const float* const* gx, int n, const float* dp, const float* dv,
                int64_t B, float* out, cudaStream_t st);
int stage_bwd_tc_multi(const ab200_drift_desc* d, const uint8_t* image, const float* y0, const float* const* a_ptrs,
                       const void* descs_v, int n_stage, const float* const* g_base, float* const* gx_out, const int32_t* n_g,
                       const int32_t* gx_src, const float* const* gx_ext, const float* dp, const float* dv, int64_t B, void* spill,
                       int blob0, int nblobs, float* g_bout, const void* const* x1_in, int save_level, float* y0_acc,
                       float* const* ga_out, cudaStream_t st);
int adjoint_gather(const ab200_drift_desc* d, const float* base, const float* const* gx, int n, const float* cpv, int64_t B, float* out,
                   const float* ga_base, const float* dp, const float* dv, float* ga_out, cudaStream_t st);
size_t wgrad_spill_bytes(int nblobs);
size_t wgrad_partial_bytes();
int wgrad_num_ctas();
float* wgrad_bout_ptr(void* partial);
int wgrad_tc(const void* spill, int nblobs, int used, void* partial, const void* const* x1_ext, int n_x1, int ntiles, int save_level,
             cudaStream_t st);
int wgrad_finalize(const void* partial, float* grad_w_flat, cudaStream_t st);
int pv_combine(const ab200_drift_desc* d, const float* y0, const float* const* a_ptrs, int n_a, float cpv, const float* cpa,
               const float* cva, int64_t B, float* out, cudaStream_t st);
int pv_combine_rowmajor(const ab200_drift_desc* d, const float* y0, const float* const* a_ptrs, int n_a, float cpv, const float* cpa,
                        const float* cva, int64_t B, float* out_rowmajor, cudaStream_t st);
int pv_combine_rowmajor_multi(const ab200_drift_desc* d, const float* y0, const float* const* a_ptrs, int n_a, int n_rows, const float* cpv,
                              const float* cpa, const float* cva, int64_t B, float* const* out_rowmajor, cudaStream_t st);
int pv_combine_bwd(const ab200_drift_desc* d, const float* g, int n_a, float cpv, const float* cpa, const float* cva, int64_t B,
                   float* G_y0, float* const* G_a, int accumulate, cudaStream_t st);

int rows_transpose(const float* src, float* dst, int64_t B, int F, int mode, cudaStream_t st);

int sde_euler_step(const float* y, const float* f, const float* g, int g_per_row, int64_t B, int D, float dt, uint64_t seed,
                   uint64_t step, float* y_out, float* xi_out, cudaStream_t st);
int grad_sumsq(const float* g, int64_t n, double* out, cudaStream_t st);
int adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps, float wd, int step,
              float max_norm, const double* sumsq, cudaStream_t st);
size_t head_workspace_bytes(int Z);
size_t head_ce_backward_workspace_bytes(int64_t M, int Z);
int head_ce_backward(const float* emb, const float* table, const int64_t* target, const float* lse, const float* g_rows,
                     const float* g_dist_rows, const float* exp_dist, const float* dist, int64_t M, int Z, int E, float tau,
                     float* d_emb_n, float* d_table_n, void* ws, size_t ws_bytes, cudaStream_t st);
int head_ce_backward_status(const void* ws, int64_t M, int Z, int* host_out, cudaStream_t st);
int head_ce_forward(const float* emb, const float* table, const int64_t* target, int64_t M, int Z, int E, float tau, float* lse,
                    float* tgt_logit, int64_t* labels, const float* dist, float* exp_dist, void* ws, size_t ws_bytes, cudaStream_t st);
int head_argmax(const float* emb, const float* table, int64_t M, int Z, int E, float tau, int64_t* labels, float* best, void* ws,
                size_t ws_bytes, cudaStream_t st);

static bool stage_shape_ok(const ab200_drift_desc* d) {
  return d && d->pos_dim == 64 && d->ctx_dim == 32 && d->hidden == 128 && d->n_res == 2 && d->res_act == 0 && d->potential == 0;
}

static bool desc_ok(const ab200_drift_desc* d) {
  return d && d->pos_dim > 0 && d->ctx_dim >= 0 && d->hidden > 0 && d->n_res >= 0 && (d->res_act == 0 || d->res_act == 1);
}
static size_t packed_bytes(const ab200_drift_desc* d) {
  const PackLayout L{d->pos_dim, d->ctx_dim, d->hidden, d->n_res};
  return align_up((size_t)L.total() * sizeof(float), 256);
}
}  // namespace ab200

using namespace ab200;

extern "C" {

int ab200_abi_version(void) { return AB200_ABI_VERSION; }

const char* ab200_status_string(int s) {
  switch (s) {
    case AB200_OK: return "ok";
    case AB200_ERR_BAD_ARG: return "bad argument";
    case AB200_ERR_UNSUPPORTED: return "drift-net shape or precision not instantiated in this build";
    case AB200_ERR_WORKSPACE: return "workspace too small";
    case AB200_ERR_CUDA: return "CUDA error";
    case AB200_ERR_NOT_MONOTONE: return "t must be strictly increasing or decreasing";
    case AB200_ERR_DT_UNDERFLOW: return "underflow in dt";
    case AB200_ERR_MAX_STEPS: return "max_num_steps exceeded";
    default: return "unknown status";
  }
}

const char* ab200_last_cuda_error(void) { return g_cuda_err; }

int64_t ab200_drift_param_count(const ab200_drift_desc* d) {
  if (!desc_ok(d)) return AB200_ERR_BAD_ARG;
  const FlatLayout F{d->pos_dim, d->ctx_dim, d->hidden, d->n_res};
  return F.total();
}

size_t ab200_rk4_workspace_bytes(const ab200_drift_desc* d, int64_t B, int32_t T, int32_t precision) {
  if (!desc_ok(d)) return 0;
  (void)B; (void)T;
  if (precision == AB200_PREC_BF16) return rk4_tc_workspace(d);
  return packed_bytes(d);
}

int ab200_rk4_forward(const ab200_drift_desc* d, const float* w_flat, const float* y0, const float* t_dev,
                      const float* t_host, int64_t B, int32_t T, float* y_path, void* workspace, size_t workspace_bytes,
                      int32_t precision, ab200_stream_t stream) {
  if (!desc_ok(d) || !w_flat || !y0 || !t_dev || !y_path || !workspace || B <= 0 || T < 1) return AB200_ERR_BAD_ARG;
  if (t_host)
    for (int i = 0; i + 1 < T; ++i)
      if (!(t_host[i + 1] > t_host[i])) return AB200_ERR_NOT_MONOTONE;
  if (precision == AB200_PREC_BF16 && rk4_tc_workspace(d) == 0) return AB200_ERR_UNSUPPORTED;
  if (workspace_bytes < ab200_rk4_workspace_bytes(d, B, T, precision)) return AB200_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == AB200_PREC_F32) {
    int rc = pack_drift(d, w_flat, (float*)workspace, st);
    if (rc) return rc;
    return rk4_forward_f32(d, (const float*)workspace, y0, t_dev, B, T, y_path, st);
  }
  if (precision == AB200_PREC_BF16) {
    if (rk4_tc_workspace(d) == 0) return AB200_ERR_UNSUPPORTED;
    return rk4_forward_tc(d, w_flat, y0, t_dev, B, T, y_path, workspace, workspace_bytes, st);
  }
  return AB200_ERR_UNSUPPORTED;
}

size_t ab200_rk4_backward_workspace_bytes(const ab200_drift_desc* d, int64_t B, int32_t T, int32_t precision) {
  if (!desc_ok(d)) return 0;
  (void)precision;
  return rk4_backward_f32_workspace(d, B, T);
}

int ab200_rk4_backward(const ab200_drift_desc* d, const float* w_flat, const float* t_dev, const float* y_path,
                       const float* grad_y_path, int64_t B, int32_t T, float* grad_y0, float* grad_w_flat, void* workspace,
                       size_t workspace_bytes, int32_t precision, ab200_stream_t stream) {
  if (!desc_ok(d) || !w_flat || !t_dev || !y_path || !grad_y_path || !grad_y0 || !grad_w_flat || !workspace || B <= 0 || T < 1)
    return AB200_ERR_BAD_ARG;
  if (workspace_bytes < ab200_rk4_backward_workspace_bytes(d, B, T, precision)) return AB200_ERR_WORKSPACE;
  if (precision != AB200_PREC_F32) return AB200_ERR_UNSUPPORTED;
  return rk4_backward_f32(d, w_flat, t_dev, y_path, grad_y_path, B, T, grad_y0, grad_w_flat, workspace, workspace_bytes,
                          (cudaStream_t)stream);
}

size_t ab200_drift_eval_workspace_bytes(const ab200_drift_desc* d, int64_t B, int32_t precision) {
  if (!desc_ok(d)) return 0;
  (void)B; (void)precision;
  return packed_bytes(d);
}

int ab200_drift_eval(const ab200_drift_desc* d, const float* w_flat, float t, const float* y, int64_t B, float* out,
                     void* workspace, size_t workspace_bytes, int32_t precision, ab200_stream_t stream) {
  if (!desc_ok(d) || !w_flat || !y || !out || !workspace || B <= 0) return AB200_ERR_BAD_ARG;
  if (workspace_bytes < packed_bytes(d)) return AB200_ERR_WORKSPACE;
  if (precision != AB200_PREC_F32) return AB200_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = pack_drift(d, w_flat, (float*)workspace, st);
  if (rc) return rc;
  return drift_eval_f32(d, (const float*)workspace, t, y, B, out, st);
}

size_t ab200_drift_vjp_workspace_bytes(const ab200_drift_desc* d, int64_t B) {
  if (!desc_ok(d) || B <= 0) return 0;
  return rk4_backward_f32_workspace(d, B, 1);
}

int ab200_drift_vjp(const ab200_drift_desc* d, const float* w_flat, float t, const float* y, const float* grad_out, int64_t B,
                    float* grad_y, float* grad_w_flat, void* workspace, size_t workspace_bytes, ab200_stream_t stream) {
  if (!desc_ok(d) || !w_flat || !y || !grad_out || !grad_y || !grad_w_flat || !workspace || B <= 0) return AB200_ERR_BAD_ARG;
  if (workspace_bytes < rk4_backward_f32_workspace(d, B, 1)) return AB200_ERR_WORKSPACE;
  return drift_vjp_f32(d, w_flat, t, y, grad_out, B, grad_y, grad_w_flat, workspace, workspace_bytes, (cudaStream_t)stream);
}

int ab200_rk_stage_combine(const float* y, const float* const* k, const float* coef_host, int32_t n_k, float dt, float* out,
                           int64_t n, ab200_stream_t stream) {
  if (!y || !out || n <= 0 || (n_k > 0 && (!k || !coef_host))) return AB200_ERR_BAD_ARG;
  return rk_stage_combine(y, k, coef_host, n_k, dt, out, n, (cudaStream_t)stream);
}

int ab200_rk_combine_errnorm(const float* y0, const float* const* k, const float* csol_host, const float* cerr_host,
                             int32_t n_k, float dt, float rtol, float atol, float* y1_out, float* sumsq, int64_t n,
                             ab200_stream_t stream) {
  if (!y0 || !k || !csol_host || !cerr_host || !sumsq || n <= 0) return AB200_ERR_BAD_ARG;
  return rk_combine_errnorm(y0, k, csol_host, cerr_host, n_k, dt, rtol, atol, y1_out, sumsq, n, (cudaStream_t)stream);
}

int ab200_rows_block(const float* src_rowmajor, float* dst_blocked, int64_t B, int32_t F, int32_t accumulate,
                     ab200_stream_t stream) {
  if (!src_rowmajor || !dst_blocked || B <= 0) return AB200_ERR_BAD_ARG;
  return rows_transpose(src_rowmajor, dst_blocked, B, F, accumulate ? 1 : 0, (cudaStream_t)stream);
}

int ab200_rows_unblock(const float* src_blocked, float* dst_rowmajor, int64_t B, int32_t F, ab200_stream_t stream) {
  if (!src_blocked || !dst_rowmajor || B <= 0) return AB200_ERR_BAD_ARG;
  return rows_transpose(src_blocked, dst_rowmajor, B, F, 2, (cudaStream_t)stream);
}

size_t ab200_stage_image_bytes(const ab200_drift_desc* d) { return stage_shape_ok(d) ? stage_tc_image_bytes() : 0; }

int ab200_stage_pack(const ab200_drift_desc* d, const float* w_flat, void* image, size_t image_bytes, ab200_stream_t stream) {
  if (!d || !w_flat || !image) return AB200_ERR_BAD_ARG;
  if (!stage_shape_ok(d)) return AB200_ERR_UNSUPPORTED;
  if (image_bytes < stage_tc_image_bytes()) return AB200_ERR_WORKSPACE;
  cudaError_t e = cudaMemsetAsync((uint8_t*)image + stage_tc_image_bytes() - 256, 0, 256, (cudaStream_t)stream);
  if (e != cudaSuccess) { set_cuda_error(e); return AB200_ERR_CUDA; }
  return stage_tc_pack(w_flat, (uint8_t*)image, (cudaStream_t)stream);
}

int ab200_stage_forward(const ab200_drift_desc* d, const void* image, const float* y0, const float* const* a,
                        const ab200_stage_desc* s, int64_t B, float* a_out, float* y_out, double* err_sumsq,
                        int32_t operand_format, ab200_stream_t stream) {
  if (!d || !image || !y0 || !s || B <= 0 || (s->n_a > 0 && !a)) return AB200_ERR_BAD_ARG;
  if (!stage_shape_ok(d)) return AB200_ERR_UNSUPPORTED;
  if ((err_sumsq && !y_out) || operand_format < 0 || operand_format > 2) return AB200_ERR_BAD_ARG;
  if (operand_format == 2) {
    float* outs[1] = {a_out};
    return stage_fwd2_tc_multi(d, (const uint8_t*)image, y0, a, s, 1, outs, B, y_out, err_sumsq, nullptr, 0, (cudaStream_t)stream);
  }
  return stage_fwd_tc(d, (const uint8_t*)image, y0, a, s, B, a_out, y_out, err_sumsq, operand_format, (cudaStream_t)stream);
}

int ab200_stage_forward_fused(const ab200_drift_desc* d, const void* image, const float* y0, const float* const* a,
                              const ab200_stage_desc* stages, int32_t n_stage, float* const* a_out, int64_t B, float* y_out,
                              double* err_sumsq, int32_t operand_format, ab200_stream_t stream) {
  if (!d || !image || !y0 || !stages || !a || B <= 0 || n_stage < 1 || n_stage > AB200_STAGE_MAX_A) return AB200_ERR_BAD_ARG;
  if (!stage_shape_ok(d)) return AB200_ERR_UNSUPPORTED;
  if ((err_sumsq && !y_out) || operand_format < 0 || operand_format > 2) return AB200_ERR_BAD_ARG;
  if (operand_format == 2)
    return stage_fwd2_tc_multi(d, (const uint8_t*)image, y0, a, stages, n_stage, a_out, B, y_out, err_sumsq, nullptr, 0, (cudaStream_t)stream);
  return stage_fwd_tc_multi(d, (const uint8_t*)image, y0, a, stages, n_stage, a_out, B, y_out, err_sumsq, operand_format,
                            (cudaStream_t)stream);
}

int ab200_stage_forward_fused_save(const ab200_drift_desc* d, const void* image, const float* y0, const float* const* a,
                                   const ab200_stage_desc* stages, int32_t n_stage, float* const* a_out, int64_t B, float* y_out,
                                   double* err_sumsq, void* const* x_outs, int32_t save_level, ab200_stream_t stream) {
  if (!d || !image || !y0 || !stages || !a || !x_outs || B <= 0 || n_stage < 1 || n_stage > AB200_STAGE_MAX_A) return AB200_ERR_BAD_ARG;
  if (!stage_shape_ok(d)) return AB200_ERR_UNSUPPORTED;
  if ((err_sumsq && !y_out) || save_level < 1 || save_level > 2) return AB200_ERR_BAD_ARG;
  for (int i = 0; i < n_stage; ++i)
    if (!x_outs[i]) return AB200_ERR_BAD_ARG;
  return stage_fwd2_tc_multi(d, (const uint8_t*)image, y0, a, stages, n_stage, a_out, B, y_out, err_sumsq, x_outs, save_level,
                             (cudaStream_t)stream);
}

size_t ab200_stage_xblob_bytes(const ab200_drift_desc* d, int64_t B, int32_t save_level) {
  if (!stage_shape_ok(d) || B <= 0 || save_level < 1 || save_level > 2) return 0;
  return wg::FwdSaveLayout{(int)((B + 127) / 128)}.total(save_level);
}

// ---- Dormand-Prince 5(4) tableau (tdq dopri5.py) and the (p0, v0, a_j) form of its linear combinations (second-order drift:
// k_j = (v_in_j, a_j), see stage.py Tableau.combo) ---------------------------------------------------------------------
namespace {
const double DP_C[7] = {0.0, 1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1.0, 1.0};
const double DP_BETA[7][6] = {
    {0, 0, 0, 0, 0, 0},
    {1.0 / 5, 0, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656, 0},
    {35.0 / 384, 0.0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84}};
const double DP_SOL[7] = {35.0 / 384, 0.0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84, 0.0};
const double DP_ERR[7] = {35.0 / 384 - 1951.0 / 21600, 0.0, 500.0 / 1113 - 22642.0 / 50085, 125.0 / 192 - 451.0 / 720,
                          -2187.0 / 6784 + 12231.0 / 42400, 11.0 / 84 - 649.0 / 6300, -1.0 / 60.0};
const double DP_MID[7] = {6025192743.0 / 30085553152.0 / 2, 0.0, 51252292925.0 / 65400821598.0 / 2, -2691868925.0 / 45128329728.0 / 2,
                          187940372067.0 / 1594534317056.0 / 2, -1776094331.0 / 19743644256.0 / 2, 11237099.0 / 235043384.0 / 2};
// y0 + dt sum_j w_j k_j over (p0, v0, a_1..a_n):  p = p0 + cpv v0 + sum cpa[l] a_l,  v = v0 + sum cva[l] a_l
void dp_combo(const double* w, int n, double dt, float* cpv, float* cpa, float* cva) {
  double sw = 0.0;
  for (int j = 0; j < n; ++j) sw += w[j];
  *cpv = (float)(dt * sw);
  for (int l = 0; l < n; ++l) {
    double acc = 0.0;
    for (int j = l + 1; j < n; ++j) acc += w[j] * (dt * DP_BETA[j][l]);     // k_j.p = v_in_j = v0 + dt sum_l beta[j][l] a_l
    cpa[l] = (float)(dt * acc);
    cva[l] = (float)(dt * w[l]);
  }
}
}  // namespace

int ab200_dopri5_attempt(const ab200_drift_desc* d, const void* image, const float* y0, float* const* a, double t0, double dt,
                         int64_t B, float* y_out, double* err_sumsq, float rtol, float atol, int32_t operand_format, void* x_blobs,
                         int32_t save_level, ab200_stream_t stream) {
  if (!d || !image || !y0 || !a || !y_out || B <= 0 || operand_format < 0 || operand_format > 2) return AB200_ERR_BAD_ARG;
  if (!stage_shape_ok(d)) return AB200_ERR_UNSUPPORTED;
  for (int i = 0; i < 7; ++i)
    if (!a[i]) return AB200_ERR_BAD_ARG;
  ab200_stage_desc st[6];
  memset(st, 0, sizeof(st));
  float* outs[6];
  const float* ins[AB200_STAGE_MAX_A];
  for (int i = 0; i < 7; ++i) ins[i] = a[i];
  for (int i = 1; i <= 6; ++i) {
    ab200_stage_desc& s = st[i - 1];
    s.n_a = i;
    dp_combo(DP_BETA[i], i, dt, &s.in_cpv, s.in_cpa, s.in_cva);
    s.t = (float)(i == 6 ? t0 + dt : t0 + DP_C[i] * dt);
    outs[i - 1] = a[i];
  }
  ab200_stage_desc& last = st[5];
  dp_combo(DP_SOL, 7, dt, &last.out_cpv, last.out_cpa, last.out_cva);
  float unused;
  dp_combo(DP_ERR, 7, dt, &unused, last.err_pa, last.err_va);
  last.rtol = rtol;
  last.atol = atol;
  if (x_blobs != nullptr && (operand_format != 2 || save_level < 1 || save_level > 2)) return AB200_ERR_UNSUPPORTED;
  if (operand_format == 2) {
    void* xo[6];
    const size_t per_stage = ab200_stage_xblob_bytes(d, B, save_level);
    for (int i = 0; i < 6; ++i) xo[i] = x_blobs ? (uint8_t*)x_blobs + (size_t)i * per_stage : nullptr;
    return stage_fwd2_tc_multi(d, (const uint8_t*)image, y0, ins, st, 6, outs, B, y_out, err_sumsq, x_blobs ? xo : nullptr,
                               x_blobs ? save_level : 0, (cudaStream_t)stream);
  }
  return stage_fwd_tc_multi(d, (const uint8_t*)image, y0, ins, st, 6, outs, B, y_out, err_sumsq, operand_format, (cudaStream_t)stream);
}

int ab200_dopri5_dense_rows(const ab200_drift_desc* d, const float* y0, const float* const* a, double dt, int32_t n_rows,
                            const double* x_host, int64_t B, float* const* out_rowmajor, ab200_stream_t stream) {
  if (!desc_ok(d) || !y0 || !a || !x_host || !out_rowmajor || B <= 0 || n_rows < 1 || n_rows > 256) return AB200_ERR_BAD_ARG;
  float cpv[256], cpa[256 * 8], cva[256 * 8];
  for (int q = 0; q < n_rows; ++q) {
    const double x = x_host[q];
    double w[7];
    for (int j = 0; j < 7; ++j) {       // interp.py: y(t0 + x dt) = y0 + dt sum_j W_j(x) k_j
      const double e1 = (j == 0), e7 = (j == 6), cs = DP_SOL[j], cm = DP_MID[j];
      const double c2 = e7 - 4 * e1 - 5 * cs + 16 * cm;
      const double c3 = 5 * e1 - 3 * e7 + 14 * cs - 32 * cm;
      const double c4 = 2 * (e7 - e1) - 8 * cs + 16 * cm;
      w[j] = x * e1 + x * x * c2 + x * x * x * c3 + x * x * x * x * c4;
    }
    dp_combo(w, 7, dt, &cpv[q], &cpa[q * 8], &cva[q * 8]);
    cpa[q * 8 + 7] = cva[q * 8 + 7] = 0.f;
  }
  return pv_combine_rowmajor_multi(d, y0, a, 7, n_rows, cpv, cpa, cva, B, out_rowmajor, (cudaStream_t)stream);
}

size_t ab200_stage_spill_bytes(const ab200_drift_desc* d, int32_t nblobs) {
  return (stage_shape_ok(d) && nblobs > 0) ? wgrad_spill_bytes(nblobs) : 0;
}
size_t ab200_wgrad_partial_bytes(const ab200_drift_desc* d) { return stage_shape_ok(d) ? wgrad_partial_bytes() : 0; }

int ab200_stage_backward(const ab200_drift_desc* d, const void* image, const float* y0, const float* const* a,
                         const ab200_stage_desc* s, int64_t B, const float* g_base, const float* const* gx, int32_t n_g,
                         const float* dp_host, const float* dv_host, float* gx_out, void* spill, size_t spill_bytes, int32_t blob0,
                         int32_t nblobs, void* partial, ab200_stream_t stream) {
  if (!d || !image || !y0 || !s || !gx_out || !spill || !partial || B <= 0 || (s->n_a > 0 && !a) ||
      (n_g > 0 && (!gx || !dp_host || !dv_host)) || (n_g == 0 && !g_base))
    return AB200_ERR_BAD_ARG;
  if (!stage_shape_ok(d)) return AB200_ERR_UNSUPPORTED;
  if (nblobs <= 0 || spill_bytes < wgrad_spill_bytes(nblobs)) return AB200_ERR_WORKSPACE;
  return stage_bwd_tc(d, (const uint8_t*)image, y0, a, s, B, g_base, gx, n_g, dp_host, dv_host, gx_out, spill, blob0, nblobs,
                      wgrad_bout_ptr(partial), (cudaStream_t)stream);
}

int ab200_stage_backward_fused(const ab200_drift_desc* d, const void* image, const float* y0, const float* const* a,
                               const ab200_stage_desc* stages, int32_t n_stage, const float* const* g_base, float* const* gx_out,
                               const int32_t* n_g, const int32_t* gx_src, const float* const* gx_ext, const float* dp_host,
                               const float* dv_host, int64_t B, void* spill, size_t spill_bytes, int32_t blob0, int32_t nblobs,
                               void* partial, const void* const* x_blobs, int32_t save_level, float* y0_accum,
                               float* const* upstream_out, ab200_stream_t stream) {
  if (!d || !image || !y0 || !a || !stages || !g_base || !gx_out || !n_g || !gx_src || !dp_host || !dv_host || !spill || !partial ||
      B <= 0 || n_stage < 1 || n_stage > AB200_STAGE_MAX_A + 1)
    return AB200_ERR_BAD_ARG;
  if (!stage_shape_ok(d)) return AB200_ERR_UNSUPPORTED;
  if (nblobs <= 0 || spill_bytes < wgrad_spill_bytes(nblobs)) return AB200_ERR_WORKSPACE;
  return stage_bwd_tc_multi(d, (const uint8_t*)image, y0, a, stages, n_stage, g_base, gx_out, n_g, gx_src, gx_ext, dp_host, dv_host, B,
                            spill, blob0, nblobs, wgrad_bout_ptr(partial), x_blobs, x_blobs ? save_level : 0, y0_accum, upstream_out,
                            (cudaStream_t)stream);
}

int ab200_pv_combine_backward_multi(const ab200_drift_desc* d, const float* const* g, int32_t n_src, const float* cpv_host,
                                    const float* cpa_host, const float* cva_host, int32_t n_a, int64_t B, float* G_y0,
                                    float* const* G_a, int32_t accumulate, int32_t rowmajor_mask, const float* add_a, int32_t add_index,
                                    ab200_stream_t stream) {
  if (!desc_ok(d) || !g || !cpv_host || !G_y0 || B <= 0 || (n_a > 0 && (!G_a || !cpa_host || !cva_host))) return AB200_ERR_BAD_ARG;
  return pv_combine_bwd_multi(d, g, n_src, cpv_host, cpa_host, cva_host, n_a, B, G_y0, G_a, accumulate, rowmajor_mask, add_a, add_index, (cudaStream_t)stream);
}

int ab200_stage_upstream(const ab200_drift_desc* d, const float* g_base, const float* const* gx, int32_t n_g, const float* dp_host,
                         const float* dv_host, int64_t B, float* g_a_out, ab200_stream_t stream) {
  if (!desc_ok(d) || !g_base || !g_a_out || B <= 0 || (n_g > 0 && (!gx || !dp_host || !dv_host))) return AB200_ERR_BAD_ARG;
  return ga_assemble(d, g_base, gx, n_g, dp_host, dv_host, B, g_a_out, (cudaStream_t)stream);
}

int ab200_adjoint_gather(const ab200_drift_desc* d, const float* base, const float* const* gx, int32_t n, const float* cpv_host,
                         int64_t B, float* out, ab200_stream_t stream) {
  if (!desc_ok(d) || !base || !out || B <= 0 || (n > 0 && (!gx || !cpv_host))) return AB200_ERR_BAD_ARG;
  return adjoint_gather(d, base, gx, n, cpv_host, B, out, nullptr, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

int ab200_adjoint_gather_upstream(const ab200_drift_desc* d, const float* base, const float* const* gx, int32_t n, const float* cpv_host,
                                  int64_t B, float* out, const float* g_base, const float* dp_host, const float* dv_host,
                                  float* g_a_out, ab200_stream_t stream) {
  if (!desc_ok(d) || !base || !out || !g_base || !g_a_out || !dp_host || !dv_host || B <= 0 || n <= 0 || !gx || !cpv_host)
    return AB200_ERR_BAD_ARG;
  return adjoint_gather(d, base, gx, n, cpv_host, B, out, g_base, dp_host, dv_host, g_a_out, (cudaStream_t)stream);
}

int ab200_wgrad_accumulate(const ab200_drift_desc* d, const void* spill, int32_t nblobs, int32_t used, void* partial,
                           const void* const* x_blobs, int32_t n_x_blobs, int32_t ntiles, int32_t save_level, ab200_stream_t stream) {
  if (!d || !spill || !partial || nblobs <= 0 || used < 0 || used > nblobs || n_x_blobs < 0 || (n_x_blobs > 0 && (!x_blobs || ntiles <= 0)))
    return AB200_ERR_BAD_ARG;
  if (!stage_shape_ok(d)) return AB200_ERR_UNSUPPORTED;
  return wgrad_tc(spill, nblobs, used, partial, x_blobs, n_x_blobs, ntiles, save_level, (cudaStream_t)stream);
}

int ab200_wgrad_finalize(const ab200_drift_desc* d, const void* partial, float* grad_w_flat, ab200_stream_t stream) {
  if (!d || !partial || !grad_w_flat) return AB200_ERR_BAD_ARG;
  if (!stage_shape_ok(d)) return AB200_ERR_UNSUPPORTED;
  return wgrad_finalize(partial, grad_w_flat, (cudaStream_t)stream);
}

int ab200_stage_status_offset(const ab200_drift_desc* d, int64_t* image_status_byte, int64_t* partial_status_byte) {
  if (!stage_shape_ok(d)) return AB200_ERR_UNSUPPORTED;
  if (image_status_byte) *image_status_byte = (int64_t)stage_tc_image_bytes() - 256;
  if (partial_status_byte) *partial_status_byte = (int64_t)wgrad_partial_bytes() - 256;
  return AB200_OK;
}

int ab200_pv_combine(const ab200_drift_desc* d, const float* y0, const float* const* a, int32_t n_a, float cpv,
                     const float* cpa_host, const float* cva_host, int64_t B, float* out, ab200_stream_t stream) {
  if (!desc_ok(d) || !y0 || !out || B <= 0 || (n_a > 0 && (!a || !cpa_host || !cva_host))) return AB200_ERR_BAD_ARG;
  return pv_combine(d, y0, a, n_a, cpv, cpa_host, cva_host, B, out, (cudaStream_t)stream);
}

int ab200_pv_combine_rowmajor(const ab200_drift_desc* d, const float* y0, const float* const* a, int32_t n_a, float cpv,
                              const float* cpa_host, const float* cva_host, int64_t B, float* out_rowmajor, ab200_stream_t stream) {
  if (!desc_ok(d) || !y0 || !out_rowmajor || B <= 0 || (n_a > 0 && (!a || !cpa_host || !cva_host))) return AB200_ERR_BAD_ARG;
  return pv_combine_rowmajor(d, y0, a, n_a, cpv, cpa_host, cva_host, B, out_rowmajor, (cudaStream_t)stream);
}

int ab200_pv_combine_rowmajor_multi(const ab200_drift_desc* d, const float* y0, const float* const* a, int32_t n_a, int32_t n_rows,
                                    const float* cpv_host, const float* cpa_host, const float* cva_host, int64_t B,
                                    float* const* out_rowmajor, ab200_stream_t stream) {
  if (!desc_ok(d) || !y0 || !out_rowmajor || !cpv_host || B <= 0 || n_rows < 1 || (n_a > 0 && (!a || !cpa_host || !cva_host)))
    return AB200_ERR_BAD_ARG;
  for (int i = 0; i < n_rows; ++i)
    if (!out_rowmajor[i]) return AB200_ERR_BAD_ARG;
  return pv_combine_rowmajor_multi(d, y0, a, n_a, n_rows, cpv_host, cpa_host, cva_host, B, out_rowmajor, (cudaStream_t)stream);
}

int ab200_pv_combine_backward(const ab200_drift_desc* d, const float* g, int32_t n_a, float cpv, const float* cpa_host,
                              const float* cva_host, int64_t B, float* G_y0, float* const* G_a, int32_t accumulate,
                              ab200_stream_t stream) {
  if (!desc_ok(d) || !g || !G_y0 || B <= 0 || (n_a > 0 && (!G_a || !cpa_host || !cva_host))) return AB200_ERR_BAD_ARG;
  return pv_combine_bwd(d, g, n_a, cpv, cpa_host, cva_host, B, G_y0, G_a, accumulate, (cudaStream_t)stream);
}

int ab200_sde_euler_step(const float* y, const float* drift, const float* diffusion, int32_t diffusion_per_row, int64_t B, int32_t D,
                         float dt, uint64_t seed, uint64_t step, float* y_out, float* xi_out, void* stream) {
  if (!y || !drift || !diffusion || !y_out || B <= 0 || D <= 0 || !(dt > 0.0f)) return AB200_ERR_BAD_ARG;
  return sde_euler_step(y, drift, diffusion, diffusion_per_row, B, D, dt, seed, step, y_out, xi_out, (cudaStream_t)stream);
}

int ab200_grad_sumsq(const float* flat_grad, int64_t n, double* sumsq_out, void* stream) {
  if (!flat_grad || !sumsq_out || n <= 0) return AB200_ERR_BAD_ARG;
  return grad_sumsq(flat_grad, n, sumsq_out, (cudaStream_t)stream);
}

int ab200_adam_step(float* flat_param, const float* flat_grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                    float beta2, float eps, float weight_decay, int32_t step, float max_grad_norm, const double* grad_sumsq,
                    void* stream) {
  if (!flat_param || !flat_grad || !exp_avg || !exp_avg_sq || n <= 0 || step < 1) return AB200_ERR_BAD_ARG;
  return adam_step(flat_param, flat_grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step, max_grad_norm, grad_sumsq,
                   (cudaStream_t)stream);
}

size_t ab200_head_workspace_bytes(int32_t Z, int32_t E) { return (Z > 0 && E == 64) ? head_workspace_bytes(Z) : 0; }

int ab200_head_argmax(const float* pred_emb, const float* class_table, int64_t M, int32_t Z, int32_t E, float tau, int64_t* labels,
                      float* best_logit, void* workspace, size_t workspace_bytes, ab200_stream_t stream) {
  if (!pred_emb || !class_table || !labels || !workspace || M <= 0 || Z <= 0 || !(tau > 0.0f)) return AB200_ERR_BAD_ARG;
  return head_argmax(pred_emb, class_table, M, Z, E, tau, labels, best_logit, workspace, workspace_bytes, (cudaStream_t)stream);
}

int ab200_head_ce_forward(const float* pred_emb, const float* class_table, const int64_t* target, int64_t M, int32_t Z, int32_t E,
                          float tau, float* lse, float* target_logit, int64_t* labels, const float* dist_mat, float* expected_dist,
                          void* workspace, size_t workspace_bytes, void* stream) {
  if (!pred_emb || !class_table || !target || !lse || !target_logit || !workspace || M <= 0 || Z <= 0 || !(tau > 0.0f) ||
      ((dist_mat != nullptr) != (expected_dist != nullptr)))
    return AB200_ERR_BAD_ARG;
  return head_ce_forward(pred_emb, class_table, target, M, Z, E, tau, lse, target_logit, labels, dist_mat, expected_dist, workspace,
                         workspace_bytes, (cudaStream_t)stream);
}

size_t ab200_head_ce_backward_workspace_bytes(int64_t M, int32_t Z, int32_t E) {
  return (M > 0 && Z > 0 && E == 64) ? head_ce_backward_workspace_bytes(M, Z) : 0;
}

int ab200_head_ce_backward(const float* pred_emb, const float* class_table, const int64_t* target, const float* lse,
                           const float* grad_rows, const float* grad_dist_rows, const float* expected_dist, const float* dist_mat,
                           int64_t M, int32_t Z, int32_t E, float tau, float* grad_emb_normalised, float* grad_table_normalised,
                           void* workspace, size_t workspace_bytes, void* stream) {
  if (!pred_emb || !class_table || !target || !lse || !grad_rows || !grad_emb_normalised || !grad_table_normalised || !workspace ||
      M <= 0 || Z <= 0 || !(tau > 0.0f))
    return AB200_ERR_BAD_ARG;
  const int nd = (grad_dist_rows != nullptr) + (expected_dist != nullptr) + (dist_mat != nullptr);
  if (nd != 0 && nd != 3) return AB200_ERR_BAD_ARG;
  return head_ce_backward(pred_emb, class_table, target, lse, grad_rows, grad_dist_rows, expected_dist, dist_mat, M, Z, E, tau,
                          grad_emb_normalised, grad_table_normalised, workspace, workspace_bytes, (cudaStream_t)stream);
}

int ab200_head_ce_backward_status(const void* workspace, int64_t M, int32_t Z, int32_t* status_host, void* stream) {
  if (!workspace || !status_host || M <= 0 || Z <= 0) return AB200_ERR_BAD_ARG;
  return head_ce_backward_status(workspace, M, Z, status_host, (cudaStream_t)stream);
}


int ab200_gat_forward(const int32_t* rowptr, const int32_t* col, int32_t Z, int32_t nnz, const float* x, int32_t F_in,
                      const float* W, const float* att_src, const float* att_dst, const float* bias, int32_t heads, int32_t F_out,
                      int32_t concat, float negative_slope, float* out, float* xw, float* a_src, float* a_dst, float* alpha,
                      ab200_stream_t stream) {
  if (!rowptr || !col || !x || !W || !att_src || !att_dst || !out || !xw || !a_src || !a_dst || !alpha || Z <= 0 || nnz < 0)
    return AB200_ERR_BAD_ARG;
  return gat_forward(rowptr, col, Z, nnz, x, F_in, W, att_src, att_dst, bias, heads, F_out, concat, negative_slope, out, xw, a_src,
                     a_dst, alpha, (cudaStream_t)stream);
}

size_t ab200_gat_backward_workspace_bytes(int32_t Z, int32_t nnz, int32_t heads, int32_t F_out) {
  if (Z <= 0 || nnz < 0 || heads <= 0 || F_out <= 0) return 0;
  return gat_backward_workspace(Z, nnz, heads, F_out);
}

int ab200_gat_backward(const int32_t* rowptr, const int32_t* col, const int32_t* rowptr_t, const int32_t* col_t, const int32_t* eid_t,
                       int32_t Z, int32_t nnz, const float* x, int32_t F_in, const float* W, const float* att_src,
                       const float* att_dst, int32_t heads, int32_t F_out, int32_t concat, float negative_slope, const float* xw,
                       const float* a_src, const float* a_dst, const float* alpha, const float* grad_out, float* grad_x,
                       float* grad_W, float* grad_att_src, float* grad_att_dst, float* grad_bias, void* workspace,
                       size_t workspace_bytes, ab200_stream_t stream) {
  if (!rowptr || !col || !rowptr_t || !col_t || !eid_t || !x || !W || !att_src || !att_dst || !xw || !a_src || !a_dst || !alpha ||
      !grad_out || !grad_W || !grad_att_src || !grad_att_dst || !workspace || Z <= 0)
    return AB200_ERR_BAD_ARG;
  return gat_backward(rowptr, col, rowptr_t, col_t, eid_t, Z, nnz, x, F_in, W, att_src, att_dst, heads, F_out, concat, negative_slope,
                      xw, a_src, a_dst, alpha, grad_out, grad_x, grad_W, grad_att_src, grad_att_dst, grad_bias, workspace,
                      workspace_bytes, (cudaStream_t)stream);
}

int ab200_debug_umma_probe(const float* A, const float* B, float* D, int32_t N, int32_t K, int32_t a_mode, int32_t b_mode,
                           int32_t* status, ab200_stream_t stream) {
  if (!A || !B || !D || !status) return AB200_ERR_BAD_ARG;
  return umma_probe(A, B, D, N, K, a_mode, b_mode, status, (cudaStream_t)stream);
}

}  // extern "C"
