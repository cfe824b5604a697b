// Embedding-space terms of the mode_sep training loss in ONE pass over (pred_emb, v) [B, T, E]:
//   mse_at_snaps            mode_sep/architecture/losses.py:24-31   at the ground-truth snaps, and again (weight w_stay_aux) at
//                           the non-snap points inside stays: mode_sep/train/train.py:126-135
//   travel_margin_loss      losses.py:56-73      hinge  m_travel - (d_prev - d_dest)  on travel points
//   travel_monotonicity     losses.py:76-115     hinges on consecutive travel points of the same segment
//   velocity regularisers   mode_sep/train/train.py:137-153   |v|^2 inside stays, (v_min - |v|)+^2 + (|v| - v_max)+^2 at interior snaps
// The reference evaluates each with boolean-mask gathers over [B, T, E] tensors (a dozen full passes and as many temporaries);
// here a row (agent, time) is read once: 16 lanes own its 64 embedding dims, class vectors are gathered with 128-bit loads, the
// six masked sums and their counts are accumulated per block and added to 12 doubles.  The backward kernel recomputes the same
// row quantities and writes d pred_emb and d v (every row, no atomics: the monotonicity pairs are gathered from both neighbours)
// and scatters d class_table with atomics.  HBM-bound: 2 x 256 B read per row forward, 4 x 256 B moved backward.
#include "common.cuh"

namespace ab200 {

constexpr int EL_E = 64, EL_LANES = EL_E / 4, EL_ROWS_PER_WARP = 32 / EL_LANES;      // 16 lanes per row, 2 rows per warp

struct EmbLossArgs {
  const float* emb;  int64_t emb_sb, emb_st;      // pred_emb [B, T, E]: element (b, t, e) at emb[b * sb + t * st + e]
  const float* v;    int64_t v_sb, v_st;          // v_t      [B, T, E]
  const float* table;                             // class_table [Z, E]
  const int64_t* y_gt;   const uint8_t* m_gt;     // y_union / is_gt_union
  const int64_t* y_stay; const uint8_t* m_stay;   // stay_loc_ids / stay_non_gt_mask
  const uint8_t* m_travel; const int64_t* prev; const int64_t* dest;
  const uint8_t* m_move;                          // gt_interior_mask
  float m_margin, eps_mono, v_min, v_max;
  int64_t B; int T; int Z;
  double* sums;                                   // forward: 12 doubles (see SUM_*)
  // backward
  const float* coef;                              // device [6]: d total / d sum_i (already divided by the counts)
  float* d_emb; float* d_v; float* d_table;       // [B, T, E] contiguous, [B, T, E] contiguous, [Z, E] (zeroed by the caller)
};
enum { SUM_MSE_GT = 0, N_GT, SUM_MSE_STAY, N_STAY, SUM_MARGIN, N_TRAVEL, SUM_AWAY, SUM_TOWARD, N_PAIR, SUM_STAYVEL, SUM_MOVEVEL, N_MOVE, N_SUMS };

// sum over the 16 lanes of a row.  The two rows of a warp take their own branches: only the row's lanes take part (`mask`),
// and all of them do, because every branch condition is a property of the row.
__device__ __forceinline__ float row_sum(float x, uint32_t mask) {
#pragma unroll
  for (int o = EL_LANES / 2; o > 0; o >>= 1) x += __shfl_xor_sync(mask, x, o);
  return x;
}
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 sub4(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ void axpy4(float4& acc, float s, float4 x) { acc.x += s * x.x; acc.y += s * x.y; acc.z += s * x.z; acc.w += s * x.w; }
__device__ __forceinline__ int64_t clamp_idx(int64_t i, int Z) { return i < 0 ? 0 : (i >= Z ? Z - 1 : i); }

// pair (t, t + 1) of the same travel segment (losses.py:95-99)
__device__ __forceinline__ bool pair_ok(const EmbLossArgs& a, int64_t r, int t) {
  return t + 1 < a.T && a.m_travel[r] && a.m_travel[r + 1] && a.prev[r] == a.prev[r + 1] && a.dest[r] == a.dest[r + 1];
}

template <bool BWD>
__global__ void __launch_bounds__(256, 4) emb_losses_kernel(const __grid_constant__ EmbLossArgs a) {
  const int lane = threadIdx.x & 31, sub = lane % EL_LANES, which = lane / EL_LANES;
  const uint32_t rmask = 0xFFFFu << (which * EL_LANES);
  const int64_t n_rows = a.B * (int64_t)a.T;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  // per-thread partial sums in fp32 (a thread sees a few hundred rows at most; the cross-thread reduction is in double): the kernel
  // is latency-bound on its gathers, so registers are occupancy (64 per thread = 32 warps per SM)
  float s_loc[N_SUMS];
#pragma unroll
  for (int i = 0; i < N_SUMS; ++i) s_loc[i] = 0.0f;
  float c_mse_gt = 0.f, c_mse_stay = 0.f, c_margin = 0.f, c_mono = 0.f, c_stayvel = 0.f, c_movevel = 0.f;
  if (BWD) {
    c_mse_gt = a.coef[0]; c_mse_stay = a.coef[1]; c_margin = a.coef[2]; c_mono = a.coef[3]; c_stayvel = a.coef[4]; c_movevel = a.coef[5];
  }

  for (int64_t rw = warp0; rw * EL_ROWS_PER_WARP < n_rows; rw += n_warps) {
    const int64_t r = rw * EL_ROWS_PER_WARP + which;
    const bool live = r < n_rows;                       // all lanes of a row agree; shuffles stay inside the row's 16 lanes
    const int64_t b = live ? r / a.T : 0;
    const int t = live ? (int)(r % a.T) : 0;
    const float* ep = a.emb + b * a.emb_sb + (int64_t)t * a.emb_st + 4 * sub;
    const float* vp = a.v + b * a.v_sb + (int64_t)t * a.v_st + 4 * sub;
    const float4 e = live ? ld4(ep) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 ge = make_float4(0.f, 0.f, 0.f, 0.f), gv = ge;
    const bool gt = live && a.m_gt[r], stay = live && a.m_stay[r], trav = live && a.m_travel[r], move = live && a.m_move[r];

    // ---- squared distance to the target class at snaps / inside stays
    if (gt || stay) {
      const int64_t y = clamp_idx(gt ? a.y_gt[r] : a.y_stay[r], a.Z);
      const float4 d = sub4(e, ld4(a.table + y * EL_E + 4 * sub));
      const float d2 = row_sum(dot4(d, d), rmask);
      if (gt) { s_loc[SUM_MSE_GT] += d2; s_loc[N_GT] += 1.0f; }
      else { s_loc[SUM_MSE_STAY] += d2; s_loc[N_STAY] += 1.0f; }
      if (BWD) {
        const float c = 2.0f * (gt ? c_mse_gt : c_mse_stay);
        axpy4(ge, c, d);
        float* q = a.d_table + y * EL_E + 4 * sub;
        atomicAdd(q, -c * d.x); atomicAdd(q + 1, -c * d.y); atomicAdd(q + 2, -c * d.z); atomicAdd(q + 3, -c * d.w);
      }
      if (gt && stay) {      // cannot happen with the reference's masks (stay_non_gt excludes snaps); kept exact anyway
        const int64_t y2 = clamp_idx(a.y_stay[r], a.Z);
        const float4 d_ = sub4(e, ld4(a.table + y2 * EL_E + 4 * sub));
        s_loc[SUM_MSE_STAY] += row_sum(dot4(d_, d_), rmask); s_loc[N_STAY] += 1.0f;
        if (BWD) {
          axpy4(ge, 2.0f * c_mse_stay, d_);
          float* q = a.d_table + y2 * EL_E + 4 * sub;
          atomicAdd(q, -2.0f * c_mse_stay * d_.x); atomicAdd(q + 1, -2.0f * c_mse_stay * d_.y);
          atomicAdd(q + 2, -2.0f * c_mse_stay * d_.z); atomicAdd(q + 3, -2.0f * c_mse_stay * d_.w);
        }
      }
    }

    // ---- travel terms: distances to the segment's origin and destination classes
    if (trav) {
      const int64_t ip = clamp_idx(a.prev[r], a.Z), id = clamp_idx(a.dest[r], a.Z);
      const float4 tp = ld4(a.table + ip * EL_E + 4 * sub), td = ld4(a.table + id * EL_E + 4 * sub);
      const float4 dp4 = sub4(e, tp), dd4 = sub4(e, td);
      const float dp = sqrtf(row_sum(dot4(dp4, dp4), rmask)), dd = sqrtf(row_sum(dot4(dd4, dd4), rmask));
      const float hinge = a.m_margin - (dp - dd);
      s_loc[N_TRAVEL] += 1.0f;
      if (hinge > 0.f) s_loc[SUM_MARGIN] += hinge;
      float g_dp = 0.f, g_dd = 0.f;                       // d total / d d_prev(t), d d_dest(t)
      if (BWD && hinge > 0.f) { g_dp -= c_margin; g_dd += c_margin; }
      // pair (t, t+1): this row is the EARLIER point
      if (pair_ok(a, r, t)) {
        const float4 en = ld4(ep + a.emb_st);
        const float4 a4 = sub4(en, tp), b4 = sub4(en, td);
        const float dpn = sqrtf(row_sum(dot4(a4, a4), rmask)), ddn = sqrtf(row_sum(dot4(b4, b4), rmask));
        const float away = dp - dpn + a.eps_mono, toward = ddn - dd + a.eps_mono;
        s_loc[N_PAIR] += 1.0f;
        if (away > 0.f) { s_loc[SUM_AWAY] += away; if (BWD) g_dp += c_mono; }
        if (toward > 0.f) { s_loc[SUM_TOWARD] += toward; if (BWD) g_dd -= c_mono; }
      }
      // pair (t-1, t): this row is the LATER point (backward only: its hinges pull on d(t) as well)
      if (BWD && t > 0 && pair_ok(a, r - 1, t - 1)) {
        const float4 em = ld4(ep - a.emb_st);
        const float4 a4 = sub4(em, tp), b4 = sub4(em, td);
        const float dpm = sqrtf(row_sum(dot4(a4, a4), rmask)), ddm = sqrtf(row_sum(dot4(b4, b4), rmask));
        if (dpm - dp + a.eps_mono > 0.f) g_dp -= c_mono;
        if (dd - ddm + a.eps_mono > 0.f) g_dd += c_mono;
      }
      if (BWD) {
        const float sp = dp > 0.f ? g_dp / dp : 0.f, sd = dd > 0.f ? g_dd / dd : 0.f;      // d |x| / d x = x / |x|
        axpy4(ge, sp, dp4);
        axpy4(ge, sd, dd4);
        if (sp != 0.f) {
          float* q = a.d_table + ip * EL_E + 4 * sub;
          atomicAdd(q, -sp * dp4.x); atomicAdd(q + 1, -sp * dp4.y); atomicAdd(q + 2, -sp * dp4.z); atomicAdd(q + 3, -sp * dp4.w);
        }
        if (sd != 0.f) {
          float* q = a.d_table + id * EL_E + 4 * sub;
          atomicAdd(q, -sd * dd4.x); atomicAdd(q + 1, -sd * dd4.y); atomicAdd(q + 2, -sd * dd4.z); atomicAdd(q + 3, -sd * dd4.w);
        }
      }
    }

    // ---- velocity regularisers
    if (stay || move) {
      const float4 vv = ld4(vp);
      const float v2 = row_sum(dot4(vv, vv), rmask);
      if (stay) {
        s_loc[SUM_STAYVEL] += v2;
        if (BWD) axpy4(gv, 2.0f * c_stayvel, vv);
      }
      if (move) {
        const float vm = sqrtf(v2);
        const float lo = fmaxf(a.v_min - vm, 0.f), hi = fmaxf(vm - a.v_max, 0.f);
        s_loc[SUM_MOVEVEL] += lo * lo + hi * hi;
        s_loc[N_MOVE] += 1.0f;
        if (BWD && vm > 0.f) axpy4(gv, c_movevel * 2.0f * (hi - lo) / vm, vv);
      }
    }
    if (BWD && live) {
      *reinterpret_cast<float4*>(a.d_emb + r * EL_E + 4 * sub) = ge;
      *reinterpret_cast<float4*>(a.d_v + r * EL_E + 4 * sub) = gv;
    }
  }

  if (!BWD) {
    // every lane of a row carries the row's value: count each row once (lane sub == 0), then warp -> block -> 12 atomics
    __shared__ double red[8][N_SUMS];
#pragma unroll
    for (int i = 0; i < N_SUMS; ++i) {
      double x = sub == 0 ? (double)s_loc[i] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      if (lane == 0) red[threadIdx.x >> 5][i] = x;
    }
    __syncthreads();
    if (threadIdx.x < N_SUMS) {
      double x = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) x += red[w][threadIdx.x];
      if (x != 0.0) atomicAdd(a.sums + threadIdx.x, x);
    }
  }
}

static int emb_grid(int64_t n_rows) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t blocks = (n_rows / EL_ROWS_PER_WARP + 7) / 8;
  const int64_t cap = (int64_t)sms * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

int emb_losses(const EmbLossArgs& a, bool backward, cudaStream_t st) {
  const int grid = emb_grid(a.B * (int64_t)a.T);
  if (backward) emb_losses_kernel<true><<<grid, 256, 0, st>>>(a);
  else emb_losses_kernel<false><<<grid, 256, 0, st>>>(a);
  return check_launch();
}

}  // namespace ab200

using namespace ab200;

extern "C" {

int ab200_emb_losses_forward(const float* pred_emb, int64_t emb_stride_b, int64_t emb_stride_t, const float* v_t, int64_t v_stride_b,
                             int64_t v_stride_t, const float* class_table, const int64_t* y_union, const uint8_t* is_gt,
                             const int64_t* y_stay, const uint8_t* stay_non_gt, const uint8_t* travel_mask, const int64_t* prev_idx,
                             const int64_t* dest_idx, const uint8_t* gt_interior, int64_t B, int32_t T, int32_t E, int32_t Z,
                             float m_travel, float epsilon_mono, float v_min_move, float v_max_move, double* sums12, void* stream) {
  if (!pred_emb || !v_t || !class_table || !y_union || !is_gt || !y_stay || !stay_non_gt || !travel_mask || !prev_idx || !dest_idx ||
      !gt_interior || !sums12 || B <= 0 || T <= 0 || Z <= 0)
    return AB200_ERR_BAD_ARG;
  if (E != EL_E || emb_stride_t % 4 || emb_stride_b % 4 || v_stride_t % 4 || v_stride_b % 4) return AB200_ERR_UNSUPPORTED;
  EmbLossArgs a{};
  a.emb = pred_emb; a.emb_sb = emb_stride_b; a.emb_st = emb_stride_t;
  a.v = v_t; a.v_sb = v_stride_b; a.v_st = v_stride_t;
  a.table = class_table; a.y_gt = y_union; a.m_gt = is_gt; a.y_stay = y_stay; a.m_stay = stay_non_gt;
  a.m_travel = travel_mask; a.prev = prev_idx; a.dest = dest_idx; a.m_move = gt_interior;
  a.m_margin = m_travel; a.eps_mono = epsilon_mono; a.v_min = v_min_move; a.v_max = v_max_move;
  a.B = B; a.T = T; a.Z = Z; a.sums = sums12;
  return emb_losses(a, false, (cudaStream_t)stream);
}

int ab200_emb_losses_backward(const float* pred_emb, int64_t emb_stride_b, int64_t emb_stride_t, const float* v_t, int64_t v_stride_b,
                              int64_t v_stride_t, const float* class_table, const int64_t* y_union, const uint8_t* is_gt,
                              const int64_t* y_stay, const uint8_t* stay_non_gt, const uint8_t* travel_mask, const int64_t* prev_idx,
                              const int64_t* dest_idx, const uint8_t* gt_interior, int64_t B, int32_t T, int32_t E, int32_t Z,
                              float m_travel, float epsilon_mono, float v_min_move, float v_max_move, const float* coef6,
                              float* grad_pred_emb, float* grad_v, float* grad_class_table, void* stream) {
  if (!pred_emb || !v_t || !class_table || !y_union || !is_gt || !y_stay || !stay_non_gt || !travel_mask || !prev_idx || !dest_idx ||
      !gt_interior || !coef6 || !grad_pred_emb || !grad_v || !grad_class_table || B <= 0 || T <= 0 || Z <= 0)
    return AB200_ERR_BAD_ARG;
  if (E != EL_E || emb_stride_t % 4 || emb_stride_b % 4 || v_stride_t % 4 || v_stride_b % 4) return AB200_ERR_UNSUPPORTED;
  EmbLossArgs a{};
  a.emb = pred_emb; a.emb_sb = emb_stride_b; a.emb_st = emb_stride_t;
  a.v = v_t; a.v_sb = v_stride_b; a.v_st = v_stride_t;
  a.table = class_table; a.y_gt = y_union; a.m_gt = is_gt; a.y_stay = y_stay; a.m_stay = stay_non_gt;
  a.m_travel = travel_mask; a.prev = prev_idx; a.dest = dest_idx; a.m_move = gt_interior;
  a.m_margin = m_travel; a.eps_mono = epsilon_mono; a.v_min = v_min_move; a.v_max = v_max_move;
  a.B = B; a.T = T; a.Z = Z; a.coef = coef6; a.d_emb = grad_pred_emb; a.d_v = grad_v; a.d_table = grad_class_table;
  return emb_losses(a, true, (cudaStream_t)stream);
}

}  // extern "C"
