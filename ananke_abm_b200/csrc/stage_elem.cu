// Elementwise companions of the stage kernels, all on the tile-blocked fp32 layout of stage_tc.cuh: linear
// combinations of the step's base state y0 = [p0, v0, h] and stage accelerations a_j (dense-output rows, step
// solutions), their adjoints, and the converters between blocked buffers and public row-major tensors.
// One pass over the state each, 128-bit accesses, HBM-bound.
//   forward :  out.p = p0 + cpv v0 + sum cpa[j] a_j ;  out.v = v0 + sum cva[j] a_j ;  out.h = h
//   adjoint :  G_y0.p (+)= g.p ; G_y0.v (+)= cpv g.p + g.v ; G_y0.h (+)= g.h ; G_a[j] (+)= cpa[j] g.p + cva[j] g.v
// (tdq: rk_common.py `_runge_kutta_step` y1 = y0 + dt sum c_sol k; interp.py `_interp_evaluate`.)
#include "common.cuh"

namespace ab200 {

constexpr int EL_MAX_A = 8;
constexpr int EL_TM = 128;
constexpr int TR_ROWS = 32;       // rows per CTA of the transposing kernels
struct ElemArgs {
  const float* y0;
  const float* a[EL_MAX_A];
  float* G_a[EL_MAX_A];
  float* out;        // forward: out (blocked [Bp][D]);  adjoint: G_y0 (blocked [Bp][D])
  const float* g;    // adjoint: upstream gradient (blocked [Bp][D])
  float cpv, cpa[EL_MAX_A], cva[EL_MAX_A];
  int n_a, accumulate;
  int ntiles;
  int P, H;
};

// work item i -> (tile, group, row): group < P/4 handles the float4 of p AND v with that index, group >= P/4 the h part
__device__ __forceinline__ void decode(int64_t i, int P4, int H4, int& tile, int& grp, int& row) {
  row = (int)(i % EL_TM);
  const int64_t q = i / EL_TM;
  grp = (int)(q % (P4 + H4));
  tile = (int)(q / (P4 + H4));
}

__global__ void __launch_bounds__(256) pv_combine_kernel(const __grid_constant__ ElemArgs a) {
  const int P4 = a.P / 4, H4 = a.H / 4, Y4 = 2 * P4 + H4;
  const int64_t n = (int64_t)a.ntiles * (P4 + H4) * EL_TM;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int tile, grp, row;
    decode(i, P4, H4, tile, grp, row);
    const float4* y4 = reinterpret_cast<const float4*>(a.y0) + (size_t)tile * Y4 * EL_TM + row;
    float4* o4 = reinterpret_cast<float4*>(a.out) + (size_t)tile * Y4 * EL_TM + row;
    if (grp >= P4) {
      const int f = 2 * P4 + (grp - P4);
      o4[(size_t)f * EL_TM] = y4[(size_t)f * EL_TM];
      continue;
    }
    const float4 p0 = y4[(size_t)grp * EL_TM], v0 = y4[(size_t)(P4 + grp) * EL_TM];
    float4 p = make_float4(p0.x + a.cpv * v0.x, p0.y + a.cpv * v0.y, p0.z + a.cpv * v0.z, p0.w + a.cpv * v0.w), v = v0;
#pragma unroll
    for (int s = 0; s < EL_MAX_A; ++s) {
      if (s < a.n_a) {
        const float4 x = reinterpret_cast<const float4*>(a.a[s])[((size_t)tile * P4 + grp) * EL_TM + row];
        const float cp = a.cpa[s], cv = a.cva[s];
        p.x += cp * x.x; p.y += cp * x.y; p.z += cp * x.z; p.w += cp * x.w;
        v.x += cv * x.x; v.y += cv * x.y; v.z += cv * x.z; v.w += cv * x.w;
      }
    }
    o4[(size_t)grp * EL_TM] = p;
    o4[(size_t)(P4 + grp) * EL_TM] = v;
  }
}

__global__ void __launch_bounds__(256) pv_combine_bwd_kernel(const __grid_constant__ ElemArgs a) {
  const int P4 = a.P / 4, H4 = a.H / 4, Y4 = 2 * P4 + H4;
  const int64_t n = (int64_t)a.ntiles * (P4 + H4) * EL_TM;
  const bool acc = a.accumulate != 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int tile, grp, row;
    decode(i, P4, H4, tile, grp, row);
    const float4* g4 = reinterpret_cast<const float4*>(a.g) + (size_t)tile * Y4 * EL_TM + row;
    float4* o4 = reinterpret_cast<float4*>(a.out) + (size_t)tile * Y4 * EL_TM + row;
    if (grp >= P4) {
      const int f = 2 * P4 + (grp - P4);
      float4 x = g4[(size_t)f * EL_TM];
      float4* q = o4 + (size_t)f * EL_TM;
      if (acc) { const float4 y = *q; x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w; }
      *q = x;
      continue;
    }
    const float4 gp = g4[(size_t)grp * EL_TM], gv = g4[(size_t)(P4 + grp) * EL_TM];
    float4 xp = gp;
    float4 xv = make_float4(a.cpv * gp.x + gv.x, a.cpv * gp.y + gv.y, a.cpv * gp.z + gv.z, a.cpv * gp.w + gv.w);
    float4* qp = o4 + (size_t)grp * EL_TM;
    float4* qv = o4 + (size_t)(P4 + grp) * EL_TM;
    if (acc) {
      const float4 y = *qp, z = *qv;
      xp.x += y.x; xp.y += y.y; xp.z += y.z; xp.w += y.w;
      xv.x += z.x; xv.y += z.y; xv.z += z.z; xv.w += z.w;
    }
    *qp = xp;
    *qv = xv;
#pragma unroll
    for (int s = 0; s < EL_MAX_A; ++s) {
      if (s < a.n_a) {
        const float cp = a.cpa[s], cv = a.cva[s];
        float4 x = make_float4(cp * gp.x + cv * gv.x, cp * gp.y + cv * gv.y, cp * gp.z + cv * gv.z, cp * gp.w + cv * gv.w);
        float4* q = reinterpret_cast<float4*>(a.G_a[s]) + ((size_t)tile * P4 + grp) * EL_TM + row;
        if (acc) { const float4 y = *q; x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w; }
        *q = x;
      }
    }
  }
}

// Several linear outputs of one step folded in ONE pass (each source: upstream gradient g_i and the output's combination):
//   G_y0.p (+)= sum_i g_i.p ; G_y0.v (+)= sum_i cpv_i g_i.p + g_i.v ; G_y0.h (+)= sum_i g_i.h ; G_a[j] (+)= sum_i cpa_i[j] g_i.p + cva_i[j] g_i.v
// (the end state of a dopri5 step and every dense-output row inside it; one read-modify-write of the accumulators
// instead of one per output)
constexpr int EL_MAX_SRC = 6;
struct MultiBwdArgs {
  const float* g[EL_MAX_SRC];
  float cpv[EL_MAX_SRC], cpa[EL_MAX_SRC][EL_MAX_A], cva[EL_MAX_SRC][EL_MAX_A];
  float* G_y0;
  float* G_a[EL_MAX_A];
  int n_src, n_a, accumulate, ntiles, P, H;
  int rowmajor_mask;      // bit s: source s is a ROW-MAJOR [B][D] row of the caller's gradient tensor (read in place: no transposed copy)
  int64_t B;
  const float* add_a;     // or null: blocked [Bp][P] added to G_a[add_idx] (the gradient a following step hands to its FSAL evaluation)
  int add_idx;
};
// float4 `f4` (of Y4 per agent) of agent (tile, row) from a blocked or a row-major source; row-major rows beyond B read as zero
__device__ __forceinline__ float4 src_f4(const float* g, bool rowmajor, int tile, int row, int f4, int Y4, int64_t B) {
  if (!rowmajor) return (reinterpret_cast<const float4*>(g) + (size_t)tile * Y4 * EL_TM + row)[(size_t)f4 * EL_TM];
  const int64_t agent = (int64_t)tile * EL_TM + row;
  if (agent >= B) return make_float4(0.f, 0.f, 0.f, 0.f);
  return __ldcs(reinterpret_cast<const float4*>(g) + agent * Y4 + f4);      // 16 B of a 640-byte row per lane: the neighbouring lanes of
                                                                            // the other feature groups pick the rest of the sector up from L2
}
// work item i -> (tile, group, row) for kernels that read ROW-MAJOR sources next to blocked ones: a warp covers 8 consecutive
// rows x 4 consecutive float4 groups, so a row-major access is 8 fully used 64-byte segments (decode(): 32 rows x 16 B = 32
// half-used sectors, every sector fetched twice) and a blocked access is 4 fully used 128-byte lines.  Needs (P4 + H4) % 4 == 0.
__device__ __forceinline__ void decode_8x4(int64_t i, int P4, int H4, int& tile, int& grp, int& row) {
  const int c = (int)(i & 3), r_lo = (int)((i >> 2) & 7);
  int64_t q = i >> 5;
  const int r_hi = (int)(q % (EL_TM / 8));
  q /= (EL_TM / 8);
  const int g_hi = (int)(q % ((P4 + H4) / 4));
  tile = (int)(q / ((P4 + H4) / 4));
  row = r_hi * 8 + r_lo;
  grp = g_hi * 4 + c;
}
__global__ void __launch_bounds__(256) pv_combine_bwd_multi_kernel(const __grid_constant__ MultiBwdArgs a) {
  const int P4 = a.P / 4, H4 = a.H / 4, Y4 = 2 * P4 + H4;
  const int64_t n = (int64_t)a.ntiles * (P4 + H4) * EL_TM;
  const bool acc = a.accumulate != 0;
  const bool lanes_8x4 = a.rowmajor_mask != 0 && (P4 % 4) == 0 && (H4 % 4) == 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int tile, grp, row;
    if (lanes_8x4) decode_8x4(i, P4, H4, tile, grp, row);
    else decode(i, P4, H4, tile, grp, row);
    const size_t t0 = (size_t)tile * Y4 * EL_TM + row;
    float4* o4 = reinterpret_cast<float4*>(a.G_y0) + t0;
    if (grp >= P4) {
      const size_t f = (size_t)(2 * P4 + (grp - P4)) * EL_TM;
      float4 x = acc ? o4[f] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int s = 0; s < EL_MAX_SRC; ++s)
        if (s < a.n_src) {
          const float4 g = src_f4(a.g[s], (a.rowmajor_mask >> s) & 1, tile, row, 2 * P4 + (grp - P4), Y4, a.B);
          x.x += g.x; x.y += g.y; x.z += g.z; x.w += g.w;
        }
      o4[f] = x;
      continue;
    }
    const size_t fp = (size_t)grp * EL_TM, fv = (size_t)(P4 + grp) * EL_TM;
    float4 gp[EL_MAX_SRC], gv[EL_MAX_SRC];
#pragma unroll
    for (int s = 0; s < EL_MAX_SRC; ++s)
      if (s < a.n_src) {
        const bool rm = (a.rowmajor_mask >> s) & 1;
        gp[s] = src_f4(a.g[s], rm, tile, row, grp, Y4, a.B);
        gv[s] = src_f4(a.g[s], rm, tile, row, P4 + grp, Y4, a.B);
      }
    float4 xp = acc ? o4[fp] : make_float4(0.f, 0.f, 0.f, 0.f), xv = acc ? o4[fv] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int s = 0; s < EL_MAX_SRC; ++s)
      if (s < a.n_src) {
        const float c = a.cpv[s];
        xp.x += gp[s].x; xp.y += gp[s].y; xp.z += gp[s].z; xp.w += gp[s].w;
        xv.x += c * gp[s].x + gv[s].x; xv.y += c * gp[s].y + gv[s].y; xv.z += c * gp[s].z + gv[s].z; xv.w += c * gp[s].w + gv[s].w;
      }
    o4[fp] = xp;
    o4[fv] = xv;
#pragma unroll 1
    for (int j = 0; j < a.n_a; ++j) {
      float4* q = reinterpret_cast<float4*>(a.G_a[j]) + ((size_t)tile * P4 + grp) * EL_TM + row;
      float4 x = acc ? *q : make_float4(0.f, 0.f, 0.f, 0.f);
      if (a.add_a != nullptr && j == a.add_idx) {
        const float4 e = (reinterpret_cast<const float4*>(a.add_a) + ((size_t)tile * P4 + grp) * EL_TM)[row];
        x.x += e.x; x.y += e.y; x.z += e.z; x.w += e.w;
      }
#pragma unroll
      for (int s = 0; s < EL_MAX_SRC; ++s)
        if (s < a.n_src) {
          const float cp = a.cpa[s][j], cv = a.cva[s][j];
          x.x += cp * gp[s].x + cv * gv[s].x; x.y += cp * gp[s].y + cv * gv[s].y;
          x.z += cp * gp[s].z + cv * gv[s].z; x.w += cp * gp[s].w + cv * gv[s].w;
        }
      *q = x;
    }
  }
}

// out = base + sum_l [gx_l.p ; cpv_l gx_l.p + gx_l.v ; gx_l.h]   (dL/dy0 of a step from the per-stage dL/d(stage input))
struct GatherArgs {
  const float* base;
  const float* gx[EL_MAX_A];
  float cpv[EL_MAX_A];
  float* out;
  // optional second product of the same pass (the gx are read once): ga_out = ga_base + sum_l dp[l] gx_l.p + dv[l] gx_l.v
  const float* ga_base;
  float* ga_out;
  float dp[EL_MAX_A], dv[EL_MAX_A];
  int n, ntiles, P, H;
};
__global__ void __launch_bounds__(256) adjoint_gather_kernel(const __grid_constant__ GatherArgs a) {
  const int P4 = a.P / 4, H4 = a.H / 4, Y4 = 2 * P4 + H4;
  const int64_t n = (int64_t)a.ntiles * (P4 + H4) * EL_TM;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int tile, grp, row;
    decode(i, P4, H4, tile, grp, row);
    const size_t t0 = (size_t)tile * Y4 * EL_TM + row;
    const float4* b4 = reinterpret_cast<const float4*>(a.base) + t0;
    float4* o4 = reinterpret_cast<float4*>(a.out) + t0;
    if (grp >= P4) {
      const size_t f = (size_t)(2 * P4 + (grp - P4)) * EL_TM;
      float4 x = b4[f];
#pragma unroll
      for (int s = 0; s < EL_MAX_A; ++s)
        if (s < a.n) {
          const float4 g = (reinterpret_cast<const float4*>(a.gx[s]) + t0)[f];
          x.x += g.x; x.y += g.y; x.z += g.z; x.w += g.w;
        }
      o4[f] = x;
      continue;
    }
    const size_t fp = (size_t)grp * EL_TM, fv = (size_t)(P4 + grp) * EL_TM;
    float4 xp = b4[fp], xv = b4[fv];
    const size_t ia = ((size_t)tile * P4 + grp) * EL_TM + row;
    float4 xa = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.ga_out != nullptr) xa = reinterpret_cast<const float4*>(a.ga_base)[ia];
#pragma unroll
    for (int s = 0; s < EL_MAX_A; ++s)
      if (s < a.n) {
        const float4* g4 = reinterpret_cast<const float4*>(a.gx[s]) + t0;
        const float4 gp = g4[fp], gv = g4[fv];
        const float c = a.cpv[s], dp = a.dp[s], dv = a.dv[s];
        xp.x += gp.x; xp.y += gp.y; xp.z += gp.z; xp.w += gp.w;
        xv.x += c * gp.x + gv.x; xv.y += c * gp.y + gv.y; xv.z += c * gp.z + gv.z; xv.w += c * gp.w + gv.w;
        xa.x += dp * gp.x + dv * gv.x; xa.y += dp * gp.y + dv * gv.y; xa.z += dp * gp.z + dv * gv.z; xa.w += dp * gp.w + dv * gv.w;
      }
    o4[fp] = xp;
    o4[fv] = xv;
    if (a.ga_out != nullptr) reinterpret_cast<float4*>(a.ga_out)[ia] = xa;
  }
}

// out (blocked [Bp][P]) = base + sum_l dp[l] gx_l.p + dv[l] gx_l.v : the upstream gradient of a stage written out as a buffer
// (needed when the stage itself is differentiated elsewhere: dopri5's FSAL evaluation belongs to the previous step)
struct AssembleArgs {
  const float* base;
  const float* gx[EL_MAX_A];
  float dp[EL_MAX_A], dv[EL_MAX_A];
  float* out;
  int n, ntiles, P, H;
};
__global__ void __launch_bounds__(256) ga_assemble_kernel(const __grid_constant__ AssembleArgs a) {
  const int P4 = a.P / 4, Y4 = 2 * P4 + a.H / 4;
  const int64_t n = (int64_t)a.ntiles * P4 * EL_TM;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(i % EL_TM);
    const int64_t q = i / EL_TM;
    const int grp = (int)(q % P4), tile = (int)(q / P4);
    float4 x = reinterpret_cast<const float4*>(a.base)[((size_t)tile * P4 + grp) * EL_TM + row];
#pragma unroll
    for (int s = 0; s < EL_MAX_A; ++s)
      if (s < a.n) {
        const float4* g4 = reinterpret_cast<const float4*>(a.gx[s]) + (size_t)tile * Y4 * EL_TM + row;
        const float4 gp = g4[(size_t)grp * EL_TM], gv = g4[(size_t)(P4 + grp) * EL_TM];
        const float dp = a.dp[s], dv = a.dv[s];
        x.x += dp * gp.x + dv * gv.x; x.y += dp * gp.y + dv * gv.y; x.z += dp * gp.z + dv * gv.z; x.w += dp * gp.w + dv * gv.w;
      }
    reinterpret_cast<float4*>(a.out)[((size_t)tile * P4 + grp) * EL_TM + row] = x;
  }
}

// ---- row-major <-> blocked ---------------------------------------------------------------------------------
// One CTA moves 32 rows x F floats through shared memory so that both the row-major side (rows contiguous) and the
// blocked side (32 consecutive agents of one float4 group contiguous) are accessed in full 128-byte lines.
//   mode 0: blocked  = row-major (padding rows zeroed)   mode 1: blocked += row-major   mode 2: row-major = blocked
__global__ void __launch_bounds__(256) rows_transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t B,
                                                             int64_t Bp, int F4, int mode) {
  extern __shared__ float4 tile_s[];     // [TR_ROWS][F4 + 1]
  const int ld = F4 + 1;
  const int64_t row0 = (int64_t)blockIdx.x * TR_ROWS;
  const int n = TR_ROWS * F4;
  if (mode != 2) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int r = i / F4, c = i % F4;
      const int64_t g = row0 + r;
      tile_s[r * ld + c] = (g < B) ? reinterpret_cast<const float4*>(src)[g * F4 + c] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int c = i / TR_ROWS, r = i % TR_ROWS;
      const int64_t g = row0 + r;
      if (g >= Bp) continue;
      float4* q = reinterpret_cast<float4*>(dst) + ((g / EL_TM) * F4 + c) * EL_TM + (g % EL_TM);
      float4 x = tile_s[r * ld + c];
      if (mode == 1) {
        if (g >= B) continue;
        const float4 y = *q;
        x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
      }
      *q = x;
    }
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int c = i / TR_ROWS, r = i % TR_ROWS;
      const int64_t g = row0 + r;
      if (g < Bp) tile_s[r * ld + c] = reinterpret_cast<const float4*>(src)[((g / EL_TM) * F4 + c) * EL_TM + (g % EL_TM)];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int r = i / F4, c = i % F4;
      const int64_t g = row0 + r;
      if (g < B) reinterpret_cast<float4*>(dst)[g * F4 + c] = tile_s[r * ld + c];
    }
  }
}

// pv_combine with a ROW-MAJOR result (the public trajectory row): blocked inputs are read in full 512 B segments, the
// 32 x F result tile is transposed through shared memory and written as contiguous rows -- saves the blocked round trip
__global__ void __launch_bounds__(256) pv_combine_rowmajor_kernel(const __grid_constant__ ElemArgs a, int64_t B) {
  extern __shared__ float4 tile_s[];     // [TR_ROWS][F4 + 1]
  const int P4 = a.P / 4, H4 = a.H / 4, F4 = 2 * P4 + H4, ld = F4 + 1;
  const int64_t row0 = (int64_t)blockIdx.x * TR_ROWS;
  const int n_work = TR_ROWS * (P4 + H4);
  for (int i = threadIdx.x; i < n_work; i += blockDim.x) {
    const int grp = i / TR_ROWS, r = i % TR_ROWS;
    const int64_t g = row0 + r;
    if (g >= B) continue;
    const int tile = (int)(g / EL_TM), row = (int)(g % EL_TM);
    const float4* y4 = reinterpret_cast<const float4*>(a.y0) + (size_t)tile * F4 * EL_TM + row;
    if (grp >= P4) {
      const int f = 2 * P4 + (grp - P4);
      tile_s[r * ld + f] = y4[(size_t)f * EL_TM];
      continue;
    }
    const float4 p0 = y4[(size_t)grp * EL_TM], v0 = y4[(size_t)(P4 + grp) * EL_TM];
    float4 p = make_float4(p0.x + a.cpv * v0.x, p0.y + a.cpv * v0.y, p0.z + a.cpv * v0.z, p0.w + a.cpv * v0.w), v = v0;
#pragma unroll
    for (int s = 0; s < EL_MAX_A; ++s) {
      if (s < a.n_a) {
        const float4 x = reinterpret_cast<const float4*>(a.a[s])[((size_t)tile * P4 + grp) * EL_TM + row];
        const float cp = a.cpa[s], cv = a.cva[s];
        p.x += cp * x.x; p.y += cp * x.y; p.z += cp * x.z; p.w += cp * x.w;
        v.x += cv * x.x; v.y += cv * x.y; v.z += cv * x.z; v.w += cv * x.w;
      }
    }
    tile_s[r * ld + grp] = p;
    tile_s[r * ld + P4 + grp] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < TR_ROWS * F4; i += blockDim.x) {
    const int r = i / F4, c = i % F4;
    const int64_t g = row0 + r;
    if (g < B) reinterpret_cast<float4*>(a.out)[g * F4 + c] = tile_s[r * ld + c];
  }
}

// Several ROW-MAJOR dense-output rows of the SAME step in one pass: y0 and the accelerations are read once, every row is
// its own linear combination (tdq interp.py `_interp_evaluate` expanded over the stage derivatives) -- a dopri5 step at
// rtol = atol = 1e-5 contains ~3 of the 97 requested rows, each of which re-read 2.4 KB per agent when launched alone.
constexpr int RM_MAX_ROWS = 4;
struct RowMajorMultiArgs {
  const float* y0;
  const float* a[EL_MAX_A];
  float* out[RM_MAX_ROWS];
  float cpv[RM_MAX_ROWS], cpa[RM_MAX_ROWS][EL_MAX_A], cva[RM_MAX_ROWS][EL_MAX_A];
  int n_a, n_rows, P, H;
};
__global__ void __launch_bounds__(256) pv_combine_rowmajor_multi_kernel(const __grid_constant__ RowMajorMultiArgs a, int64_t B) {
  extern __shared__ float4 tile_s[];     // [n_rows][TR_ROWS][F4 + 1]
  const int P4 = a.P / 4, H4 = a.H / 4, F4 = 2 * P4 + H4, ld = F4 + 1;
  const int64_t row0 = (int64_t)blockIdx.x * TR_ROWS;
  const int n_work = TR_ROWS * (P4 + H4);
  for (int i = threadIdx.x; i < n_work; i += blockDim.x) {
    const int grp = i / TR_ROWS, r = i % TR_ROWS;
    const int64_t g = row0 + r;
    if (g >= B) continue;
    const int tile = (int)(g / EL_TM), row = (int)(g % EL_TM);
    const float4* y4 = reinterpret_cast<const float4*>(a.y0) + (size_t)tile * F4 * EL_TM + row;
    if (grp >= P4) {
      const int f = 2 * P4 + (grp - P4);
      const float4 hv = y4[(size_t)f * EL_TM];
      for (int q = 0; q < a.n_rows; ++q) tile_s[(q * TR_ROWS + r) * ld + f] = hv;
      continue;
    }
    const float4 p0 = y4[(size_t)grp * EL_TM], v0 = y4[(size_t)(P4 + grp) * EL_TM];
    float4 x[EL_MAX_A];
#pragma unroll
    for (int s = 0; s < EL_MAX_A; ++s)
      if (s < a.n_a) x[s] = reinterpret_cast<const float4*>(a.a[s])[((size_t)tile * P4 + grp) * EL_TM + row];
#pragma unroll 1
    for (int q = 0; q < a.n_rows; ++q) {
      const float c = a.cpv[q];
      float4 p = make_float4(p0.x + c * v0.x, p0.y + c * v0.y, p0.z + c * v0.z, p0.w + c * v0.w), v = v0;
#pragma unroll
      for (int s = 0; s < EL_MAX_A; ++s) {
        if (s < a.n_a) {
          const float cp = a.cpa[q][s], cv = a.cva[q][s];
          p.x += cp * x[s].x; p.y += cp * x[s].y; p.z += cp * x[s].z; p.w += cp * x[s].w;
          v.x += cv * x[s].x; v.y += cv * x[s].y; v.z += cv * x[s].z; v.w += cv * x[s].w;
        }
      }
      tile_s[(q * TR_ROWS + r) * ld + grp] = p;
      tile_s[(q * TR_ROWS + r) * ld + P4 + grp] = v;
    }
  }
  __syncthreads();
  for (int q = 0; q < a.n_rows; ++q) {
    float4* o = reinterpret_cast<float4*>(a.out[q]);
    for (int i = threadIdx.x; i < TR_ROWS * F4; i += blockDim.x) {
      const int r = i / F4, c = i % F4;
      const int64_t g = row0 + r;
      if (g < B) o[g * F4 + c] = tile_s[(q * TR_ROWS + r) * ld + c];
    }
  }
}

static int launch_cfg(int64_t n) {
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = 148 * 16;
  return (int)(blocks < cap ? blocks : cap);
}

int pv_combine(const ab200_drift_desc* d, const float* y0, const float* const* a_ptrs, int n_a, float cpv, const float* cpa,
               const float* cva, int64_t B, float* out, cudaStream_t st) {
  if (n_a < 0 || n_a > EL_MAX_A || d->pos_dim % 4 || d->ctx_dim % 4) return AB200_ERR_BAD_ARG;
  ElemArgs k{};
  k.y0 = y0; k.out = out; k.cpv = cpv; k.n_a = n_a; k.P = d->pos_dim; k.H = d->ctx_dim;
  k.ntiles = (int)((B + EL_TM - 1) / EL_TM);
  for (int i = 0; i < n_a; ++i) { k.a[i] = a_ptrs[i]; k.cpa[i] = cpa[i]; k.cva[i] = cva[i]; }
  pv_combine_kernel<<<launch_cfg((int64_t)k.ntiles * EL_TM * ((k.P + k.H) / 4)), 256, 0, st>>>(k);
  return check_launch();
}

int pv_combine_rowmajor(const ab200_drift_desc* d, const float* y0, const float* const* a_ptrs, int n_a, float cpv, const float* cpa,
                        const float* cva, int64_t B, float* out_rowmajor, cudaStream_t st) {
  if (n_a < 0 || n_a > EL_MAX_A || d->pos_dim % 4 || d->ctx_dim % 4) return AB200_ERR_BAD_ARG;
  ElemArgs k{};
  k.y0 = y0; k.out = out_rowmajor; k.cpv = cpv; k.n_a = n_a; k.P = d->pos_dim; k.H = d->ctx_dim;
  k.ntiles = (int)((B + EL_TM - 1) / EL_TM);
  for (int i = 0; i < n_a; ++i) { k.a[i] = a_ptrs[i]; k.cpa[i] = cpa[i]; k.cva[i] = cva[i]; }
  const int F4 = (2 * k.P + k.H) / 4;
  const size_t smem = (size_t)TR_ROWS * (F4 + 1) * sizeof(float4);
  pv_combine_rowmajor_kernel<<<(int)((B + TR_ROWS - 1) / TR_ROWS), 256, smem, st>>>(k, B);
  return check_launch();
}

// n_rows row-major outputs out[q] = combination q (cpv[q], cpa[q * 8 + j], cva[q * 8 + j]) of the same (y0, a[]); any n_rows >= 1
int pv_combine_rowmajor_multi(const ab200_drift_desc* d, const float* y0, const float* const* a_ptrs, int n_a, int n_rows, const float* cpv,
                              const float* cpa, const float* cva, int64_t B, float* const* out_rowmajor, cudaStream_t st) {
  if (n_a < 0 || n_a > EL_MAX_A || n_rows < 1 || d->pos_dim % 4 || d->ctx_dim % 4) return AB200_ERR_BAD_ARG;
  const int F4 = (2 * d->pos_dim + d->ctx_dim) / 4;
  for (int r0 = 0; r0 < n_rows; r0 += RM_MAX_ROWS) {
    RowMajorMultiArgs k{};
    k.y0 = y0; k.n_a = n_a; k.P = d->pos_dim; k.H = d->ctx_dim;
    k.n_rows = (n_rows - r0 < RM_MAX_ROWS) ? n_rows - r0 : RM_MAX_ROWS;
    for (int i = 0; i < n_a; ++i) k.a[i] = a_ptrs[i];
    for (int q = 0; q < k.n_rows; ++q) {
      k.out[q] = out_rowmajor[r0 + q];
      k.cpv[q] = cpv[r0 + q];
      for (int i = 0; i < n_a; ++i) { k.cpa[q][i] = cpa[(r0 + q) * EL_MAX_A + i]; k.cva[q][i] = cva[(r0 + q) * EL_MAX_A + i]; }
    }
    const size_t smem = (size_t)k.n_rows * TR_ROWS * (F4 + 1) * sizeof(float4);
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(pv_combine_rowmajor_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) { set_cuda_error(e); return AB200_ERR_CUDA; }
    }
    pv_combine_rowmajor_multi_kernel<<<(int)((B + TR_ROWS - 1) / TR_ROWS), 256, smem, st>>>(k, B);
    const int rc = check_launch();
    if (rc) return rc;
  }
  return AB200_OK;
}

int pv_combine_bwd(const ab200_drift_desc* d, const float* g, int n_a, float cpv, const float* cpa, const float* cva, int64_t B,
                   float* G_y0, float* const* G_a, int accumulate, cudaStream_t st) {
  if (n_a < 0 || n_a > EL_MAX_A || d->pos_dim % 4 || d->ctx_dim % 4) return AB200_ERR_BAD_ARG;
  ElemArgs k{};
  k.g = g; k.out = G_y0; k.cpv = cpv; k.n_a = n_a; k.P = d->pos_dim; k.H = d->ctx_dim; k.accumulate = accumulate;
  k.ntiles = (int)((B + EL_TM - 1) / EL_TM);
  for (int i = 0; i < n_a; ++i) { k.G_a[i] = G_a[i]; k.cpa[i] = cpa[i]; k.cva[i] = cva[i]; }
  pv_combine_bwd_kernel<<<launch_cfg((int64_t)k.ntiles * EL_TM * ((k.P + k.H) / 4)), 256, 0, st>>>(k);
  return check_launch();
}

int adjoint_gather(const ab200_drift_desc* d, const float* base, const float* const* gx, int n, const float* cpv, int64_t B, float* out,
                   const float* ga_base, const float* dp, const float* dv, float* ga_out, cudaStream_t st) {
  if (n < 0 || n > EL_MAX_A || d->pos_dim % 4 || d->ctx_dim % 4) return AB200_ERR_BAD_ARG;
  GatherArgs k{};
  k.base = base; k.out = out; k.n = n; k.P = d->pos_dim; k.H = d->ctx_dim;
  k.ga_base = ga_base; k.ga_out = ga_out;
  k.ntiles = (int)((B + EL_TM - 1) / EL_TM);
  for (int i = 0; i < n; ++i) { k.gx[i] = gx[i]; k.cpv[i] = cpv[i]; k.dp[i] = ga_out ? dp[i] : 0.f; k.dv[i] = ga_out ? dv[i] : 0.f; }
  adjoint_gather_kernel<<<launch_cfg((int64_t)k.ntiles * EL_TM * ((k.P + k.H) / 4)), 256, 0, st>>>(k);
  return check_launch();
}

int pv_combine_bwd_multi(const ab200_drift_desc* d, const float* const* g, int n_src, const float* cpv, const float* cpa, const float* cva,
                         int n_a, int64_t B, float* G_y0, float* const* G_a, int accumulate, int rowmajor_mask, const float* add_a, int add_idx,
                         cudaStream_t st) {
  if (add_a != nullptr && (add_idx < 0 || add_idx >= n_a)) return AB200_ERR_BAD_ARG;
  if (n_src < 1 || n_src > EL_MAX_SRC || n_a < 0 || n_a > EL_MAX_A || d->pos_dim % 4 || d->ctx_dim % 4) return AB200_ERR_BAD_ARG;
  MultiBwdArgs k{};
  k.G_y0 = G_y0; k.n_src = n_src; k.n_a = n_a; k.accumulate = accumulate; k.P = d->pos_dim; k.H = d->ctx_dim;
  k.rowmajor_mask = rowmajor_mask; k.B = B;
  k.add_a = add_a; k.add_idx = add_idx;
  k.ntiles = (int)((B + EL_TM - 1) / EL_TM);
  for (int s = 0; s < n_src; ++s) {
    k.g[s] = g[s];
    k.cpv[s] = cpv[s];
    for (int j = 0; j < n_a; ++j) { k.cpa[s][j] = cpa[s * EL_MAX_A + j]; k.cva[s][j] = cva[s * EL_MAX_A + j]; }
  }
  for (int j = 0; j < n_a; ++j) k.G_a[j] = G_a[j];
  pv_combine_bwd_multi_kernel<<<launch_cfg((int64_t)k.ntiles * EL_TM * ((k.P + k.H) / 4)), 256, 0, st>>>(k);
  return check_launch();
}

int ga_assemble(const ab200_drift_desc* d, const float* base, const float* const* gx, int n, const float* dp, const float* dv,
                int64_t B, float* out, cudaStream_t st) {
  if (n < 0 || n > EL_MAX_A || d->pos_dim % 4 || d->ctx_dim % 4) return AB200_ERR_BAD_ARG;
  AssembleArgs k{};
  k.base = base; k.out = out; k.n = n; k.P = d->pos_dim; k.H = d->ctx_dim;
  k.ntiles = (int)((B + EL_TM - 1) / EL_TM);
  for (int i = 0; i < n; ++i) { k.gx[i] = gx[i]; k.dp[i] = dp[i]; k.dv[i] = dv[i]; }
  ga_assemble_kernel<<<launch_cfg((int64_t)k.ntiles * EL_TM * (k.P / 4)), 256, 0, st>>>(k);
  return check_launch();
}

int rows_transpose(const float* src, float* dst, int64_t B, int F, int mode, cudaStream_t st) {
  if (F % 4 || F <= 0 || F > 1024 || mode < 0 || mode > 2) return AB200_ERR_BAD_ARG;
  const int64_t Bp = (B + EL_TM - 1) / EL_TM * EL_TM;
  const int F4 = F / 4;
  const int64_t rows = (mode == 0) ? Bp : B;
  const int grid = (int)((rows + TR_ROWS - 1) / TR_ROWS);
  const size_t smem = (size_t)TR_ROWS * (F4 + 1) * sizeof(float4);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(rows_transpose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_cuda_error(e); return AB200_ERR_CUDA; }
  }
  rows_transpose_kernel<<<grid, 256, smem, st>>>(src, dst, B, Bp, F4, mode);
  return check_launch();
}

}  // namespace ab200
