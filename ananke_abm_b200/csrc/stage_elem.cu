// Elementwise companions of the stage kernels: linear combinations of the step's base state y0 = [p0, v0, h] and
// stage accelerations a_j (dense-output rows, step solutions) and their adjoints.  One pass over the state each,
// 128-bit accesses, HBM-bound.
//   forward :  out.p = p0 + cpv v0 + sum cpa[j] a_j ;  out.v = v0 + sum cva[j] a_j ;  out.h = h
//   adjoint :  G_y0.p (+)= g.p ; G_y0.v (+)= cpv g.p + g.v ; G_y0.h (+)= g.h ; G_a[j] (+)= cpa[j] g.p + cva[j] g.v
// (tdq: rk_common.py `_runge_kutta_step` y1 = y0 + dt sum c_sol k; interp.py `_interp_evaluate`.)
#include "common.cuh"

namespace ab200 {

constexpr int EL_MAX_A = 8;
struct ElemArgs {
  const float* y0;
  const float* a[EL_MAX_A];
  float* G_a[EL_MAX_A];
  float* out;        // forward: out [B][D];  adjoint: G_y0 [B][D]
  const float* g;    // adjoint: upstream gradient [B][D]
  float cpv, cpa[EL_MAX_A], cva[EL_MAX_A];
  int n_a, accumulate;
  int64_t B;
  int P, H;
};

__global__ void __launch_bounds__(256) pv_combine_kernel(const __grid_constant__ ElemArgs a) {
  const int P4 = a.P / 4, H4 = a.H / 4, D = 2 * a.P + a.H;
  const int64_t n = a.B * (P4 + H4);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / (P4 + H4);
    const int c = (int)(i % (P4 + H4));
    const float* yr = a.y0 + b * D;
    float* o = a.out + b * D;
    if (c >= P4) {
      reinterpret_cast<float4*>(o + 2 * a.P)[c - P4] = reinterpret_cast<const float4*>(yr + 2 * a.P)[c - P4];
      continue;
    }
    const float4 p0 = reinterpret_cast<const float4*>(yr)[c], v0 = reinterpret_cast<const float4*>(yr + a.P)[c];
    float4 p = make_float4(p0.x + a.cpv * v0.x, p0.y + a.cpv * v0.y, p0.z + a.cpv * v0.z, p0.w + a.cpv * v0.w), v = v0;
#pragma unroll
    for (int s = 0; s < EL_MAX_A; ++s) {
      if (s < a.n_a) {
        const float4 x = reinterpret_cast<const float4*>(a.a[s] + b * a.P)[c];
        const float cp = a.cpa[s], cv = a.cva[s];
        p.x += cp * x.x; p.y += cp * x.y; p.z += cp * x.z; p.w += cp * x.w;
        v.x += cv * x.x; v.y += cv * x.y; v.z += cv * x.z; v.w += cv * x.w;
      }
    }
    reinterpret_cast<float4*>(o)[c] = p;
    reinterpret_cast<float4*>(o + a.P)[c] = v;
  }
}

__global__ void __launch_bounds__(256) pv_combine_bwd_kernel(const __grid_constant__ ElemArgs a) {
  const int P4 = a.P / 4, H4 = a.H / 4, D = 2 * a.P + a.H;
  const int64_t n = a.B * (P4 + H4);
  const bool acc = a.accumulate != 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / (P4 + H4);
    const int c = (int)(i % (P4 + H4));
    const float* gr = a.g + b * D;
    float* o = a.out + b * D;
    if (c >= P4) {
      float4 x = reinterpret_cast<const float4*>(gr + 2 * a.P)[c - P4];
      float4* q = reinterpret_cast<float4*>(o + 2 * a.P) + (c - P4);
      if (acc) { const float4 y = *q; x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w; }
      *q = x;
      continue;
    }
    const float4 gp = reinterpret_cast<const float4*>(gr)[c], gv = reinterpret_cast<const float4*>(gr + a.P)[c];
    float4 xp = gp;
    float4 xv = make_float4(a.cpv * gp.x + gv.x, a.cpv * gp.y + gv.y, a.cpv * gp.z + gv.z, a.cpv * gp.w + gv.w);
    float4* qp = reinterpret_cast<float4*>(o) + c;
    float4* qv = reinterpret_cast<float4*>(o + a.P) + c;
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
        float4* q = reinterpret_cast<float4*>(a.G_a[s] + b * a.P) + c;
        if (acc) { const float4 y = *q; x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w; }
        *q = x;
      }
    }
  }
}

static int launch_cfg(int64_t B, int P, int H) {
  const int64_t n = B * ((P + H) / 4);
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = 148 * 16;
  return (int)(blocks < cap ? blocks : cap);
}

int pv_combine(const ab200_drift_desc* d, const float* y0, const float* const* a_ptrs, int n_a, float cpv, const float* cpa,
               const float* cva, int64_t B, float* out, cudaStream_t st) {
  if (n_a < 0 || n_a > EL_MAX_A || d->pos_dim % 4 || d->ctx_dim % 4) return AB200_ERR_BAD_ARG;
  ElemArgs k{};
  k.y0 = y0; k.out = out; k.cpv = cpv; k.n_a = n_a; k.B = B; k.P = d->pos_dim; k.H = d->ctx_dim;
  for (int i = 0; i < n_a; ++i) { k.a[i] = a_ptrs[i]; k.cpa[i] = cpa[i]; k.cva[i] = cva[i]; }
  pv_combine_kernel<<<launch_cfg(B, k.P, k.H), 256, 0, st>>>(k);
  return check_launch();
}

int pv_combine_bwd(const ab200_drift_desc* d, const float* g, int n_a, float cpv, const float* cpa, const float* cva, int64_t B,
                   float* G_y0, float* const* G_a, int accumulate, cudaStream_t st) {
  if (n_a < 0 || n_a > EL_MAX_A || d->pos_dim % 4 || d->ctx_dim % 4) return AB200_ERR_BAD_ARG;
  ElemArgs k{};
  k.g = g; k.out = G_y0; k.cpv = cpv; k.n_a = n_a; k.B = B; k.P = d->pos_dim; k.H = d->ctx_dim; k.accumulate = accumulate;
  for (int i = 0; i < n_a; ++i) { k.G_a[i] = G_a[i]; k.cpa[i] = cpa[i]; k.cva[i] = cva[i]; }
  pv_combine_bwd_kernel<<<launch_cfg(B, k.P, k.H), 256, 0, st>>>(k);
  return check_launch();
}

}  // namespace ab200
