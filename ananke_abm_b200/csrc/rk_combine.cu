// Generic-func path: fused Runge-Kutta stage combine and (for adaptive pairs) the error-norm partial.
// One pass over the state per call: reads y0 and the n_k stage derivatives once, writes one vector.
// Purely HBM-bound: (n_k + 2) * 4 bytes per element; grid = multiple of the SM count, 128-bit accesses.
#include "common.cuh"

namespace ab200 {

constexpr int MAXK = 8;
struct KPtrs {
  const float* k[MAXK];
  float c[MAXK];
  float e[MAXK];
};

template <int VEC>
__global__ void __launch_bounds__(256) stage_combine_kernel(const float* __restrict__ y, KPtrs kp, int n_k, float dt,
                                                            float* __restrict__ out, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * VEC;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC; i < n; i += stride) {
    if (VEC == 4 && i + 3 < n) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j = 0; j < n_k; ++j) {
        const float4 kv = *reinterpret_cast<const float4*>(kp.k[j] + i);
        const float c = kp.c[j] * dt;           // tdq: k[..., :i+1] * (beta_i * dt), summed over stages
        acc.x = fadd(acc.x, fmul(kv.x, c)); acc.y = fadd(acc.y, fmul(kv.y, c));
        acc.z = fadd(acc.z, fmul(kv.z, c)); acc.w = fadd(acc.w, fmul(kv.w, c));
      }
      const float4 yv = *reinterpret_cast<const float4*>(y + i);
      *reinterpret_cast<float4*>(out + i) = make_float4(fadd(yv.x, acc.x), fadd(yv.y, acc.y), fadd(yv.z, acc.z), fadd(yv.w, acc.w));
    } else {
      for (int64_t q = i; q < n && q < i + VEC; ++q) {
        float acc = 0.f;
        for (int j = 0; j < n_k; ++j) acc = fadd(acc, fmul(kp.k[j][q], kp.c[j] * dt));
        out[q] = fadd(y[q], acc);
      }
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(256) combine_errnorm_kernel(const float* __restrict__ y0, KPtrs kp, int n_k, float dt,
                                                              float rtol, float atol, float* __restrict__ y1_out,
                                                              float* __restrict__ sumsq, int64_t n) {
  float local = 0.f;
  auto one = [&](float a, float s, float e) -> float {      // element: solution value + squared error ratio
    const float b = fadd(a, s);
    const float tol = atol + rtol * fmaxf(fabsf(a), fabsf(b));
    const float r = e / tol;
    local = fmaf(r, r, local);
    return b;
  };
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * VEC;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC; i < n; i += stride) {
    if (VEC == 4 && i + 3 < n) {
      // all stage loads are issued before the first is consumed (8 independent 128-bit loads in flight per thread)
      float4 kv[MAXK];
#pragma unroll
      for (int j = 0; j < MAXK; ++j)
        if (j < n_k) kv[j] = *reinterpret_cast<const float4*>(kp.k[j] + i);
      const float4 a = *reinterpret_cast<const float4*>(y0 + i);
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f), e = s;
#pragma unroll
      for (int j = 0; j < MAXK; ++j)
        if (j < n_k) {
          const float c = kp.c[j] * dt, ce = kp.e[j] * dt;
          s.x = fadd(s.x, fmul(kv[j].x, c)); s.y = fadd(s.y, fmul(kv[j].y, c));
          s.z = fadd(s.z, fmul(kv[j].z, c)); s.w = fadd(s.w, fmul(kv[j].w, c));
          e.x = fadd(e.x, fmul(kv[j].x, ce)); e.y = fadd(e.y, fmul(kv[j].y, ce));
          e.z = fadd(e.z, fmul(kv[j].z, ce)); e.w = fadd(e.w, fmul(kv[j].w, ce));
        }
      const float4 b = make_float4(one(a.x, s.x, e.x), one(a.y, s.y, e.y), one(a.z, s.z, e.z), one(a.w, s.w, e.w));
      if (y1_out) *reinterpret_cast<float4*>(y1_out + i) = b;
    } else {
      for (int64_t q = i; q < n && q < i + VEC; ++q) {
        float s = 0.f, e = 0.f;
        for (int j = 0; j < n_k; ++j) {
          const float kv = kp.k[j][q];
          s = fadd(s, fmul(kv, kp.c[j] * dt));
          e = fadd(e, fmul(kv, kp.e[j] * dt));
        }
        const float b = one(y0[q], s, e);
        if (y1_out) y1_out[q] = b;
      }
    }
  }
  // warp -> block -> one atomic per block
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = part[threadIdx.x];
    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffu, v, o);
    if (threadIdx.x == 0) atomicAdd(sumsq, v);
  }
}

static int grid_for(int64_t n, int per_thread) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t blocks = (n + 256 * (int64_t)per_thread - 1) / (256 * (int64_t)per_thread);
  const int64_t cap = (int64_t)sms * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

int rk_stage_combine(const float* y, const float* const* k, const float* coef, int n_k, float dt, float* out, int64_t n,
                     cudaStream_t st) {
  if (n_k < 0 || n_k > MAXK) return AB200_ERR_BAD_ARG;
  KPtrs kp{};
  bool aligned = (((uintptr_t)y | (uintptr_t)out) & 15) == 0;
  for (int j = 0; j < n_k; ++j) { kp.k[j] = k[j]; kp.c[j] = coef[j]; aligned = aligned && (((uintptr_t)k[j]) & 15) == 0; }
  if (aligned) stage_combine_kernel<4><<<grid_for(n, 4), 256, 0, st>>>(y, kp, n_k, dt, out, n);
  else stage_combine_kernel<1><<<grid_for(n, 1), 256, 0, st>>>(y, kp, n_k, dt, out, n);
  return check_launch();
}

int rk_combine_errnorm(const float* y0, const float* const* k, const float* csol, const float* cerr, int n_k, float dt,
                       float rtol, float atol, float* y1_out, float* sumsq, int64_t n, cudaStream_t st) {
  if (n_k < 0 || n_k > MAXK) return AB200_ERR_BAD_ARG;
  KPtrs kp{};
  bool aligned = (((uintptr_t)y0 | (uintptr_t)y1_out) & 15) == 0;
  for (int j = 0; j < n_k; ++j) {
    kp.k[j] = k[j]; kp.c[j] = csol[j]; kp.e[j] = cerr[j];
    aligned = aligned && (((uintptr_t)k[j]) & 15) == 0;
  }
  if (aligned) combine_errnorm_kernel<4><<<grid_for(n, 4), 256, 0, st>>>(y0, kp, n_k, dt, rtol, atol, y1_out, sumsq, n);
  else combine_errnorm_kernel<1><<<grid_for(n, 1), 256, 0, st>>>(y0, kp, n_k, dt, rtol, atol, y1_out, sumsq, n);
  return check_launch();
}

}  // namespace ab200
