// Euler-Maruyama step for the SDE branch: the reference calls
//   sdeint(sde, y0, ts, method="euler", dt=0.01)     latent_ode/architecture/model.py:192-194, mode_sep/architecture/model.py:158-182
// with diagonal Ito noise that acts on the state part only (g = sigma on [p, v] / state, 0 on the context h).
//   y_out = y + f * dt + g * sqrt(dt) * xi,      xi ~ N(0, 1)
// torchsde's Brownian interval cannot be reproduced, so the noise has its OWN specification, restated in
// oracle/sde_oracle.py: xi for element group q = (row * D + d) / 4 of step n comes from Philox4x32-10 with
// counter = (q_lo, q_hi, n_lo, n_hi), key = (seed_lo, seed_hi); the four 32-bit outputs make two Box-Muller pairs
// (u = (x + 0.5) * 2^-32;  z0 = sqrt(-2 ln u1) cos(2 pi u2), z1 = ... sin ...).  Counter-based: any step of any agent
// can be regenerated independently (sharding over ranks does not change an agent's noise).
// One pass over the state, 128-bit accesses; D must be a multiple of 4.
#include "common.cuh"

namespace ab200 {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ void box_muller(uint32_t x0, uint32_t x1, float& z0, float& z1) {
  const float u1 = ((float)x0 + 0.5f) * 2.3283064365386963e-10f;     // (0, 1]: float rounding may reach 1, never 0
  const float u2 = ((float)x1 + 0.5f) * 2.3283064365386963e-10f;
  const float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincosf(6.283185307179586f * u2, &s, &c);
  z0 = r * c; z1 = r * s;
}

// g: [D] (g_per_row == 0, the same diffusion for every agent) or [B][D]
__global__ void __launch_bounds__(256) sde_euler_step_kernel(const float* __restrict__ y, const float* __restrict__ f,
                                                             const float* __restrict__ g, int g_per_row, int64_t B, int D, float dt,
                                                             float sqrt_dt, uint64_t seed, uint64_t step, float* __restrict__ y_out,
                                                             float* __restrict__ xi_out) {
  const int64_t n4 = B * (int64_t)D / 4;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n4; q += (int64_t)gridDim.x * blockDim.x) {
    uint32_t rnd[4];
    philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)step, (uint32_t)(step >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), rnd);
    float z[4];
    box_muller(rnd[0], rnd[1], z[0], z[1]);
    box_muller(rnd[2], rnd[3], z[2], z[3]);
    const float4 yv = reinterpret_cast<const float4*>(y)[q];
    const float4 fv = reinterpret_cast<const float4*>(f)[q];
    const int64_t e = q * 4;
    const float4 gv = *reinterpret_cast<const float4*>(g + (g_per_row ? e : e % D));
    float4 o;
    o.x = yv.x + fv.x * dt + gv.x * sqrt_dt * z[0];
    o.y = yv.y + fv.y * dt + gv.y * sqrt_dt * z[1];
    o.z = yv.z + fv.z * dt + gv.z * sqrt_dt * z[2];
    o.w = yv.w + fv.w * dt + gv.w * sqrt_dt * z[3];
    reinterpret_cast<float4*>(y_out)[q] = o;
    if (xi_out != nullptr) reinterpret_cast<float4*>(xi_out)[q] = make_float4(z[0], z[1], z[2], z[3]);
  }
}

int sde_euler_step(const float* y, const float* f, const float* g, int g_per_row, int64_t B, int D, float dt, uint64_t seed,
                   uint64_t step, float* y_out, float* xi_out, cudaStream_t st) {
  if ((D & 3) != 0) return AB200_ERR_UNSUPPORTED;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t n4 = B * (int64_t)D / 4;
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
  if (blocks < 1) blocks = 1;
  sde_euler_step_kernel<<<(int)blocks, 256, 0, st>>>(y, f, g, g_per_row, B, D, dt, sqrtf(dt), seed, step, y_out, xi_out);
  return check_launch();
}

}  // namespace ab200
