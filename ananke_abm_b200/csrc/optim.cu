// Optimiser step over the flat fp32 buffers the gradient all-reduce already uses:
//   torch.nn.utils.clip_grad_norm_(params, max_norm) ; torch.optim.Adam(params, lr, weight_decay).step()
//   (mode_sep/train/train.py:68,163-164) as two launches and no host synchronisation: the squared gradient norm is
//   reduced on the device and the update kernel derives the clip coefficient from it.
// HBM-bound: reads p, g, m, v and writes p, m, v once (28 B per parameter).
#include "common.cuh"

namespace ab200 {

__global__ void __launch_bounds__(256) grad_sumsq_kernel(const float* __restrict__ g, int64_t n, double* __restrict__ out) {
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = g[i];
    acc += (double)x * (double)x;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += part[i];
    atomicAdd(out, s);
  }
}

// torch.optim.Adam (single-tensor path, amsgrad = False, maximize = False), same operation order in fp32:
//   g += wd * p ; m = lerp(m, g, 1 - b1) ; v = v * b2 + (1 - b2) g g ; p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
__global__ void __launch_bounds__(256) adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                                                        float wd, float bc1, float bc2_sqrt, float max_norm,
                                                        const double* __restrict__ sumsq) {
  float clip = 1.0f;
  if (max_norm > 0.0f && sumsq != nullptr) {
    const float total = (float)sqrt(*sumsq);
    clip = fminf(max_norm / (total + 1e-6f), 1.0f);          // clip_grad_norm_: clamp(max_norm / (norm + 1e-6), max = 1)
  }
  const float step_size = lr / bc1;        // torch: step_size = lr / bias_correction1 (a Python double, applied as a float scalar)
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float pi = p[i];
    float gi = g[i] * clip;
    if (wd != 0.0f) gi = fmaf(wd, pi, gi);
    float mi = m[i];
    mi = mi + (1.0f - b1) * (gi - mi);
    const float vi = v[i] * b2 + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - step_size * (mi / denom);
  }
}

static int grid_1d(int64_t n) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t b = (n + 255) / 256;
  const int64_t cap = (int64_t)sms * 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

int grad_sumsq(const float* g, int64_t n, double* out, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(double), st);
  if (e != cudaSuccess) { set_cuda_error(e); return AB200_ERR_CUDA; }
  grad_sumsq_kernel<<<grid_1d(n), 256, 0, st>>>(g, n, out);
  return check_launch();
}

int adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps, float wd, int step,
              float max_norm, const double* sumsq, cudaStream_t st) {
  // bias corrections in double on the host, as torch's Python front end computes them
  const double bc1 = 1.0 - pow((double)b1, (double)step);
  const double bc2_sqrt = sqrt(1.0 - pow((double)b2, (double)step));
  adam_step_kernel<<<grid_1d(n), 256, 0, st>>>(p, g, m, v, n, lr, b1, b2, eps, wd, (float)bc1, (float)bc2_sqrt, max_norm, sumsq);
  return check_launch();
}

}  // namespace ab200
