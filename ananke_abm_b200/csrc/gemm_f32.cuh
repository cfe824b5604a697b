// fp32 FFMA tile GEMMs shared by the strict-precision forward and backward kernels.
//
// Activations live transposed in shared memory, [feature][agent] with row stride XS floats, so that a
// thread reads the 4 agents it owns as one 128-bit load and every k-step is a broadcast-friendly
// rank-1 update.  Weights [K][N] are streamed from L2 in KC-row chunks with cp.async double buffering.
#pragma once
#include "common.cuh"

namespace ab200 {

constexpr int NT = 256;   // threads per CTA in all fp32 kernels
constexpr int KC = 32;    // k rows per staged weight chunk

template <int N>
__device__ __forceinline__ void stage_weights(float* sW, const float* __restrict__ gW, int k0, int rows) {
  const int n4 = rows * N / 4;
  const float4* src = reinterpret_cast<const float4*>(gW + (size_t)k0 * N);
  float4* dst = reinterpret_cast<float4*>(sW);
  for (int i = threadIdx.x; i < n4; i += NT) cp_async16(dst + i, src + i);
}

// Thread mapping for a TMv x N output tile: tm = tid % (TMv/4) owns agents 4tm..4tm+3,
// tn = tid / (TMv/4) owns CN = max(1, N/TGN) consecutive columns; threads with tn*CN >= N idle.
template <int TMv, int N>
struct TileMap {
  static constexpr int TGM = TMv / 4;
  static constexpr int TGN = NT / TGM;
  static constexpr int CN = (N / TGN) > 0 ? (N / TGN) : 1;
  __device__ static __forceinline__ int tm() { return threadIdx.x % TGM; }
  __device__ static __forceinline__ int tn() { return threadIdx.x / TGM; }
  __device__ static __forceinline__ bool active() { return tn() * CN < N; }
};

// acc[4][CN] = sum_k X^T[k][agents] * W[k][cols]     (sX: [K][XS] shared; gW: [K][N] global, contiguous)
// Ends with a __syncthreads(): on return every thread is done reading sX and the staging buffer.
template <int TMv, int XS, int K, int N>
__device__ __forceinline__ void gemm_tile(const float* __restrict__ gW, const float* sX, float* sW,
                                          float (&acc)[4][TileMap<TMv, N>::CN]) {
  using M = TileMap<TMv, N>;
  constexpr int CN = M::CN;
  static_assert(N % 4 == 0 && K % 4 == 0 && TMv % 4 == 0, "shape");
  const int tm = M::tm(), tn = M::tn();
  const bool on = M::active();
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < CN; ++j) acc[i][j] = 0.0f;

  constexpr int NCH = (K + KC - 1) / KC;
  stage_weights<N>(sW, gW, 0, (K < KC ? K : KC));
  cp_async_commit();
#pragma unroll 1
  for (int c = 0; c < NCH; ++c) {
    const int k0 = c * KC;
    const int rows = (K - k0 < KC) ? (K - k0) : KC;
    float* cur = sW + (c & 1) * (KC * N);
    if (c + 1 < NCH) {
      const int k1 = k0 + KC;
      stage_weights<N>(sW + ((c + 1) & 1) * (KC * N), gW, k1, (K - k1 < KC) ? (K - k1) : KC);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (on) {
      const float* xr = sX + (size_t)k0 * XS + 4 * tm;
      const float* wr = cur + tn * CN;
#pragma unroll 8
      for (int kk = 0; kk < rows; ++kk) {
        const float4 xa = *reinterpret_cast<const float4*>(xr + kk * XS);
        float w[CN];
        if constexpr (CN % 4 == 0) {
#pragma unroll
          for (int j = 0; j < CN; j += 4) {
            const float4 t4 = *reinterpret_cast<const float4*>(wr + kk * N + j);
            w[j] = t4.x; w[j + 1] = t4.y; w[j + 2] = t4.z; w[j + 3] = t4.w;
          }
        } else if constexpr (CN % 2 == 0) {
#pragma unroll
          for (int j = 0; j < CN; j += 2) {
            const float2 t2 = *reinterpret_cast<const float2*>(wr + kk * N + j);
            w[j] = t2.x; w[j + 1] = t2.y;
          }
        } else {
#pragma unroll
          for (int j = 0; j < CN; ++j) w[j] = wr[kk * N + j];
        }
#pragma unroll
        for (int j = 0; j < CN; ++j) {
          acc[0][j] = fmaf(xa.x, w[j], acc[0][j]);
          acc[1][j] = fmaf(xa.y, w[j], acc[1][j]);
          acc[2][j] = fmaf(xa.z, w[j], acc[2][j]);
          acc[3][j] = fmaf(xa.w, w[j], acc[3][j]);
        }
      }
    }
    __syncthreads();
  }
}

template <int ACT>
__device__ __forceinline__ float act_fn(float x) {
  if (ACT == 0) return fmaxf(x, 0.0f);
  return tanhf(x);
}
// derivative of the activation expressed through its OUTPUT value
template <int ACT>
__device__ __forceinline__ float act_grad_from_out(float out) {
  if (ACT == 0) return out > 0.0f ? 1.0f : 0.0f;
  return 1.0f - out * out;
}

}  // namespace ab200
