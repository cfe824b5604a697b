// Tensor-core fused RK4 (3/8 rule) trajectory kernel: tcgen05.mma (fp16 x fp16 -> fp32 in TMEM; fp16 = 11-bit significand,
// same throughput as bf16 and 8x less rounding noise -- every value of this net is far inside the fp16 range), one CTA per SM,
// one 128-agent tile at a time, the WHOLE trajectory of the tile on chip.
//
//   shared memory : all six weight matrices of the drift net as 16-bit (fp16) values in the canonical K-major (8x8 core matrix,
//                   un-swizzled) UMMA layout, resident for the kernel's lifetime (184 KiB), + fp32 biases
//   tensor memory : ACC  [  0,128)  fp32 accumulator of the current layer           (lane = agent row)
//                   ACT0 [128,192)  16-bit A operand: stage input [p,v] / residual stream z   (2 values / column)
//                   HCTX [192,208)  16-bit A operand: the agent's static context h (K-steps 8,9 of layer 1)
//                   ACT1 [208,272)  16-bit A operand: residual-inner activation u
//   registers     : the fp32 Runge-Kutta state of the thread's (agent, 32-dim slice): p0, v0 and two stage
//                   combinations -- the ODE state never round-trips through HBM inside a step
// Every layer is: 8 (or 10) tcgen05.mma issued by one thread with A read from TMEM and B from shared memory,
// tcgen05.commit -> mbarrier, then all 8 warps run the epilogue (tcgen05.ld, bias + activation in fp32, pack to
// bf16, tcgen05.st straight into the next layer's A operand).  HBM traffic is the algorithmic minimum.
#include <cuda_fp16.h>
#include "common.cuh"
#include "umma.cuh"
#include "stage_tc.cuh"      // pack2 / un_lo / un_hi (operand-format helpers)

namespace ab200 {
using namespace umma;

constexpr int TC_THREADS = 256;
constexpr int TC_M = 128;          // agents per tile == UMMA M

// canonical un-swizzled K-major image of a [N][K] bf16 matrix: core matrices of 8 rows x 8 k
__host__ __device__ constexpr uint32_t tc_lbo(int N) { return (uint32_t)N * 16u; }   // K-adjacent core matrices
constexpr uint32_t TC_SBO = 128u;                                                        // N-adjacent core matrices

template <int P, int H, int HID, int NRES>
struct TcLayout {
  static constexpr int K1 = 2 * P + H;
  // byte offsets of the weight images inside the packed buffer / shared memory
  static constexpr uint32_t off_w1 = 0;
  static constexpr uint32_t sz_w1 = (uint32_t)HID * K1 * 2;
  static constexpr uint32_t off_res = off_w1 + sz_w1;
  static constexpr uint32_t sz_hh = (uint32_t)HID * HID * 2;
  static constexpr uint32_t off_wo = off_res + 2u * NRES * sz_hh;
  static constexpr uint32_t sz_wo = (uint32_t)P * HID * 2;
  static constexpr uint32_t w_bytes = off_wo + sz_wo;
  // fp32 vectors: b1, wsin, wcos, {bA, bB} x NRES, bO
  static constexpr uint32_t off_vec = w_bytes;
  static constexpr uint32_t n_vec = (uint32_t)(3 + 2 * NRES) * HID + P;
  static constexpr uint32_t total_bytes = off_vec + n_vec * 4;
  __host__ __device__ static constexpr uint32_t off_wa(int r) { return off_res + (uint32_t)(2 * r) * sz_hh; }
  __host__ __device__ static constexpr uint32_t off_wb(int r) { return off_res + (uint32_t)(2 * r + 1) * sz_hh; }
};

// ---- prepack: torch-layout fp32 weights -> bf16 UMMA images + fp32 vectors ---------------------------------
template <int P, int H, int HID, int NRES>
__global__ void pack_tc_kernel(const float* __restrict__ w, uint8_t* __restrict__ out) {
  using TL = TcLayout<P, H, HID, NRES>;
  const FlatLayout F{P, H, HID, NRES};
  const int IN = 2 * P + H + 2;
  const int64_t n_w1 = (int64_t)HID * TL::K1, n_hh = (int64_t)HID * HID, n_wo = (int64_t)P * HID;
  const int64_t n_mats = n_w1 + 2 * NRES * n_hh + n_wo;
  const int64_t total = n_mats + TL::n_vec;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    if (i < n_w1) {
      const int n = (int)(i / TL::K1), k = (int)(i % TL::K1);
      *reinterpret_cast<__half*>(out + TL::off_w1 + off_kmajor_noswz(n, k, tc_lbo(HID), TC_SBO)) =
          __float2half_rn(w[F.off_win() + (int64_t)n * IN + k]);
    } else if (i < n_w1 + 2 * NRES * n_hh) {
      const int64_t q = i - n_w1;
      const int m = (int)(q / n_hh);            // 0..2*NRES-1 : a0,b0,a1,b1,...
      const int n = (int)((q % n_hh) / HID), k = (int)(q % HID);
      const int64_t src = (m & 1) ? F.off_wb(m >> 1) : F.off_wa(m >> 1);
      *reinterpret_cast<__half*>(out + TL::off_res + (uint32_t)m * TL::sz_hh + off_kmajor_noswz(n, k, tc_lbo(HID), TC_SBO)) =
          __float2half_rn(w[src + (int64_t)n * HID + k]);
    } else if (i < n_mats) {
      const int64_t q = i - n_w1 - 2 * NRES * n_hh;
      const int n = (int)(q / HID), k = (int)(q % HID);
      *reinterpret_cast<__half*>(out + TL::off_wo + off_kmajor_noswz(n, k, tc_lbo(P), TC_SBO)) =
          __float2half_rn(w[F.off_wout() + (int64_t)n * HID + k]);
    } else {
      const int64_t q = i - n_mats;
      float* vec = reinterpret_cast<float*>(out + TL::off_vec);
      float v;
      if (q < HID) v = w[F.off_bin() + q];
      else if (q < 2 * HID) v = w[F.off_win() + (q - HID) * IN + 2 * P + H];
      else if (q < 3 * HID) v = w[F.off_win() + (q - 2 * HID) * IN + 2 * P + H + 1];
      else if (q < (3 + 2 * NRES) * HID) {
        const int m = (int)((q - 3 * HID) / HID), n = (int)((q - 3 * HID) % HID);
        v = w[((m & 1) ? F.off_bb(m >> 1) : F.off_ba(m >> 1)) + n];
      } else v = w[F.off_bout() + (q - (3 + 2 * NRES) * HID)];
      vec[q] = v;
    }
  }
}

struct TcArgs {
  const uint8_t* pk;     // packed image (TcLayout)
  const float* y0;       // [B][D]
  const float* t;        // [T]
  float* y_path;         // [T][B][D]
  int64_t B;
  int T;
  int ntiles;
  float period;
  int* status;           // device int: set to 1 if an MMA barrier timed out
};

template <int ACT>
__device__ __forceinline__ float tc_act(float x) {
  if (ACT == 0) return fmaxf(x, 0.0f);
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int P, int H, int HID, int NRES, int ACT>
__global__ void __launch_bounds__(TC_THREADS, 1) rk4_tc_kernel(TcArgs a) {
  static_assert(P == 64 && HID == 128 && H == 32, "v1 instantiation: mode_sep shapes");
  using TL = TcLayout<P, H, HID, NRES>;
  constexpr int D = 2 * P + H;
  constexpr int PD = P / 2;                  // state dims owned by one thread (its column half)
  constexpr uint32_t C_ACC = 0, C_ACT0 = 128, C_H = 192, C_ACT1 = 208;

  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const float* sVec = reinterpret_cast<const float*>(smem + TL::off_vec);
  const float* sB1 = sVec;
  const float* sWsin = sVec + HID;
  const float* sWcos = sVec + 2 * HID;
  const float* sBres = sVec + 3 * HID;       // [2*NRES][HID]
  const float* sBO = sVec + (3 + 2 * NRES) * HID;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, hf = warp >> 2;
  const int row = q * 32 + lane;             // agent row inside the tile == TMEM lane

  // ---- one-time setup: weights -> smem, TMEM, barrier
  {
    const int4* src = reinterpret_cast<const int4*>(a.pk);
    int4* dst = reinterpret_cast<int4*>(smem);
    for (uint32_t i = tid; i < TL::total_bytes / 16; i += TC_THREADS) dst[i] = src[i];
  }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
  const uint32_t t_acc = tmem + lane_sel + C_ACC, t_act0 = tmem + lane_sel + C_ACT0, t_h = tmem + lane_sel + C_H,
                 t_act1 = tmem + lane_sel + C_ACT1;
  const uint32_t sbase = smem_u32(smem);
  uint32_t phase = 0;
  bool alive = true;

  // issue one layer: D[128 x N] = A(tmem cols a_col.., K) * W^T, then commit to the barrier
  auto issue_layer = [&](uint32_t a_col, uint32_t w_off, int K, int N) {
    if (tid == 0) {
      tc_fence_after();
      const uint32_t idesc = make_idesc_bf16(TC_M, N, false, false, false, /*half_ops=*/true);
      const uint32_t lbo = tc_lbo(N);
#pragma unroll 1
      for (int ks = 0; ks < K / 16; ++ks) {
        const uint64_t bdesc = make_smem_desc(sbase + w_off + (uint32_t)ks * 2u * lbo, lbo, TC_SBO, SWZ_NONE);
        mma_ts(tmem + C_ACC, tmem + a_col + (uint32_t)ks * 8u, bdesc, idesc, ks > 0 ? 1u : 0u);
      }
      mma_commit(&bar);
    }
  };
  // make this thread's TMEM writes / reads visible-ordered, then let thread 0 issue, then wait for completion
  auto run_layer = [&](uint32_t a_col, uint32_t w_off, int K, int N) {
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    issue_layer(a_col, w_off, K, N);
    __syncwarp();      // lanes 1..31 of the issuing warp must not enter the (sleeping) try_wait before lane 0 has issued
    if (alive && !mbar_wait(&bar, phase)) { alive = false; *a.status = 1; }
    phase ^= 1;
    __syncwarp();
    tc_fence_after();
  };

  // hidden-layer epilogue for this thread's 64 columns: out = act(acc + bias [+ residual]) -> bf16 -> TMEM
  auto epilogue_hidden = [&](const float* bias, const float* bias2, float s2, const float* bias3, float s3, uint32_t t_out,
                             bool residual, uint32_t t_res, bool relu_only) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int n0 = hf * 64 + c * 32;
      uint32_t r[32];
      tmem_ld32(t_acc + (uint32_t)n0, r);
      uint32_t zr[16];
      if (residual) tmem_ld16(t_res + (uint32_t)(n0 / 2), zr);
      tmem_ld_wait();
      uint32_t o[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float x0 = __uint_as_float(r[2 * j]) + bias[n0 + 2 * j];
        float x1 = __uint_as_float(r[2 * j + 1]) + bias[n0 + 2 * j + 1];
        if (bias2 != nullptr) {
          x0 += s2 * bias2[n0 + 2 * j] + s3 * bias3[n0 + 2 * j];
          x1 += s2 * bias2[n0 + 2 * j + 1] + s3 * bias3[n0 + 2 * j + 1];
        }
        if (residual) {
          x0 += stc::un_lo<true>(zr[j]);
          x1 += stc::un_hi<true>(zr[j]);
        }
        if (relu_only) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
        else { x0 = tc_act<ACT>(x0); x1 = tc_act<ACT>(x1); }
        o[j] = stc::pack2<true>(x0, x1);
      }
      tmem_st16(t_out + (uint32_t)(n0 / 2), o);
    }
  };

  // one drift evaluation: stage input already in ACT0 (p,v) -> this thread's acceleration slice (PD dims)
  auto drift = [&](float ts, float (&acc_out)[PD]) {
    float sn, cs;
    time_features(ts, a.period, sn, cs);
    run_layer(C_ACT0, TL::off_w1, TL::K1, HID);
    epilogue_hidden(sB1, sWsin, sn, sWcos, cs, t_act0, false, 0, true);
#pragma unroll 1
    for (int r = 0; r < NRES; ++r) {
      run_layer(C_ACT0, TL::off_wa(r), HID, HID);
      epilogue_hidden(sBres + (2 * r) * HID, nullptr, 0.f, nullptr, 0.f, t_act1, false, 0, false);
      run_layer(C_ACT1, TL::off_wb(r), HID, HID);
      epilogue_hidden(sBres + (2 * r + 1) * HID, nullptr, 0.f, nullptr, 0.f, t_act0, true, t_act0, false);
    }
    run_layer(C_ACT0, TL::off_wo, HID, P);
    {
      uint32_t r[32];
      tmem_ld32(t_acc + (uint32_t)(hf * PD), r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < PD; ++j) acc_out[j] = __uint_as_float(r[j]) + sBO[hf * PD + j];
    }
  };

  auto write_stage_input = [&](const float (&pin)[PD], const float (&vin)[PD]) {
    uint32_t o[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] = stc::pack2<true>(pin[2 * j], pin[2 * j + 1]);
    tmem_st16(t_act0 + (uint32_t)(hf * PD / 2), o);
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] = stc::pack2<true>(vin[2 * j], vin[2 * j + 1]);
    tmem_st16(t_act0 + (uint32_t)(P / 2 + hf * PD / 2), o);
  };

#pragma unroll 1
  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    const int64_t g = (int64_t)tile * TC_M + row;
    const bool valid = g < a.B;
    float p0[PD], v0[PD];
    {
      const float4* yr = reinterpret_cast<const float4*>(a.y0 + (valid ? g : 0) * D);
#pragma unroll
      for (int j = 0; j < PD / 4; ++j) {
        const float4 pv = valid ? yr[(hf * PD) / 4 + j] : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 vv = valid ? yr[(P + hf * PD) / 4 + j] : make_float4(0.f, 0.f, 0.f, 0.f);
        p0[4 * j] = pv.x; p0[4 * j + 1] = pv.y; p0[4 * j + 2] = pv.z; p0[4 * j + 3] = pv.w;
        v0[4 * j] = vv.x; v0[4 * j + 1] = vv.y; v0[4 * j + 2] = vv.z; v0[4 * j + 3] = vv.w;
      }
      // static context h: 16 dims per thread-half -> 8 packed columns of HCTX; also row 0 of the trajectory
      float hh[H / 2];
#pragma unroll
      for (int j = 0; j < H / 8; ++j) {
        const float4 x = valid ? yr[(2 * P + hf * (H / 2)) / 4 + j] : make_float4(0.f, 0.f, 0.f, 0.f);
        hh[4 * j] = x.x; hh[4 * j + 1] = x.y; hh[4 * j + 2] = x.z; hh[4 * j + 3] = x.w;
      }
      uint32_t o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = stc::pack2<true>(hh[2 * j], hh[2 * j + 1]);
      tmem_st8(t_h + (uint32_t)(hf * (H / 4)), o);
    }

    auto store_row = [&](int trow) {
      if (!valid) return;
      float4* dst = reinterpret_cast<float4*>(a.y_path + ((size_t)trow * a.B + g) * D);
      const float4* src = reinterpret_cast<const float4*>(a.y0 + g * D);
#pragma unroll
      for (int j = 0; j < PD / 4; ++j) {
        dst[(hf * PD) / 4 + j] = make_float4(p0[4 * j], p0[4 * j + 1], p0[4 * j + 2], p0[4 * j + 3]);
        dst[(P + hf * PD) / 4 + j] = make_float4(v0[4 * j], v0[4 * j + 1], v0[4 * j + 2], v0[4 * j + 3]);
      }
#pragma unroll
      for (int j = 0; j < H / 8; ++j) dst[(2 * P + hf * (H / 2)) / 4 + j] = src[(2 * P + hf * (H / 2)) / 4 + j];
    };
    store_row(0);

#pragma unroll 1
    for (int step = 0; step + 1 < a.T; ++step) {
      const float t0 = a.t[step], t1 = a.t[step + 1];
      const float dt = t1 - t0;
      const float third = 0.333333343267440796f, two_thirds = 0.666666686534881592f;
      float a1[PD], a2[PD], a3[PD], pin[PD], vin[PD];

      write_stage_input(p0, v0);
      drift(t0, a1);
#pragma unroll
      for (int j = 0; j < PD; ++j) {
        pin[j] = p0[j] + dt * third * v0[j];
        vin[j] = v0[j] + dt * third * a1[j];
      }
      write_stage_input(pin, vin);
      drift(t0 + dt * third, a2);
#pragma unroll
      for (int j = 0; j < PD; ++j) {
        pin[j] = p0[j] + dt * (two_thirds * v0[j] + dt * third * a1[j]);
        vin[j] = v0[j] + dt * (a2[j] - third * a1[j]);
      }
      write_stage_input(pin, vin);
      drift(t0 + dt * two_thirds, a3);
#pragma unroll
      for (int j = 0; j < PD; ++j) {
        pin[j] = p0[j] + dt * (v0[j] + dt * (a2[j] - two_thirds * a1[j]));
        vin[j] = v0[j] + dt * (a1[j] - a2[j] + a3[j]);
        const float fv = a1[j] + 3.0f * (a2[j] + a3[j]);
        const float gp = a1[j] + 2.0f * a2[j] + a3[j];
        a1[j] = fv;      // a1 <- k1 + 3(k2 + k3) (v part);  a2 <- a1 + 2 a2 + a3 (p part)
        a2[j] = gp;
      }
      write_stage_input(pin, vin);
      drift(t1, a3);     // a3 <- a4
#pragma unroll
      for (int j = 0; j < PD; ++j) {
        p0[j] = p0[j] + dt * v0[j] + dt * dt * 0.125f * a2[j];
        v0[j] = v0[j] + dt * 0.125f * (a1[j] + a3[j]);
      }
      store_row(step + 1);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// ---- host side -----------------------------------------------------------------------------------------
size_t rk4_tc_workspace(const ab200_drift_desc* d) {
  if (d->pos_dim == 64 && d->ctx_dim == 32 && d->hidden == 128 && d->n_res == 2)
    return align_up(TcLayout<64, 32, 128, 2>::total_bytes, 256) + 256;
  return 0;
}

int rk4_forward_tc(const ab200_drift_desc* d, const float* w_flat, const float* y0, const float* t_dev, int64_t B, int T,
                   float* y_path, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!(d->pos_dim == 64 && d->ctx_dim == 32 && d->hidden == 128 && d->n_res == 2 && d->potential == 0)) return AB200_ERR_UNSUPPORTED;
  using TL = TcLayout<64, 32, 128, 2>;
  const size_t need = rk4_tc_workspace(d);
  if (ws_bytes < need) return AB200_ERR_WORKSPACE;
  uint8_t* pk = (uint8_t*)ws;
  int* status = (int*)(pk + align_up(TL::total_bytes, 256));
  cudaError_t e = cudaMemsetAsync(status, 0, sizeof(int), st);
  if (e != cudaSuccess) { set_cuda_error(e); return AB200_ERR_CUDA; }
  pack_tc_kernel<64, 32, 128, 2><<<96, 256, 0, st>>>(w_flat, pk);
  int rc = check_launch();
  if (rc) return rc;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int ntiles = (int)((B + TC_M - 1) / TC_M);
  const int grid = ntiles < sms ? ntiles : sms;
  TcArgs args{pk, y0, t_dev, y_path, B, T, ntiles, d->time_period, status};
  const size_t smem = TL::total_bytes;
  if (d->res_act == 0) {
    auto kern = rk4_tc_kernel<64, 32, 128, 2, 0>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_cuda_error(e); return AB200_ERR_CUDA; }
    kern<<<grid, TC_THREADS, smem, st>>>(args);
  } else {
    auto kern = rk4_tc_kernel<64, 32, 128, 2, 1>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_cuda_error(e); return AB200_ERR_CUDA; }
    kern<<<grid, TC_THREADS, smem, st>>>(args);
  }
  return check_launch();
}

}  // namespace ab200
