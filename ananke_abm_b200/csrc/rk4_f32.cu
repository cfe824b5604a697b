// Strict-fp32 fused RK4 (3/8 rule) trajectory kernel for the second-order residual-MLP drift.
//
// One CTA owns TM=64 agents for the WHOLE trajectory: agents are independent and share the time grid
// (mode_sep/architecture/model.py:184-191), so the state never leaves the SM between steps.  HBM traffic
// is the algorithmic minimum: y0 read once, one trajectory row written per step.  The six GEMMs of one
// drift evaluation run on the fp32 FFMA pipe with fp32 accumulation (this is the 1e-5 parity path; the
// tensor-core path lives in rk4_tc.cu).  Weights are streamed k-chunk by k-chunk from L2 with cp.async
// double buffering; activations live transposed ([feature][agent]) in shared memory; the Runge-Kutta
// stage algebra lives in registers of the thread that owns the (agent, column) element.
//
// Thread mapping of every GEMM (TM x N tile, 256 threads): tm = tid & 15 owns agents 4tm..4tm+3,
// tn = tid >> 4 owns columns tn*CN .. tn*CN+CN-1 with CN = N/16.
#include "gemm_f32.cuh"

namespace ab200 {

constexpr int TM = 64;    // agents per CTA
constexpr int XS = TM;    // row stride of transposed activation buffers

struct Rk4Args {
  const float* pk;        // packed weights (PackLayout)
  const float* y0;        // [B][D]
  const float* t;         // [T] device
  float* y_path;          // [T][B][D]   (MODE 0)   or out [B][D] (MODE 1: single drift evaluation at t_eval)
  int64_t B;
  int T;
  float t_eval;
  float period;
  int pot_a, pot_b;
  float pot_strength;
};

// MODE 0: whole RK4 trajectory.  MODE 1: one drift evaluation f(t_eval, y0) -> y_path.
template <int P, int H, int HID, int NRES, int ACT, int POT, int MODE>
__global__ void __launch_bounds__(NT, 1) rk4_f32_kernel(Rk4Args a) {
  constexpr int D = 2 * P + H;
  constexpr int CP = P / 16;          // output columns per thread in the P-wide layers
  constexpr int CH = HID / 16;        // output columns per thread in the hidden layers
  constexpr int XR = (2 * P > HID) ? 2 * P : HID;   // rows of the shared X/U buffer
  static_assert(P % 16 == 0 && HID % 16 == 0 && H % 4 == 0, "shape");
  const PackLayout L{P, H, HID, NRES};

  extern __shared__ __align__(16) float smem[];
  float* sXU = smem;                       // [XR][XS]  stage input (p,v) / residual-inner activations
  float* sZ = sXU + XR * XS;               // [HID][XS] hidden state
  float* sCH = sZ + HID * XS;              // [HID][XS] per-agent bias: b_in + W_h h
  float* sHt = sCH + HID * XS;             // [H][XS]   h transposed
  float* sW = sHt + H * XS;                // [2][KC][max N]
  float* sPot = sW + 2 * KC * ((HID > P) ? HID : P);   // [2][XS]  p[pot_a], p[pot_b] of the stage input

  const int tid = threadIdx.x, tm = tid & 15, tn = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * TM;       // first agent of this CTA

  // ---- load h (transposed) --------------------------------------------------------------------
  for (int i = tid; i < TM * H; i += NT) {
    const int m = i / H, j = i % H;
    const int64_t g = m0 + m;
    sHt[j * XS + m] = (g < a.B) ? a.y0[g * D + 2 * P + j] : 0.0f;
  }
  // ---- this thread's slice of the state: agents 4tm..4tm+3, columns tn*CP.. ----------------------
  float p0[4][CP], v0[4][CP];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t g = m0 + 4 * tm + i;
#pragma unroll
    for (int j = 0; j < CP; ++j) {
      p0[i][j] = (g < a.B) ? a.y0[g * D + tn * CP + j] : 0.0f;
      v0[i][j] = (g < a.B) ? a.y0[g * D + P + tn * CP + j] : 0.0f;
    }
  }
  __syncthreads();

  // ---- per-agent bias CH = b_in + W_h^T h (h is constant along the trajectory: dh/dt = 0) ------
  {
    float acc[4][CH];
    gemm_tile<TM, XS, H, HID>(a.pk + L.off_WH(), sHt, sW, acc);
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int n = tn * CH + j;
      const float b = a.pk[L.off_bin() + n];
      *reinterpret_cast<float4*>(&sCH[n * XS + 4 * tm]) =
          make_float4(acc[0][j] + b, acc[1][j] + b, acc[2][j] + b, acc[3][j] + b);
    }
  }

  auto write_stage_input = [&](const float (&pin)[4][CP], const float (&vin)[4][CP]) {
#pragma unroll
    for (int j = 0; j < CP; ++j) {
      const int n = tn * CP + j;
      *reinterpret_cast<float4*>(&sXU[n * XS + 4 * tm]) = make_float4(pin[0][j], pin[1][j], pin[2][j], pin[3][j]);
      *reinterpret_cast<float4*>(&sXU[(P + n) * XS + 4 * tm]) = make_float4(vin[0][j], vin[1][j], vin[2][j], vin[3][j]);
      if (POT) {
        if (n == a.pot_a) *reinterpret_cast<float4*>(&sPot[4 * tm]) = make_float4(pin[0][j], pin[1][j], pin[2][j], pin[3][j]);
        if (n == a.pot_b) *reinterpret_cast<float4*>(&sPot[XS + 4 * tm]) = make_float4(pin[0][j], pin[1][j], pin[2][j], pin[3][j]);
      }
    }
  };

  // one drift evaluation: reads the stage input from sXU, returns this thread's acceleration slice
  auto drift = [&](float ts, float (&acc_out)[4][CP]) {
    float sn, cs;
    time_features(ts, a.period, sn, cs);
    {
      float acc[4][CH];
      gemm_tile<TM, XS, 2 * P, HID>(a.pk + L.off_W0(), sXU, sW, acc);
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        const int n = tn * CH + j;
        const float tb = fmaf(sn, a.pk[L.off_wsin() + n], cs * a.pk[L.off_wcos() + n]);
        const float4 ch = *reinterpret_cast<const float4*>(&sCH[n * XS + 4 * tm]);
        float4 z;
        z.x = fmaxf(acc[0][j] + (ch.x + tb), 0.0f);
        z.y = fmaxf(acc[1][j] + (ch.y + tb), 0.0f);
        z.z = fmaxf(acc[2][j] + (ch.z + tb), 0.0f);
        z.w = fmaxf(acc[3][j] + (ch.w + tb), 0.0f);
        *reinterpret_cast<float4*>(&sZ[n * XS + 4 * tm]) = z;
      }
    }
    __syncthreads();
#pragma unroll 1
    for (int r = 0; r < NRES; ++r) {
      {
        float acc[4][CH];
        gemm_tile<TM, XS, HID, HID>(a.pk + L.off_WA(r), sZ, sW, acc);
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          const int n = tn * CH + j;
          const float b = a.pk[L.off_bA(r) + n];
          float4 u;
          u.x = act_fn<ACT>(acc[0][j] + b); u.y = act_fn<ACT>(acc[1][j] + b);
          u.z = act_fn<ACT>(acc[2][j] + b); u.w = act_fn<ACT>(acc[3][j] + b);
          *reinterpret_cast<float4*>(&sXU[n * XS + 4 * tm]) = u;
        }
      }
      __syncthreads();
      {
        float acc[4][CH];
        gemm_tile<TM, XS, HID, HID>(a.pk + L.off_WB(r), sXU, sW, acc);
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          const int n = tn * CH + j;
          const float b = a.pk[L.off_bB(r) + n];
          float4 z = *reinterpret_cast<const float4*>(&sZ[n * XS + 4 * tm]);
          z.x = act_fn<ACT>(z.x + (acc[0][j] + b)); z.y = act_fn<ACT>(z.y + (acc[1][j] + b));
          z.z = act_fn<ACT>(z.z + (acc[2][j] + b)); z.w = act_fn<ACT>(z.w + (acc[3][j] + b));
          *reinterpret_cast<float4*>(&sZ[n * XS + 4 * tm]) = z;
        }
      }
      __syncthreads();
    }
    gemm_tile<TM, XS, HID, P>(a.pk + L.off_WO(), sZ, sW, acc_out);
#pragma unroll
    for (int j = 0; j < CP; ++j) {
      const int n = tn * CP + j;
      const float b = a.pk[L.off_bO() + n];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc_out[i][j] += b;
      if (POT) {
        // -d/dp of sum (sig(p_a) + sig(p_b) - 1)^2   (latent_ode/architecture/model.py:56-74,93-95)
        if (n == a.pot_a || n == a.pot_b) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float sa = 1.0f / (1.0f + expf(-sPot[4 * tm + i]));
            const float sb = 1.0f / (1.0f + expf(-sPot[XS + 4 * tm + i]));
            const float r2 = 2.0f * (sa + sb - 1.0f);
            const float s = (n == a.pot_a) ? sa : sb;
            acc_out[i][j] += a.pot_strength * (-r2 * s * (1.0f - s));
          }
        }
      }
    }
  };

  auto store_row = [&](float* row_base /* [B][D] */, const float (&pp)[4][CP], const float (&vv)[4][CP], bool with_h) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t g = m0 + 4 * tm + i;
      if (g < a.B) {
#pragma unroll
        for (int j = 0; j < CP; ++j) {
          row_base[g * D + tn * CP + j] = pp[i][j];
          row_base[g * D + P + tn * CP + j] = vv[i][j];
        }
      }
    }
    if (with_h) {
      for (int i = tid; i < TM * H; i += NT) {
        const int m = i / H, j = i % H;
        const int64_t g = m0 + m;
        if (g < a.B) row_base[g * D + 2 * P + j] = sHt[j * XS + m];
      }
    }
  };

  if (MODE == 1) {
    write_stage_input(p0, v0);
    __syncthreads();
    float acc[4][CP];
    drift(a.t_eval, acc);
    // f = [v, accel, 0]
    float zero[4][CP];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < CP; ++j) zero[i][j] = 0.0f;
    (void)zero;
    store_row(a.y_path, v0, acc, false);
    for (int i = tid; i < TM * H; i += NT) {
      const int m = i / H, j = i % H;
      const int64_t g = m0 + m;
      if (g < a.B) a.y_path[g * D + 2 * P + j] = 0.0f;
    }
    return;
  }

  store_row(a.y_path, p0, v0, true);          // row 0 = y0
  const float third = 0.333333343267440796f;   // float(1/3), float(2/3) as torch casts the Python scalars
  const float two_thirds = 0.666666686534881592f;

#pragma unroll 1
  for (int step = 0; step + 1 < a.T; ++step) {
    const float t0 = a.t[step], t1 = a.t[step + 1];
    const float dt = fsub(t1, t0);
    float a1[4][CP], a2[4][CP], a3[4][CP], a4[4][CP];
    float pin[4][CP], vin[4][CP];

    // stage 1: k1 = f(t0, y0)
    write_stage_input(p0, v0);
    __syncthreads();
    drift(t0, a1);
    // stage 2: k2 = f(t0 + dt/3, y0 + dt*k1/3)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < CP; ++j) {
        pin[i][j] = fadd(p0[i][j], fmul(fmul(dt, v0[i][j]), third));
        vin[i][j] = fadd(v0[i][j], fmul(fmul(dt, a1[i][j]), third));
      }
    write_stage_input(pin, vin);
    __syncthreads();
    drift(fadd(t0, fmul(dt, third)), a2);
    // stage 3: k3 = f(t0 + 2dt/3, y0 + dt*(k2 - k1/3));  k2 = (v-part of stage-2 input, a2)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < CP; ++j) {
        const float k2p = fadd(v0[i][j], fmul(fmul(dt, a1[i][j]), third));
        pin[i][j] = fadd(p0[i][j], fmul(dt, fsub(k2p, fmul(v0[i][j], third))));
        vin[i][j] = fadd(v0[i][j], fmul(dt, fsub(a2[i][j], fmul(a1[i][j], third))));
      }
    write_stage_input(pin, vin);
    __syncthreads();
    drift(fadd(t0, fmul(dt, two_thirds)), a3);
    // stage 4: k4 = f(t1, y0 + dt*(k1 - k2 + k3))
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < CP; ++j) {
        const float k2p = fadd(v0[i][j], fmul(fmul(dt, a1[i][j]), third));
        const float k3p = fadd(v0[i][j], fmul(dt, fsub(a2[i][j], fmul(a1[i][j], third))));
        pin[i][j] = fadd(p0[i][j], fmul(dt, fadd(fsub(v0[i][j], k2p), k3p)));
        vin[i][j] = fadd(v0[i][j], fmul(dt, fadd(fsub(a1[i][j], a2[i][j]), a3[i][j])));
      }
    write_stage_input(pin, vin);
    __syncthreads();
    drift(t1, a4);
    // y1 = y0 + (k1 + 3*(k2 + k3) + k4) * dt * 0.125
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < CP; ++j) {
        const float k2p = fadd(v0[i][j], fmul(fmul(dt, a1[i][j]), third));
        const float k3p = fadd(v0[i][j], fmul(dt, fsub(a2[i][j], fmul(a1[i][j], third))));
        const float k4p = fadd(v0[i][j], fmul(dt, fadd(fsub(a1[i][j], a2[i][j]), a3[i][j])));
        const float dp = fmul(fmul(fadd(fadd(v0[i][j], fmul(3.0f, fadd(k2p, k3p))), k4p), dt), 0.125f);
        const float dv = fmul(fmul(fadd(fadd(a1[i][j], fmul(3.0f, fadd(a2[i][j], a3[i][j]))), a4[i][j]), dt), 0.125f);
        p0[i][j] = fadd(p0[i][j], dp);
        v0[i][j] = fadd(v0[i][j], dv);
      }
    store_row(a.y_path + (size_t)(step + 1) * a.B * D, p0, v0, true);
  }
}

// ---- weight packing -----------------------------------------------------------------------------------
__global__ void pack_drift_kernel(const float* __restrict__ w, float* __restrict__ pk, int P, int H, int HID, int NRES) {
  const FlatLayout F{P, H, HID, NRES};
  const PackLayout L{P, H, HID, NRES};
  const int IN = 2 * P + H + 2;
  const int64_t total = L.total();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float v;
    if (i < L.off_WH()) {                       // W0[k][n] = w_in[n][k], k < 2P
      const int k = (int)(i / HID), n = (int)(i % HID);
      v = w[F.off_win() + (int64_t)n * IN + k];
    } else if (i < L.off_wsin()) {              // WH[j][n] = w_in[n][2P+j]
      const int64_t r = i - L.off_WH();
      const int j = (int)(r / HID), n = (int)(r % HID);
      v = w[F.off_win() + (int64_t)n * IN + 2 * P + j];
    } else if (i < L.off_wcos()) {
      const int n = (int)(i - L.off_wsin());
      v = w[F.off_win() + (int64_t)n * IN + 2 * P + H];
    } else if (i < L.off_bin()) {
      const int n = (int)(i - L.off_wcos());
      v = w[F.off_win() + (int64_t)n * IN + 2 * P + H + 1];
    } else if (i < L.off_res(0)) {
      v = w[F.off_bin() + (i - L.off_bin())];
    } else if (i < L.off_WO()) {
      const int64_t per = 2 * (int64_t)HID * HID + 2 * HID;
      const int r = (int)((i - L.off_res(0)) / per);
      const int64_t q = (i - L.off_res(0)) % per;
      if (q < (int64_t)HID * HID) {             // WA[k][n] = wa[n][k]
        const int k = (int)(q / HID), n = (int)(q % HID);
        v = w[F.off_wa(r) + (int64_t)n * HID + k];
      } else if (q < (int64_t)HID * HID + HID) {
        v = w[F.off_ba(r) + (q - (int64_t)HID * HID)];
      } else if (q < 2 * (int64_t)HID * HID + HID) {
        const int64_t q2 = q - (int64_t)HID * HID - HID;
        const int k = (int)(q2 / HID), n = (int)(q2 % HID);
        v = w[F.off_wb(r) + (int64_t)n * HID + k];
      } else {
        v = w[F.off_bb(r) + (q - 2 * (int64_t)HID * HID - HID)];
      }
    } else if (i < L.off_bO()) {                // WO[k][n] = w_out[n][k]
      const int64_t r = i - L.off_WO();
      const int k = (int)(r / P), n = (int)(r % P);
      v = w[F.off_wout() + (int64_t)n * HID + k];
    } else {
      v = w[F.off_bout() + (i - L.off_bO())];
    }
    pk[i] = v;
  }
}

template <int P, int H, int HID, int NRES, int ACT, int POT, int MODE>
static int launch_f32(const ab200_drift_desc* d, const Rk4Args& args, cudaStream_t st) {
  constexpr int XR = (2 * P > HID) ? 2 * P : HID;
  constexpr int WN = (HID > P) ? HID : P;
  const size_t smem = sizeof(float) * ((size_t)XR * XS + 2 * (size_t)HID * XS + (size_t)H * XS + 2 * KC * WN + 2 * XS);
  auto kern = rk4_f32_kernel<P, H, HID, NRES, ACT, POT, MODE>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_cuda_error(e); return AB200_ERR_CUDA; }
  const int64_t grid = (args.B + TM - 1) / TM;
  kern<<<(unsigned)grid, NT, smem, st>>>(args);
  return check_launch();
}

// dispatch on the two shapes the reference instantiates
template <int MODE>
static int dispatch_f32(const ab200_drift_desc* d, const Rk4Args& args, cudaStream_t st) {
  if (d->pos_dim == 64 && d->ctx_dim == 32 && d->hidden == 128 && d->n_res == 2 && d->res_act == 0 && d->potential == 0)
    return launch_f32<64, 32, 128, 2, 0, 0, MODE>(d, args, st);
  if (d->pos_dim == 16 && d->ctx_dim == 32 && d->hidden == 128 && d->n_res == 2 && d->res_act == 1 && d->potential == 1)
    return launch_f32<16, 32, 128, 2, 1, 1, MODE>(d, args, st);
  if (d->pos_dim == 16 && d->ctx_dim == 32 && d->hidden == 128 && d->n_res == 2 && d->res_act == 1 && d->potential == 0)
    return launch_f32<16, 32, 128, 2, 1, 0, MODE>(d, args, st);
  return AB200_ERR_UNSUPPORTED;
}

int pack_drift(const ab200_drift_desc* d, const float* w_flat, float* packed, cudaStream_t st) {
  const PackLayout L{d->pos_dim, d->ctx_dim, d->hidden, d->n_res};
  const int64_t total = L.total();
  const int threads = 256;
  const int blocks = (int)((total + threads - 1) / threads);
  pack_drift_kernel<<<blocks, threads, 0, st>>>(w_flat, packed, d->pos_dim, d->ctx_dim, d->hidden, d->n_res);
  return check_launch();
}

int rk4_forward_f32(const ab200_drift_desc* d, const float* packed, const float* y0, const float* t_dev, int64_t B,
                    int T, float* y_path, cudaStream_t st) {
  Rk4Args a{packed, y0, t_dev, y_path, B, T, 0.0f, d->time_period, d->pot_idx_a, d->pot_idx_b, d->pot_strength};
  return dispatch_f32<0>(d, a, st);
}

int drift_eval_f32(const ab200_drift_desc* d, const float* packed, float t, const float* y, int64_t B, float* out,
                   cudaStream_t st) {
  Rk4Args a{packed, y, nullptr, out, B, 1, t, d->time_period, d->pot_idx_a, d->pot_idx_b, d->pot_strength};
  return dispatch_f32<1>(d, a, st);
}

}  // namespace ab200
