// Strict-fp32 discrete adjoint of the fused RK4 (3/8 rule) trajectory: the gradient that reverse-mode
// autograd through every solver op produces in the reference (mode_sep/train/train.py:162), computed
// in one persistent kernel.
//
// A CTA owns TMB=32 agents at a time and sweeps their trajectory backwards.  Per step it re-evaluates
// stages 1..3 from the saved y_path row (to recover the stage inputs), then for stages 4,3,2,1 runs the
// drift net forward with activations kept in shared memory and back-propagates through it:
//   dgrad  GEMMs use torch's own [out][in] weight storage as the k-major operand (no transposed copy),
//   wgrad  outer products are reduced over the tile's agents in registers and accumulated into a
//          CTA-PRIVATE gradient buffer in global memory with plain vector read-modify-writes (no atomics;
//          148 x 372 KiB stays L2 resident), summed across CTAs by a small second kernel.
// HBM traffic per agent-step: y_path row (p,v) read + grad_y_path row read = the algorithmic 2x(2P)x4 B.
#include "gemm_f32.cuh"

namespace ab200 {

constexpr int TMB = 32;   // agents per tile
constexpr int XSB = 36;   // padded row stride (floats): conflict-light for both GEMM and outer-product reads

// dst[k][n] (+)= sum_m xT[k][m] * dT[n][m] over the tile's agents.  Thread (tk = tid>>4, tn = tid&15)
// owns k in {tk + 16 i}, n in {tn + 16 j}; the private global tile is updated with plain RMW.
template <int K, int N>
__device__ __forceinline__ void wgrad_accumulate(const float* xT, const float* dT, float* __restrict__ gdst) {
  constexpr int RK = (K + 15) / 16, RN = (N + 15) / 16;
  const int tn = threadIdx.x & 15, tk = threadIdx.x >> 4;
  float acc[RK][RN];
#pragma unroll
  for (int i = 0; i < RK; ++i)
#pragma unroll
    for (int j = 0; j < RN; ++j) acc[i][j] = 0.0f;
#pragma unroll 2
  for (int m = 0; m < TMB; m += 4) {
    float4 xv[RK], dv[RN];
#pragma unroll
    for (int i = 0; i < RK; ++i) {
      const int k = tk + 16 * i;
      xv[i] = (k < K) ? *reinterpret_cast<const float4*>(xT + k * XSB + m) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < RN; ++j) {
      const int n = tn + 16 * j;
      dv[j] = (n < N) ? *reinterpret_cast<const float4*>(dT + n * XSB + m) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < RK; ++i)
#pragma unroll
      for (int j = 0; j < RN; ++j) {
        acc[i][j] = fmaf(xv[i].x, dv[j].x, acc[i][j]);
        acc[i][j] = fmaf(xv[i].y, dv[j].y, acc[i][j]);
        acc[i][j] = fmaf(xv[i].z, dv[j].z, acc[i][j]);
        acc[i][j] = fmaf(xv[i].w, dv[j].w, acc[i][j]);
      }
  }
#pragma unroll
  for (int i = 0; i < RK; ++i) {
    const int k = tk + 16 * i;
    if (k < K) {
#pragma unroll
      for (int j = 0; j < RN; ++j) {
        const int n = tn + 16 * j;
        if (n < N) gdst[(size_t)k * N + n] += acc[i][j];
      }
    }
  }
}

// dst[n] += scale * sum_m dT[n][m]
template <int N>
__device__ __forceinline__ void bias_accumulate(const float* dT, float* __restrict__ gdst, float scale) {
  for (int n = threadIdx.x; n < N; n += NT) {
    float s = 0.0f;
#pragma unroll
    for (int m = 0; m < TMB; m += 4) {
      const float4 v = *reinterpret_cast<const float4*>(dT + n * XSB + m);
      s += (v.x + v.y) + (v.z + v.w);
    }
    gdst[n] += scale * s;
  }
}

struct BwdArgs {
  const float* pk;      // packed forward weights (PackLayout, [K][N])
  const float* wflat;   // torch layout ([out][in]) -- the dgrad operand of the hidden/out layers
  const float* w0d;     // [HID][2P]  (p,v) columns of w_in, contiguous
  const float* whd;     // [HID][H]   h columns of w_in, contiguous
  const float* t;       // [T]
  const float* y_path;  // [T][B][D]
  const float* gy;      // [T][B][D]  dL/dy_path
  float* gy0;           // [B][D]
  float* priv;          // [gridDim.x][PackLayout.total()] private weight-gradient accumulators (zeroed)
  int64_t B;
  int T;
  int ntiles;
  float period;
  int pot_a, pot_b;
  float pot_strength;
  float t_eval;         // MODE 1: the time of the single evaluation
};

// MODE 0: discrete adjoint of the whole rk4 trajectory (above).
// MODE 1: vector-Jacobian product of ONE drift evaluation f(t, y) = [v, net(p, v, h, t), 0]: y_path = y [B][D],
//         gy = upstream gradient g [B][D];  gy0 = J^T g = [J_p^T g_v, g_p + J_v^T g_v, J_h^T g_v]  and the weight
//         gradients of that evaluation -- what autograd needs to differentiate a solver step written in PyTorch ops
//         (the reference's own training path, mode_sep/train/train.py:162, latent_ode/train/train.py:73) and what the
//         continuous adjoint evaluates (latent_ode/architecture/ode_components.py:50).
template <int P, int H, int HID, int NRES, int ACT, int POT, int MODE>
__global__ void __launch_bounds__(NT, 1) rk4_bwd_f32_kernel(BwdArgs a) {
  constexpr int D = 2 * P + H;
  using MapP = TileMap<TMB, P>;
  using MapH = TileMap<TMB, HID>;
  using MapX = TileMap<TMB, 2 * P>;
  using MapC = TileMap<TMB, H>;
  constexpr int CP = MapP::CN, CHd = MapH::CN, CX = MapX::CN, CC = MapC::CN;
  constexpr int GR = (2 * P > HID) ? 2 * P : HID;
  constexpr int WN = (HID > 2 * P) ? HID : 2 * P;
  const PackLayout L{P, H, HID, NRES};
  const FlatLayout F{P, H, HID, NRES};

  extern __shared__ __align__(16) float smem[];
  float* sX = smem;                                  // [2P+H][XSB]  stage input (p,v) and h
  float* sZ = sX + (2 * P + H) * XSB;                // [NRES+1][HID][XSB]
  float* sU = sZ + (NRES + 1) * HID * XSB;           // [NRES][HID][XSB]
  float* sCH = sU + NRES * HID * XSB;                // [HID][XSB]
  float* sDA = sCH + HID * XSB;                      // [HID][XSB]
  float* sDB = sDA + HID * XSB;                      // [GR][XSB]   (also the input-gradient buffer)
  float* sDO = sDB + GR * XSB;                       // [P][XSB]
  float* sGH = sDO + P * XSB;                        // [H][XSB]
  float* sPot = sGH + H * XSB;                       // [2][XSB]
  float* sW = sPot + 2 * XSB;                        // [2][KC][WN]

  const int tid = threadIdx.x;
  const int tm = MapP::tm(), tn = MapP::tn();
  const bool own = MapP::active();                   // this thread owns (agents 4tm.., columns tn*CP..) of p and v
  float* priv = a.priv + (size_t)blockIdx.x * L.total();
  const float third = 0.333333343267440796f, two_thirds = 0.666666686534881592f;

  auto write_stage_input = [&](const float (&pin)[4][CP], const float (&vin)[4][CP]) {
    if (own) {
#pragma unroll
      for (int j = 0; j < CP; ++j) {
        const int n = tn * CP + j;
        *reinterpret_cast<float4*>(&sX[n * XSB + 4 * tm]) = make_float4(pin[0][j], pin[1][j], pin[2][j], pin[3][j]);
        *reinterpret_cast<float4*>(&sX[(P + n) * XSB + 4 * tm]) = make_float4(vin[0][j], vin[1][j], vin[2][j], vin[3][j]);
        if (POT) {
          if (n == a.pot_a) *reinterpret_cast<float4*>(&sPot[4 * tm]) = make_float4(pin[0][j], pin[1][j], pin[2][j], pin[3][j]);
          if (n == a.pot_b) *reinterpret_cast<float4*>(&sPot[XSB + 4 * tm]) = make_float4(pin[0][j], pin[1][j], pin[2][j], pin[3][j]);
        }
      }
    }
  };

  // forward through the drift net keeping every post-activation; returns this thread's acceleration slice
  auto mlp_forward = [&](float ts, float (&aout)[4][CP]) {
    float sn, cs;
    time_features(ts, a.period, sn, cs);
    {
      float acc[4][CHd];
      gemm_tile<TMB, XSB, 2 * P, HID>(a.pk + L.off_W0(), sX, sW, acc);
      const int hm = MapH::tm(), hn = MapH::tn();
#pragma unroll
      for (int j = 0; j < CHd; ++j) {
        const int n = hn * CHd + j;
        const float tb = fmaf(sn, a.pk[L.off_wsin() + n], cs * a.pk[L.off_wcos() + n]);
        const float4 ch = *reinterpret_cast<const float4*>(&sCH[n * XSB + 4 * hm]);
        float4 z;
        z.x = fmaxf(acc[0][j] + (ch.x + tb), 0.0f); z.y = fmaxf(acc[1][j] + (ch.y + tb), 0.0f);
        z.z = fmaxf(acc[2][j] + (ch.z + tb), 0.0f); z.w = fmaxf(acc[3][j] + (ch.w + tb), 0.0f);
        *reinterpret_cast<float4*>(&sZ[n * XSB + 4 * hm]) = z;
      }
    }
    __syncthreads();
#pragma unroll 1
    for (int r = 0; r < NRES; ++r) {
      float* zin = sZ + r * HID * XSB;
      float* zout = sZ + (r + 1) * HID * XSB;
      float* u = sU + r * HID * XSB;
      const int hm = MapH::tm(), hn = MapH::tn();
      {
        float acc[4][CHd];
        gemm_tile<TMB, XSB, HID, HID>(a.pk + L.off_WA(r), zin, sW, acc);
#pragma unroll
        for (int j = 0; j < CHd; ++j) {
          const int n = hn * CHd + j;
          const float b = a.pk[L.off_bA(r) + n];
          *reinterpret_cast<float4*>(&u[n * XSB + 4 * hm]) = make_float4(
              act_fn<ACT>(acc[0][j] + b), act_fn<ACT>(acc[1][j] + b), act_fn<ACT>(acc[2][j] + b), act_fn<ACT>(acc[3][j] + b));
        }
      }
      __syncthreads();
      {
        float acc[4][CHd];
        gemm_tile<TMB, XSB, HID, HID>(a.pk + L.off_WB(r), u, sW, acc);
#pragma unroll
        for (int j = 0; j < CHd; ++j) {
          const int n = hn * CHd + j;
          const float b = a.pk[L.off_bB(r) + n];
          const float4 z = *reinterpret_cast<const float4*>(&zin[n * XSB + 4 * hm]);
          *reinterpret_cast<float4*>(&zout[n * XSB + 4 * hm]) = make_float4(
              act_fn<ACT>(z.x + (acc[0][j] + b)), act_fn<ACT>(z.y + (acc[1][j] + b)),
              act_fn<ACT>(z.z + (acc[2][j] + b)), act_fn<ACT>(z.w + (acc[3][j] + b)));
        }
      }
      __syncthreads();
    }
    gemm_tile<TMB, XSB, HID, P>(a.pk + L.off_WO(), sZ + NRES * HID * XSB, sW, aout);
    if (own) {
#pragma unroll
      for (int j = 0; j < CP; ++j) {
        const int n = tn * CP + j;
        const float b = a.pk[L.off_bO() + n];
#pragma unroll
        for (int i = 0; i < 4; ++i) aout[i][j] += b;
        if (POT) {
          if (n == a.pot_a || n == a.pot_b) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float sa = 1.0f / (1.0f + expf(-sPot[4 * tm + i]));
              const float sb = 1.0f / (1.0f + expf(-sPot[XSB + 4 * tm + i]));
              const float r2 = 2.0f * (sa + sb - 1.0f);
              const float s = (n == a.pot_a) ? sa : sb;
              aout[i][j] += a.pot_strength * (-r2 * s * (1.0f - s));
            }
          }
        }
      }
    }
  };

  // back-propagate `delta` (= dL/d accel, owner layout) through the net whose activations mlp_forward just
  // saved.  Accumulates weight gradients into `priv`, dL/dh into sGH, returns dL/d(p_in), dL/d(v_in).
  auto mlp_backward = [&](float ts, const float (&delta)[4][CP], float (&xp)[4][CP], float (&xv)[4][CP]) {
    float sn, cs;
    time_features(ts, a.period, sn, cs);
    const int hm = MapH::tm(), hn = MapH::tn();
    if (own) {
#pragma unroll
      for (int j = 0; j < CP; ++j)
        *reinterpret_cast<float4*>(&sDO[(tn * CP + j) * XSB + 4 * tm]) =
            make_float4(delta[0][j], delta[1][j], delta[2][j], delta[3][j]);
    }
    __syncthreads();
    // ---- output layer
    wgrad_accumulate<HID, P>(sZ + NRES * HID * XSB, sDO, priv + L.off_WO());
    bias_accumulate<P>(sDO, priv + L.off_bO(), 1.0f);
    {
      float acc[4][CHd];
      gemm_tile<TMB, XSB, P, HID>(a.wflat + F.off_wout(), sDO, sW, acc);
      const float* zl = sZ + NRES * HID * XSB;
#pragma unroll
      for (int j = 0; j < CHd; ++j) {
        const int n = hn * CHd + j;
        const float4 z = *reinterpret_cast<const float4*>(&zl[n * XSB + 4 * hm]);
        float4 g;
        if (NRES > 0) {
          g = make_float4(acc[0][j] * act_grad_from_out<ACT>(z.x), acc[1][j] * act_grad_from_out<ACT>(z.y),
                          acc[2][j] * act_grad_from_out<ACT>(z.z), acc[3][j] * act_grad_from_out<ACT>(z.w));
        } else {
          g = make_float4(acc[0][j] * act_grad_from_out<0>(z.x), acc[1][j] * act_grad_from_out<0>(z.y),
                          acc[2][j] * act_grad_from_out<0>(z.z), acc[3][j] * act_grad_from_out<0>(z.w));
        }
        *reinterpret_cast<float4*>(&sDA[n * XSB + 4 * hm]) = g;
      }
    }
    __syncthreads();
    // ---- residual blocks, last to first.  sDA holds dL/d(pre-activation of the block output).
#pragma unroll 1
    for (int r = NRES - 1; r >= 0; --r) {
      const float* zin = sZ + r * HID * XSB;
      const float* u = sU + r * HID * XSB;
      wgrad_accumulate<HID, HID>(u, sDA, priv + L.off_WB(r));
      bias_accumulate<HID>(sDA, priv + L.off_bB(r), 1.0f);
      {
        float acc[4][CHd];
        gemm_tile<TMB, XSB, HID, HID>(a.wflat + F.off_wb(r), sDA, sW, acc);
#pragma unroll
        for (int j = 0; j < CHd; ++j) {
          const int n = hn * CHd + j;
          const float4 uo = *reinterpret_cast<const float4*>(&u[n * XSB + 4 * hm]);
          *reinterpret_cast<float4*>(&sDB[n * XSB + 4 * hm]) =
              make_float4(acc[0][j] * act_grad_from_out<ACT>(uo.x), acc[1][j] * act_grad_from_out<ACT>(uo.y),
                          acc[2][j] * act_grad_from_out<ACT>(uo.z), acc[3][j] * act_grad_from_out<ACT>(uo.w));
        }
      }
      __syncthreads();
      wgrad_accumulate<HID, HID>(zin, sDB, priv + L.off_WA(r));
      bias_accumulate<HID>(sDB, priv + L.off_bA(r), 1.0f);
      {
        float acc[4][CHd];
        gemm_tile<TMB, XSB, HID, HID>(a.wflat + F.off_wa(r), sDB, sW, acc);
#pragma unroll
        for (int j = 0; j < CHd; ++j) {
          const int n = hn * CHd + j;
          const float4 zi = *reinterpret_cast<const float4*>(&zin[n * XSB + 4 * hm]);
          const float4 sk = *reinterpret_cast<const float4*>(&sDA[n * XSB + 4 * hm]);
          float4 g;
          if (r > 0) {
            g = make_float4((sk.x + acc[0][j]) * act_grad_from_out<ACT>(zi.x), (sk.y + acc[1][j]) * act_grad_from_out<ACT>(zi.y),
                            (sk.z + acc[2][j]) * act_grad_from_out<ACT>(zi.z), (sk.w + acc[3][j]) * act_grad_from_out<ACT>(zi.w));
          } else {   // zin is the ReLU output of the input layer
            g = make_float4((sk.x + acc[0][j]) * act_grad_from_out<0>(zi.x), (sk.y + acc[1][j]) * act_grad_from_out<0>(zi.y),
                            (sk.z + acc[2][j]) * act_grad_from_out<0>(zi.z), (sk.w + acc[3][j]) * act_grad_from_out<0>(zi.w));
          }
          *reinterpret_cast<float4*>(&sDA[n * XSB + 4 * hm]) = g;
        }
      }
      __syncthreads();
    }
    // ---- input layer: sDA = dL/d(pre-activation of layer 0)
    wgrad_accumulate<2 * P, HID>(sX, sDA, priv + L.off_W0());
    wgrad_accumulate<H, HID>(sX + 2 * P * XSB, sDA, priv + L.off_WH());
    bias_accumulate<HID>(sDA, priv + L.off_bin(), 1.0f);
    bias_accumulate<HID>(sDA, priv + L.off_wsin(), sn);
    bias_accumulate<HID>(sDA, priv + L.off_wcos(), cs);
    {
      float acc[4][CX];
      gemm_tile<TMB, XSB, HID, 2 * P>(a.w0d, sDA, sW, acc);
      if (MapX::active()) {
        const int xm = MapX::tm(), xn = MapX::tn();
#pragma unroll
        for (int j = 0; j < CX; ++j)
          *reinterpret_cast<float4*>(&sDB[(xn * CX + j) * XSB + 4 * xm]) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
      }
    }
    {
      float acc[4][CC];
      gemm_tile<TMB, XSB, HID, H>(a.whd, sDA, sW, acc);
      if (MapC::active()) {
        const int cm = MapC::tm(), cn = MapC::tn();
#pragma unroll
        for (int j = 0; j < CC; ++j) {
          float4 g = *reinterpret_cast<const float4*>(&sGH[(cn * CC + j) * XSB + 4 * cm]);
          g.x += acc[0][j]; g.y += acc[1][j]; g.z += acc[2][j]; g.w += acc[3][j];
          *reinterpret_cast<float4*>(&sGH[(cn * CC + j) * XSB + 4 * cm]) = g;
        }
      }
    }
    __syncthreads();
    if (own) {
#pragma unroll
      for (int j = 0; j < CP; ++j) {
        const int n = tn * CP + j;
        const float4 gp = *reinterpret_cast<const float4*>(&sDB[n * XSB + 4 * tm]);
        const float4 gv = *reinterpret_cast<const float4*>(&sDB[(P + n) * XSB + 4 * tm]);
        xp[0][j] = gp.x; xp[1][j] = gp.y; xp[2][j] = gp.z; xp[3][j] = gp.w;
        xv[0][j] = gv.x; xv[1][j] = gv.y; xv[2][j] = gv.z; xv[3][j] = gv.w;
        if (POT) {
          if (n == a.pot_a || n == a.pot_b) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int m = 4 * tm + i;
              const float sa = 1.0f / (1.0f + expf(-sPot[m]));
              const float sb = 1.0f / (1.0f + expf(-sPot[XSB + m]));
              const float da = sa * (1.0f - sa), db = sb * (1.0f - sb), r = sa + sb - 1.0f;
              const float g_a = sDO[a.pot_a * XSB + m], g_b = sDO[a.pot_b * XSB + m];
              // corr_a = -2 r da, corr_b = -2 r db
              float add;
              if (n == a.pot_a) add = g_a * (-2.0f * (da * da + r * da * (1.0f - 2.0f * sa))) + g_b * (-2.0f * da * db);
              else add = g_a * (-2.0f * db * da) + g_b * (-2.0f * (db * db + r * db * (1.0f - 2.0f * sb)));
              xp[i][j] += a.pot_strength * add;
            }
          }
        }
      }
    }
    __syncthreads();
  };

#pragma unroll 1
  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    const int64_t m0 = (int64_t)tile * TMB;
    __syncthreads();
    // h rows of the stage-input buffer, dL/dh accumulator
    for (int i = tid; i < TMB * H; i += NT) {
      const int m = i / H, j = i % H;
      const int64_t g = m0 + m;
      sX[(2 * P + j) * XSB + m] = (g < a.B) ? a.y_path[g * D + 2 * P + j] : 0.0f;
      sGH[j * XSB + m] = 0.0f;
    }
    __syncthreads();
    {
      float acc[4][CHd];
      gemm_tile<TMB, XSB, H, HID>(a.pk + L.off_WH(), sX + 2 * P * XSB, sW, acc);
      const int hm = MapH::tm(), hn = MapH::tn();
#pragma unroll
      for (int j = 0; j < CHd; ++j) {
        const int n = hn * CHd + j;
        const float b = a.pk[L.off_bin() + n];
        *reinterpret_cast<float4*>(&sCH[n * XSB + 4 * hm]) = make_float4(acc[0][j] + b, acc[1][j] + b, acc[2][j] + b, acc[3][j] + b);
      }
    }
    __syncthreads();

    auto load_pv = [&](const float* base, float (&pp)[4][CP], float (&vv)[4][CP]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t g = m0 + 4 * tm + i;
#pragma unroll
        for (int j = 0; j < CP; ++j) {
          const bool ok = own && g < a.B;
          pp[i][j] = ok ? base[g * D + tn * CP + j] : 0.0f;
          vv[i][j] = ok ? base[g * D + P + tn * CP + j] : 0.0f;
        }
      }
    };

    if constexpr (MODE == 1) {
      float p0[4][CP], v0[4][CP], gp[4][CP], gv[4][CP], acc_out[4][CP], xp[4][CP], xv[4][CP];
      load_pv(a.y_path, p0, v0);
      load_pv(a.gy, gp, gv);
      write_stage_input(p0, v0);
      __syncthreads();
      mlp_forward(a.t_eval, acc_out);
      mlp_backward(a.t_eval, gv, xp, xv);
      if (own) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int64_t g = m0 + 4 * tm + i;
          if (g < a.B) {
#pragma unroll
            for (int j = 0; j < CP; ++j) {
              a.gy0[g * D + tn * CP + j] = xp[i][j];
              a.gy0[g * D + P + tn * CP + j] = gp[i][j] + xv[i][j];
            }
          }
        }
      }
      __syncthreads();
      for (int i = tid; i < TMB * H; i += NT) {
        const int m = i / H, j = i % H;
        const int64_t g = m0 + m;
        if (g < a.B) a.gy0[g * D + 2 * P + j] = sGH[j * XSB + m];
      }
      continue;
    }

    float lp[4][CP], lv[4][CP];
    load_pv(a.gy + (size_t)(a.T - 1) * a.B * D, lp, lv);

#pragma unroll 1
    for (int step = a.T - 2; step >= 0; --step) {
      const float t0 = a.t[step], t1 = a.t[step + 1];
      const float dt = fsub(t1, t0);
      const float c8 = dt * 0.125f, c38 = 3.0f * dt * 0.125f, d3 = dt * third;
      float p0[4][CP], v0[4][CP], a1[4][CP], a2[4][CP], a3[4][CP], a4[4][CP];
      float pin[4][CP], vin[4][CP];
      load_pv(a.y_path + (size_t)step * a.B * D, p0, v0);

      // ---- recompute the stage inputs (same arithmetic as the forward kernel)
      write_stage_input(p0, v0);
      __syncthreads();
      mlp_forward(t0, a1);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < CP; ++j) {
          pin[i][j] = fadd(p0[i][j], fmul(fmul(dt, v0[i][j]), third));
          vin[i][j] = fadd(v0[i][j], fmul(fmul(dt, a1[i][j]), third));
        }
      write_stage_input(pin, vin);
      __syncthreads();
      mlp_forward(fadd(t0, fmul(dt, third)), a2);
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
      mlp_forward(fadd(t0, fmul(dt, two_thirds)), a3);

      float g4p[4][CP], g4v[4][CP], g3p[4][CP], g3v[4][CP], g2p[4][CP], g2v[4][CP];
      float del[4][CP], xp[4][CP], xv[4][CP];

      // ---- stage 4: u4 = y0 + dt (k1 - k2 + k3);  dL/dk4 = (dt/8) lambda
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < CP; ++j) {
          const float k2p = fadd(v0[i][j], fmul(fmul(dt, a1[i][j]), third));
          const float k3p = fadd(v0[i][j], fmul(dt, fsub(a2[i][j], fmul(a1[i][j], third))));
          pin[i][j] = fadd(p0[i][j], fmul(dt, fadd(fsub(v0[i][j], k2p), k3p)));
          vin[i][j] = fadd(v0[i][j], fmul(dt, fadd(fsub(a1[i][j], a2[i][j]), a3[i][j])));
          del[i][j] = c8 * lv[i][j];
        }
      write_stage_input(pin, vin);
      __syncthreads();
      mlp_forward(t1, a4);
      mlp_backward(t1, del, xp, xv);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < CP; ++j) { g4p[i][j] = xp[i][j]; g4v[i][j] = c8 * lp[i][j] + xv[i][j]; }

      // ---- stage 3: u3 = y0 + dt (k2 - k1/3);  dL/dk3 = (3dt/8) lambda + dt g4
      float dkp[4][CP];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < CP; ++j) {
          const float k2p = fadd(v0[i][j], fmul(fmul(dt, a1[i][j]), third));
          pin[i][j] = fadd(p0[i][j], fmul(dt, fsub(k2p, fmul(v0[i][j], third))));
          vin[i][j] = fadd(v0[i][j], fmul(dt, fsub(a2[i][j], fmul(a1[i][j], third))));
          del[i][j] = c38 * lv[i][j] + dt * g4v[i][j];
          dkp[i][j] = c38 * lp[i][j] + dt * g4p[i][j];
        }
      write_stage_input(pin, vin);
      __syncthreads();
      mlp_forward(fadd(t0, fmul(dt, two_thirds)), a4);
      mlp_backward(fadd(t0, fmul(dt, two_thirds)), del, xp, xv);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < CP; ++j) { g3p[i][j] = xp[i][j]; g3v[i][j] = dkp[i][j] + xv[i][j]; }

      // ---- stage 2: u2 = y0 + (dt/3) k1;  dL/dk2 = (3dt/8) lambda - dt g4 + dt g3
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < CP; ++j) {
          pin[i][j] = fadd(p0[i][j], fmul(fmul(dt, v0[i][j]), third));
          vin[i][j] = fadd(v0[i][j], fmul(fmul(dt, a1[i][j]), third));
          del[i][j] = c38 * lv[i][j] + dt * (g3v[i][j] - g4v[i][j]);
          dkp[i][j] = c38 * lp[i][j] + dt * (g3p[i][j] - g4p[i][j]);
        }
      write_stage_input(pin, vin);
      __syncthreads();
      mlp_forward(fadd(t0, fmul(dt, third)), a4);
      mlp_backward(fadd(t0, fmul(dt, third)), del, xp, xv);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < CP; ++j) { g2p[i][j] = xp[i][j]; g2v[i][j] = dkp[i][j] + xv[i][j]; }

      // ---- stage 1: u1 = y0;  dL/dk1 = (dt/8) lambda + dt g4 - (dt/3) g3 + (dt/3) g2
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < CP; ++j) {
          del[i][j] = c8 * lv[i][j] + dt * g4v[i][j] + d3 * (g2v[i][j] - g3v[i][j]);
          dkp[i][j] = c8 * lp[i][j] + dt * g4p[i][j] + d3 * (g2p[i][j] - g3p[i][j]);
        }
      write_stage_input(p0, v0);
      __syncthreads();
      mlp_forward(t0, a4);
      mlp_backward(t0, del, xp, xv);

      // ---- adjoint at t[step]: lambda += g1 + g2 + g3 + g4 + dL/dy_path[step]
      float gp[4][CP], gv[4][CP];
      load_pv(a.gy + (size_t)step * a.B * D, gp, gv);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < CP; ++j) {
          lp[i][j] += (xp[i][j] + g2p[i][j]) + (g3p[i][j] + g4p[i][j]) + gp[i][j];
          lv[i][j] += ((dkp[i][j] + xv[i][j]) + g2v[i][j]) + (g3v[i][j] + g4v[i][j]) + gv[i][j];
        }
    }

    // ---- dL/dy0
    if (own) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t g = m0 + 4 * tm + i;
        if (g < a.B) {
#pragma unroll
          for (int j = 0; j < CP; ++j) {
            a.gy0[g * D + tn * CP + j] = lp[i][j];
            a.gy0[g * D + P + tn * CP + j] = lv[i][j];
          }
        }
      }
    }
    __syncthreads();
    for (int i = tid; i < TMB * H; i += NT) {
      const int m = i / H, j = i % H;
      const int64_t g = m0 + m;
      if (g < a.B) {
        float s = sGH[j * XSB + m];
        for (int tt = 0; tt < a.T; ++tt) s += a.gy[((size_t)tt * a.B + g) * D + 2 * P + j];
        a.gy0[g * D + 2 * P + j] = s;
      }
    }
  }
}

// ---- sum the CTA-private gradient buffers and scatter into torch's flat [out][in] layout ---------------
__global__ void reduce_unpack_kernel(const float* __restrict__ priv, int nbuf, float* __restrict__ gw, int P, int H, int HID,
                                     int NRES) {
  const FlatLayout F{P, H, HID, NRES};
  const PackLayout L{P, H, HID, NRES};
  const int IN = 2 * P + H + 2;
  const int64_t total = F.total();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t src;
    if (i < F.off_bin()) {
      const int n = (int)(i / IN), k = (int)(i % IN);
      if (k < 2 * P) src = L.off_W0() + (int64_t)k * HID + n;
      else if (k < 2 * P + H) src = L.off_WH() + (int64_t)(k - 2 * P) * HID + n;
      else if (k == 2 * P + H) src = L.off_wsin() + n;
      else src = L.off_wcos() + n;
    } else if (i < F.off_res(0)) {
      src = L.off_bin() + (i - F.off_bin());
    } else if (i < F.off_wout()) {
      const int64_t per = 2 * (int64_t)HID * HID + 2 * HID;
      const int r = (int)((i - F.off_res(0)) / per);
      const int64_t q = (i - F.off_res(0)) % per;
      if (q < (int64_t)HID * HID) {
        const int n = (int)(q / HID), k = (int)(q % HID);
        src = L.off_WA(r) + (int64_t)k * HID + n;
      } else if (q < (int64_t)HID * HID + HID) {
        src = L.off_bA(r) + (q - (int64_t)HID * HID);
      } else if (q < 2 * (int64_t)HID * HID + HID) {
        const int64_t q2 = q - (int64_t)HID * HID - HID;
        const int n = (int)(q2 / HID), k = (int)(q2 % HID);
        src = L.off_WB(r) + (int64_t)k * HID + n;
      } else {
        src = L.off_bB(r) + (q - 2 * (int64_t)HID * HID - HID);
      }
    } else if (i < F.off_bout()) {
      const int64_t q = i - F.off_wout();
      const int n = (int)(q / HID), k = (int)(q % HID);
      src = L.off_WO() + (int64_t)k * P + n;
    } else {
      src = L.off_bO() + (i - F.off_bout());
    }
    float s = 0.0f;
    const int64_t stride = L.total();
    for (int b = 0; b < nbuf; ++b) s += priv[(size_t)b * stride + src];
    gw[i] = s;
  }
}

// [HID][2P] and [HID][H] contiguous copies of the (p,v) and h columns of w_in (dgrad operands of layer 0)
__global__ void pack_w0d_kernel(const float* __restrict__ w, float* __restrict__ w0d, float* __restrict__ whd, int P, int H, int HID) {
  const int IN = 2 * P + H + 2;
  const int64_t n0 = (int64_t)HID * 2 * P, n1 = (int64_t)HID * H;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n0 + n1; i += (int64_t)gridDim.x * blockDim.x) {
    if (i < n0) {
      const int n = (int)(i / (2 * P)), k = (int)(i % (2 * P));
      w0d[i] = w[(int64_t)n * IN + k];
    } else {
      const int64_t q = i - n0;
      const int n = (int)(q / H), j = (int)(q % H);
      whd[q] = w[(int64_t)n * IN + 2 * P + j];
    }
  }
}

int pack_drift(const ab200_drift_desc* d, const float* w_flat, float* packed, cudaStream_t st);

static int bwd_grid(int64_t B) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t ntiles = (B + TMB - 1) / TMB;
  return (int)(ntiles < sms ? ntiles : sms);
}

struct BwdWs {
  size_t off_pk, off_w0d, off_whd, off_priv, total;
};
static BwdWs bwd_ws(const ab200_drift_desc* d, int64_t B) {
  const PackLayout L{d->pos_dim, d->ctx_dim, d->hidden, d->n_res};
  BwdWs w;
  w.off_pk = 0;
  w.off_w0d = align_up((size_t)L.total() * 4, 256);
  w.off_whd = w.off_w0d + align_up((size_t)d->hidden * 2 * d->pos_dim * 4, 256);
  w.off_priv = w.off_whd + align_up((size_t)d->hidden * d->ctx_dim * 4, 256);
  w.total = w.off_priv + align_up((size_t)bwd_grid(B) * L.total() * 4, 256);
  return w;
}

size_t rk4_backward_f32_workspace(const ab200_drift_desc* d, int64_t B, int T) {
  (void)T;
  return bwd_ws(d, B).total;
}

template <int P, int H, int HID, int NRES, int ACT, int POT, int MODE = 0>
static int launch_bwd(const BwdArgs& args, int grid, cudaStream_t st) {
  constexpr int GR = (2 * P > HID) ? 2 * P : HID;
  constexpr int WN = (HID > 2 * P) ? HID : 2 * P;
  const size_t rows = (size_t)(2 * P + H) + (size_t)(NRES + 1) * HID + (size_t)NRES * HID + HID + HID + GR + P + H + 2;
  const size_t smem = sizeof(float) * (rows * XSB + 2 * (size_t)KC * WN);
  auto kern = rk4_bwd_f32_kernel<P, H, HID, NRES, ACT, POT, MODE>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_cuda_error(e); return AB200_ERR_CUDA; }
  kern<<<grid, NT, smem, st>>>(args);
  return check_launch();
}

template <int MODE>
static int backward_f32_common(const ab200_drift_desc* d, const float* w_flat, const float* t_dev, float t_eval, const float* y_path,
                               const float* grad_y_path, int64_t B, int T, float* grad_y0, float* grad_w_flat, void* ws, size_t ws_bytes,
                               cudaStream_t st) {
  const BwdWs w = bwd_ws(d, B);
  if (ws_bytes < w.total) return AB200_ERR_WORKSPACE;
  char* base = (char*)ws;
  float* pk = (float*)(base + w.off_pk);
  float* w0d = (float*)(base + w.off_w0d);
  float* whd = (float*)(base + w.off_whd);
  float* priv = (float*)(base + w.off_priv);
  const int grid = bwd_grid(B);
  const PackLayout L{d->pos_dim, d->ctx_dim, d->hidden, d->n_res};
  int rc = pack_drift(d, w_flat, pk, st);
  if (rc) return rc;
  pack_w0d_kernel<<<64, 256, 0, st>>>(w_flat, w0d, whd, d->pos_dim, d->ctx_dim, d->hidden);
  if ((rc = check_launch())) return rc;
  cudaError_t e = cudaMemsetAsync(priv, 0, (size_t)grid * L.total() * 4, st);
  if (e != cudaSuccess) { set_cuda_error(e); return AB200_ERR_CUDA; }

  BwdArgs a{pk, w_flat, w0d, whd, t_dev, y_path, grad_y_path, grad_y0, priv, B, T, (int)((B + TMB - 1) / TMB),
            d->time_period, d->pot_idx_a, d->pot_idx_b, d->pot_strength, t_eval};
  if (d->pos_dim == 64 && d->ctx_dim == 32 && d->hidden == 128 && d->n_res == 2 && d->res_act == 0 && d->potential == 0)
    rc = launch_bwd<64, 32, 128, 2, 0, 0, MODE>(a, grid, st);
  else if (d->pos_dim == 16 && d->ctx_dim == 32 && d->hidden == 128 && d->n_res == 2 && d->res_act == 1 && d->potential == 1)
    rc = launch_bwd<16, 32, 128, 2, 1, 1, MODE>(a, grid, st);
  else if (d->pos_dim == 16 && d->ctx_dim == 32 && d->hidden == 128 && d->n_res == 2 && d->res_act == 1 && d->potential == 0)
    rc = launch_bwd<16, 32, 128, 2, 1, 0, MODE>(a, grid, st);
  else
    return AB200_ERR_UNSUPPORTED;
  if (rc) return rc;
  const int64_t total = FlatLayout{d->pos_dim, d->ctx_dim, d->hidden, d->n_res}.total();
  reduce_unpack_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(priv, grid, grad_w_flat, d->pos_dim, d->ctx_dim, d->hidden,
                                                                     d->n_res);
  return check_launch();
}

int rk4_backward_f32(const ab200_drift_desc* d, const float* w_flat, const float* t_dev, const float* y_path,
                     const float* grad_y_path, int64_t B, int T, float* grad_y0, float* grad_w_flat, void* ws, size_t ws_bytes,
                     cudaStream_t st) {
  return backward_f32_common<0>(d, w_flat, t_dev, 0.0f, y_path, grad_y_path, B, T, grad_y0, grad_w_flat, ws, ws_bytes, st);
}

// J^T g of one drift evaluation at (t, y): grad_y [B][D] and grad_w_flat (both OVERWRITTEN); same workspace as the rk4 adjoint
int drift_vjp_f32(const ab200_drift_desc* d, const float* w_flat, float t, const float* y, const float* g, int64_t B, float* grad_y,
                  float* grad_w_flat, void* ws, size_t ws_bytes, cudaStream_t st) {
  return backward_f32_common<1>(d, w_flat, nullptr, t, y, g, B, 1, grad_y, grad_w_flat, ws, ws_bytes, st);
}

}  // namespace ab200
