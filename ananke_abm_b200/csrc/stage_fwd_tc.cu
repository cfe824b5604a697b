// Tensor-core forward STAGE kernel: one drift evaluation at a Runge-Kutta stage input, for every agent.
//
//   stage input   p = p0 + cpv v0 + sum_j cpa[j] a_j ,  v = v0 + sum_j cva[j] a_j      (fp32, from HBM/L2)
//   drift net     6 tcgen05 layers (bf16 x bf16 -> fp32 in TMEM), bias + time features inside the MMA
//   outputs       a_out = net(p, v, h, t)                                           [B][64]   (optional)
//                 y_out = [p0 + ocpv v0 + sum ocpa a_j + ocpa[n] a_out,  v0 + sum ocva a_j + ocva[n] a_out, h]   (optional)
//                 err   += sum_i ( e_i / (atol + rtol max(|y0_i|, |y_out_i|)) )^2,   e = sum_j (epa|eva)[j] a_j (+ a_out)
// which covers every stage of torchdiffeq's fixed-grid rk4 (rk_common.py rk4_alt_step_func) and adaptive dopri5
// (rk_common.py _runge_kutta_step + misc.py _compute_error_ratio): the host passes the tableau as `Combo`s.
// Two 128-agent tiles ("slots") are in flight per CTA; see stage_tc.cuh.
#include <cuda_fp16.h>
#include <stdlib.h>
#include "stage_tc.cuh"

namespace ab200 {
using namespace stc;

#ifdef AB200_STAGE_TRACE
extern "C" int ab200_debug_stage_trace(long long* host_out, int* counts) {
  cudaMemcpyFromSymbol(host_out, stc::g_stage_trace, sizeof(long long) * 8192);
  cudaMemcpyFromSymbol(counts, stc::g_stage_trace_n, sizeof(int) * 2);
  int z[2] = {0, 0};
  cudaMemcpyToSymbol(stc::g_stage_trace_n, z, sizeof(z));
  return 0;
}
#endif

// ---- prepack: torch-layout fp32 weights -> bf16 UMMA image with bias / time-feature K extensions ---------------
template <bool HALF>
__global__ void stage_pack_kernel(const float* __restrict__ w, uint8_t* __restrict__ out) {
  const FlatLayout F{P, H, HID, NRES};
  const int IN = 2 * P + H + 2;
  const int n_w1 = HID * K1, n_hh = HID * KH, n_wo = P * KH;
  const int total = n_w1 + 2 * NRES * n_hh + n_wo;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int n, k, N, Kmain;
    uint32_t base;
    const float* wsrc;
    const float* bsrc;
    int ldw;
    if (i < n_w1) {
      n = i / K1; k = i % K1; N = HID; Kmain = 2 * P + H; base = OFF_W1;
      wsrc = w + F.off_win(); bsrc = w + F.off_bin(); ldw = IN;
    } else if (i < n_w1 + 2 * NRES * n_hh) {
      const int q = i - n_w1, m = q / n_hh, r = q % n_hh;
      n = r / KH; k = r % KH; N = HID; Kmain = HID; base = off_hh(m);
      wsrc = w + ((m & 1) ? F.off_wb(m >> 1) : F.off_wa(m >> 1));
      bsrc = w + ((m & 1) ? F.off_bb(m >> 1) : F.off_ba(m >> 1));
      ldw = HID;
    } else {
      const int q = i - n_w1 - 2 * NRES * n_hh;
      n = q / KH; k = q % KH; N = P; Kmain = HID; base = OFF_WO;
      wsrc = w + F.off_wout(); bsrc = w + F.off_bout(); ldw = HID;
    }
    float v = 0.0f;
    if (k < Kmain) {
      v = wsrc[(size_t)n * ldw + k];
    } else {
      const int e = k - Kmain;
      auto hi = [](float x) { return HALF ? __half2float(__float2half_rn(x)) : __bfloat162float(__float2bfloat16_rn(x)); };
      if (base == OFF_W1 && e < 6) {
        const float wt = wsrc[(size_t)n * ldw + Kmain + (e < 3 ? 0 : 1)];     // sin column, cos column of w_in
        const int r = e % 3;
        v = (r == 1) ? wt - hi(wt) : wt;                                      // hi, lo, hi
      } else if (e == 6) {
        v = bsrc[n];
      } else if (e == 7) {
        v = bsrc[n] - hi(bsrc[n]);
      }
    }
    if (HALF) *reinterpret_cast<__half*>(out + base + off_kmajor_noswz(n, k, lbo(N), SBO)) = __float2half_rn(v);
    else *reinterpret_cast<__nv_bfloat16*>(out + base + off_kmajor_noswz(n, k, lbo(N), SBO)) = __float2bfloat16_rn(v);
  }
}

int stage_flags() {
  static int f = -1;
  if (f < 0) {
    const char* e = getenv("AB200_STAGE_FLAGS");
    int v = e ? atoi(e) : 0;   // measured on B200: neither the issue mutex nor the bulk L2 prefetch pays (profiles/r01_stage_notes.md)
    // bits 16 / 32 skip the backward kernel's stores (timing experiments, results INVALID): honoured only together with
    // AB200_STAGE_TIMING_ONLY=1 so that a stray environment variable cannot silently corrupt gradients
    const char* u = getenv("AB200_STAGE_TIMING_ONLY");
    if (!(u && atoi(u) == 1)) v &= (15 | 128);      // bit 128 (L2 eviction hints) changes no result
    f = v;
  }
  return f;
}

size_t stage_tc_image_bytes() { return STATUS_OFFSET + 256; }   // bf16 image, fp16 image, split-activation image + table, status word

int stage_tc_pack(const float* w_flat, uint8_t* image, cudaStream_t st) {
  stage_pack_kernel<false><<<148, 256, 0, st>>>(w_flat, image);
  stage_pack_kernel<true><<<148, 256, 0, st>>>(w_flat, image + IMG_STRIDE);
  const int rc = stage_fwd2_pack(w_flat, image + IMG2_OFFSET, st);
  if (rc) return rc;
  return check_launch();
}

struct StageParams {          // one stage of a fused sequence
  int n_a;                    // reads a[0 .. n_a)
  Combo in;                   // stage input
  float t;
  float* a_out;               // blocked [Bp][64] or null
  float* y_out;               // blocked [Bp][160] or null
  Combo out;                  // y_out combination; index n_a of cpa/cva multiplies this stage's own output
  Combo err;                  // error combination (cpv unused); index n_a multiplies this stage's own output
  int want_err;
};

struct StageFwdArgs {
  const uint8_t* wimg;
  const float* y0;            // blocked [Bp][160]
  const float* a[MAX_A];      // blocked [Bp][64] each; a stage may read what an EARLIER stage of this launch wrote
  int n_stage;
  StageParams st[MAX_A];
  double* err_sumsq;          // or null
  float rtol, atol, period;
  int64_t B;
  int ntiles;
  int flags;
  int* status;
};

// All stages of one solver step (or attempt) for a tile before moving to the next tile: the accelerations a stage needs
// were written by the SAME thread a few microseconds earlier and are still in L2, so HBM sees y0 once, each a_j once
// (its write) and the step result once -- instead of re-reading every earlier a_j from HBM in every stage launch.
template <bool HALF>
__global__ void __launch_bounds__(THREADS, 1) stage_fwd_tc_kernel(const __grid_constant__ StageFwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[NSLOT + 1];
  __shared__ uint32_t tmem_base_s;
  __shared__ double err_red[THREADS / 32];
  __shared__ int issue_lock;
  SlotCtx c = stage_setup(smem, a.wimg, bars, &tmem_base_s, &issue_lock, a.status, a.flags);
  const uint32_t tmem_base = tmem_base_s;
  double err_local = 0.0;

#pragma unroll 1
  for (int it = 0;; ++it) {
    const int tile = slot_tile(it, a.ntiles, c.slot);
    if (tile < 0) break;
    const bool valid = (int64_t)tile * TM + c.row < a.B;      // padding rows hold zeros and are never stored to
    STAGE_TRACE(c, 9);
    // context h -> HB once per tile (constant along the step)
    {
      uint32_t o[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 x = ldro(blk4(a.y0, tile, YF4, 2 * AF4 + c.hf * 4 + j, c.row));
        o[2 * j] = pack2<HALF>(x.x, x.y);
        o[2 * j + 1] = pack2<HALF>(x.z, x.w);
      }
      tmem_st8(c.tmem + c.lane_sel + C_HB + (uint32_t)(c.hf * 8), o);
    }

#pragma unroll 1
    for (int si = 0; si < a.n_stage; ++si) {
      const StageParams& sp = a.st[si];
      const int n_a = sp.n_a;
      // ---- stage input -> ACT (16-bit), time/bias block -> TB          (all buffers blocked, see stage_tc.cuh)
      // Both 16-dim halves of p and of v are built together: every source (y0, then each a_s) is ONE batch of 8-16
      // independent 128-bit loads, so the stage input costs n_a + 1 memory round trips instead of 2 * (n_a + 1).
      {
        const int f0 = c.hf * 8;             // this thread's 8 float4 groups (32 dims) of p and of v
        float pin[32], vin[32];
        const float cpv = sp.in.cpv;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 pv = ldro(blk4(a.y0, tile, YF4, f0 + j, c.row));
          const float4 vv = ldro(blk4(a.y0, tile, YF4, AF4 + f0 + j, c.row));
          pin[4 * j] = pv.x + cpv * vv.x; pin[4 * j + 1] = pv.y + cpv * vv.y;
          pin[4 * j + 2] = pv.z + cpv * vv.z; pin[4 * j + 3] = pv.w + cpv * vv.w;
          vin[4 * j] = vv.x; vin[4 * j + 1] = vv.y; vin[4 * j + 2] = vv.z; vin[4 * j + 3] = vv.w;
        }
#pragma unroll 1
        for (int s = 0; s < n_a; ++s) {
          const float cp = sp.in.cpa[s], cv = sp.in.cva[s];
          float4 x[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = *blk4(a.a[s], tile, AF4, f0 + j, c.row);   // coherent: may have been written by this launch
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            pin[4 * j] += cp * x[j].x; pin[4 * j + 1] += cp * x[j].y; pin[4 * j + 2] += cp * x[j].z; pin[4 * j + 3] += cp * x[j].w;
            vin[4 * j] += cv * x[j].x; vin[4 * j + 1] += cv * x[j].y; vin[4 * j + 2] += cv * x[j].z; vin[4 * j + 3] += cv * x[j].w;
          }
        }
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = pack2<HALF>(pin[16 * ch + 2 * j], pin[16 * ch + 2 * j + 1]);
          tmem_st8(c.tmem + c.lane_sel + C_ACT + (uint32_t)((f0 + ch * 4) * 2), o);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = pack2<HALF>(vin[16 * ch + 2 * j], vin[16 * ch + 2 * j + 1]);
          tmem_st8(c.tmem + c.lane_sel + C_ACT + (uint32_t)(P / 2 + (f0 + ch * 4) * 2), o);
        }
      }
      write_time_block<HALF>(c, sp.t, a.period);
      STAGE_TRACE(c, 10);

      // ---- drift net
      uint32_t z[32];
      run_layer<false, (2 * P + H) / 16, true, HID, HID, false, HALF>(c, C_ACT, OFF_W1);
      epi_relu<true, HALF>(c, z);
#pragma unroll 1
      for (int r = 0; r < NRES; ++r) {
        uint32_t dummy[32];
        run_layer<false, HID / 16, true, HID, HID, false, HALF>(c, C_ACT, off_hh(2 * r));
        epi_relu<false, HALF>(c, dummy);
        run_layer<false, HID / 16, true, HID, HID, false, HALF>(c, C_ACT, off_hh(2 * r + 1));
        epi_residual<HALF>(c, z);
      }
      run_layer<false, HID / 16, true, P, P, false, HALF>(c, C_ACT, OFF_WO);

      // ---- output epilogue: this thread's 32 acceleration dims (float4 groups hf*8 ..)
      STAGE_TRACE(c, 11);
      const int f0 = c.hf * 8;
      const bool want_y = sp.y_out != nullptr;
      const bool want_err = want_y && sp.want_err != 0;
      float oc = 0.f, ov = 0.f, ecp = 0.f, ecv = 0.f, ocpv = 0.f;
      if (want_y) { oc = sp.out.cpa[n_a]; ov = sp.out.cva[n_a]; ecp = sp.err.cpa[n_a]; ecv = sp.err.cva[n_a]; ocpv = sp.out.cpv; }
#pragma unroll 1
      for (int qd = 0; qd < 2; ++qd) {      // 4 float4 groups (16 dims) per pass: 4-8 loads in flight per source
        uint32_t r[16];
        tmem_ld16(c.tmem + c.lane_sel + C_ACC + (uint32_t)(c.hf * 32 + qd * 16), r);
        tmem_ld_wait();
        const int fq = f0 + qd * 4;
        if (sp.a_out != nullptr) {      // padding rows of the last tile are written as zeros: output buffers need no initialisation
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *blk4(sp.a_out, tile, AF4, fq + j, c.row) = valid ? make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                                            __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]))
                                                              : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (want_y) {
          float po[16], vo[16], ep[16], ev[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 p0 = ldro(blk4(a.y0, tile, YF4, fq + j, c.row));
            const float4 v0 = ldro(blk4(a.y0, tile, YF4, AF4 + fq + j, c.row));
            const float pb[4] = {p0.x, p0.y, p0.z, p0.w}, vb[4] = {v0.x, v0.y, v0.z, v0.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float ao = __uint_as_float(r[4 * j + e]);
              po[4 * j + e] = pb[e] + ocpv * vb[e] + oc * ao;
              vo[4 * j + e] = vb[e] + ov * ao;
              ep[4 * j + e] = ecp * ao;
              ev[4 * j + e] = ecv * ao;
            }
          }
#pragma unroll 1
          for (int s = 0; s < n_a; ++s) {
            float4 x[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) x[j] = *blk4(a.a[s], tile, AF4, fq + j, c.row);
            const float cp = sp.out.cpa[s], cv = sp.out.cva[s], xp = sp.err.cpa[s], xv = sp.err.cva[s];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float xs[4] = {x[j].x, x[j].y, x[j].z, x[j].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                po[4 * j + e] += cp * xs[e];
                vo[4 * j + e] += cv * xs[e];
                ep[4 * j + e] += xp * xs[e];
                ev[4 * j + e] += xv * xs[e];
              }
            }
          }
          {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              *blk4(sp.y_out, tile, YF4, fq + j, c.row) =
                  valid ? make_float4(po[4 * j], po[4 * j + 1], po[4 * j + 2], po[4 * j + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
              *blk4(sp.y_out, tile, YF4, AF4 + fq + j, c.row) =
                  valid ? make_float4(vo[4 * j], vo[4 * j + 1], vo[4 * j + 2], vo[4 * j + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (valid && want_err) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {       // y0 again (L1/L2 hit): keeping it live through the source loop costs 32 registers
                const float4 p0 = ldro(blk4(a.y0, tile, YF4, fq + j, c.row));
                const float4 v0 = ldro(blk4(a.y0, tile, YF4, AF4 + fq + j, c.row));
                const float pb[4] = {p0.x, p0.y, p0.z, p0.w}, vb[4] = {v0.x, v0.y, v0.z, v0.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float tp = a.atol + a.rtol * fmaxf(fabsf(pb[e]), fabsf(po[4 * j + e]));
                  const float tv = a.atol + a.rtol * fmaxf(fabsf(vb[e]), fabsf(vo[4 * j + e]));
                  const float qp = ep[4 * j + e] / tp, qv = ev[4 * j + e] / tv;
                  err_local += (double)(qp * qp + qv * qv);
                }
              }
            }
          }
        }
      }
      if (want_y) {   // context h rides along unchanged (dh/dt = 0); padding rows: zeros
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 hv = ldro(blk4(a.y0, tile, YF4, 2 * AF4 + c.hf * 4 + j, c.row));
          *blk4(sp.y_out, tile, YF4, 2 * AF4 + c.hf * 4 + j, c.row) = valid ? hv : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
  }

  if (a.err_sumsq != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) err_local += __shfl_xor_sync(0xffffffffu, err_local, o);
    if ((threadIdx.x & 31) == 0) err_red[threadIdx.x >> 5] = err_local;
  }
  stage_teardown(tmem_base);
  if (a.err_sumsq != nullptr && threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < THREADS / 32; ++i) s += err_red[i];
    if (*reinterpret_cast<volatile int*>(a.status) != 0) s = __longlong_as_double(0x7ff8000000000000LL);   // a bounded wait expired: poison the norm
    atomicAdd(a.err_sumsq, s);
  }
}

// ---- host side -----------------------------------------------------------------------------------------------
struct StageFwdHost {   // mirrors ab200_stage_desc in the public header
  int32_t n_a;
  float in_cpv, in_cpa[MAX_A], in_cva[MAX_A];
  float t;
  float out_cpv, out_cpa[MAX_A + 1], out_cva[MAX_A + 1];
  float err_pa[MAX_A + 1], err_va[MAX_A + 1];
  float rtol, atol;
};

// n_stage fused stages: stage s reads a_ptrs[0 .. descs[s].n_a), writes a_outs[s] (may be null); the LAST stage may also
// write y_out (+ the error norm) with its out / err combinations.
int stage_fwd_tc_multi(const ab200_drift_desc* d, const uint8_t* image, const float* y0, const float* const* a_ptrs, const void* descs_v,
                       int n_stage, float* const* a_outs, int64_t B, float* y_out, double* err_sumsq, int half_ops, cudaStream_t st) {
  const StageFwdHost* hs = reinterpret_cast<const StageFwdHost*>(descs_v);
  if (n_stage < 1 || n_stage > MAX_A) return AB200_ERR_BAD_ARG;
  StageFwdArgs k{};
  k.wimg = image + (half_ops ? IMG_STRIDE : 0);
  k.y0 = y0;
  int max_a = 0;
  for (int s = 0; s < n_stage; ++s) {
    const StageFwdHost& h = hs[s];
    if (h.n_a < 0 || h.n_a > MAX_A) return AB200_ERR_BAD_ARG;
    max_a = h.n_a > max_a ? h.n_a : max_a;
    StageParams& sp = k.st[s];
    sp.n_a = h.n_a;
    sp.in.cpv = h.in_cpv;
    sp.out.cpv = h.out_cpv;
    sp.err.cpv = 0.f;
    for (int i = 0; i < MAX_A; ++i) { sp.in.cpa[i] = h.in_cpa[i]; sp.in.cva[i] = h.in_cva[i]; }
    for (int i = 0; i <= MAX_A; ++i) {
      sp.out.cpa[i] = h.out_cpa[i]; sp.out.cva[i] = h.out_cva[i];
      sp.err.cpa[i] = h.err_pa[i]; sp.err.cva[i] = h.err_va[i];
    }
    sp.t = h.t;
    sp.a_out = a_outs ? a_outs[s] : nullptr;
    sp.y_out = (s == n_stage - 1) ? y_out : nullptr;
    sp.want_err = (s == n_stage - 1 && err_sumsq != nullptr) ? 1 : 0;
  }
  for (int i = 0; i < MAX_A; ++i) k.a[i] = (i < max_a) ? a_ptrs[i] : nullptr;
  k.n_stage = n_stage;
  k.period = d->time_period;
  k.err_sumsq = err_sumsq;
  k.rtol = hs[n_stage - 1].rtol;
  k.atol = hs[n_stage - 1].atol;
  k.B = B;
  k.ntiles = (int)((B + TM - 1) / TM);
  k.flags = stage_flags();
  k.status = stage_status_ptr(image);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = k.ntiles < sms ? k.ntiles : sms;      // a partial wave uses one slot per CTA first (slot_tile)
  auto kern = half_ops ? stage_fwd_tc_kernel<true> : stage_fwd_tc_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)W_BYTES);
  if (e != cudaSuccess) { set_cuda_error(e); return AB200_ERR_CUDA; }
  kern<<<grid, THREADS, W_BYTES, st>>>(k);
  return check_launch();
}

int stage_fwd_tc(const ab200_drift_desc* d, const uint8_t* image, const float* y0, const float* const* a_ptrs, const void* desc_v,
                 int64_t B, float* a_out, float* y_out, double* err_sumsq, int half_ops, cudaStream_t st) {
  float* outs[1] = {a_out};
  return stage_fwd_tc_multi(d, image, y0, a_ptrs, desc_v, 1, outs, B, y_out, err_sumsq, half_ops, st);
}

}  // namespace ab200
