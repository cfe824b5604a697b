// Split-activation forward STAGE kernel (operand_format 2, "fp16x2"): the drift evaluation of stage_fwd_tc.cu with
// every ACTIVATION entering the tensor core as a two-term IEEE fp16 split  x = hi + lo  (hi = fp16(x), lo = fp16(x - hi)),
// i.e. two tcgen05.mma per K-step against the same fp16 weight tile; biases and the sin/cos time-feature columns are
// added in fp32 by the epilogue.  Products of fp16 numbers are exact in the fp32 accumulator, so the only rounding left
// in an evaluation is the (fixed, state-independent) fp16 rounding of the WEIGHTS: the net that is integrated is a
// slightly perturbed but smooth function, and an adaptive solver sees no evaluation noise.
//
// Why it exists: dopri5's embedded error estimate  dt * sum_j c_err[j] k_j  differentiates the stage evaluations.  With
// activations rounded to 11 bits (stage_fwd_tc.cu, fp16) the estimate is noise-limited below rtol ~ 1e-4: at the
// reference's rtol = atol = 1e-5 (mode_sep/config.py:27-28) the solver took 79 accepted + 21 rejected steps per day
// where the fp32 reference takes 33 (bf16 activations: 469).  Rounding the weights alone leaves the step sequence of
// the fp32 solver unchanged (tests/test_gpu_stage.py, profiles/r02_*).
//
// Tensor memory per slot (256 columns, two slots per CTA): two 128-column regions X = [0,128) and Y = [128,256) used
// in ping-pong.  A layer reads its A operand from one region and accumulates into the other; the epilogue converts the
// fp32 accumulator IN PLACE into the next layer's operand: the 16 accumulator columns of features 16g .. 16g+15 become
// 8 columns of packed hi pairs followed by 8 columns of packed lo pairs (exactly the two A tiles of K-step g).
// The context h (constant along a trajectory: dh/dt = 0) enters layer 1 as a hi and a lo fp16 tile from shared memory.
#include <cuda_fp16.h>
#include "stage_tc.cuh"
#include "wgrad_layout.cuh"

namespace ab200 {
using namespace stc;

namespace f2 {
constexpr int KIN = 2 * P + H;                                        // 160 input features that go through the MMA
constexpr uint32_t OFF_W1 = 0, SZ_W1 = (uint32_t)HID * KIN * 2;       // [128][160]
constexpr uint32_t OFF_RES = OFF_W1 + SZ_W1, SZ_HH = (uint32_t)HID * HID * 2;   // 4 x [128][128]
constexpr uint32_t OFF_WO = OFF_RES + 2u * NRES * SZ_HH, SZ_WO = (uint32_t)P * HID * 2;   // [64][128]
constexpr uint32_t W2_BYTES = OFF_WO + SZ_WO;                         // 188,416
__host__ __device__ constexpr uint32_t off_hh(int m) { return OFF_RES + (uint32_t)m * SZ_HH; }
// fp32 table behind the matrices (floats): w_sin[128] w_cos[128] b_in[128] | bA0 bB0 bA1 bB1 [128 each] | b_out[64]
constexpr int T_WSIN = 0, T_WCOS = HID, T_BIN = 2 * HID, T_BHH = 3 * HID, T_BOUT = T_BHH + 2 * NRES * HID, T_FLOATS = T_BOUT + P;
constexpr uint32_t OFF_TAB = W2_BYTES, TAB_BYTES = (uint32_t)T_FLOATS * 4;     // 3,840
constexpr uint32_t IMG_BYTES = OFF_TAB + TAB_BYTES;                   // 192,256: what ab200_stage_pack writes / the kernel copies
static_assert(IMG_BYTES == IMG2_BYTES, "stage_tc.cuh IMG2_BYTES");
// shared memory behind the image: per-slot h tiles (canonical K-major A operands, 128 agents x 32: hi then lo) and per-slot
// layer-1 vector
constexpr uint32_t OFF_HT = IMG_BYTES, SZ_HT1 = (uint32_t)TM * H * 2, SZ_HT = 2 * SZ_HT1;   // 8,192 per term
constexpr uint32_t OFF_CT = OFF_HT + NSLOT * SZ_HT;                   // [slot][128] floats: w_sin sin + w_cos cos + b_in
constexpr uint32_t SMEM_BYTES = OFF_CT + NSLOT * HID * 4;             // 226,048 (+ ~1 KB static) of the 232,448 available
static_assert(OFF_HT % 128 == 0, "operand tile alignment");
constexpr uint32_t RX = 0, RY = 128;                                  // tensor-memory regions of a slot
constexpr uint32_t HT_LBO = (uint32_t)TM * 16u;                       // K-adjacent core matrices of the h tile
}  // namespace f2

// ---- prepack ---------------------------------------------------------------------------------------------------
__global__ void stage_fwd2_pack_kernel(const float* __restrict__ w, uint8_t* __restrict__ out) {
  const FlatLayout F{P, H, HID, NRES};
  const int IN = 2 * P + H + 2;
  const int n_w1 = HID * f2::KIN, n_hh = HID * HID, n_wo = P * HID;
  const int n_mat = n_w1 + 2 * NRES * n_hh + n_wo;
  const int total = n_mat + f2::T_FLOATS;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    if (i >= n_mat) {
      const int e = i - n_mat;
      float v;
      if (e < f2::T_WCOS) v = w[F.off_win() + (size_t)e * IN + 2 * P + H];
      else if (e < f2::T_BIN) v = w[F.off_win() + (size_t)(e - f2::T_WCOS) * IN + 2 * P + H + 1];
      else if (e < f2::T_BHH) v = w[F.off_bin() + (e - f2::T_BIN)];
      else if (e < f2::T_BOUT) {
        const int m = (e - f2::T_BHH) / HID, n = (e - f2::T_BHH) % HID;
        v = w[((m & 1) ? F.off_bb(m >> 1) : F.off_ba(m >> 1)) + n];
      } else v = w[F.off_bout() + (e - f2::T_BOUT)];
      reinterpret_cast<float*>(out + f2::OFF_TAB)[e] = v;
      continue;
    }
    int n, k, N;
    uint32_t base;
    float v;
    if (i < n_w1) {
      n = i / f2::KIN; k = i % f2::KIN; N = HID; base = f2::OFF_W1;
      v = w[F.off_win() + (size_t)n * IN + k];
    } else if (i < n_w1 + 2 * NRES * n_hh) {
      const int q = i - n_w1, m = q / n_hh, r = q % n_hh;
      n = r / HID; k = r % HID; N = HID; base = f2::off_hh(m);
      v = w[((m & 1) ? F.off_wb(m >> 1) : F.off_wa(m >> 1)) + (size_t)n * HID + k];
    } else {
      const int q = i - n_w1 - 2 * NRES * n_hh;
      n = q / HID; k = q % HID; N = P; base = f2::OFF_WO;
      v = w[F.off_wout() + (size_t)n * HID + k];
    }
    *reinterpret_cast<__half*>(out + base + off_kmajor_noswz(n, k, lbo(N), SBO)) = __float2half_rn(v);
  }
}

int stage_fwd2_pack(const float* w_flat, uint8_t* image2, cudaStream_t st) {
  stage_fwd2_pack_kernel<<<148, 256, 0, st>>>(w_flat, image2);
  return check_launch();
}

// ---- one layer -------------------------------------------------------------------------------------------------
// The MMA stream of one layer, issued by ONE thread (out of line and rolled: see issue_mmas in stage_tc.cuh).
//   A: `nks` K-steps in tensor memory, K-step ks = hi tile at a0 + 16 ks and lo tile at a0 + 16 ks + 8;
//      then, when hdesc != 0, two K-steps of the h tiles in shared memory (hi tile at hdesc, lo tile SZ_HT1 behind it).
static __device__ __noinline__ void issue_mmas2(uint32_t acc, uint32_t a0, uint64_t d0, uint32_t step16, uint32_t idesc, int nks,
                                                uint64_t hdesc, uint64_t* bar) {
#pragma unroll 1
  for (int ks = 0; ks < nks; ++ks) {
    const uint64_t bd = d0 + (uint64_t)((uint32_t)ks * step16);
    mma_ts(acc, a0 + (uint32_t)ks * 16u, bd, idesc, ks > 0 ? 1u : 0u);
    mma_ts(acc, a0 + (uint32_t)ks * 16u + 8u, bd, idesc, 1u);
  }
  if (hdesc != 0) {
#pragma unroll 1
    for (int j = 0; j < H / 16; ++j) {
      const uint64_t ad = hdesc + (uint64_t)((uint32_t)j * ((2u * f2::HT_LBO) >> 4));
      const uint64_t bd = d0 + (uint64_t)((uint32_t)(nks + j) * step16);
      mma_ss(acc, ad, bd, idesc, 1u);
      mma_ss(acc, ad + (uint64_t)(f2::SZ_HT1 >> 4), bd, idesc, 1u);
    }
  }
  mma_commit(bar);
}

// ACC[region d_reg][128 x N] = A(region a_reg) * W^T (K-major image at w_off, N_IMG rows); `after_issue`: see run_layer (stage_tc.cuh)
template <int NKS, bool HPART, int N_IMG, int N, class Hook = NoHook>
__device__ __forceinline__ void run_layer2(SlotCtx& c, uint32_t a_reg, uint32_t d_reg, uint32_t w_off, Hook after_issue = Hook()) {
  STAGE_TRACE(c, 1);
  tmem_st_wait();
  tc_fence_before();
  if (HPART) fence_async_smem();        // the h tile was written with st.shared: make it visible to the tensor core's proxy
  slot_sync(c.slot);
  STAGE_TRACE(c, 3);
  if (c.stid == 0) {
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc_bf16(TM, N, false, false, false, true);
    constexpr uint32_t L = lbo(N_IMG);
    constexpr uint32_t step16 = (2u * L) >> 4;
    const uint64_t d0 = make_smem_desc(c.sbase + w_off, L, SBO, SWZ_NONE);
    const uint64_t hd = HPART ? make_smem_desc(c.sbase + f2::OFF_HT + (uint32_t)c.slot * f2::SZ_HT, f2::HT_LBO, SBO, SWZ_NONE) : 0ull;
    issue_mmas2(c.tmem + d_reg, c.tmem + a_reg, d0, step16, idesc, NKS, hd, c.bar);
  }
  STAGE_TRACE(c, 4);
  __syncwarp();
  after_issue();
  if (c.alive && !wait_mma(c.bar, c.phase)) { c.alive = false; *c.status = 1; }
  c.phase ^= 1;
  __syncwarp();
  tc_fence_after();
  STAGE_TRACE(c, 5);
}

// Hidden-layer epilogue on this thread's 64 columns (hf * 64 ..) of region `reg`, in place:
//   x = acc + bias [+ z] ;  x = relu(x) ;  [z = x] ;  columns <- (hi pairs | lo pairs) per 16-feature group
template <bool RES, bool KEEP>
__device__ __forceinline__ void epi2(const SlotCtx& c, uint32_t reg, const float* __restrict__ bias, float (&z)[64]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int col = c.hf * 64 + q * 16;
    uint32_t r[16];
    tmem_ld16(c.tmem + c.lane_sel + reg + (uint32_t)col, r);
    tmem_ld_wait();
    uint32_t o[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 b2 = *reinterpret_cast<const float2*>(bias + col + 2 * j);      // same address for the whole warp: broadcast
      float x0 = __uint_as_float(r[2 * j]) + b2.x, x1 = __uint_as_float(r[2 * j + 1]) + b2.y;
      if (RES) { x0 += z[q * 16 + 2 * j]; x1 += z[q * 16 + 2 * j + 1]; }
      x0 = fmaxf(x0, 0.0f); x1 = fmaxf(x1, 0.0f);
      if (KEEP) { z[q * 16 + 2 * j] = x0; z[q * 16 + 2 * j + 1] = x1; }
      const uint32_t hi = pack2<true>(x0, x1);
      o[j] = hi;
      o[8 + j] = pack2<true>(x0 - un_lo<true>(hi), x1 - un_hi<true>(hi));
    }
    tmem_st16(c.tmem + c.lane_sel + reg + (uint32_t)col, o);
  }
}

// Saving for the backward pass (SAVE_ACTS kernels), executed while the tensor core runs the layer that CONSUMES the operand: the
// hi words of this thread's 64 columns of region `reg` (fp16 pairs; the lo words only matter below bf16 resolution) are read back
// from tensor memory, converted to the bf16 blob of the weight-gradient kernel (`blob` = this tile's blob + row * 16; a mixed
// fp16 x bf16 tcgen05.mma is an illegal instruction on sm_100a, so the conversion cannot be skipped) and their non-zero pattern
// goes out as the ReLU mask (two words, the bit layout of stage_bwd_tc.cu).  In the epilogue itself the same work lengthened the
// slot's critical path by a third (profiles/r02_stage_source_stalls.txt).
__device__ __forceinline__ void save_hidden(const SlotCtx& c, uint32_t reg, uint8_t* blob, uint2* mask_out) {
  uint32_t mw[2] = {0u, 0u};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t r[8];
    tmem_ld8(c.tmem + c.lane_sel + reg + (uint32_t)(c.hf * 64 + q * 16), r);
    tmem_ld_wait();
    uint32_t w[8];
    uint32_t m = 0u;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      w[j] = pack_bf16(un_lo<true>(r[j]), un_hi<true>(r[j]));
      // post-ReLU halves are >= +0: adding 0x7FFF sets bit 15 / 31 iff the low / high half is non-zero (no carry across)
      m |= ((r[j] + 0x7FFF7FFFu) >> ((q & 1) * 8 + j)) & (0x80008000u >> ((q & 1) * 8 + j));
    }
    mw[q >> 1] |= m;
    __stcs(reinterpret_cast<uint4*>(blob + (size_t)(c.hf * 8 + q * 2) * wg::FG_BYTES), make_uint4(w[0], w[1], w[2], w[3]));
    __stcs(reinterpret_cast<uint4*>(blob + (size_t)(c.hf * 8 + q * 2 + 1) * wg::FG_BYTES), make_uint4(w[4], w[5], w[6], w[7]));
  }
  __stcs(mask_out, make_uint2(mw[0], mw[1]));
}

// the same for the stage input (operand of layer 1): p / v hi words from region X, h hi words from this slot's shared-memory
// tile, the time-feature group.  Feature groups: p 4 hf + {0..3}, v 8 + 4 hf + {0..3}, h 16 + 2 hf + {0, 1}, [sin, cos, 1, 0..] 20, zeros 21
__device__ __forceinline__ void save_input(const SlotCtx& c, uint8_t* xb, const uint8_t* htile, float t, float period) {
  auto cvt4 = [](uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3) {
    return make_uint4(pack_bf16(un_lo<true>(a0), un_hi<true>(a0)), pack_bf16(un_lo<true>(a1), un_hi<true>(a1)),
                      pack_bf16(un_lo<true>(a2), un_hi<true>(a2)), pack_bf16(un_lo<true>(a3), un_hi<true>(a3)));
  };
#pragma unroll
  for (int pv = 0; pv < 2; ++pv) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      uint32_t r[8];
      tmem_ld8(c.tmem + c.lane_sel + f2::RX + (uint32_t)(pv * P + c.hf * 32 + q * 16), r);
      tmem_ld_wait();
      uint8_t* g0 = xb + (size_t)(pv * (P / 8) + c.hf * 4 + q * 2) * wg::FG_BYTES;
      __stcs(reinterpret_cast<uint4*>(g0), cvt4(r[0], r[1], r[2], r[3]));
      __stcs(reinterpret_cast<uint4*>(g0 + wg::FG_BYTES), cvt4(r[4], r[5], r[6], r[7]));
    }
  }
#pragma unroll
  for (int kc = 0; kc < 2; ++kc) {
    const uint4 v = *reinterpret_cast<const uint4*>(htile + off_kmajor_noswz(c.row, c.hf * 16 + kc * 8, f2::HT_LBO, SBO));
    __stcs(reinterpret_cast<uint4*>(xb + (size_t)(2 * P / 8 + c.hf * 2 + kc) * wg::FG_BYTES), cvt4(v.x, v.y, v.z, v.w));
  }
  if (c.hf == 0) {
    float sn, co;
    time_features(t, period, sn, co);
    __stcs(reinterpret_cast<uint4*>(xb + (size_t)((2 * P + H) / 8) * wg::FG_BYTES), make_uint4(pack_bf16(sn, co), pack_bf16(1.0f, 0.0f), 0u, 0u));
    __stcs(reinterpret_cast<uint4*>(xb + (size_t)((2 * P + H) / 8 + 1) * wg::FG_BYTES), make_uint4(0u, 0u, 0u, 0u));
  }
}

// 16 features -> (8 hi columns | 8 lo columns) at tensor-memory column `col` of this thread's lane; `xg` != null: they also go
// out as two feature groups of the bf16 X blob (xg = address of the first group for this row)
__device__ __forceinline__ void st_split16(const SlotCtx& c, uint32_t col, const float* x, uint8_t* xg = nullptr) {
  uint32_t o[16];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t hi = pack2<true>(x[2 * j], x[2 * j + 1]);
    o[j] = hi;
    o[8 + j] = pack2<true>(x[2 * j] - un_lo<true>(hi), x[2 * j + 1] - un_hi<true>(hi));
  }
  tmem_st16(c.tmem + c.lane_sel + col, o);
  if (xg != nullptr) {
    __stcs(reinterpret_cast<uint4*>(xg), make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7])));
    __stcs(reinterpret_cast<uint4*>(xg + wg::FG_BYTES),
           make_uint4(pack_bf16(x[8], x[9]), pack_bf16(x[10], x[11]), pack_bf16(x[12], x[13]), pack_bf16(x[14], x[15])));
  }
}

struct Stage2Params {         // one stage of a fused sequence (same meaning as StageParams in stage_fwd_tc.cu)
  int n_a;
  Combo in;
  float t;
  float* a_out;
  float* y_out;
  Combo out;
  Combo err;
  int want_err;
  uint8_t* x1_out;            // or null: what this stage saves for the backward pass (wg::FwdSaveLayout: X blob; + activations and masks
                              // when the launch runs the SAVE_ACTS kernel)
};

struct StageFwd2Args {
  const uint8_t* wimg;        // split-activation image (matrices + fp32 table)
  const float* y0;
  const float* a[MAX_A];
  int n_stage;
  Stage2Params st[MAX_A];
  double* err_sumsq;
  float rtol, atol, period;
  int64_t B;
  int ntiles;
  int flags;
  int* status;
};

template <bool SAVE_ACTS, bool L2POL = false>
__global__ void __launch_bounds__(THREADS, 1) stage_fwd2_tc_kernel(const __grid_constant__ StageFwd2Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[NSLOT + 1];
  __shared__ uint32_t tmem_base_s;
  __shared__ double err_red[THREADS / 32];
  __shared__ int issue_lock;
  SlotCtx c = stage_setup<f2::IMG_BYTES>(smem, a.wimg, bars, &tmem_base_s, &issue_lock, a.status, a.flags);
  const uint32_t tmem_base = tmem_base_s;
  const float* tab = reinterpret_cast<const float*>(smem + f2::OFF_TAB);
  float* ct = reinterpret_cast<float*>(smem + f2::OFF_CT) + c.slot * HID;
  uint8_t* htile = smem + f2::OFF_HT + (uint32_t)c.slot * f2::SZ_HT;
  double err_local = 0.0;
  uint64_t pol_keep = 0, pol_drop = 0;
  if (L2POL) { pol_keep = l2_policy_keep(); pol_drop = l2_policy_drop(); }

#pragma unroll 1
  for (int it = 0;; ++it) {
    const int tile = slot_tile(it, a.ntiles, c.slot);
    if (tile < 0) break;
    const bool valid = (int64_t)tile * TM + c.row < a.B;      // padding rows hold zeros and are never stored to
    STAGE_TRACE(c, 9);
    // context h -> this slot's shared-memory operand tile, once per tile (canonical K-major: 8-row x 16-byte cores)
    {
#pragma unroll
      for (int kc = 0; kc < 2; ++kc) {      // this thread's two 8-feature cores: h dims hf*16 + kc*8 ..
        const float4 x0 = ldro(blk4(a.y0, tile, YF4, 2 * AF4 + c.hf * 4 + kc * 2, c.row));
        const float4 x1 = ldro(blk4(a.y0, tile, YF4, 2 * AF4 + c.hf * 4 + kc * 2 + 1, c.row));
        const uint4 v = make_uint4(pack2<true>(x0.x, x0.y), pack2<true>(x0.z, x0.w), pack2<true>(x1.x, x1.y), pack2<true>(x1.z, x1.w));
        const uint4 l = make_uint4(pack2<true>(x0.x - un_lo<true>(v.x), x0.y - un_hi<true>(v.x)), pack2<true>(x0.z - un_lo<true>(v.y), x0.w - un_hi<true>(v.y)),
                                   pack2<true>(x1.x - un_lo<true>(v.z), x1.y - un_hi<true>(v.z)), pack2<true>(x1.z - un_lo<true>(v.w), x1.w - un_hi<true>(v.w)));
        const uint32_t off = off_kmajor_noswz(c.row, c.hf * 16 + kc * 8, f2::HT_LBO, SBO);
        *reinterpret_cast<uint4*>(htile + off) = v;
        *reinterpret_cast<uint4*>(htile + f2::SZ_HT1 + off) = l;
      }
    }

#pragma unroll 1
    for (int si = 0; si < a.n_stage; ++si) {
      const Stage2Params& sp = a.st[si];
      const int n_a = sp.n_a;
      // ---- stage input -> region X as hi/lo splits; layer-1 time/bias vector -> ct
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
          for (int j = 0; j < 8; ++j) {      // coherent: may have been written by this launch
            const float4* ptr = blk4(a.a[s], tile, AF4, f0 + j, c.row);
            x[j] = L2POL ? ld_l2hint(ptr, pol_keep) : *ptr;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            pin[4 * j] += cp * x[j].x; pin[4 * j + 1] += cp * x[j].y; pin[4 * j + 2] += cp * x[j].z; pin[4 * j + 3] += cp * x[j].w;
            vin[4 * j] += cv * x[j].x; vin[4 * j + 1] += cv * x[j].y; vin[4 * j + 2] += cv * x[j].z; vin[4 * j + 3] += cv * x[j].w;
          }
        }
        // p features 32 hf .. -> columns 32 hf .. ; v features 64 + 32 hf .. -> columns 64 + 32 hf ..
        // The backward pass needs exactly this stage input as a bf16 operand image (weight-gradient B operand; at save level 1 also
        // the recompute A operand): written here it costs 352 B per agent-stage of a launch that uses little of the DRAM bandwidth,
        // and saves the backward kernel re-reading y0 and up to six a_j (1.5 KB) and spilling the blob.
        // Feature groups: p 4 hf + {0..3}, v 8 + 4 hf + {0..3}, h 16 + 2 hf + {0, 1}, [sin, cos, 1, 0..] 20, zeros 21.
        // (SAVE_ACTS kernels write it after the issue of layer 1 instead, from the operand itself: save_input)
        uint8_t* xb = (!SAVE_ACTS && sp.x1_out != nullptr) ? sp.x1_out + (size_t)tile * wg::X1_BYTES + (size_t)c.row * 16 : nullptr;
        auto grp = [&](int g) -> uint8_t* { return xb ? xb + (size_t)g * wg::FG_BYTES : nullptr; };
        st_split16(c, f2::RX + (uint32_t)(c.hf * 32), pin, grp(c.hf * 4));
        st_split16(c, f2::RX + (uint32_t)(c.hf * 32 + 16), pin + 16, grp(c.hf * 4 + 2));
        st_split16(c, f2::RX + (uint32_t)(P + c.hf * 32), vin, grp(P / 8 + c.hf * 4));
        st_split16(c, f2::RX + (uint32_t)(P + c.hf * 32 + 16), vin + 16, grp(P / 8 + c.hf * 4 + 2));
        if (xb != nullptr) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {      // h groups 16 + 2 hf + q (16 context dims per thread; L1 / L2 hits)
            const float4 x0 = ldro(blk4(a.y0, tile, YF4, 2 * AF4 + c.hf * 4 + 2 * q, c.row));
            const float4 x1 = ldro(blk4(a.y0, tile, YF4, 2 * AF4 + c.hf * 4 + 2 * q + 1, c.row));
            __stcs(reinterpret_cast<uint4*>(grp(2 * P / 8 + c.hf * 2 + q)),
                   make_uint4(pack_bf16(x0.x, x0.y), pack_bf16(x0.z, x0.w), pack_bf16(x1.x, x1.y), pack_bf16(x1.z, x1.w)));
          }
          if (c.hf == 0) {
            float sn, co;
            time_features(sp.t, a.period, sn, co);
            __stcs(reinterpret_cast<uint4*>(grp((2 * P + H) / 8)), make_uint4(pack_bf16(sn, co), pack_bf16(1.0f, 0.0f), 0u, 0u));
            __stcs(reinterpret_cast<uint4*>(grp((2 * P + H) / 8 + 1)), make_uint4(0u, 0u, 0u, 0u));
          }
        }
      }
      if (c.stid < HID) {
        float s, co;
        time_features(sp.t, a.period, s, co);
        ct[c.stid] = fmaf(tab[f2::T_WSIN + c.stid], s, fmaf(tab[f2::T_WCOS + c.stid], co, tab[f2::T_BIN + c.stid]));
      }
      STAGE_TRACE(c, 10);

      // ---- drift net (regions alternate: X -> Y -> X -> Y -> X -> Y -> X)
      float z[64];
      const wg::FwdSaveLayout FS{a.ntiles};      // SAVE_ACTS launches save every stage (the host checks x1_out != null)
      const bool save = SAVE_ACTS && !(c.flags & 64);      // flag 64: timing experiment without the saving work (results invalid)
      auto hidden_hook = [&](int l, uint32_t reg) {
        return [=, &c]() {
          if (save) save_hidden(c, reg, sp.x1_out + FS.act(l, tile) + (size_t)c.row * 16,
                                reinterpret_cast<uint2*>(sp.x1_out + FS.mask(tile)) + (l * 2 * TM + c.hf * TM + c.row));
        };
      };
      run_layer2<2 * P / 16, true, HID, HID>(c, f2::RX, f2::RY, f2::OFF_W1, [&]() {
        if (save) save_input(c, sp.x1_out + FS.x1(tile) + (size_t)c.row * 16, htile, sp.t, a.period);
      });
      epi2<false, true>(c, f2::RY, ct, z);
      run_layer2<HID / 16, false, HID, HID>(c, f2::RY, f2::RX, f2::off_hh(0), hidden_hook(0, f2::RY));
      epi2<false, false>(c, f2::RX, tab + f2::T_BHH, z);
      run_layer2<HID / 16, false, HID, HID>(c, f2::RX, f2::RY, f2::off_hh(1), hidden_hook(1, f2::RX));
      epi2<true, true>(c, f2::RY, tab + f2::T_BHH + HID, z);
      run_layer2<HID / 16, false, HID, HID>(c, f2::RY, f2::RX, f2::off_hh(2), hidden_hook(2, f2::RY));
      epi2<false, false>(c, f2::RX, tab + f2::T_BHH + 2 * HID, z);
      run_layer2<HID / 16, false, HID, HID>(c, f2::RX, f2::RY, f2::off_hh(3), hidden_hook(3, f2::RX));
      epi2<true, false>(c, f2::RY, tab + f2::T_BHH + 3 * HID, z);
      run_layer2<HID / 16, false, P, P>(c, f2::RY, f2::RX, f2::OFF_WO, hidden_hook(4, f2::RY));

      // ---- output epilogue: this thread's 32 acceleration dims (float4 groups hf*8 ..)
      STAGE_TRACE(c, 11);
      const int f0 = c.hf * 8;
      const bool want_y = sp.y_out != nullptr;
      const bool want_err = want_y && sp.want_err != 0;
      float oc = 0.f, ov = 0.f, ecp = 0.f, ecv = 0.f, ocpv = 0.f;
      if (want_y) { oc = sp.out.cpa[n_a]; ov = sp.out.cva[n_a]; ecp = sp.err.cpa[n_a]; ecv = sp.err.cva[n_a]; ocpv = sp.out.cpv; }
#pragma unroll 1
      for (int qd = 0; qd < 2; ++qd) {      // 4 float4 groups (16 dims) per pass
        uint32_t r[16];
        tmem_ld16(c.tmem + c.lane_sel + f2::RX + (uint32_t)(c.hf * 32 + qd * 16), r);
        tmem_ld_wait();
        const int fq = f0 + qd * 4;
        float ao[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 b4 = *reinterpret_cast<const float4*>(tab + f2::T_BOUT + 4 * (fq + j));
          ao[4 * j] = __uint_as_float(r[4 * j]) + b4.x; ao[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b4.y;
          ao[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b4.z; ao[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b4.w;
        }
        if (sp.a_out != nullptr) {      // padding rows of the last tile are written as zeros: output buffers need no initialisation
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 v4 = valid ? make_float4(ao[4 * j], ao[4 * j + 1], ao[4 * j + 2], ao[4 * j + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (L2POL && si + 1 < a.n_stage) st_l2hint(blk4(sp.a_out, tile, AF4, fq + j, c.row), v4, pol_keep);
            else *blk4(sp.a_out, tile, AF4, fq + j, c.row) = v4;
          }
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
              const float av = ao[4 * j + e];
              po[4 * j + e] = pb[e] + ocpv * vb[e] + oc * av;
              vo[4 * j + e] = vb[e] + ov * av;
              ep[4 * j + e] = ecp * av;
              ev[4 * j + e] = ecv * av;
            }
          }
#pragma unroll 1
          for (int s = 0; s < n_a; ++s) {
            float4 x[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {      // the launch's last use of a_j (only the final stage produces y_out): demote the line
              const float4* ptr = blk4(a.a[s], tile, AF4, fq + j, c.row);
              x[j] = L2POL ? ld_l2hint(ptr, pol_drop) : *ptr;
            }
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
struct StageFwd2Host {   // mirrors ab200_stage_desc in the public header
  int32_t n_a;
  float in_cpv, in_cpa[MAX_A], in_cva[MAX_A];
  float t;
  float out_cpv, out_cpa[MAX_A + 1], out_cva[MAX_A + 1];
  float err_pa[MAX_A + 1], err_va[MAX_A + 1];
  float rtol, atol;
};
static_assert(sizeof(StageFwd2Host) == sizeof(ab200_stage_desc), "stage descriptor layout");

int stage_fwd2_tc_multi(const ab200_drift_desc* d, const uint8_t* image, const float* y0, const float* const* a_ptrs, const void* descs_v,
                        int n_stage, float* const* a_outs, int64_t B, float* y_out, double* err_sumsq, void* const* x1_outs, int save_level,
                        cudaStream_t st) {
  const StageFwd2Host* hs = reinterpret_cast<const StageFwd2Host*>(descs_v);
  if (n_stage < 1 || n_stage > MAX_A) return AB200_ERR_BAD_ARG;
  StageFwd2Args k{};
  k.wimg = image + IMG2_OFFSET;
  k.y0 = y0;
  int max_a = 0;
  for (int s = 0; s < n_stage; ++s) {
    const StageFwd2Host& h = hs[s];
    if (h.n_a < 0 || h.n_a > MAX_A) return AB200_ERR_BAD_ARG;
    max_a = h.n_a > max_a ? h.n_a : max_a;
    Stage2Params& sp = k.st[s];
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
    sp.x1_out = x1_outs ? (uint8_t*)x1_outs[s] : nullptr;
    if (save_level >= 2 && sp.x1_out == nullptr) return AB200_ERR_BAD_ARG;      // the activation-saving kernel saves every stage
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
  const int grid = k.ntiles < sms ? k.ntiles : sms;      // a partial wave uses one slot per CTA first
  auto kern = save_level >= 2 ? stage_fwd2_tc_kernel<true> : stage_fwd2_tc_kernel<false>;
  if (k.flags & 128) kern = save_level >= 2 ? stage_fwd2_tc_kernel<true, true> : stage_fwd2_tc_kernel<false, true>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f2::SMEM_BYTES);
  if (e != cudaSuccess) { set_cuda_error(e); return AB200_ERR_CUDA; }
  kern<<<grid, THREADS, f2::SMEM_BYTES, st>>>(k);
  return check_launch();
}

}  // namespace ab200
