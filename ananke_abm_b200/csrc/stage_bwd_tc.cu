// Tensor-core backward STAGE kernel: the vector-Jacobian product of one drift evaluation (see stage_fwd_tc.cu).
//
// Given dL/da_out for a stage whose input was  p = p0 + cpv v0 + sum cpa[j] a_j,  v = v0 + sum cva[j] a_j, it
//   1. rebuilds the stage input and re-runs the five hidden layers on the tensor core, keeping only the ReLU masks
//      (bits) and the residual stream in registers,
//   2. back-propagates through the net with six dgrad GEMMs that read the SAME resident weight image MN-major
//      (W^T without a transposed copy),
//   3. WRITES dL/d(stage input) = gx = [g_p, g_v, g_h] (no read-modify-write anywhere: the upstream gradient of a
//      stage is itself assembled in the prologue as a linear combination
//         dL/da_out = g_base + sum_l dp[l] gx_l.p + dv[l] gx_l.v
//      of the step-level term g_base and the gx of the LATER stages that consumed a_out, mirroring how the forward
//      stage input is assembled from earlier stages; `ab200_adjoint_gather` folds all gx into dL/dy0 at the end),
//   4. writes every layer's (input activation, output gradient) pair as bf16 "blobs" already laid out as the
//      canonical MN-major UMMA operand image (K = agent), which wgrad_tc.cu streams straight into shared memory.
// All per-agent fp32 buffers are tile-blocked (stage_tc.cuh), so every global access of a warp is one contiguous
// 512-byte segment.  The weight gradient cannot be fused here: its fp32 accumulators (95,168 floats = 372 KiB)
// exceed the 256 KiB of tensor memory of one SM, see DESIGN.md.
#include "stage_tc.cuh"
#include "wgrad_layout.cuh"

namespace ab200 {
using namespace stc;

#ifdef AB200_STAGE_TRACE
extern "C" int ab200_debug_stage_trace_bwd(long long* host_out, int* counts) {
  cudaMemcpyFromSymbol(host_out, stc::g_stage_trace, sizeof(long long) * 8192);
  cudaMemcpyFromSymbol(counts, stc::g_stage_trace_n, sizeof(int) * 2);
  int z[2] = {0, 0};
  cudaMemcpyToSymbol(stc::g_stage_trace_n, z, sizeof(z));
  return 0;
}
#endif

struct BwdStageParams {       // one stage of a fused (latest-first) sequence
  int n_a;                    // the stage input combined a[0 .. n_a)
  Combo in;
  float t;
  const float* g_base;        // blocked [Bp][64] or null: step-level part of dL/da_out
  int n_g;                    // later stages (earlier entries of this launch, or buffers of earlier launches) that read a_out
  const float* gx[MAX_A];     // their dL/d(stage input), blocked [Bp][160]
  float dp[MAX_A], dv[MAX_A]; // and coefficients (that stage's in.cpa / in.cva entry for a_out)
  float* gx_out;              // blocked [Bp][160]: [g_p, g_v, g_h] of this stage (written)
  int blob0;                  // blob index of tile 0 for this stage
  const uint8_t* x1_in;       // what the forward launch saved for this stage (wg::FwdSaveLayout) or null: rebuild the input from y0 / a_j
  int saved_acts;             // 1: x1_in also holds the hidden activations and the ReLU masks -> nothing is recomputed or re-spilled
  // GATHER entry (gx_out == null; last of a launch; no GEMMs, no blobs): from the p / v / h parts of its sources' gx it writes
  //   ga_out (if non-null, fp32 blocked [Bp][64]) = g_base + sum_l dp[l] gx_l.p + dv[l] gx_l.v     (the gradient handed to the FSAL
  //                                                                                                evaluation of the previous step)
  //   y0_acc (StageBwdArgs, if non-null)         += sum_l [gx_l.p ; cpv_src[l] gx_l.p + gx_l.v ; gx_l.h]     (dL/dy0 of the step)
  // The sources were written by this same thread moments ago, so they are read from L2, not from DRAM.
  int gather;
  float* ga_out;
  float cpv_src[MAX_A];
};

struct StageBwdArgs {
  const uint8_t* wimg;
  const float* y0;            // blocked [Bp][160]
  const float* a[MAX_A];      // blocked [Bp][64]
  int n_stage;
  BwdStageParams st[MAX_A + 1];
  float period;
  uint8_t* spill;             // blob buffer (SpillLayout)
  float* g_bout;              // [64] atomically accumulated column sums of dL/da_out (bias gradient of the output layer)
  float* y0_acc;              // or null: blocked [Bp][160], updated in place by the gather entry
  int64_t B;
  int ntiles;
  int nblobs;                 // blobs the spill buffer was sized for
  int flags;
  int* status;
};

// store packed pairs as feature groups of a blob: NG groups starting at fg0, row = agent
template <int NG>
__device__ __forceinline__ void spill_groups(uint8_t* blob, int fg0, int row, const uint32_t* o) {
  if (blob == nullptr) return;     // AB200_STAGE_FLAGS bit 16: timing experiment without the blob spill
#pragma unroll
  for (int q = 0; q < NG; ++q)
    __stcs(reinterpret_cast<uint4*>(blob + (size_t)(fg0 + q) * wg::FG_BYTES + (size_t)row * 16),
           make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]));     // streaming: blobs must not evict what the next stage re-reads
}

// forward hidden epilogue with mask capture and spill.  RES: residual add of z.  Result -> ACT (if TO_ACT), z (if KEEP),
// blob.  The last hidden layer must NOT touch ACT: its columns are about to receive the output-layer gradient from
// the other column-half's warp, and nothing orders the two stores.
template <bool RES, bool KEEP, bool TO_ACT = true>
__device__ __forceinline__ void bwd_fwd_epi(const SlotCtx& c, uint32_t (&z)[32], uint32_t (&mask)[2], uint8_t* blob) {
  // 16-column passes (not 32): the backward kernel lives at the 128-register cap and every live register counts
  mask[0] = mask[1] = 0u;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t r[16];
    tmem_ld16(c.tmem + c.lane_sel + C_ACC + (uint32_t)(c.hf * 64 + q * 16), r);
    tmem_ld_wait();
    uint32_t o[8];
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int zi = q * 8 + j;
      float x0 = __uint_as_float(r[2 * j]), x1 = __uint_as_float(r[2 * j + 1]);
      if (RES) { x0 += bf16lo(z[zi]); x1 += bf16hi(z[zi]); }
      o[j] = pack_relu_bf16(x0, x1);
      // ReLU mask bits on the whole 32-bit word: bit 15 / 31 of t is set iff the low / high bf16 is non-zero;
      // pair jj = (q % 2) * 8 + j of mask word q / 2 lands at bits (15 - jj) and (31 - jj)
      const uint32_t t = ((o[j] & 0x7FFF7FFFu) + 0x7FFF7FFFu) & 0x80008000u;
      m |= t >> ((q & 1) * 8 + j);
      if (KEEP) z[zi] = o[j];
    }
    mask[q >> 1] |= m;
    if (TO_ACT) tmem_st8(c.tmem + c.lane_sel + C_ACT + (uint32_t)(c.hf * 32 + q * 8), o);
    spill_groups<2>(blob, c.hf * 8 + q * 2, c.row, o);
  }
}

// backward hidden epilogue: g = (acc [+ skip]) * mask -> ACT, blob; SKIP_IN adds the skip-path gradient held in
// `gs`, SKIP_OUT stores the result into `gs` (it is the gradient of the residual stream one block further down).
template <bool SKIP_IN, bool SKIP_OUT>
__device__ __forceinline__ void bwd_bwd_epi(const SlotCtx& c, uint32_t (&gs)[32], const uint32_t (&mask)[2], uint8_t* blob) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t r[16];
    tmem_ld16(c.tmem + c.lane_sel + C_ACC + (uint32_t)(c.hf * 64 + q * 16), r);
    tmem_ld_wait();
    uint32_t o[8];
    const uint32_t m = mask[q >> 1];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int zi = q * 8 + j, jj = (q & 1) * 8 + j;
      float x0 = __uint_as_float(r[2 * j]), x1 = __uint_as_float(r[2 * j + 1]);
      if (SKIP_IN) { x0 += bf16lo(gs[zi]); x1 += bf16hi(gs[zi]); }
      x0 = (m & (0x8000u >> jj)) ? x0 : 0.0f;
      x1 = (m & (0x80000000u >> jj)) ? x1 : 0.0f;
      o[j] = pack_bf16(x0, x1);
      if (SKIP_OUT) gs[zi] = o[j];
    }
    tmem_st8(c.tmem + c.lane_sel + C_ACT + (uint32_t)(c.hf * 32 + q * 8), o);
    spill_groups<2>(blob, c.hf * 8 + q * 2, c.row, o);
  }
}

// dL/da_out of a stage for this thread's 32 columns: g_base + sum_l dp[l] gx_l.p + dv[l] gx_l.v
template <bool L2POL>
__device__ __forceinline__ void upstream_gather(const BwdStageParams& sp, int tile, const SlotCtx& c, float (&gv)[32], uint64_t pol_keep) {
  // all loads of one source are issued before any is consumed: a later stage's gx usually comes from L2, and a
  // load -> FMA -> load chain per float4 (as a naive loop nest gives) exposes that latency 8 x n_g times per tile
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sp.g_base != nullptr) x = ldro(blk4(sp.g_base, tile, AF4, c.hf * 8 + j, c.row));     // padding rows are zero
    gv[4 * j] = x.x; gv[4 * j + 1] = x.y; gv[4 * j + 2] = x.z; gv[4 * j + 3] = x.w;
  }
#pragma unroll 1
  for (int s = 0; s < sp.n_g; ++s) {
    const float dp = sp.dp[s], dv = sp.dv[s];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float4 gp[4], gq[4];     // coherent loads: an earlier stage of THIS launch (same thread) may have written these
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4* pp = blk4(sp.gx[s], tile, YF4, c.hf * 8 + half * 4 + j, c.row);
        const float4* pq = blk4(sp.gx[s], tile, YF4, AF4 + c.hf * 8 + half * 4 + j, c.row);
        gp[j] = L2POL ? ld_l2hint(pp, pol_keep) : *pp;
        gq[j] = L2POL ? ld_l2hint(pq, pol_keep) : *pq;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int b = (half * 4 + j) * 4;
        gv[b] += dp * gp[j].x + dv * gq[j].x; gv[b + 1] += dp * gp[j].y + dv * gq[j].y;
        gv[b + 2] += dp * gp[j].z + dv * gq[j].z; gv[b + 3] += dp * gp[j].w + dv * gq[j].w;
      }
    }
  }
}

template <bool L2POL>
__global__ void __launch_bounds__(THREADS, 1) stage_bwd_tc_kernel(const __grid_constant__ StageBwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[NSLOT + 1];
  __shared__ uint32_t tmem_base_s;
  __shared__ int issue_lock;
  SlotCtx c = stage_setup(smem, a.wimg, bars, &tmem_base_s, &issue_lock, a.status, a.flags);
  const uint32_t tmem_base = tmem_base_s;
  const wg::SpillLayout S{a.nblobs};
  const int lane = threadIdx.x & 31;
  const bool no_spill = (c.flags & 16) != 0, no_gx = (c.flags & 32) != 0;     // timing experiments only (results invalid)
  auto blob_at = [&](size_t off) -> uint8_t* { return no_spill ? nullptr : a.spill + off; };
  uint64_t pol_keep = 0, pol_drop = 0;      // L2POL: gx tiles of the launch kept in L2 until the gather entry has folded them
  if (L2POL) { pol_keep = l2_policy_keep(); pol_drop = l2_policy_drop(); }

#pragma unroll 1
  for (int it = 0;; ++it) {
    const int tile = slot_tile(it, a.ntiles, c.slot);
    if (tile < 0) break;
    const bool valid = (int64_t)tile * TM + c.row < a.B && !no_gx;      // padding rows: zeros in, nothing stored
#pragma unroll 1
   for (int si = 0; si < a.n_stage; ++si) {                   // all stages of the step for this tile: later stages' gx come from L2
    const BwdStageParams& sp = a.st[si];
    const int blob = sp.blob0 + tile;
    STAGE_TRACE(c, 9);
    if (sp.gather) {
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {      // 16 of this thread's 32 columns of p and of v per pass (register budget)
        const int f0 = c.hf * 8 + half * 4;
        float xa[16], xp[16], xv[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4 b = make_float4(0.f, 0.f, 0.f, 0.f), p = b, v = b;
          if (sp.g_base != nullptr && sp.ga_out != nullptr) b = ldro(blk4(sp.g_base, tile, AF4, f0 + j, c.row));
          if (a.y0_acc != nullptr) { p = *blk4(a.y0_acc, tile, YF4, f0 + j, c.row); v = *blk4(a.y0_acc, tile, YF4, AF4 + f0 + j, c.row); }
          xa[4 * j] = b.x; xa[4 * j + 1] = b.y; xa[4 * j + 2] = b.z; xa[4 * j + 3] = b.w;
          xp[4 * j] = p.x; xp[4 * j + 1] = p.y; xp[4 * j + 2] = p.z; xp[4 * j + 3] = p.w;
          xv[4 * j] = v.x; xv[4 * j + 1] = v.y; xv[4 * j + 2] = v.z; xv[4 * j + 3] = v.w;
        }
#pragma unroll 1
        for (int s = 0; s < sp.n_g; ++s) {
          const float dp = sp.dp[s], dv = sp.dv[s], cs = sp.cpv_src[s];
          float4 gp[4], gq[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4* pp = blk4(sp.gx[s], tile, YF4, f0 + j, c.row);
            const float4* pq = blk4(sp.gx[s], tile, YF4, AF4 + f0 + j, c.row);
            gp[j] = L2POL ? ld_l2hint(pp, pol_drop) : *pp;      // last use of the tile's gx: demote
            gq[j] = L2POL ? ld_l2hint(pq, pol_drop) : *pq;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float pe[4] = {gp[j].x, gp[j].y, gp[j].z, gp[j].w}, ve[4] = {gq[j].x, gq[j].y, gq[j].z, gq[j].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              xa[4 * j + e] += dp * pe[e] + dv * ve[e];
              xp[4 * j + e] += pe[e];
              xv[4 * j + e] += cs * pe[e] + ve[e];
            }
          }
        }
        if (valid) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (sp.ga_out != nullptr)
              *blk4(sp.ga_out, tile, AF4, f0 + j, c.row) = make_float4(xa[4 * j], xa[4 * j + 1], xa[4 * j + 2], xa[4 * j + 3]);
            if (a.y0_acc != nullptr) {
              *blk4(a.y0_acc, tile, YF4, f0 + j, c.row) = make_float4(xp[4 * j], xp[4 * j + 1], xp[4 * j + 2], xp[4 * j + 3]);
              *blk4(a.y0_acc, tile, YF4, AF4 + f0 + j, c.row) = make_float4(xv[4 * j], xv[4 * j + 1], xv[4 * j + 2], xv[4 * j + 3]);
            }
          }
        }
      }
      if (a.y0_acc != nullptr) {      // context part: 16 of 32 columns per thread
        float4 xh[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) xh[j] = *blk4(a.y0_acc, tile, YF4, 2 * AF4 + c.hf * 4 + j, c.row);
#pragma unroll 1
        for (int s = 0; s < sp.n_g; ++s) {
          float4 g[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) g[j] = *blk4(sp.gx[s], tile, YF4, 2 * AF4 + c.hf * 4 + j, c.row);
#pragma unroll
          for (int j = 0; j < 4; ++j) { xh[j].x += g[j].x; xh[j].y += g[j].y; xh[j].z += g[j].z; xh[j].w += g[j].w; }
        }
        if (valid) {
#pragma unroll
          for (int j = 0; j < 4; ++j) *blk4(a.y0_acc, tile, YF4, 2 * AF4 + c.hf * 4 + j, c.row) = xh[j];
        }
      }
      continue;
    }

    uint32_t z[32];
    uint32_t m_z0[2], m_u0[2], m_z1[2], m_u1[2], m_z2[2];
    if (sp.saved_acts) {
      // The training forward saved every layer input of this evaluation as blobs and the ReLU masks as bits: the recompute
      // half of this kernel (stage input, five GEMMs, five epilogues, 1.6 KB of spills per agent) reduces to loading 40 bytes.
      const uint2* mp = reinterpret_cast<const uint2*>(sp.x1_in + wg::FwdSaveLayout{a.ntiles}.mask(tile)) + (c.hf * TM + c.row);
      const uint2 q0 = __ldg(mp), q1 = __ldg(mp + 256), q2 = __ldg(mp + 512), q3 = __ldg(mp + 768), q4 = __ldg(mp + 1024);
      m_z0[0] = q0.x; m_z0[1] = q0.y; m_u0[0] = q1.x; m_u0[1] = q1.y; m_z1[0] = q2.x; m_z1[1] = q2.y;
      m_u1[0] = q3.x; m_u1[1] = q3.y; m_z2[0] = q4.x; m_z2[1] = q4.y;
    } else {
    // ---- stage input -> ACT / HB / TB and the X blob (features: p 0..63, v 64..127, h 128..159, sin, cos, 1)
    if (sp.x1_in != nullptr) {
      // The forward launch of this step already wrote the stage input as the bf16 X blob (the very operand image the recompute
      // and the weight-gradient kernel need): 10 independent 16-byte loads per thread instead of y0 plus up to six a_j
      // (1.5 KB -> 0.35 KB per agent-stage), and nothing to spill.
      const uint8_t* xin = sp.x1_in + (size_t)tile * wg::X1_BYTES + (size_t)c.row * 16;
      uint4 pg[4], vg[4], hg[2];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        pg[q] = __ldg(reinterpret_cast<const uint4*>(xin + (size_t)(c.hf * 4 + q) * wg::FG_BYTES));
        vg[q] = __ldg(reinterpret_cast<const uint4*>(xin + (size_t)(P / 8 + c.hf * 4 + q) * wg::FG_BYTES));
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) hg[q] = __ldg(reinterpret_cast<const uint4*>(xin + (size_t)(2 * P / 8 + c.hf * 2 + q) * wg::FG_BYTES));
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const uint32_t op[8] = {pg[2 * ch].x, pg[2 * ch].y, pg[2 * ch].z, pg[2 * ch].w, pg[2 * ch + 1].x, pg[2 * ch + 1].y, pg[2 * ch + 1].z, pg[2 * ch + 1].w};
        const uint32_t ov[8] = {vg[2 * ch].x, vg[2 * ch].y, vg[2 * ch].z, vg[2 * ch].w, vg[2 * ch + 1].x, vg[2 * ch + 1].y, vg[2 * ch + 1].z, vg[2 * ch + 1].w};
        tmem_st8(c.tmem + c.lane_sel + C_ACT + (uint32_t)((c.hf * 8 + ch * 4) * 2), op);
        tmem_st8(c.tmem + c.lane_sel + C_ACT + (uint32_t)(P / 2 + (c.hf * 8 + ch * 4) * 2), ov);
      }
      const uint32_t oh[8] = {hg[0].x, hg[0].y, hg[0].z, hg[0].w, hg[1].x, hg[1].y, hg[1].z, hg[1].w};
      tmem_st8(c.tmem + c.lane_sel + C_HB + (uint32_t)(c.hf * 8), oh);
      write_time_block(c, sp.t, a.period);
    } else {
    uint8_t* xb = blob_at(S.x1(blob));
    {   // both 16-dim halves together: one batch of independent 128-bit loads per source (n_a + 1 round trips, not 2x that)
      const int f0 = c.hf * 8;
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
      for (int s = 0; s < sp.n_a; ++s) {
        const float cp = sp.in.cpa[s], cv = sp.in.cva[s];
        float4 x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = ldro(blk4(a.a[s], tile, AF4, f0 + j, c.row));
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          pin[4 * j] += cp * x[j].x; pin[4 * j + 1] += cp * x[j].y; pin[4 * j + 2] += cp * x[j].z; pin[4 * j + 3] += cp * x[j].w;
          vin[4 * j] += cv * x[j].x; vin[4 * j + 1] += cv * x[j].y; vin[4 * j + 2] += cv * x[j].z; vin[4 * j + 3] += cv * x[j].w;
        }
      }
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const int fc = f0 + ch * 4;
        uint32_t o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = pack_bf16(pin[16 * ch + 2 * j], pin[16 * ch + 2 * j + 1]);
        tmem_st8(c.tmem + c.lane_sel + C_ACT + (uint32_t)(fc * 2), o);
        spill_groups<2>(xb, fc / 2, c.row, o);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = pack_bf16(vin[16 * ch + 2 * j], vin[16 * ch + 2 * j + 1]);
        tmem_st8(c.tmem + c.lane_sel + C_ACT + (uint32_t)(P / 2 + fc * 2), o);
        spill_groups<2>(xb, P / 8 + fc / 2, c.row, o);
      }
    }
    {
      uint32_t o[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 x = ldro(blk4(a.y0, tile, YF4, 2 * AF4 + c.hf * 4 + j, c.row));
        o[2 * j] = pack_bf16(x.x, x.y);
        o[2 * j + 1] = pack_bf16(x.z, x.w);
      }
      tmem_st8(c.tmem + c.lane_sel + C_HB + (uint32_t)(c.hf * 8), o);
      spill_groups<2>(xb, 2 * P / 8 + c.hf * 2, c.row, o);
    }
    write_time_block(c, sp.t, a.period);
    if (c.hf == 0) {   // time-feature / bias feature groups of the X blob: [sin, cos, 1, 0...], [0...]
      float s, co;
      time_features(sp.t, a.period, s, co);
      const uint32_t o[8] = {pack_bf16(s, co), pack_bf16(1.0f, 0.0f), 0u, 0u, 0u, 0u, 0u, 0u};
      spill_groups<2>(xb, (2 * P + H) / 8, c.row, o);
    }
    }

    STAGE_TRACE(c, 10);
    // ---- forward recompute (hidden layers only), masks + blobs
    run_layer<false, (2 * P + H) / 16, true, HID, HID, true>(c, C_ACT, OFF_W1);
    bwd_fwd_epi<false, true>(c, z, m_z0, blob_at(S.act(0, blob)));
    run_layer<false, HID / 16, true, HID, HID, true>(c, C_ACT, off_hh(0));
    bwd_fwd_epi<false, false>(c, z, m_u0, blob_at(S.act(1, blob)));
    run_layer<false, HID / 16, true, HID, HID, true>(c, C_ACT, off_hh(1));
    bwd_fwd_epi<true, true>(c, z, m_z1, blob_at(S.act(2, blob)));
    run_layer<false, HID / 16, true, HID, HID, true>(c, C_ACT, off_hh(2));
    bwd_fwd_epi<false, false>(c, z, m_u1, blob_at(S.act(3, blob)));
    run_layer<false, HID / 16, true, HID, HID, true>(c, C_ACT, off_hh(3));
    bwd_fwd_epi<true, false, false>(c, z, m_z2, blob_at(S.act(4, blob)));
    }

    STAGE_TRACE(c, 12);
    // ---- upstream gradient of the output layer -> ACT (K = 64) + gO blob + bias column sums
    {
      uint32_t o[16];
      float gv[32];
      upstream_gather<L2POL>(sp, tile, c, gv, pol_keep);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o[2 * j] = pack_bf16(gv[4 * j], gv[4 * j + 1]);
        o[2 * j + 1] = pack_bf16(gv[4 * j + 2], gv[4 * j + 3]);
      }
      tmem_st16(c.tmem + c.lane_sel + C_ACT + (uint32_t)(c.hf * 16), o);
      spill_groups<4>(blob_at(S.go(blob)), c.hf * 4, c.row, o);
      // column sums over the warp's 32 agents: transpose-reduce, lane j ends up with column j
#pragma unroll
      for (int w = 16; w >= 1; w >>= 1) {
#pragma unroll
        for (int j = 0; j < w; ++j) {
          const bool up = (lane & w) != 0;
          const float send = up ? gv[j] : gv[j + w];
          const float keep = up ? gv[j + w] : gv[j];
          gv[j] = keep + __shfl_xor_sync(0xffffffffu, send, w);
        }
      }
      atomicAdd(a.g_bout + c.hf * 32 + lane, gv[0]);
    }

    STAGE_TRACE(c, 13);
    // ---- backward through the net (dgrad GEMMs on the MN-major view of the weight image)
    run_layer<true, P / 16, false, P, HID, true>(c, C_ACT, OFF_WO);                       // g_z2 = gO W_O
    bwd_bwd_epi<false, true>(c, z, m_z2, blob_at(S.grad(4, blob)));                // gB1 = g_z2 * [z2 > 0]  (skip -> z)
    run_layer<true, HID / 16, false, HID, HID, true>(c, C_ACT, off_hh(3));                // g_u1 = gB1 W_B1
    {
      uint32_t dummy[32];
      bwd_bwd_epi<false, false>(c, dummy, m_u1, blob_at(S.grad(3, blob)));         // gA1
    }
    run_layer<true, HID / 16, false, HID, HID, true>(c, C_ACT, off_hh(2));                // gA1 W_A1 (+ skip)
    bwd_bwd_epi<true, true>(c, z, m_z1, blob_at(S.grad(2, blob)));                 // gB0
    run_layer<true, HID / 16, false, HID, HID, true>(c, C_ACT, off_hh(1));
    {
      uint32_t dummy[32];
      bwd_bwd_epi<false, false>(c, dummy, m_u0, blob_at(S.grad(1, blob)));         // gA0
    }
    run_layer<true, HID / 16, false, HID, HID, true>(c, C_ACT, off_hh(0));
    bwd_bwd_epi<true, false>(c, z, m_z0, blob_at(S.grad(0, blob)));                // g1
    run_layer<true, HID / 16, false, HID, 2 * P + H, true>(c, C_ACT, OFF_W1);             // g_x[128 x 160] = g1 W_1[:, :160]

    STAGE_TRACE(c, 11);
    // ---- dL/d(stage input) -> gx_out (tcgen05.ld is warp-collective: issued by every lane; only lanes that own a real
    //      agent store)
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {     // 16 dims per pass
      const int f0 = c.hf * 8 + ch * 4;
      uint32_t rp[16], rv[16];
      tmem_ld16(c.tmem + c.lane_sel + C_ACC + (uint32_t)(f0 * 4), rp);
      tmem_ld16(c.tmem + c.lane_sel + C_ACC + (uint32_t)(P + f0 * 4), rv);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 vp = make_float4(__uint_as_float(rp[4 * j]), __uint_as_float(rp[4 * j + 1]), __uint_as_float(rp[4 * j + 2]),
                                        __uint_as_float(rp[4 * j + 3]));
          const float4 vv = make_float4(__uint_as_float(rv[4 * j]), __uint_as_float(rv[4 * j + 1]), __uint_as_float(rv[4 * j + 2]),
                                        __uint_as_float(rv[4 * j + 3]));
          if (L2POL) {
            st_l2hint(blk4(sp.gx_out, tile, YF4, f0 + j, c.row), vp, pol_keep);
            st_l2hint(blk4(sp.gx_out, tile, YF4, AF4 + f0 + j, c.row), vv, pol_keep);
          } else {
            *blk4(sp.gx_out, tile, YF4, f0 + j, c.row) = vp;
            *blk4(sp.gx_out, tile, YF4, AF4 + f0 + j, c.row) = vv;
          }
        }
      }
    }
    {
      uint32_t rh[16];
      tmem_ld16(c.tmem + c.lane_sel + C_ACC + (uint32_t)(2 * P + c.hf * 16), rh);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *blk4(sp.gx_out, tile, YF4, 2 * AF4 + c.hf * 4 + j, c.row) =
              make_float4(__uint_as_float(rh[4 * j]), __uint_as_float(rh[4 * j + 1]), __uint_as_float(rh[4 * j + 2]),
                          __uint_as_float(rh[4 * j + 3]));
      }
    }
   }
  }
  stage_teardown(tmem_base);
}

struct StageBwdHost {   // mirrors the head of ab200_stage_desc
  int32_t n_a;
  float in_cpv, in_cpa[MAX_A], in_cva[MAX_A];
  float t;
  float rest[4 * (MAX_A + 1) + 3];      // out_* / err_* / rtol / atol of ab200_stage_desc (unused here)
};
static_assert(sizeof(StageBwdHost) == sizeof(ab200_stage_desc), "stage descriptor layout");

// Fused sequence, latest stage first.  Stage s: upstream gradient = g_base[s] + sum over its n_g[s] sources; a source is
// gx_src[s][l] >= 0 -> the gx_out of entry gx_src[s][l] of THIS sequence, or < 0 -> external buffer gx_ext[-1 - idx].
int stage_bwd_tc_multi(const ab200_drift_desc* d, const uint8_t* image, const float* y0, const float* const* a_ptrs,
                       const void* descs_v, int n_stage, const float* const* g_base, float* const* gx_out, const int32_t* n_g,
                       const int32_t* gx_src, const float* const* gx_ext, const float* dp, const float* dv, int64_t B, void* spill,
                       int blob0, int nblobs, float* g_bout, const void* const* x1_in, int save_level, float* y0_acc,
                       float* const* ga_out, cudaStream_t st) {
  const StageBwdHost* hs = reinterpret_cast<const StageBwdHost*>(descs_v);
  if (n_stage < 1 || n_stage > MAX_A + 1) return AB200_ERR_BAD_ARG;
  StageBwdArgs k{};
  k.wimg = image;
  k.y0 = y0;
  k.ntiles = (int)((B + TM - 1) / TM);
  int max_a = 0;
  for (int s = 0; s < n_stage; ++s) {
    const StageBwdHost& h = hs[s];
    const bool gather = gx_out[s] == nullptr;      // the gather entry comes last and writes no blobs
    float* ga = (gather && ga_out != nullptr) ? ga_out[s] : nullptr;
    if (gather && (s != n_stage - 1 || s == 0 || (ga == nullptr && y0_acc == nullptr))) return AB200_ERR_BAD_ARG;
    if (!gather && s >= MAX_A) return AB200_ERR_BAD_ARG;
    if (h.n_a < 0 || h.n_a > MAX_A || n_g[s] < 0 || n_g[s] > MAX_A) return AB200_ERR_BAD_ARG;
    if (n_g[s] == 0 && !g_base[s]) return AB200_ERR_BAD_ARG;
    max_a = h.n_a > max_a ? h.n_a : max_a;
    BwdStageParams& sp = k.st[s];
    sp.n_a = h.n_a;
    sp.in.cpv = h.in_cpv;
    for (int i = 0; i < MAX_A; ++i) { sp.in.cpa[i] = h.in_cpa[i]; sp.in.cva[i] = h.in_cva[i]; }
    sp.t = h.t;
    sp.g_base = g_base[s];
    sp.n_g = n_g[s];
    for (int l = 0; l < n_g[s]; ++l) {
      const int src = gx_src[s * MAX_A + l];
      if (src >= s) return AB200_ERR_BAD_ARG;                      // only stages processed earlier in this sequence
      if (gather && src < 0) return AB200_ERR_BAD_ARG;            // a gather entry folds stages of its own launch
      sp.gx[l] = src >= 0 ? gx_out[src] : gx_ext[-1 - src];
      sp.cpv_src[l] = src >= 0 ? hs[src].in_cpv : 0.f;
      sp.dp[l] = dp[s * MAX_A + l];
      sp.dv[l] = dv[s * MAX_A + l];
    }
    sp.gx_out = gx_out[s];
    sp.gather = gather ? 1 : 0;
    sp.ga_out = ga;
    sp.blob0 = blob0 + s * k.ntiles;
    sp.x1_in = x1_in ? (const uint8_t*)x1_in[s] : nullptr;
    sp.saved_acts = (sp.x1_in != nullptr && save_level >= 2) ? 1 : 0;
  }
  const int n_real = n_stage - (gx_out[n_stage - 1] == nullptr ? 1 : 0);
  if (y0_acc != nullptr && n_real == n_stage) return AB200_ERR_BAD_ARG;      // y0_accum is the gather entry's job
  if (blob0 < 0 || blob0 + n_real * k.ntiles > nblobs) return AB200_ERR_BAD_ARG;
  for (int i = 0; i < MAX_A; ++i) k.a[i] = (i < max_a) ? a_ptrs[i] : nullptr;
  k.n_stage = n_stage;
  k.period = d->time_period;
  k.spill = (uint8_t*)spill;
  k.g_bout = g_bout;
  k.y0_acc = y0_acc;
  k.B = B;
  k.nblobs = nblobs;
  k.flags = stage_flags();
  k.status = stage_status_ptr(image);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = k.ntiles < sms ? k.ntiles : sms;      // a partial wave uses one slot per CTA first (slot_tile)
  // L2 eviction hints on the launch's gx tiles (stage_tc.cuh): measured 1,231 -> 1,204 us per fused 6-stage launch over 250,112 agents,
  // identical results; on by default, AB200_STAGE_FLAGS bit 4 turns them off
  auto kern = (k.flags & 4) ? stage_bwd_tc_kernel<false> : stage_bwd_tc_kernel<true>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)W_BYTES);
  if (e != cudaSuccess) { set_cuda_error(e); return AB200_ERR_CUDA; }
  kern<<<grid, THREADS, W_BYTES, st>>>(k);
  return check_launch();
}

int stage_bwd_tc(const ab200_drift_desc* d, const uint8_t* image, const float* y0, const float* const* a_ptrs, const void* desc_v,
                 int64_t B, const float* g_base, const float* const* gx_ptrs, int n_g, const float* dp, const float* dv, float* gx_out,
                 void* spill, int blob0, int nblobs, float* g_bout, cudaStream_t st) {
  if (n_g < 0 || n_g > MAX_A) return AB200_ERR_BAD_ARG;
  const float* gb[1] = {g_base};
  float* go[1] = {gx_out};
  int32_t ng[1] = {n_g};
  int32_t src[MAX_A];
  float dpa[MAX_A], dva[MAX_A];
  for (int l = 0; l < MAX_A; ++l) { src[l] = -1 - l; dpa[l] = l < n_g ? dp[l] : 0.f; dva[l] = l < n_g ? dv[l] : 0.f; }
  return stage_bwd_tc_multi(d, image, y0, a_ptrs, desc_v, 1, gb, go, ng, src, gx_ptrs, dpa, dva, B, spill, blob0, nblobs, g_bout, nullptr, 0, nullptr, nullptr, st);
}

}  // namespace ab200
