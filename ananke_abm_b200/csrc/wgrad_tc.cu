// Weight-gradient kernel of the tensor-core backward pass: dW_l = sum over agents of g_l^T a_l for the six layers,
// streamed from the bf16 blobs stage_bwd_tc.cu wrote (wgrad_layout.cuh).  HBM-bound by construction: each blob pair
// (64-76 KiB) feeds 8 tcgen05 K-steps (512 cycles), the copy takes ~5x longer.
//
// Warp-specialised, one CTA per SM, each CTA owns a contiguous range of blobs and walks the six layer pairs over it:
//   warp 0  producer : cp.async.bulk (1-D TMA) blob pair -> shared-memory ring, completion on `full` mbarriers
//   warp 1  MMA      : operands are MN-major images with K = agent; both A (gradient, M = out feature) and B
//                      (activation, N = in feature) come from shared memory; fp32 accumulators stay in TMEM for the
//                      CTA's whole blob range; bias gradients ride along as an N = 16 GEMM against a constant block
//                      whose first feature is 1; tcgen05.commit frees the ring slot
//   warps 2-5 epilogue: TMEM -> += CTA-private fp32 partial buffer (plain coalesced RMW, no atomics); a second
//                      accumulator set lets the next pair's MMAs overlap the drain
// `wgrad_finalize` sums the partials over CTAs and writes the gradient in torch's parameter layout.
#include "common.cuh"
#include "umma.cuh"
#include "wgrad_layout.cuh"

namespace ab200 {
using namespace umma;
using namespace wg;

constexpr int WG_THREADS = 192;
constexpr long long WG_WAIT_CYCLES = 400000000LL;    // ~0.2 s: bounded so that a bug cannot hang the device
constexpr int WG_NS = 2;                                   // ring slots
constexpr uint32_t WG_SLOT = HID_BYTES + X1_BYTES;         // 77,824: A blob then B blob
constexpr uint32_t WG_ONES = 2 * FG_BYTES;                 // 16 features x 128 agents
constexpr uint32_t WG_SMEM = WG_NS * WG_SLOT + WG_ONES;
constexpr uint32_t WG_SET_COLS = 256, WG_BIAS_COL = 192;

struct WgradArgs {
  const uint8_t* spill;
  float* partial;      // [gridDim.x][PART_TOTAL]
  int nblobs;          // capacity of the spill buffer (layout stride)
  int used;            // blobs actually filled
  int* status;
  // X blobs (pair 0's B operand) written by the FORWARD launch of the step instead of the backward kernel: one base pointer per ring
  // position (= blob / ntiles, the stage's place in the backward launch), tile-indexed; null = the spill buffer's own x1 region
  const uint8_t* x1_ext[8];
  int ntiles;
  int save_level;      // 2: those buffers (wg::FwdSaveLayout) also hold the hidden activation blobs: every layer INPUT comes from them
};

struct PairDesc { size_t a_off, b_off; uint32_t a_bytes, b_bytes; int N; bool bias; int part_off, part_ld; };

__device__ __forceinline__ PairDesc pair_desc(const SpillLayout& S, int pair, int blob) {
  PairDesc d;
  if (pair == 0) {
    d = {S.grad(0, blob), S.x1(blob), HID_BYTES, X1_BYTES, (int)X1_FEATS, false, PART_W1, PART_W1_N};
  } else if (pair <= 4) {
    d = {S.grad(pair, blob), S.act(pair - 1, blob), HID_BYTES, HID_BYTES, 128, true, PART_HH + (pair - 1) * PART_HH_SZ, PART_HH_N};
  } else {
    d = {S.act(4, blob), S.go(blob), HID_BYTES, GO_BYTES, 64, false, PART_WO, PART_WO_N};
  }
  return d;
}

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ WgradArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[WG_NS], empty[WG_NS], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* ones = smem + WG_NS * WG_SLOT;
  const SpillLayout S{a.nblobs};

  // this CTA's blob range
  const int per = a.used / gridDim.x, rem = a.used % gridDim.x;
  const int b0 = blockIdx.x * per + min((int)blockIdx.x, rem);
  const int b1 = b0 + per + ((int)blockIdx.x < rem ? 1 : 0);

  for (uint32_t i = tid; i < WG_ONES / 4; i += WG_THREADS) {
    // feature 0 of feature group 0 is 1.0 (bf16 0x3F80) for every agent, everything else 0
    const uint32_t byte = i * 4;
    reinterpret_cast<uint32_t*>(ones)[i] = (byte < FG_BYTES && (byte & 15u) == 0u) ? 0x00003F80u : 0u;
  }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  if (tid == 0) {
    for (int i = 0; i < WG_NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 128); }
    mbar_fence_init();
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const int nb = b1 - b0;

  if (nb > 0) {
    if (warp == 0) {
      if (lane == 0) {
        int it = 0;
        bool ok = true;
        for (int pair = 0; pair < 6 && ok; ++pair) {
          for (int b = b0; b < b1 && ok; ++b, ++it) {
            const int slot = it % WG_NS;
            const uint32_t par = (uint32_t)((it / WG_NS) & 1);
            if (!mbar_wait(&empty[slot], par ^ 1u, WG_WAIT_CYCLES)) { *a.status = 2; ok = false; break; }
            const PairDesc d = pair_desc(S, pair, b);
            uint8_t* dst = smem + slot * WG_SLOT;
            mbar_arrive_expect_tx(&full[slot], d.a_bytes + d.b_bytes);
            const uint8_t* asrc = a.spill + d.a_off;
            const uint8_t* bsrc = a.spill + d.b_off;
            if (a.ntiles > 0) {      // layer inputs saved by the forward launch (pair 0: X; pairs 1..4: B = act(pair - 1); pair 5: A = act(4))
              const int pos = b / a.ntiles, tile = b - pos * a.ntiles;
              const uint8_t* ext = pos < 8 ? a.x1_ext[pos] : nullptr;
              if (ext != nullptr) {
                const FwdSaveLayout FS{a.ntiles};
                if (pair == 0) bsrc = ext + FS.x1(tile);
                else if (a.save_level >= 2) {
                  if (pair <= 4) bsrc = ext + FS.act(pair - 1, tile);
                  else asrc = ext + FS.act(4, tile);
                }
              }
            }
            bulk_g2s(dst, asrc, d.a_bytes, &full[slot]);
            bulk_g2s(dst + HID_BYTES, bsrc, d.b_bytes, &full[slot]);
          }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        int it = 0;
        const uint32_t ones_addr = smem_u32(ones);
        bool ok = true;
        for (int pair = 0; pair < 6 && ok; ++pair) {
          const int set = pair & 1;
          const uint32_t use = (uint32_t)(pair >> 1);            // how many times this set was used before
          if (!mbar_wait(&acc_empty[set], (use & 1u) ^ 1u, WG_WAIT_CYCLES)) { *a.status = 3; ok = false; break; }
          tc_fence_after();
          const PairDesc d = pair_desc(S, pair, b0);
          const uint32_t idesc = make_idesc_bf16(128, d.N, false, true, true);
          const uint32_t idesc_b = make_idesc_bf16(128, 16, false, true, true);
          const uint32_t acc = tmem + (uint32_t)set * WG_SET_COLS;
          for (int b = b0; b < b1 && ok; ++b, ++it) {
            const int slot = it % WG_NS;
            const uint32_t par = (uint32_t)((it / WG_NS) & 1);
            if (!mbar_wait(&full[slot], par, WG_WAIT_CYCLES)) { *a.status = 4; ok = false; break; }
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + slot * WG_SLOT), sb = sa + HID_BYTES;
#pragma unroll 1
            for (int ks = 0; ks < TM / 16; ++ks) {
              const uint64_t ad = make_smem_desc(sa + (uint32_t)ks * 256u, 128u, FG_BYTES, SWZ_NONE);
              const uint64_t bd = make_smem_desc(sb + (uint32_t)ks * 256u, 128u, FG_BYTES, SWZ_NONE);
              const uint32_t accum = (b > b0 || ks > 0) ? 1u : 0u;
              mma_ss(acc, ad, bd, idesc, accum);
              if (d.bias)
                mma_ss(acc + WG_BIAS_COL, ad, make_smem_desc(ones_addr + (uint32_t)ks * 256u, 128u, FG_BYTES, SWZ_NONE), idesc_b, accum);
            }
            mma_commit(&empty[slot]);           // slot reusable once these MMAs have read it
          }
          mma_commit(&acc_full[set]);
        }
      }
    } else {
      // epilogue warps: TMEM lanes 32 * (warp % 4)
      const int q = warp & 3;
      const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
      const int m = q * 32 + lane;             // accumulator row (feature index of the A operand)
      float* part = a.partial + (size_t)blockIdx.x * PART_TOTAL;
      for (int pair = 0; pair < 6; ++pair) {
        const int set = pair & 1;
        const uint32_t use = (uint32_t)(pair >> 1);
        if (!mbar_wait(&acc_full[set], use & 1u, WG_WAIT_CYCLES)) { *a.status = 5; break; }
        tc_fence_after();
        const PairDesc d = pair_desc(S, pair, b0);
        const uint32_t acc = tmem + (uint32_t)set * WG_SET_COLS + lane_sel;
        float* prow = part + d.part_off + (size_t)m * d.part_ld;
        for (int c0 = 0; c0 < d.N; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(acc + (uint32_t)c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float4* p4 = reinterpret_cast<float4*>(prow + c0 + 4 * j);
            float4 x = *p4;
            x.x += __uint_as_float(r[4 * j]); x.y += __uint_as_float(r[4 * j + 1]);
            x.z += __uint_as_float(r[4 * j + 2]); x.w += __uint_as_float(r[4 * j + 3]);
            *p4 = x;
          }
        }
        if (d.bias) {
          uint32_t r[16];
          tmem_ld16(acc + WG_BIAS_COL, r);
          tmem_ld_wait();
          prow[128] += __uint_as_float(r[0]);
        }
        tc_fence_before();
        mbar_arrive(&acc_empty[set]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// grad_w_flat (torch parameter order) = sum over CTAs of the partials, re-laid-out
__global__ void wgrad_finalize_kernel(const float* __restrict__ partial, int ncta, const float* __restrict__ g_bout,
                                      float* __restrict__ gw) {
  constexpr int P = 64, H = 32, HID = 128, NRES = 2;
  const FlatLayout F{P, H, HID, NRES};
  const int IN = 2 * P + H + 2;
  const int total = (int)F.total();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int src = -1;
    if (i < F.off_bin()) {
      const int m = i / IN, k = i % IN;
      src = PART_W1 + m * PART_W1_N + k;                 // columns 160/161 hold the sin/cos features
    } else if (i < F.off_res(0)) {
      src = PART_W1 + (i - (int)F.off_bin()) * PART_W1_N + (2 * P + H + 2);
    } else if (i < F.off_wout()) {
      const int q = i - (int)F.off_res(0);
      const int per = 2 * HID * HID + 2 * HID;
      const int r = q / per, o = q % per;
      int mat, m, k;
      if (o < HID * HID) { mat = 2 * r; m = o / HID; k = o % HID; }
      else if (o < HID * HID + HID) { mat = 2 * r; m = o - HID * HID; k = 128; }
      else if (o < 2 * HID * HID + HID) { mat = 2 * r + 1; const int oo = o - HID * HID - HID; m = oo / HID; k = oo % HID; }
      else { mat = 2 * r + 1; m = o - 2 * HID * HID - HID; k = 128; }
      src = PART_HH + mat * PART_HH_SZ + m * PART_HH_N + k;
    } else if (i < F.off_bout()) {
      const int q = i - (int)F.off_wout();
      const int o = q / HID, k = q % HID;                 // w_out[o][k]  <-  partial WO[k][o]
      src = PART_WO + k * PART_WO_N + o;
    }
    float s = 0.0f;
    if (src >= 0) {
      for (int c = 0; c < ncta; ++c) s += partial[(size_t)c * PART_TOTAL + src];
    } else {
      s = g_bout[i - (int)F.off_bout()];
    }
    gw[i] = s;
  }
}

// ---- host side -----------------------------------------------------------------------------------------------
size_t wgrad_spill_bytes(int nblobs) { return SpillLayout{nblobs}.total(); }
int wgrad_num_ctas() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}
size_t wgrad_partial_bytes() { return ((size_t)wgrad_num_ctas() * PART_TOTAL + 64) * sizeof(float) + 256; }

// partial buffer: [ncta][PART_TOTAL] floats, then g_bout[64], then a status word
int wgrad_tc(const void* spill, int nblobs, int used, void* partial, const void* const* x1_ext, int n_x1, int ntiles, int save_level,
             cudaStream_t st) {
  if (used <= 0) return AB200_OK;
  const int ncta = wgrad_num_ctas();
  WgradArgs k{(const uint8_t*)spill, (float*)partial, nblobs, used, reinterpret_cast<int*>((float*)partial + (size_t)ncta * PART_TOTAL + 64), {}, 0, 0};
  if (x1_ext != nullptr && n_x1 > 0 && ntiles > 0) {
    if (n_x1 > 8) return AB200_ERR_BAD_ARG;
    for (int i = 0; i < n_x1; ++i) k.x1_ext[i] = (const uint8_t*)x1_ext[i];
    k.ntiles = ntiles;
    k.save_level = save_level;
  }
  cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WG_SMEM);
  if (e != cudaSuccess) { set_cuda_error(e); return AB200_ERR_CUDA; }
  wgrad_tc_kernel<<<ncta, WG_THREADS, WG_SMEM, st>>>(k);
  return check_launch();
}

float* wgrad_bout_ptr(void* partial) { return (float*)partial + (size_t)wgrad_num_ctas() * PART_TOTAL; }

int wgrad_finalize(const void* partial, float* grad_w_flat, cudaStream_t st) {
  const int ncta = wgrad_num_ctas();
  wgrad_finalize_kernel<<<148, 256, 0, st>>>((const float*)partial, ncta, (const float*)partial + (size_t)ncta * PART_TOTAL, grad_w_flat);
  return check_launch();
}

}  // namespace ab200
