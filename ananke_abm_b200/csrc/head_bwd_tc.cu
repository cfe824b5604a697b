// Backward of the fused cross-entropy head on tensor cores: the [M, Z] probability matrix is
// recomputed tile by tile and consumed on chip, like the forward.
//
//   logit[m,z] = cos(e_m, t_z) / tau,   dlogit[m,z] = g_m (softmax[m,z] - [z = y_m])         (g_m = upstream row gradient)
//   d e^_m = (1/tau) sum_z dlogit[m,z] t^_z            d t^_z = (1/tau) sum_m dlogit[m,z] e^_m        (^ = normalised)
//
// ONE kernel serves both sums.  "X" is the operand whose gradient is produced (128 x-rows per CTA tile, resident), "Y"
// is streamed in 128-row chunks; both come as the split-bf16 UMMA images of head_tc.cu (K = 3 x 64):
//   S[x,y]  = X' Y'^T                      tcgen05 SS, fp32 accumulator in TMEM (two buffers)
//   dS[x,y] = g (2^(S*c - lse2) - hit)     epilogue warps, written back to TMEM as the hi/lo bf16 A operand
//   dX[x,:] += dS[x,:] Y^[:, :]            tcgen05 TS, B = the MN-major view of the same Y image (N = 64, K = 128 y)
// rows-outer pass (X = pred_emb rows, Y = zones) gives d e^; zones-outer pass (X = zones, Y = rows) gives d t^.  Every
// X tile is owned by one CTA: plain stores, no atomics, deterministic.
//   warp 0 : TMA producer (X tile, Y ring of 3)     warp 1 : MMA issuer
//   warps 2-9 : epilogue, thread = (x row, half of the chunk's 128 columns)
// Work item = (X tile, range of Y chunks): when there are few X tiles (zones-outer: Z / 128 = 79 at configs[2]) the Y
// stream is split so that every SM has work; the partial dX of the splits are added in a fixed order afterwards.
#include "common.cuh"
#include "umma.cuh"

namespace ab200 {
using namespace umma;

constexpr int HB_E = 64, HB_K = 3 * HB_E, HB_T = 128;
constexpr uint32_t HB_IMG = HB_T * HB_K * 2;                // 49,152 B per 128-row image
constexpr uint32_t HB_SEG = HB_T * HB_E * 2;                // 16,384 B per K segment
constexpr int HB_NS = 3;
constexpr uint32_t HB_SMEM = (1 + HB_NS) * HB_IMG;
constexpr int HB_THREADS = 320, HB_EPI = 256;
constexpr uint32_t HB_LBO = 128u * 16u, HB_SBO = 128u;
constexpr uint32_t HB_C_S = 0, HB_C_DSH = 256, HB_C_DSL = 320, HB_C_DX = 384;
constexpr long long HB_WAIT = 400000000LL;

// normalise rows and write split-bf16 images; seg order {hi, lo, hi} (table, as head_pack_table_kernel) or {hi, hi, lo} (rows)
__global__ void __launch_bounds__(128) head_pack_rows_kernel(const float* __restrict__ src, int64_t n_rows, int rows_order,
                                                             uint8_t* __restrict__ img) {
  const int64_t chunk = blockIdx.x;
  const int n = threadIdx.x;
  const int64_t r = chunk * HB_T + n;
  float v[HB_E];
  float ss = 0.0f;
#pragma unroll
  for (int k = 0; k < HB_E; ++k) {
    v[k] = r < n_rows ? src[r * HB_E + k] : 0.0f;
    ss += v[k] * v[k];
  }
  const float inv = 1.0f / (sqrtf(ss) + 1e-8f);
  uint8_t* blob = img + (size_t)chunk * HB_IMG;
  const int s_lo = rows_order ? 2 : 1, s_hi2 = rows_order ? 1 : 2;
#pragma unroll
  for (int k = 0; k < HB_E; ++k) {
    const float t = v[k] * inv;
    const __nv_bfloat16 hi = __float2bfloat16_rn(t);
    const __nv_bfloat16 lo = __float2bfloat16_rn(t - __bfloat162float(hi));
    *reinterpret_cast<__nv_bfloat16*>(blob + off_kmajor_noswz(n, k, HB_LBO, HB_SBO)) = hi;
    *reinterpret_cast<__nv_bfloat16*>(blob + off_kmajor_noswz(n, s_lo * HB_E + k, HB_LBO, HB_SBO)) = lo;
    *reinterpret_cast<__nv_bfloat16*>(blob + off_kmajor_noswz(n, s_hi2 * HB_E + k, HB_LBO, HB_SBO)) = hi;
  }
}

struct HeadBwdArgs {
  const uint8_t* ximg;      // [nx][48 KiB]
  const uint8_t* yimg;      // [ny][48 KiB]
  int nx, ny;               // tiles / chunks
  int64_t NX, NY;           // true row counts
  int rows_outer;           // 1: x = (agent,time) row, y = zone;  0: x = zone, y = row
  int y_lo_seg;             // K segment of the Y image that holds the lo part (1 for the table image, 2 for the row image)
  const float* lse;         // per ROW (x if rows_outer else y)
  const float* g;           // per ROW upstream gradient
  const int64_t* target;    // per ROW
  // expected-distance term (all three set or all null): dlogit += g2 * softmax * (dist[target, z] - edist)
  const float* g2;          // per ROW upstream gradient of the expected distance
  const float* edist;       // per ROW expected distance (forward output)
  const float* dist;        // [Z][Z]
  int Z;
  float inv_tau;
  float* dx;                // [NX][64]                       (nsplit == 1)
  float* partial;           // [nsplit][nx * 128][64]         (nsplit > 1)
  int nsplit;               // Y chunks are divided into nsplit contiguous ranges; work item = tile * nsplit + split
  int* status;
};

__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void hb_item(const HeadBwdArgs& a, int item, int& tile, int& c_begin, int& c_end, int& split) {
  tile = item / a.nsplit;
  split = item - tile * a.nsplit;
  c_begin = (int)(((int64_t)split * a.ny) / a.nsplit);
  c_end = (int)(((int64_t)(split + 1) * a.ny) / a.nsplit);
}

// ROWS: rows-outer pass (x = row, y = zone) or zones-outer; DIST: with the expected-distance term.  Both are template
// parameters: the per-element epilogue is the kernel's limiter (16 instructions per element), runtime switches cost ~70 %.
template <bool ROWS, bool DIST>
__global__ void __launch_bounds__(HB_THREADS, 1) head_ce_bwd_kernel(const __grid_constant__ HeadBwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[HB_NS], empty[HB_NS], acc_full[2], acc_empty[2], x_full, x_free, ds_full, ds_free, dx_full, dx_empty;
  __shared__ uint32_t tmem_base_s;
  __shared__ float y_lse2[2][HB_T], y_g[2][HB_T];
  __shared__ int y_tgt[2][HB_T];
  __shared__ float y_g2[2][HB_T], y_e[2][HB_T];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* sX = smem;
  uint8_t* sY = smem + HB_IMG;

  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  if (tid == 0) {
    for (int i = 0; i < HB_NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], HB_EPI); }
    mbar_init(&x_full, 1); mbar_init(&x_free, 1);
    mbar_init(&ds_full, HB_EPI); mbar_init(&ds_free, 1);
    mbar_init(&dx_full, 1); mbar_init(&dx_empty, HB_EPI);
    mbar_fence_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const int n_items = a.nx * a.nsplit;
  const int my_tiles = (n_items > (int)blockIdx.x) ? (n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;     // work items of this CTA

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      bool ok = true;
      for (int t = 0; t < my_tiles && ok; ++t) {
        int tile, c_begin, c_end, split;
        hb_item(a, blockIdx.x + t * gridDim.x, tile, c_begin, c_end, split);
        if (t > 0 && !mbar_wait(&x_free, (uint32_t)((t - 1) & 1), HB_WAIT)) { *a.status = 1; break; }
        mbar_arrive_expect_tx(&x_full, HB_IMG);
        bulk_g2s(sX, a.ximg + (size_t)tile * HB_IMG, HB_IMG, &x_full);
        for (int c = c_begin; c < c_end; ++c, ++it) {
          const int slot = it % HB_NS;
          if (!mbar_wait(&empty[slot], (uint32_t)(((it / HB_NS) & 1) ^ 1), HB_WAIT)) { *a.status = 2; ok = false; break; }
          mbar_arrive_expect_tx(&full[slot], HB_IMG);
          bulk_g2s(sY + slot * HB_IMG, a.yimg + (size_t)c * HB_IMG, HB_IMG, &full[slot]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(HB_T, HB_T);
      const uint32_t idesc_d = make_idesc_bf16(HB_T, HB_E, false, false, /*b_mn=*/true);
      const uint64_t xdesc0 = make_smem_desc(smem_u32(sX), HB_LBO, HB_SBO, SWZ_NONE);
      constexpr uint32_t step_s = (2u * HB_LBO) >> 4;       // K-major: one K-step = two 8-feature groups
      constexpr uint32_t step_d = (2u * HB_SBO) >> 4;       // MN-major view: one K-step = two 8-row groups
      int nseq = 0;          // global chunk counter of this CTA (ring slot, accumulator buffer, dS hand-over)
      bool ok = true;
      auto issue_s = [&](int n) -> bool {
        const int slot = n % HB_NS, buf = n & 1;
        if (!mbar_wait(&acc_empty[buf], (uint32_t)(((n >> 1) & 1) ^ 1), HB_WAIT)) { *a.status = 4; return false; }
        if (!mbar_wait(&full[slot], (uint32_t)((n / HB_NS) & 1), HB_WAIT)) { *a.status = 5; return false; }
        tc_fence_after();
        const uint64_t ydesc0 = make_smem_desc(smem_u32(sY + slot * HB_IMG), HB_LBO, HB_SBO, SWZ_NONE);
#pragma unroll
        for (int ks = 0; ks < HB_K / 16; ++ks)
          mma_ss(tmem + HB_C_S + (uint32_t)buf * HB_T, xdesc0 + (uint64_t)(ks * step_s), ydesc0 + (uint64_t)(ks * step_s), idesc_s, ks > 0 ? 1u : 0u);
        mma_commit(&acc_full[buf]);
        return true;
      };
      for (int t = 0; t < my_tiles && ok; ++t) {
        int tile, c_begin, c_end, split;
        hb_item(a, blockIdx.x + t * gridDim.x, tile, c_begin, c_end, split);
        const int ny = c_end - c_begin;
        if (!mbar_wait(&x_full, (uint32_t)(t & 1), HB_WAIT)) { *a.status = 3; break; }
        tc_fence_after();
        if (!issue_s(nseq)) break;
        if (ny == 1) mma_commit(&x_free);
        for (int c = 0; c < ny; ++c, ++nseq) {
          if (c + 1 < ny) {
            if (!issue_s(nseq + 1)) { ok = false; break; }
            if (c + 1 == ny - 1) mma_commit(&x_free);       // every S MMA of this X tile has been issued
          }
          if (!mbar_wait(&ds_full, (uint32_t)(nseq & 1), HB_WAIT)) { *a.status = 6; ok = false; break; }
          if (c == 0 && t > 0 && !mbar_wait(&dx_empty, (uint32_t)((t - 1) & 1), HB_WAIT)) { *a.status = 7; ok = false; break; }
          tc_fence_after();
          const int slot = nseq % HB_NS;
          const uint32_t ybase = smem_u32(sY + slot * HB_IMG);
          const uint64_t yh = make_smem_desc(ybase, HB_SBO, HB_LBO, SWZ_NONE);                                 // hi features, MN-major view
          const uint64_t yl = make_smem_desc(ybase + (uint32_t)a.y_lo_seg * HB_SEG, HB_SBO, HB_LBO, SWZ_NONE);  // lo features
#pragma unroll
          for (int ks = 0; ks < HB_T / 16; ++ks) {
            const uint32_t ah = tmem + HB_C_DSH + (uint32_t)ks * 8u, al = tmem + HB_C_DSL + (uint32_t)ks * 8u;
            mma_ts(tmem + HB_C_DX, ah, yh + (uint64_t)(ks * step_d), idesc_d, (c > 0 || ks > 0) ? 1u : 0u);
            mma_ts(tmem + HB_C_DX, ah, yl + (uint64_t)(ks * step_d), idesc_d, 1u);
            mma_ts(tmem + HB_C_DX, al, yh + (uint64_t)(ks * step_d), idesc_d, 1u);
          }
          mma_commit(&empty[slot]);
          mma_commit(&ds_free);
          if (c == ny - 1) mma_commit(&dx_full);
        }
      }
    }
  } else {
    const int q = warp & 3;                           // TMEM lane quarter this warp may touch
    const int hh = (warp - 2) >> 2;                   // which 64 of the chunk's 128 columns
    const int row = q * 32 + lane;                    // x row inside the tile = TMEM lane
    const int et = tid - 64;                          // 0..255; the first 128 stage the per-y entries
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const float sc2 = a.inv_tau * 1.4426950408889634f;
    int nseq = 0;
    bool dead = false;
    for (int t = 0; t < my_tiles && !dead; ++t) {
      int tile, c_begin, c_end, split;
      hb_item(a, blockIdx.x + t * gridDim.x, tile, c_begin, c_end, split);
      const int64_t xg = (int64_t)tile * HB_T + row;
      const bool xvalid = xg < a.NX;
      float x_lse2 = 0.0f, x_g = 0.0f;
      int x_tgt = -1;
      constexpr bool want_d = DIST;
      float x_g2 = 0.0f, x_e = 0.0f;
      const float* drow = nullptr;
      if (ROWS && xvalid) {
        x_lse2 = a.lse[xg] * 1.4426950408889634f; x_g = a.g[xg] * a.inv_tau; x_tgt = (int)a.target[xg];
        if (want_d) {
          x_g2 = a.g2[xg] * a.inv_tau; x_e = a.edist[xg];
          const int tc = (x_tgt < 0 || x_tgt >= a.Z) ? 0 : x_tgt;
          drow = a.dist + (size_t)tc * a.Z;
        }
      }
      const int xi = (int)xg;                         // zones-outer: the zone index of this row (Z < 2^31)
      for (int c = c_begin; c < c_end; ++c, ++nseq) {
        const int buf = nseq & 1;
        const int64_t ybase = (int64_t)c * HB_T;
        if (!ROWS) {          // per-row quantities of this Y chunk -> shared memory (read by every x thread)
          if (et < HB_T) {
            const int64_t yr = ybase + et;
            const bool yv = yr < a.NY;
            y_lse2[buf][et] = yv ? a.lse[yr] * 1.4426950408889634f : 0.0f;
            y_g[buf][et] = yv ? a.g[yr] * a.inv_tau : 0.0f;            // 0 on padding rows: their dS vanishes
            y_tgt[buf][et] = yv ? (int)a.target[yr] : -1;
            if (want_d) { y_g2[buf][et] = yv ? a.g2[yr] * a.inv_tau : 0.0f; y_e[buf][et] = yv ? a.edist[yr] : 0.0f; }
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        if (!mbar_wait(&acc_full[buf], (uint32_t)((nseq >> 1) & 1), HB_WAIT)) { *a.status = 8; dead = true; break; }
        tc_fence_after();
        const int n_valid = (int)min((int64_t)HB_T, a.NY - ybase);     // columns of this chunk that are real y rows
#pragma unroll 1
        for (int c0 = hh * 64; c0 < hh * 64 + 64; c0 += 32) {
          uint32_t r[32], hi[16], lo[16];
          tmem_ld32(tmem + lane_sel + HB_C_S + (uint32_t)(buf * HB_T + c0), r);
          tmem_ld_wait();
          const int t_loc = ROWS ? x_tgt - (int)ybase - c0 : 0;          // column of the target inside this group
          const int lim = n_valid - c0;                                          // columns < lim are valid
          const bool all_valid = lim >= 32;                                      // uniform: only the last chunk is ragged
          float dd[32];
          if (ROWS && want_d) {
            const bool vec = all_valid && (a.Z & 3) == 0 && drow != nullptr;
            if (vec) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 t4 = __ldg(reinterpret_cast<const float4*>(drow + ybase + c0) + q);
                dd[4 * q] = t4.x; dd[4 * q + 1] = t4.y; dd[4 * q + 2] = t4.z; dd[4 * q + 3] = t4.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) dd[j] = (drow != nullptr && j < lim) ? __ldg(drow + ybase + c0 + j) : 0.0f;
            }
          }
          int prev_t = -2;            // zones-outer: consecutive rows of a target-sorted chunk share dist[target, zone]
          float prev_d = 0.0f;
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) {
            float d[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int j = 2 * jj + u;
              const float L = __uint_as_float(r[j]);
              float v;
              if (ROWS) {
                const float p = ex2_fast(fmaf(L, sc2, -x_lse2));
                v = x_g * (p - (j == t_loc ? 1.0f : 0.0f));
                if (want_d) v = fmaf(x_g2 * p, dd[j] - x_e, v);
                if (!all_valid) v = j < lim ? v : 0.0f;
              } else {
                const float p = ex2_fast(fmaf(L, sc2, -y_lse2[buf][c0 + j]));
                const int yt = y_tgt[buf][c0 + j];
                v = y_g[buf][c0 + j] * (p - (yt == xi ? 1.0f : 0.0f));
                if (want_d && xvalid) {
                  if (yt != prev_t) {            // uniform across the warp (yt comes from shared memory)
                    const int tc = (yt < 0 || yt >= a.Z) ? 0 : yt;
                    prev_d = __ldg(a.dist + (size_t)tc * a.Z + xi);
                    prev_t = yt;
                  }
                  v = fmaf(y_g2[buf][c0 + j] * p, prev_d - y_e[buf][c0 + j], v);
                }
              }
              d[u] = v;
            }
            hi[jj] = pack_bf16(d[0], d[1]);
            const float h0 = __uint_as_float(hi[jj] << 16), h1 = __uint_as_float(hi[jj] & 0xffff0000u);
            lo[jj] = pack_bf16(d[0] - h0, d[1] - h1);
          }
          if (c0 == hh * 64 && nseq > 0) {
            // the dS operand is single-buffered: the previous chunk's dX MMAs must have read it.  Waiting HERE, with the
            // first 32 columns already computed in registers, overlaps half of the epilogue with those MMAs.
            if (!mbar_wait(&ds_free, (uint32_t)((nseq - 1) & 1), HB_WAIT)) { *a.status = 9; dead = true; }
            tc_fence_after();
          }
          tmem_st16(tmem + lane_sel + HB_C_DSH + (uint32_t)(c0 / 2), hi);
          tmem_st16(tmem + lane_sel + HB_C_DSL + (uint32_t)(c0 / 2), lo);
        }
        if (dead) break;
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&ds_full);
        mbar_arrive(&acc_empty[buf]);
      }
      if (dead) break;
      if (!mbar_wait(&dx_full, (uint32_t)(t & 1), HB_WAIT)) { *a.status = 10; break; }
      tc_fence_after();
      {
        uint32_t r[32];
        tmem_ld32(tmem + lane_sel + HB_C_DX + (uint32_t)(hh * 32), r);
        tmem_ld_wait();
        float* base = a.nsplit > 1 ? a.partial + ((size_t)split * a.nx * HB_T + (size_t)xg) * HB_E : a.dx + (size_t)xg * HB_E;
        if (xvalid || a.nsplit > 1) {      // the partial buffer is padded to whole tiles
          float4* o = reinterpret_cast<float4*>(base + hh * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            o[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
        }
      }
      tc_fence_before();
      mbar_arrive(&dx_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// dx[x][f] = sum over splits (fixed order) of partial[split][x][f]
__global__ void head_reduce_splits_kernel(const float* __restrict__ partial, int nsplit, int64_t rows_padded, int64_t n_rows,
                                          float* __restrict__ dx) {
  const int64_t n = n_rows * HB_E;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.0f;
    for (int k = 0; k < nsplit; ++k) s += partial[(size_t)k * rows_padded * HB_E + i];
    dx[i] = s;
  }
}

// ---- host side --------------------------------------------------------------------------------------------------
static int hb_tiles(int64_t n) { return (int)((n + HB_T - 1) / HB_T); }
static int hb_nsplit(int nx, int ny, int sms) {       // enough work items for every SM when there are few X tiles
  if (nx >= 4 * sms) return 1;
  int want = (8 * sms + nx - 1) / nx;
  if (want > ny) want = ny;
  return want < 1 ? 1 : want;
}
constexpr int HB_SMS_ASSUMED = 160;                   // workspace sizing only (an upper bound on the split count)
static size_t hb_partial_bytes(int64_t NX, int64_t NY) {
  const int nx = hb_tiles(NX), ny = hb_tiles(NY);
  const int ns = hb_nsplit(nx, ny, HB_SMS_ASSUMED);
  return ns > 1 ? (size_t)ns * nx * HB_T * HB_E * sizeof(float) : 0;
}

size_t head_ce_backward_workspace_bytes(int64_t M, int Z) {
  const size_t part = hb_partial_bytes(M, Z) > hb_partial_bytes(Z, M) ? hb_partial_bytes(M, Z) : hb_partial_bytes(Z, M);
  return (size_t)hb_tiles(M) * HB_IMG + (size_t)hb_tiles(Z) * HB_IMG + 256 + part;
}

// d emb^ [M][64] and d table^ [Z][64] (gradients w.r.t. the NORMALISED vectors; the caller applies x / (|x| + 1e-8)).
int head_ce_backward(const float* emb, const float* table, const int64_t* target, const float* lse, const float* g_rows,
                     const float* g_dist_rows, const float* exp_dist, const float* dist, int64_t M, int Z, int E, float tau,
                     float* d_emb_n, float* d_table_n, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (E != HB_E) return AB200_ERR_UNSUPPORTED;
  if (ws_bytes < head_ce_backward_workspace_bytes(M, Z)) return AB200_ERR_WORKSPACE;
  const int nm = hb_tiles(M), nz = hb_tiles(Z);
  uint8_t* eimg = (uint8_t*)ws;
  uint8_t* timg = eimg + (size_t)nm * HB_IMG;
  int* status = (int*)(timg + (size_t)nz * HB_IMG);
  cudaError_t e = cudaMemsetAsync(status, 0, sizeof(int), st);
  if (e != cudaSuccess) { set_cuda_error(e); return AB200_ERR_CUDA; }
  head_pack_rows_kernel<<<nm, 128, 0, st>>>(emb, M, 1, eimg);
  head_pack_rows_kernel<<<nz, 128, 0, st>>>(table, Z, 0, timg);
  int rc = check_launch();
  if (rc) return rc;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  for (const void* fn : {(const void*)head_ce_bwd_kernel<true, true>, (const void*)head_ce_bwd_kernel<true, false>,
                         (const void*)head_ce_bwd_kernel<false, true>, (const void*)head_ce_bwd_kernel<false, false>}) {
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HB_SMEM);
    if (e != cudaSuccess) { set_cuda_error(e); return AB200_ERR_CUDA; }
  }
  float* partial = (float*)((uint8_t*)status + 256);
  auto pass = [&](const uint8_t* ximg, const uint8_t* yimg, int nx, int ny, int64_t NX, int64_t NY, int rows_outer, int y_lo_seg,
                  float* dx) -> int {
    int ns = hb_nsplit(nx, ny, sms);
    const int ns_cap = hb_nsplit(nx, ny, HB_SMS_ASSUMED);
    if (ns > ns_cap) ns = ns_cap;                      // never beyond what the workspace was sized for
    HeadBwdArgs k{ximg, yimg, nx, ny, NX, NY, rows_outer, y_lo_seg, lse, g_rows, target, g_dist_rows, exp_dist, dist, Z, 1.0f / tau, dx, partial, ns, status};
    const int items = nx * ns;
    const int grid = items < sms ? items : sms;
    const bool with_dist = dist != nullptr;
    if (rows_outer) {
      if (with_dist) head_ce_bwd_kernel<true, true><<<grid, HB_THREADS, HB_SMEM, st>>>(k);
      else head_ce_bwd_kernel<true, false><<<grid, HB_THREADS, HB_SMEM, st>>>(k);
    } else {
      if (with_dist) head_ce_bwd_kernel<false, true><<<grid, HB_THREADS, HB_SMEM, st>>>(k);
      else head_ce_bwd_kernel<false, false><<<grid, HB_THREADS, HB_SMEM, st>>>(k);
    }
    int r = check_launch();
    if (r || ns == 1) return r;
    head_reduce_splits_kernel<<<sms * 4, 256, 0, st>>>(partial, ns, (int64_t)nx * HB_T, NX, dx);
    return check_launch();
  };
  if ((rc = pass(eimg, timg, nm, nz, M, (int64_t)Z, 1, 1, d_emb_n))) return rc;      // rows outer: x = rows (d emb^), y = zones
  return pass(timg, eimg, nz, nm, (int64_t)Z, M, 0, 2, d_table_n);                   // zones outer: x = zones (d table^), y = rows
}

int head_ce_backward_status(const void* ws, int64_t M, int Z, int* host_out, cudaStream_t st) {
  const uint8_t* p = (const uint8_t*)ws + (size_t)hb_tiles(M) * HB_IMG + (size_t)hb_tiles(Z) * HB_IMG;
  cudaError_t e = cudaMemcpyAsync(host_out, p, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { set_cuda_error(e); return AB200_ERR_CUDA; }
  return 0;
}

}  // namespace ab200
