// Fused classification head + label prediction: never materialises the [B, T, Z] logits tensor.
//
//   reference   emb_norm = pred_emb / (|pred_emb| + 1e-8) ; table_norm = class_table / (|class_table| + 1e-8)
//               logits = einsum("bte,ze->btz", emb_norm, table_norm) / tau            mode_sep/architecture/model.py:196-199
//               labels = logits.argmax(-1)                                             mode_sep/inference/inference.py:57,63
//   at configs[2] scale logits would be 1M x 97 x 10k x 4 B = 3.9 TB.
//
// One CTA per 128 rows (a row = one (agent, time) pair), persistent over row tiles; the zone table is streamed in
// 128-zone chunks:
//   warp 0      TMA producer : cp.async.bulk of the prepacked table chunk (48 KiB UMMA image) into a 3-slot ring
//   warp 1      MMA issuer   : D[128 x 128] = A'[128 x 192] B'[128 x 192]^T on tcgen05, two TMEM accumulators ping-pong
//   warps 2-5   epilogue     : tcgen05.ld, running top-2 (value, zone) per row; build the A' tile of the next row tile
// Exactness: operands are 2-term bf16 splits (A' = [e_hi|e_hi|e_lo], B' = [t_hi|t_lo|t_hi], K = 3 x 64: products exact
// to ~2^-16), the tensor cores only NOMINATE the two best zones per row; both are re-scored with fp32 FMAs on the fp32
// normalised vectors and the winner (lower index on ties, like torch.argmax) is the label -- the reference's fp32 argmax
// unless more than two zones lie within 1e-5 of the maximum.
#include "common.cuh"
#include "umma.cuh"

namespace ab200 {
using namespace umma;

constexpr int HD_E = 64, HD_K = 3 * HD_E, HD_TM = 128, HD_TN = 128;
constexpr uint32_t HD_A_BYTES = HD_TM * HD_K * 2;          // 49,152
constexpr uint32_t HD_B_BYTES = HD_TN * HD_K * 2;          // 49,152 per table chunk
constexpr int HD_NS = 3;
constexpr uint32_t HD_SMEM = HD_A_BYTES + HD_NS * HD_B_BYTES;
constexpr int HD_THREADS = 320, HD_EPI = 256;       // warps 0/1: producer / MMA issuer; warps 2-9: epilogue (row x column half)
constexpr uint32_t HD_LBO = 128u * 16u, HD_SBO = 128u;     // [128 rows][K] K-major, un-swizzled
constexpr long long HD_WAIT = 400000000LL;

// ---- table prep: normalise, keep fp32 copy, write split-bf16 UMMA images chunk by chunk ----------------------------
__global__ void __launch_bounds__(128) head_pack_table_kernel(const float* __restrict__ table, int Z, float* __restrict__ tn,
                                                              uint8_t* __restrict__ img) {
  const int chunk = blockIdx.x, n = threadIdx.x, z = chunk * HD_TN + n;
  float v[HD_E];
  float ss = 0.0f;
#pragma unroll
  for (int k = 0; k < HD_E; ++k) {
    v[k] = z < Z ? table[(size_t)z * HD_E + k] : 0.0f;
    ss += v[k] * v[k];
  }
  const float inv = 1.0f / (sqrtf(ss) + 1e-8f);
  uint8_t* blob = img + (size_t)chunk * HD_B_BYTES;
#pragma unroll
  for (int k = 0; k < HD_E; ++k) {
    const float t = v[k] * inv;
    tn[(size_t)z * HD_E + k] = t;
    const __nv_bfloat16 hi = __float2bfloat16_rn(t);
    const __nv_bfloat16 lo = __float2bfloat16_rn(t - __bfloat162float(hi));
    *reinterpret_cast<__nv_bfloat16*>(blob + off_kmajor_noswz(n, k, HD_LBO, HD_SBO)) = hi;
    *reinterpret_cast<__nv_bfloat16*>(blob + off_kmajor_noswz(n, HD_E + k, HD_LBO, HD_SBO)) = lo;
    *reinterpret_cast<__nv_bfloat16*>(blob + off_kmajor_noswz(n, 2 * HD_E + k, HD_LBO, HD_SBO)) = hi;
  }
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct HeadArgs {
  const float* emb;       // [M][64] un-normalised pred_emb
  const float* tn;        // [Zp][64] fp32 normalised table
  const uint8_t* img;     // [nchunk][48 KiB]
  int64_t M;
  int Z, nchunk;
  int ntiles;
  float inv_tau;
  int64_t* labels;        // [M] or null
  float* best;            // [M] or null: the winning logit
  int* status;
  // cross-entropy mode (all three set, or all null): lse[m] = log sum_z exp(logit[m,z]); tgt_logit[m] = logit[m, target[m]]
  const int64_t* target;  // [M], values outside [0, Z) are read as zone 0 (the caller masks those rows)
  float* lse;
  float* tgt_logit;
  // expected distance (optional, with the cross-entropy mode): exp_dist[m] = sum_z softmax[m,z] dist[target[m], z]
  const float* dist;      // [Z][Z] row-major
  float* exp_dist;        // [M]
};

__global__ void __launch_bounds__(HD_THREADS, 1) head_argmax_kernel(const __grid_constant__ HeadArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[HD_NS], empty[HD_NS], acc_full[2], acc_empty[2], a_ready, a_free;
  __shared__ uint32_t tmem_base_s;
  __shared__ float x_b1[HD_TM], x_b2[HD_TM], x_m[HD_TM], x_s[HD_TM], x_d[HD_TM];     // second column half -> first: top-2 and log-sum-exp state
  __shared__ int x_i1[HD_TM], x_i2[HD_TM];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* sA = smem;
  uint8_t* sB = smem + HD_A_BYTES;

  if (warp == 0) tmem_alloc<256>(&tmem_base_s);
  if (tid == 0) {
    for (int i = 0; i < HD_NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], HD_EPI); }
    mbar_init(&a_ready, 128);
    mbar_init(&a_free, 1);
    mbar_fence_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const int my_tiles = (a.ntiles > (int)blockIdx.x) ? (a.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      bool ok = true;
      for (int t = 0; t < my_tiles && ok; ++t)
        for (int c = 0; c < a.nchunk; ++c, ++it) {
          const int slot = it % HD_NS;
          if (!mbar_wait(&empty[slot], (uint32_t)(((it / HD_NS) & 1) ^ 1), HD_WAIT)) { *a.status = 2; ok = false; break; }
          mbar_arrive_expect_tx(&full[slot], HD_B_BYTES);
          bulk_g2s(sB + slot * HD_B_BYTES, a.img + (size_t)c * HD_B_BYTES, HD_B_BYTES, &full[slot]);
        }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(HD_TM, HD_TN);
      const uint64_t adesc0 = make_smem_desc(smem_u32(sA), HD_LBO, HD_SBO, SWZ_NONE);
      constexpr uint32_t step16 = (2u * HD_LBO) >> 4;
      int it = 0, nacc = 0;
      bool ok = true;
      for (int t = 0; t < my_tiles && ok; ++t) {
        if (!mbar_wait(&a_ready, (uint32_t)(t & 1), HD_WAIT)) { *a.status = 3; break; }
        tc_fence_after();
        for (int c = 0; c < a.nchunk; ++c, ++it, ++nacc) {
          const int slot = it % HD_NS, buf = nacc & 1;
          if (!mbar_wait(&acc_empty[buf], (uint32_t)(((nacc >> 1) & 1) ^ 1), HD_WAIT)) { *a.status = 4; ok = false; break; }
          if (!mbar_wait(&full[slot], (uint32_t)((it / HD_NS) & 1), HD_WAIT)) { *a.status = 5; ok = false; break; }
          tc_fence_after();
          const uint64_t bdesc0 = make_smem_desc(smem_u32(sB + slot * HD_B_BYTES), HD_LBO, HD_SBO, SWZ_NONE);
#pragma unroll
          for (int ks = 0; ks < HD_K / 16; ++ks)
            mma_ss(tmem + (uint32_t)buf * HD_TN, adesc0 + (uint64_t)(ks * step16), bdesc0 + (uint64_t)(ks * step16), idesc, ks > 0 ? 1u : 0u);
          mma_commit(&empty[slot]);
          mma_commit(&acc_full[buf]);
        }
        mma_commit(&a_free);      // the A' tile may be overwritten once every MMA of this row tile has read it
      }
    }
  } else {
    // ---- epilogue warps: thread = (row, half of the chunk's 128 zone columns); two warps per scheduler hide each
    //      other's dependent-issue latency (one warp per scheduler left the SM idle two cycles out of three)
    const int q = warp & 3;
    const int hh = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    int nacc = 0;
    for (int t = 0; t < my_tiles; ++t) {
      const int tile = blockIdx.x + t * gridDim.x;
      const int64_t m = (int64_t)tile * HD_TM + row;
      const bool valid = m < a.M;
      // -- A' tile: normalised row, split into hi / lo bf16
      float inv = 0.0f;
      {
        const float4* er = reinterpret_cast<const float4*>(a.emb + (valid ? m : 0) * HD_E);
        float ss = 0.0f;
#pragma unroll
        for (int j = 0; j < HD_E / 4; ++j) {
          const float4 x = valid ? er[j] : make_float4(0.f, 0.f, 0.f, 0.f);
          ss += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
        }
        inv = 1.0f / (sqrtf(ss) + 1e-8f);
        if (hh == 0) {
        if (t > 0 && !mbar_wait(&a_free, (uint32_t)((t - 1) & 1), HD_WAIT)) { *a.status = 6; break; }
#pragma unroll
        for (int j = 0; j < HD_E / 8; ++j) {     // 8 features -> one 16-byte core-matrix row, for each of the three K segments
          const float4 x0 = valid ? er[2 * j] : make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 x1 = valid ? er[2 * j + 1] : make_float4(0.f, 0.f, 0.f, 0.f);
          const float e[8] = {x0.x * inv, x0.y * inv, x0.z * inv, x0.w * inv, x1.x * inv, x1.y * inv, x1.z * inv, x1.w * inv};
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            hi[p] = pack_bf16(e[2 * p], e[2 * p + 1]);
            const float h0 = __uint_as_float(hi[p] << 16), h1 = __uint_as_float(hi[p] & 0xffff0000u);
            lo[p] = pack_bf16(e[2 * p] - h0, e[2 * p + 1] - h1);
          }
          const uint32_t o = off_kmajor_noswz(row, 8 * j, HD_LBO, HD_SBO);
          *reinterpret_cast<uint4*>(sA + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(sA + o + (HD_E / 8) * HD_LBO) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(sA + o + 2 * (HD_E / 8) * HD_LBO) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        fence_async_smem();
        mbar_arrive(&a_ready);
        }
      }
      // -- stream the zone chunks: running top-2 of this row
      float b1 = -INFINITY, b2 = -INFINITY;
      int i1 = 0, i2 = 0;
      bool dead = false;
      const bool ce = a.lse != nullptr;
      const bool need_top = a.labels != nullptr || a.best != nullptr;      // pure loss evaluation skips the top-2 tracking
      const float sc2 = a.inv_tau * 1.4426950408889634f;     // logits in log2 units
      float run_m = -INFINITY, run_s = 0.0f;                 // running max (cosine units) and sum of exp(logit - max)
      float run_d = 0.0f;                                    // sum of exp(logit - max) * dist[target, z]
      const bool want_d = ce && a.dist != nullptr;
      const float* drow = nullptr;                           // this row's line of the distance matrix
      if (want_d && valid) {
        int64_t tg = a.target[m];
        if (tg < 0 || tg >= a.Z) tg = 0;
        drow = a.dist + (size_t)tg * a.Z;
      }
      for (int c = 0; c < a.nchunk; ++c, ++nacc) {
        const int buf = nacc & 1;
        if (!mbar_wait(&acc_full[buf], (uint32_t)((nacc >> 1) & 1), HD_WAIT)) { *a.status = 7; dead = true; break; }
        tc_fence_after();
        const int zbase = c * HD_TN;
#pragma unroll 1
        for (int c0 = hh * (HD_TN / 2); c0 < (hh + 1) * (HD_TN / 2); c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tmem + lane_sel + (uint32_t)(buf * HD_TN + c0), r);
          tmem_ld_wait();
          float mx = -INFINITY;
          if (zbase + c0 + 32 <= a.Z) {        // whole group valid (all but the table's last chunk)
#pragma unroll
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, (zbase + c0 + j < a.Z) ? __uint_as_float(r[j]) : -INFINITY);
          }
          if (ce && mx > -INFINITY) {       // streaming log-sum-exp over the zones (split-bf16 logits: ~2^-16 relative)
            if (mx > run_m) { const float f = exp2f((run_m - mx) * sc2); run_s *= f; run_d *= f; run_m = mx; }
            const float off = -run_m * sc2;
            if (want_d) {       // same sweep, weighted by the target's distance row (rows sorted by target share it: L1 hits)
              float pd[4] = {0.0f, 0.0f, 0.0f, 0.0f}, ps[4] = {0.0f, 0.0f, 0.0f, 0.0f};
              if (zbase + c0 + 32 <= a.Z && (a.Z & 3) == 0) {      // whole group valid, distance row 16-byte aligned
                float4 dv[8];
#pragma unroll
                for (int q = 0; q < 8; ++q)
                  dv[q] = drow != nullptr ? __ldg(reinterpret_cast<const float4*>(drow + zbase + c0) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  const float dd[4] = {dv[q].x, dv[q].y, dv[q].z, dv[q].w};
#pragma unroll
                  for (int u = 0; u < 4; ++u) {
                    const float e = ex2_approx(fmaf(__uint_as_float(r[4 * q + u]), sc2, off));
                    ps[u] += e;
                    pd[u] = fmaf(e, dd[u], pd[u]);
                  }
                }
              } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int z = zbase + c0 + j;
                if (z < a.Z) {
                  const float e = ex2_approx(fmaf(__uint_as_float(r[j]), sc2, off));
                  ps[j & 3] += e;
                  pd[j & 3] = fmaf(e, drow != nullptr ? __ldg(drow + z) : 0.0f, pd[j & 3]);
                }
              }
              }
              run_s += (ps[0] + ps[1]) + (ps[2] + ps[3]);
              run_d += (pd[0] + pd[1]) + (pd[2] + pd[3]);
            } else {
            float part[4] = {0.0f, 0.0f, 0.0f, 0.0f};          // four independent chains
            if (zbase + c0 + 32 <= a.Z) {
#pragma unroll
              for (int j = 0; j < 32; ++j) part[j & 3] += ex2_approx(fmaf(__uint_as_float(r[j]), sc2, off));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (zbase + c0 + j < a.Z) part[j & 3] += ex2_approx(fmaf(__uint_as_float(r[j]), sc2, off));
            }
            run_s += (part[0] + part[1]) + (part[2] + part[3]);
            }
          }
          if (need_top && mx > b2) {        // per lane rare after the first chunks, but a warp enters when ANY of its rows does
            // The group maximum is known: locate it (lowest column on ties), take the runner-up of the remaining columns,
            // and merge the pair into the running top-2 -- the same result as a left-to-right scan with strict '>',
            // at ~4 instead of ~8 instructions per column.
            const int nv = min(32, a.Z - (zbase + c0));           // valid columns of this group (>= 1 here)
            int jm = 0;
#pragma unroll
            for (int j = 31; j >= 0; --j)
              if (j < nv && __uint_as_float(r[j]) == mx) jm = j;
            float m2 = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j) m2 = fmaxf(m2, (j != jm && j < nv) ? __uint_as_float(r[j]) : -INFINITY);
            const int zm = zbase + c0 + jm;
            if (mx > b1) { b2 = b1; i2 = i1; b1 = mx; i1 = zm; }
            else { b2 = mx; i2 = zm; }
            if (m2 > b2) {                  // rarer still: the group's runner-up qualifies too (it can only take second place)
              int j2 = 0;
#pragma unroll
              for (int j = 31; j >= 0; --j)
                if (j != jm && j < nv && __uint_as_float(r[j]) == m2) j2 = j;
              b2 = m2; i2 = zbase + c0 + j2;
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&acc_empty[buf]);
      }
      if (dead) break;
      // -- merge the two column halves of the row: top-2 of four candidates (lower zone index wins ties, as a single
      //    left-to-right sweep would have it), log-sum-exp states combined on the common maximum
      if (hh == 1) { x_b1[row] = b1; x_i1[row] = i1; x_b2[row] = b2; x_i2[row] = i2; x_m[row] = run_m; x_s[row] = run_s; x_d[row] = run_d; }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (hh == 0) {
        const float c1 = x_b1[row], c2 = x_b2[row];
        const int j1 = x_i1[row], j2 = x_i2[row];
        // "better" = larger value, lower zone index on ties (what one left-to-right sweep over all zones would keep)
        auto better = [](float va, int ia, float vb, int ib) { return va > vb || (va == vb && ia < ib); };
        if (better(c1, j1, b1, i1)) {
          const bool keep_b1 = better(b1, i1, c2, j2);
          b2 = keep_b1 ? b1 : c2; i2 = keep_b1 ? i1 : j2;
          b1 = c1; i1 = j1;
        } else if (better(c1, j1, b2, i2)) { b2 = c1; i2 = j1; }
        if (ce) {
          const float om = x_m[row], os = x_s[row];
          const float mn = fmaxf(run_m, om);
          if (mn > -INFINITY) {
            const float f0 = run_m > -INFINITY ? exp2f((run_m - mn) * sc2) : 0.0f, f1 = om > -INFINITY ? exp2f((om - mn) * sc2) : 0.0f;
            run_s = run_s * f0 + os * f1;
            run_d = run_d * f0 + x_d[row] * f1;
            run_m = mn;
          }
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");       // the exchange buffers are free for the next row tile
      // -- exact fp32 re-score of the two nominees
      if (valid && hh == 0) {
        const float4* er = reinterpret_cast<const float4*>(a.emb + m * HD_E);
        const float4* t1 = reinterpret_cast<const float4*>(a.tn + (size_t)i1 * HD_E);
        const float4* t2 = reinterpret_cast<const float4*>(a.tn + (size_t)i2 * HD_E);
        float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
        for (int j = 0; j < HD_E / 4; ++j) {
          const float4 x = er[j], u = t1[j], w = t2[j];
          const float e0 = x.x * inv, e1 = x.y * inv, e2 = x.z * inv, e3 = x.w * inv;
          s1 = fmaf(e0, u.x, s1); s1 = fmaf(e1, u.y, s1); s1 = fmaf(e2, u.z, s1); s1 = fmaf(e3, u.w, s1);
          s2 = fmaf(e0, w.x, s2); s2 = fmaf(e1, w.y, s2); s2 = fmaf(e2, w.z, s2); s2 = fmaf(e3, w.w, s2);
        }
        const bool second = (a.Z > 1) && (s2 > s1 || (s2 == s1 && i2 < i1));
        if (a.labels != nullptr) a.labels[m] = second ? i2 : i1;
        if (a.best != nullptr) a.best[m] = (second ? s2 : s1) * a.inv_tau;
        if (ce) {
          int64_t tg = a.target[m];
          if (tg < 0 || tg >= a.Z) tg = 0;
          const float4* tt = reinterpret_cast<const float4*>(a.tn + (size_t)tg * HD_E);
          float st = 0.0f;
#pragma unroll
          for (int j = 0; j < HD_E / 4; ++j) {
            const float4 x = er[j], u = tt[j];
            st = fmaf(x.x * inv, u.x, st); st = fmaf(x.y * inv, u.y, st); st = fmaf(x.z * inv, u.z, st); st = fmaf(x.w * inv, u.w, st);
          }
          a.lse[m] = run_m * a.inv_tau + logf(run_s);
          a.tgt_logit[m] = st * a.inv_tau;
          if (want_d) a.exp_dist[m] = run_d / run_s;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

// ---- host side --------------------------------------------------------------------------------------------------
static int head_chunks(int Z) { return (Z + HD_TN - 1) / HD_TN; }
size_t head_workspace_bytes(int Z) {
  const size_t nc = head_chunks(Z);
  return align_up(nc * HD_TN * HD_E * sizeof(float), 256) + nc * HD_B_BYTES + 256;
}

static int head_launch(const float* emb, const float* table, int64_t M, int Z, int E, float tau, int64_t* labels, float* best,
                       const int64_t* target, float* lse, float* tgt_logit, const float* dist, float* exp_dist, void* ws,
                       size_t ws_bytes, cudaStream_t st);

int head_argmax(const float* emb, const float* table, int64_t M, int Z, int E, float tau, int64_t* labels, float* best, void* ws,
                size_t ws_bytes, cudaStream_t st) {
  return head_launch(emb, table, M, Z, E, tau, labels, best, nullptr, nullptr, nullptr, nullptr, nullptr, ws, ws_bytes, st);
}

// Cross-entropy forward without the [M, Z] logits: per row the log-sum-exp over all zones and the target's logit
// (loss_row = lse - tgt_logit; masking / averaging is the caller's), optionally the argmax labels in the same pass.
int head_ce_forward(const float* emb, const float* table, const int64_t* target, int64_t M, int Z, int E, float tau, float* lse,
                    float* tgt_logit, int64_t* labels, const float* dist, float* exp_dist, void* ws, size_t ws_bytes, cudaStream_t st) {
  return head_launch(emb, table, M, Z, E, tau, labels, nullptr, target, lse, tgt_logit, dist, exp_dist, ws, ws_bytes, st);
}

static int head_launch(const float* emb, const float* table, int64_t M, int Z, int E, float tau, int64_t* labels, float* best,
                       const int64_t* target, float* lse, float* tgt_logit, const float* dist, float* exp_dist, void* ws,
                       size_t ws_bytes, cudaStream_t st) {
  if (E != HD_E) return AB200_ERR_UNSUPPORTED;
  if (ws_bytes < head_workspace_bytes(Z)) return AB200_ERR_WORKSPACE;
  const int nc = head_chunks(Z);
  float* tn = (float*)ws;
  uint8_t* img = (uint8_t*)ws + align_up((size_t)nc * HD_TN * HD_E * sizeof(float), 256);
  int* status = (int*)(img + (size_t)nc * HD_B_BYTES);
  cudaError_t e = cudaMemsetAsync(status, 0, sizeof(int), st);
  if (e != cudaSuccess) { set_cuda_error(e); return AB200_ERR_CUDA; }
  head_pack_table_kernel<<<nc, 128, 0, st>>>(table, Z, tn, img);
  int rc = check_launch();
  if (rc) return rc;
  HeadArgs k{emb, tn, img, M, Z, nc, (int)((M + HD_TM - 1) / HD_TM), 1.0f / tau, labels, best, status, target, lse, tgt_logit, dist, exp_dist};
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = k.ntiles < sms ? k.ntiles : sms;
  e = cudaFuncSetAttribute(head_argmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HD_SMEM);
  if (e != cudaSuccess) { set_cuda_error(e); return AB200_ERR_CUDA; }
  head_argmax_kernel<<<grid, HD_THREADS, HD_SMEM, st>>>(k);
  return check_launch();
}

}  // namespace ab200
