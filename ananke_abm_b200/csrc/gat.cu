// Graph attention over the zone graph (PyG GATConv semantics): fused edge-softmax
// SDDMM + SpMM over a CSR sorted by destination, one warp per destination row.
//
//   project   : xw[z, h, :] = W_h x_z ;  a_src[z,h] = <att_src[h], xw[z,h]> ; a_dst[z,h] = <att_dst[h], xw[z,h]>
//   aggregate : e_ij = leaky_relu(a_src[j,h] + a_dst[i,h]) ; alpha = softmax_j(e_ij) over the in-edges of i
//               out[i,h,:] = sum_j alpha_ij xw[j,h,:]   (+ bias; heads concatenated or averaged)
// Each lane owns 4 consecutive channels (one 128-bit gather per neighbour row); the softmax statistics of a head are
// recomputed redundantly by the lanes of that head (degree ~7), so the SDDMM scores never leave registers.
// The whole working set at SA1 scale (Z=10k, nnz=70k, 64 channels: ~6 MB) is L2 resident; these kernels are
// bandwidth/latency bound and are reported against the HBM roofline separately from the agent-side kernels.
#include "common.cuh"

namespace ab200 {

__device__ __forceinline__ float lrelu(float x, float s) { return x > 0.f ? x : s * x; }

// one warp per node; HF = heads * F_out channels
__global__ void __launch_bounds__(256) gat_project_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                          const float* __restrict__ att_src, const float* __restrict__ att_dst,
                                                          int Z, int F_in, int heads, int F_out, float* __restrict__ xw,
                                                          float* __restrict__ a_src, float* __restrict__ a_dst) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= Z) return;
  const int HF = heads * F_out;
  const float* xr = x + (size_t)warp * F_in;
  for (int c0 = 0; c0 < HF; c0 += 32) {
    const int c = c0 + lane;
    float acc = 0.f;
    if (c < HF) {
      const float* wr = W + (size_t)c * F_in;
      for (int k = 0; k < F_in; ++k) acc = fmaf(xr[k], wr[k], acc);
      xw[(size_t)warp * HF + c] = acc;
    }
    // per-head dot products with the attention vectors: segmented reduction over F_out consecutive lanes/channels
    float ps = (c < HF) ? acc * att_src[c] : 0.f;
    float pd = (c < HF) ? acc * att_dst[c] : 0.f;
    if (F_out >= 32) {
      for (int o = 16; o > 0; o >>= 1) { ps += __shfl_xor_sync(0xffffffffu, ps, o); pd += __shfl_xor_sync(0xffffffffu, pd, o); }
      if (lane == 0 && c < HF) { atomicAdd(&a_src[(size_t)warp * heads + c / F_out], ps); atomicAdd(&a_dst[(size_t)warp * heads + c / F_out], pd); }
    } else {
      for (int o = F_out >> 1; o > 0; o >>= 1) { ps += __shfl_xor_sync(0xffffffffu, ps, o); pd += __shfl_xor_sync(0xffffffffu, pd, o); }
      if ((lane % F_out) == 0 && c < HF) { a_src[(size_t)warp * heads + c / F_out] = ps; a_dst[(size_t)warp * heads + c / F_out] = pd; }
    }
  }
}

// one warp per destination row; lane owns channels 4*lane .. 4*lane+3 (+128 per pass)
__global__ void __launch_bounds__(256) gat_aggregate_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, int Z,
                                                            const float* __restrict__ xw, const float* __restrict__ a_src,
                                                            const float* __restrict__ a_dst, const float* __restrict__ bias,
                                                            int heads, int F_out, int concat, float slope, float* __restrict__ out,
                                                            float* __restrict__ alpha) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= Z) return;
  const int HF = heads * F_out;
  const int e0 = rowptr[i], e1 = rowptr[i + 1];
  for (int c0 = 0; c0 < HF; c0 += 128) {
    const int c = c0 + 4 * lane;
    const bool on = c < HF;
    const int h = on ? c / F_out : 0;
    const float ad = a_dst[(size_t)i * heads + h];
    float m = -INFINITY;
    for (int e = e0; e < e1; ++e) m = fmaxf(m, lrelu(a_src[(size_t)col[e] * heads + h] + ad, slope));
    float s = 0.f;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = e0; e < e1; ++e) {
      const int j = col[e];
      const float w = __expf(lrelu(a_src[(size_t)j * heads + h] + ad, slope) - m);
      s += w;
      if (on) {
        const float4 v = *reinterpret_cast<const float4*>(xw + (size_t)j * HF + c);
        acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
      }
    }
    const float inv = (e1 > e0) ? 1.f / s : 0.f;
    if (on) {
      // normalised attention coefficients, kept for the backward pass: one writer per (edge, head)
      if ((c % F_out) == 0)
        for (int e = e0; e < e1; ++e)
          alpha[(size_t)e * heads + h] = __expf(lrelu(a_src[(size_t)col[e] * heads + h] + ad, slope) - m) * inv;
      acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
      if (concat) {
        const float4 b = bias ? *reinterpret_cast<const float4*>(bias + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(out + (size_t)i * HF + c) = make_float4(acc.x + b.x, acc.y + b.y, acc.z + b.z, acc.w + b.w);
      } else {
        const float sc = 1.f / heads;
        const int f = c % F_out;
        atomicAdd(out + (size_t)i * F_out + f + 0, acc.x * sc);
        atomicAdd(out + (size_t)i * F_out + f + 1, acc.y * sc);
        atomicAdd(out + (size_t)i * F_out + f + 2, acc.z * sc);
        atomicAdd(out + (size_t)i * F_out + f + 3, acc.w * sc);
      }
    }
  }
}

__global__ void gat_mean_init_kernel(float* __restrict__ out, const float* __restrict__ bias, int Z, int F_out) {
  const int64_t n = (int64_t)Z * F_out;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = bias ? bias[i % F_out] : 0.f;
}

// ---- backward ---------------------------------------------------------------------------------------------
// B1: per destination i (warp per row).  g[i,h,:] = dL/d out (per head, pre-bias);  de_ij = alpha_ij (dalpha_ij - c_i) lrelu'
//     with dalpha_ij = <g[i,h], xw[j,h]>,  c_i = sum_k alpha_ik dalpha_ik.  Writes de[e,h] and d a_dst[i,h].
__global__ void __launch_bounds__(256) gat_bwd_dst_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, int Z,
                                                          const float* __restrict__ xw, const float* __restrict__ a_src,
                                                          const float* __restrict__ a_dst, const float* __restrict__ alpha,
                                                          const float* __restrict__ gout, int heads, int F_out, int concat, float slope,
                                                          float* __restrict__ de, float* __restrict__ d_adst) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= Z) return;
  const int HF = heads * F_out;
  const int e0 = rowptr[i], e1 = rowptr[i + 1];
  const int lph = F_out / 4;                     // lanes per head
  for (int c0 = 0; c0 < HF; c0 += 128) {
    const int c = c0 + 4 * lane;
    const bool on = c < HF;
    const int h = on ? c / F_out : 0;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (on) {
      if (concat) g = *reinterpret_cast<const float4*>(gout + (size_t)i * HF + c);
      else {
        const float sc = 1.f / heads;
        const float4 t = *reinterpret_cast<const float4*>(gout + (size_t)i * F_out + (c % F_out));
        g = make_float4(t.x * sc, t.y * sc, t.z * sc, t.w * sc);
      }
    }
    // pass 1: c_i = sum_k alpha_ik <g, xw_k>
    float ci = 0.f;
    for (int e = e0; e < e1; ++e) {
      float d = 0.f;
      if (on) {
        const float4 v = *reinterpret_cast<const float4*>(xw + (size_t)col[e] * HF + c);
        d = g.x * v.x + g.y * v.y + g.z * v.z + g.w * v.w;
      }
      for (int o = lph >> 1; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      if (on) ci += alpha[(size_t)e * heads + h] * d;
    }
    float dad = 0.f;
    const float ad = a_dst[(size_t)i * heads + h];
    for (int e = e0; e < e1; ++e) {
      const int j = col[e];
      float d = 0.f;
      if (on) {
        const float4 v = *reinterpret_cast<const float4*>(xw + (size_t)j * HF + c);
        d = g.x * v.x + g.y * v.y + g.z * v.z + g.w * v.w;
      }
      for (int o = lph >> 1; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      if (on) {
        const float pre = a_src[(size_t)j * heads + h] + ad;
        const float dev = alpha[(size_t)e * heads + h] * (d - ci) * (pre > 0.f ? 1.f : slope);
        dad += dev;
        if ((c % F_out) == 0) de[(size_t)e * heads + h] = dev;
      }
    }
    if (on && (c % F_out) == 0) d_adst[(size_t)i * heads + h] = dad;
  }
}

// B2: per source j (warp per row of the transposed CSR): dxw[j,h,:] = sum_i alpha_ij g[i,h,:] + d_asrc att_src + d_adst att_dst
__global__ void __launch_bounds__(256) gat_bwd_src_kernel(const int* __restrict__ rowptr_t, const int* __restrict__ col_t,
                                                          const int* __restrict__ eid_t, int Z, const float* __restrict__ alpha,
                                                          const float* __restrict__ de, const float* __restrict__ gout,
                                                          const float* __restrict__ d_adst, const float* __restrict__ att_src,
                                                          const float* __restrict__ att_dst, int heads, int F_out, int concat,
                                                          float* __restrict__ dxw, float* __restrict__ d_asrc) {
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (j >= Z) return;
  const int HF = heads * F_out;
  const int e0 = rowptr_t[j], e1 = rowptr_t[j + 1];
  for (int c0 = 0; c0 < HF; c0 += 128) {
    const int c = c0 + 4 * lane;
    if (c >= HF) continue;
    const int h = c / F_out;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float das = 0.f;
    for (int e = e0; e < e1; ++e) {
      const int i = col_t[e], eid = eid_t[e];
      const float al = alpha[(size_t)eid * heads + h];
      float4 g;
      if (concat) g = *reinterpret_cast<const float4*>(gout + (size_t)i * HF + c);
      else {
        const float sc = 1.f / heads;
        const float4 t = *reinterpret_cast<const float4*>(gout + (size_t)i * F_out + (c % F_out));
        g = make_float4(t.x * sc, t.y * sc, t.z * sc, t.w * sc);
      }
      acc.x = fmaf(al, g.x, acc.x); acc.y = fmaf(al, g.y, acc.y); acc.z = fmaf(al, g.z, acc.z); acc.w = fmaf(al, g.w, acc.w);
      das += de[(size_t)eid * heads + h];
    }
    const float dad = d_adst[(size_t)j * heads + h];
    const float4 as = *reinterpret_cast<const float4*>(att_src + c);
    const float4 at = *reinterpret_cast<const float4*>(att_dst + c);
    acc.x += das * as.x + dad * at.x; acc.y += das * as.y + dad * at.y;
    acc.z += das * as.z + dad * at.z; acc.w += das * as.w + dad * at.w;
    *reinterpret_cast<float4*>(dxw + (size_t)j * HF + c) = acc;
    if ((c % F_out) == 0) d_asrc[(size_t)j * heads + h] = das;
  }
}

// B3: parameter gradients (reductions over zones).  One block per output channel c (of HF):
//   dW[c, k] = sum_z dxw[z,c] x[z,k] ; d att_src[c] = sum_z d_asrc[z,h] xw[z,c] ; d att_dst[c] likewise ; d bias
__global__ void __launch_bounds__(256) gat_bwd_param_kernel(const float* __restrict__ x, const float* __restrict__ xw,
                                                            const float* __restrict__ dxw, const float* __restrict__ d_asrc,
                                                            const float* __restrict__ d_adst, const float* __restrict__ gout, int Z,
                                                            int F_in, int heads, int F_out, int concat, float* __restrict__ dW,
                                                            float* __restrict__ datt_src, float* __restrict__ datt_dst,
                                                            float* __restrict__ dbias) {
  const int c = blockIdx.x, HF = heads * F_out, h = c / F_out;
  extern __shared__ float red[];      // [(F_in + 3)][256]
  const int nq = F_in + 3;
  float accq[36];
  for (int q = 0; q < nq; ++q) accq[q] = 0.f;
  for (int z = threadIdx.x; z < Z; z += blockDim.x) {
    const float d = dxw[(size_t)z * HF + c];
    for (int k = 0; k < F_in; ++k) accq[k] = fmaf(d, x[(size_t)z * F_in + k], accq[k]);
    const float v = xw[(size_t)z * HF + c];
    accq[F_in] = fmaf(d_asrc[(size_t)z * heads + h], v, accq[F_in]);
    accq[F_in + 1] = fmaf(d_adst[(size_t)z * heads + h], v, accq[F_in + 1]);
    if (concat) accq[F_in + 2] += gout[(size_t)z * HF + c];
    else if (c < F_out) accq[F_in + 2] += gout[(size_t)z * F_out + c];
  }
  for (int q = 0; q < nq; ++q) red[q * 256 + threadIdx.x] = accq[q];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s)
      for (int q = 0; q < nq; ++q) red[q * 256 + threadIdx.x] += red[q * 256 + threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    for (int k = 0; k < F_in; ++k) dW[(size_t)c * F_in + k] = red[k * 256];
    datt_src[c] = red[F_in * 256];
    datt_dst[c] = red[(F_in + 1) * 256];
    if (dbias && (concat || c < F_out)) dbias[c] = red[(F_in + 2) * 256];
  }
}

// dx[z,k] = sum_c dxw[z,c] W[c,k]   (only when the caller wants gradients w.r.t. the zone features)
__global__ void gat_bwd_input_kernel(const float* __restrict__ dxw, const float* __restrict__ W, int Z, int F_in, int HF,
                                     float* __restrict__ dx) {
  const int64_t n = (int64_t)Z * F_in;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int z = (int)(i / F_in), k = (int)(i % F_in);
    float acc = 0.f;
    for (int c = 0; c < HF; ++c) acc = fmaf(dxw[(size_t)z * HF + c], W[(size_t)c * F_in + k], acc);
    dx[i] = acc;
  }
}


// =========================================================================================================
// Sub-warp kernels for the shapes the zone graph uses (heads * F_out = 8 .. 128 channels, F_in <= 8).
// LPR = channels / 4 lanes own one row (each lane 4 consecutive channels), so a warp works on 32 / LPR rows at once
// and no lane idles.  The forward is ONE kernel in the reordered (linear) form
//     out[i,h,:] = W_h (sum_j alpha_ijh x_j),   a_src[j,h] = (W_h^T att_src_h) . x_j
// i.e. neighbours are gathered in the F_in-dimensional input space (28 B per edge instead of a 256 B projected row);
// xw, a_src, a_dst and alpha are still written for the backward pass (one coalesced store each).
// =========================================================================================================
constexpr int GAT_FIN = 8;

template <int LPR>
__global__ void __launch_bounds__(256) gat_fwd_fused_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, int Z,
                                                            const float* __restrict__ x, int F_in, const float* __restrict__ W,
                                                            const float* __restrict__ att_src, const float* __restrict__ att_dst,
                                                            const float* __restrict__ bias, int heads, int F_out, int concat,
                                                            float slope, float* __restrict__ out, float* __restrict__ xw,
                                                            float* __restrict__ a_src, float* __restrict__ a_dst,
                                                            float* __restrict__ alpha) {
  constexpr int RPW = 32 / LPR, HF = 4 * LPR;
  __shared__ float u_s[32 * GAT_FIN], u_d[32 * GAT_FIN];     // W_h^T att_src_h, W_h^T att_dst_h
  for (int idx = threadIdx.x; idx < heads * GAT_FIN; idx += blockDim.x) {
    const int h = idx / GAT_FIN, k = idx % GAT_FIN;
    float s = 0.f, d = 0.f;
    if (k < F_in)
      for (int f = 0; f < F_out; ++f) {
        const float w = W[(size_t)(h * F_out + f) * F_in + k];
        s = fmaf(att_src[h * F_out + f], w, s);
        d = fmaf(att_dst[h * F_out + f], w, d);
      }
    u_s[idx] = s; u_d[idx] = d;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane / LPR, sl = lane % LPR;
  const int c = 4 * sl, h = c / F_out, lph = F_out / 4;
  const bool leader = (sl % lph) == 0;
  float wr[4][GAT_FIN], us[GAT_FIN];
#pragma unroll
  for (int k = 0; k < GAT_FIN; ++k) {
    us[k] = u_s[h * GAT_FIN + k];
#pragma unroll
    for (int q = 0; q < 4; ++q) wr[q][k] = k < F_in ? W[(size_t)(c + q) * F_in + k] : 0.f;
  }
  const float4 as4 = *reinterpret_cast<const float4*>(att_src + c), ad4 = *reinterpret_cast<const float4*>(att_dst + c);
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias) b4 = *reinterpret_cast<const float4*>(bias + (concat ? c : c % F_out));
  const int nwarp = (gridDim.x * blockDim.x) >> 5;
  for (int rg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; rg * RPW < Z; rg += nwarp) {
    const int i = rg * RPW + g;
    const bool on = i < Z;
    // ---- own row: projection (kept for the backward pass) and the two attention logits
    float xi[GAT_FIN];
#pragma unroll
    for (int k = 0; k < GAT_FIN; ++k) xi[k] = (on && k < F_in) ? x[(size_t)i * F_in + k] : 0.f;
    float p[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < GAT_FIN; ++k) acc = fmaf(xi[k], wr[q][k], acc);
      p[q] = acc;
    }
    if (on) *reinterpret_cast<float4*>(xw + (size_t)i * HF + c) = make_float4(p[0], p[1], p[2], p[3]);
    float ps = p[0] * as4.x + p[1] * as4.y + p[2] * as4.z + p[3] * as4.w;
    float pd = p[0] * ad4.x + p[1] * ad4.y + p[2] * ad4.z + p[3] * ad4.w;
    for (int o = lph >> 1; o > 0; o >>= 1) { ps += __shfl_xor_sync(0xffffffffu, ps, o); pd += __shfl_xor_sync(0xffffffffu, pd, o); }
    if (on && leader) { a_src[(size_t)i * heads + h] = ps; a_dst[(size_t)i * heads + h] = pd; }
    const float ad = pd;
    int e0 = 0, deg = 0;
    if (on) { e0 = rowptr[i]; deg = rowptr[i + 1] - e0; }
    int maxdeg = deg;
    for (int o = 16; o > 0; o >>= 1) maxdeg = max(maxdeg, __shfl_xor_sync(0xffffffffu, maxdeg, o));
    // ---- pass 1: running max and sum of the edge scores (input-space gathers only)
    float m = -INFINITY, ssum = 0.f;
    for (int b0 = 0; b0 < maxdeg; b0 += LPR) {
      const int jm = (b0 + sl < deg) ? col[e0 + b0 + sl] : -1;
#pragma unroll
      for (int t = 0; t < LPR; ++t) {
        if (b0 + t >= maxdeg) break;
        const int j = __shfl_sync(0xffffffffu, jm, g * LPR + t);
        if (j >= 0) {
          float pre = ad;
#pragma unroll
          for (int k = 0; k < GAT_FIN; ++k)
            if (k < F_in) pre = fmaf(us[k], x[(size_t)j * F_in + k], pre);
          const float sc = lrelu(pre, slope), mn = fmaxf(m, sc);
          ssum = ssum * __expf(m - mn) + __expf(sc - mn);
          m = mn;
        }
      }
    }
    const float inv = deg > 0 ? 1.f / ssum : 0.f;
    // ---- pass 2: normalised coefficients (stored) and the input-space aggregate
    float agg[GAT_FIN];
#pragma unroll
    for (int k = 0; k < GAT_FIN; ++k) agg[k] = 0.f;
    for (int b0 = 0; b0 < maxdeg; b0 += LPR) {
      const int jm = (b0 + sl < deg) ? col[e0 + b0 + sl] : -1;
#pragma unroll
      for (int t = 0; t < LPR; ++t) {
        if (b0 + t >= maxdeg) break;
        const int j = __shfl_sync(0xffffffffu, jm, g * LPR + t);
        if (j >= 0) {
          float xj[GAT_FIN];
          float pre = ad;
#pragma unroll
          for (int k = 0; k < GAT_FIN; ++k) {
            xj[k] = k < F_in ? x[(size_t)j * F_in + k] : 0.f;
            pre = fmaf(us[k], xj[k], pre);
          }
          const float w = __expf(lrelu(pre, slope) - m) * inv;
#pragma unroll
          for (int k = 0; k < GAT_FIN; ++k) agg[k] = fmaf(w, xj[k], agg[k]);
          if (leader) alpha[(size_t)(e0 + b0 + t) * heads + h] = w;
        }
      }
    }
    float o4[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < GAT_FIN; ++k) acc = fmaf(agg[k], wr[q][k], acc);
      o4[q] = acc;
    }
    if (concat) {
      if (on) *reinterpret_cast<float4*>(out + (size_t)i * HF + c) = make_float4(o4[0] + b4.x, o4[1] + b4.y, o4[2] + b4.z, o4[3] + b4.w);
    } else {
      // mean over heads: lanes that own the same channels of different heads sit lph * 2^r lanes apart
      for (int o = lph; o < LPR; o <<= 1) {
#pragma unroll
        for (int q = 0; q < 4; ++q) o4[q] += __shfl_xor_sync(0xffffffffu, o4[q], o);
      }
      const float sc = 1.f / heads;
      if (on && h == 0) *reinterpret_cast<float4*>(out + (size_t)i * F_out + c) =
          make_float4(o4[0] * sc + b4.x, o4[1] * sc + b4.y, o4[2] * sc + b4.z, o4[3] * sc + b4.w);
    }
  }
}

// B1 on sub-warps: dalpha_ij = <g_ih, W_h x_j> = (W_h^T g_ih) . x_j, so the destination pass gathers 28 B per edge too
template <int LPR>
__global__ void __launch_bounds__(256) gat_bwd_dst_sub_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, int Z,
                                                              const float* __restrict__ x, int F_in, const float* __restrict__ W,
                                                              const float* __restrict__ a_src, const float* __restrict__ a_dst,
                                                              const float* __restrict__ alpha, const float* __restrict__ gout,
                                                              int heads, int F_out, int concat, float slope, float* __restrict__ de,
                                                              float* __restrict__ d_adst) {
  constexpr int RPW = 32 / LPR, HF = 4 * LPR;
  const int lane = threadIdx.x & 31, g = lane / LPR, sl = lane % LPR;
  const int c = 4 * sl, h = c / F_out, lph = F_out / 4;
  const bool leader = (sl % lph) == 0;
  float wr[4][GAT_FIN];
#pragma unroll
  for (int k = 0; k < GAT_FIN; ++k)
#pragma unroll
    for (int q = 0; q < 4; ++q) wr[q][k] = k < F_in ? W[(size_t)(c + q) * F_in + k] : 0.f;
  const int nwarp = (gridDim.x * blockDim.x) >> 5;
  for (int rg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; rg * RPW < Z; rg += nwarp) {
    const int i = rg * RPW + g;
    const bool on = i < Z;
    float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (on) {
      if (concat) gv = *reinterpret_cast<const float4*>(gout + (size_t)i * HF + c);
      else {
        const float sc = 1.f / heads;
        const float4 t = *reinterpret_cast<const float4*>(gout + (size_t)i * F_out + (c % F_out));
        gv = make_float4(t.x * sc, t.y * sc, t.z * sc, t.w * sc);
      }
    }
    float qv[GAT_FIN];
#pragma unroll
    for (int k = 0; k < GAT_FIN; ++k) {
      float v = gv.x * wr[0][k] + gv.y * wr[1][k] + gv.z * wr[2][k] + gv.w * wr[3][k];
      for (int o = lph >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      qv[k] = v;
    }
    int e0 = 0, deg = 0;
    if (on) { e0 = rowptr[i]; deg = rowptr[i + 1] - e0; }
    int maxdeg = deg;
    for (int o = 16; o > 0; o >>= 1) maxdeg = max(maxdeg, __shfl_xor_sync(0xffffffffu, maxdeg, o));
    const float ad = on ? a_dst[(size_t)i * heads + h] : 0.f;
    float ci = 0.f;
    for (int b0 = 0; b0 < maxdeg; b0 += LPR) {
      const int jm = (b0 + sl < deg) ? col[e0 + b0 + sl] : -1;
#pragma unroll
      for (int t = 0; t < LPR; ++t) {
        if (b0 + t >= maxdeg) break;
        const int j = __shfl_sync(0xffffffffu, jm, g * LPR + t);
        if (j >= 0) {
          float d = 0.f;
#pragma unroll
          for (int k = 0; k < GAT_FIN; ++k)
            if (k < F_in) d = fmaf(qv[k], x[(size_t)j * F_in + k], d);
          ci = fmaf(alpha[(size_t)(e0 + b0 + t) * heads + h], d, ci);
        }
      }
    }
    float dad = 0.f;
    for (int b0 = 0; b0 < maxdeg; b0 += LPR) {
      const int jm = (b0 + sl < deg) ? col[e0 + b0 + sl] : -1;
#pragma unroll
      for (int t = 0; t < LPR; ++t) {
        if (b0 + t >= maxdeg) break;
        const int j = __shfl_sync(0xffffffffu, jm, g * LPR + t);
        if (j >= 0) {
          float d = 0.f;
#pragma unroll
          for (int k = 0; k < GAT_FIN; ++k)
            if (k < F_in) d = fmaf(qv[k], x[(size_t)j * F_in + k], d);
          const size_t eh = (size_t)(e0 + b0 + t) * heads + h;
          const float pre = a_src[(size_t)j * heads + h] + ad;
          const float dev = alpha[eh] * (d - ci) * (pre > 0.f ? 1.f : slope);
          dad += dev;
          if (leader) de[eh] = dev;
        }
      }
    }
    if (on && leader) d_adst[(size_t)i * heads + h] = dad;
  }
}

// B2 on sub-warps (transposed CSR): same outputs as gat_bwd_src_kernel
template <int LPR>
__global__ void __launch_bounds__(256) gat_bwd_src_sub_kernel(const int* __restrict__ rowptr_t, const int* __restrict__ col_t,
                                                              const int* __restrict__ eid_t, int Z, const float* __restrict__ alpha,
                                                              const float* __restrict__ de, const float* __restrict__ gout,
                                                              const float* __restrict__ d_adst, const float* __restrict__ att_src,
                                                              const float* __restrict__ att_dst, int heads, int F_out, int concat,
                                                              float* __restrict__ dxw, float* __restrict__ d_asrc) {
  constexpr int RPW = 32 / LPR, HF = 4 * LPR;
  const int lane = threadIdx.x & 31, g = lane / LPR, sl = lane % LPR;
  const int c = 4 * sl, h = c / F_out, lph = F_out / 4;
  const bool leader = (sl % lph) == 0;
  const float4 as4 = *reinterpret_cast<const float4*>(att_src + c), at4 = *reinterpret_cast<const float4*>(att_dst + c);
  const float sc = concat ? 1.f : 1.f / heads;
  const int nwarp = (gridDim.x * blockDim.x) >> 5;
  for (int rg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; rg * RPW < Z; rg += nwarp) {
    const int j = rg * RPW + g;
    const bool on = j < Z;
    int e0 = 0, deg = 0;
    if (on) { e0 = rowptr_t[j]; deg = rowptr_t[j + 1] - e0; }
    int maxdeg = deg;
    for (int o = 16; o > 0; o >>= 1) maxdeg = max(maxdeg, __shfl_xor_sync(0xffffffffu, maxdeg, o));
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float das = 0.f;
    for (int b0 = 0; b0 < maxdeg; b0 += LPR) {
      const bool has = b0 + sl < deg;
      const int im = has ? col_t[e0 + b0 + sl] : -1;
      const int em = has ? eid_t[e0 + b0 + sl] : 0;
      constexpr int GRP = LPR < 4 ? LPR : 4;       // loads of a group are independent: issued together
      for (int t0 = 0; t0 < LPR; t0 += GRP) {
        if (b0 + t0 >= maxdeg) break;
        int ii[GRP], ee[GRP];
        float al[GRP], dd[GRP];
        float4 gv[GRP];
#pragma unroll
        for (int u = 0; u < GRP; ++u) {
          ii[u] = __shfl_sync(0xffffffffu, im, g * LPR + t0 + u);
          ee[u] = __shfl_sync(0xffffffffu, em, g * LPR + t0 + u);
        }
#pragma unroll
        for (int u = 0; u < GRP; ++u) {
          al[u] = 0.f; dd[u] = 0.f; gv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ii[u] >= 0) {
            al[u] = alpha[(size_t)ee[u] * heads + h] * sc;
            dd[u] = de[(size_t)ee[u] * heads + h];
            gv[u] = *reinterpret_cast<const float4*>(gout + (concat ? (size_t)ii[u] * HF + c : (size_t)ii[u] * F_out + (c % F_out)));
          }
        }
#pragma unroll
        for (int u = 0; u < GRP; ++u) {
          acc.x = fmaf(al[u], gv[u].x, acc.x); acc.y = fmaf(al[u], gv[u].y, acc.y);
          acc.z = fmaf(al[u], gv[u].z, acc.z); acc.w = fmaf(al[u], gv[u].w, acc.w);
          das += dd[u];
        }
      }
    }
    if (on) {
      const float dad = d_adst[(size_t)j * heads + h];
      acc.x += das * as4.x + dad * at4.x; acc.y += das * as4.y + dad * at4.y;
      acc.z += das * as4.z + dad * at4.z; acc.w += das * as4.w + dad * at4.w;
      *reinterpret_cast<float4*>(dxw + (size_t)j * HF + c) = acc;
      if (leader) d_asrc[(size_t)j * heads + h] = das;
    }
  }
}

// B3, coalesced: thread = (zone lane, channel); every block walks a strided set of zones over ALL channels (row reads
// of dxw / xw / gout are contiguous), keeps F_in + 3 partial sums per thread and leaves one partial row per block;
// gat_param_finalize_kernel adds the blocks in a fixed order (deterministic, no atomics).
template <int HF>
__global__ void __launch_bounds__(256) gat_bwd_param_tiled_kernel(const float* __restrict__ x, const float* __restrict__ xw,
                                                                  const float* __restrict__ dxw, const float* __restrict__ d_asrc,
                                                                  const float* __restrict__ d_adst, const float* __restrict__ gout,
                                                                  int Z, int F_in, int heads, int F_out, int concat,
                                                                  float* __restrict__ partial) {
  constexpr int ZL = 256 / HF, NQ = GAT_FIN + 3;
  const int c = threadIdx.x % HF, zl = threadIdx.x / HF, h = c / F_out;
  float acc[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) acc[q] = 0.f;
  const int zstep = gridDim.x * ZL;
  for (int z0 = blockIdx.x * ZL + zl; z0 < Z; z0 += 4 * zstep) {
    float d[4], v[4], ds[4], dd[4], gb[4], xv[4][GAT_FIN];
#pragma unroll
    for (int u = 0; u < 4; ++u) {                 // all loads of four zones in flight before the first FMA
      const int z = z0 + u * zstep;
      const bool ok = z < Z;
      d[u] = ok ? dxw[(size_t)z * HF + c] : 0.f;
      v[u] = ok ? xw[(size_t)z * HF + c] : 0.f;
      ds[u] = ok ? d_asrc[(size_t)z * heads + h] : 0.f;
      dd[u] = ok ? d_adst[(size_t)z * heads + h] : 0.f;
      gb[u] = 0.f;
      if (ok) { if (concat) gb[u] = gout[(size_t)z * HF + c]; else if (c < F_out) gb[u] = gout[(size_t)z * F_out + c]; }
#pragma unroll
      for (int k = 0; k < GAT_FIN; ++k) xv[u][k] = (ok && k < F_in) ? x[(size_t)z * F_in + k] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int k = 0; k < GAT_FIN; ++k) acc[k] = fmaf(d[u], xv[u][k], acc[k]);
      acc[GAT_FIN] = fmaf(ds[u], v[u], acc[GAT_FIN]);
      acc[GAT_FIN + 1] = fmaf(dd[u], v[u], acc[GAT_FIN + 1]);
      acc[GAT_FIN + 2] += gb[u];
    }
  }
  __shared__ float red[NQ][256];
#pragma unroll
  for (int q = 0; q < NQ; ++q) red[q][threadIdx.x] = acc[q];
  __syncthreads();
  if (zl == 0) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      float s = 0.f;
      for (int l = 0; l < ZL; ++l) s += red[q][l * HF + c];
      partial[((size_t)blockIdx.x * NQ + q) * HF + c] = s;
    }
  }
}

__global__ void gat_param_finalize_kernel(const float* __restrict__ partial, int nblk, int HF, int F_in, int F_out, int concat,
                                          float* __restrict__ dW, float* __restrict__ datt_src, float* __restrict__ datt_dst,
                                          float* __restrict__ dbias) {
  constexpr int NQ = GAT_FIN + 3;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= NQ * HF) return;
  const int q = idx / HF, c = idx % HF;
  float s = 0.f;
  for (int b = 0; b < nblk; ++b) s += partial[((size_t)b * NQ + q) * HF + c];
  if (q < GAT_FIN) { if (q < F_in) dW[(size_t)c * F_in + q] = s; }
  else if (q == GAT_FIN) datt_src[c] = s;
  else if (q == GAT_FIN + 1) datt_dst[c] = s;
  else if (dbias && (concat || c < F_out)) dbias[c] = s;
}

// ---------------------------------------------------------------------------------------------------------
// Lane = (row, head): `heads` lanes share a destination row, each owns the FO channels of one head.  Nothing about a
// head's edge softmax is computed twice (the sub-warp kernels above repeat it in every lane of the head), no shuffle
// sits in the edge loops, and rows of different degree simply diverge at the loop tail.  W lives in shared memory,
// one padded block per head ([FO][8] + 4 floats: 128-bit rows, heads on different banks).
// ---------------------------------------------------------------------------------------------------------
template <int FO>
__global__ void __launch_bounds__(256, 3) gat_fwd_rowhead_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, int Z,
                                                              const float* __restrict__ x, int F_in, const float* __restrict__ W,
                                                              const float* __restrict__ att_src, const float* __restrict__ att_dst,
                                                              const float* __restrict__ bias, int heads, int concat, float slope,
                                                              float* __restrict__ out, float* __restrict__ xw,
                                                              float* __restrict__ a_src, float* __restrict__ a_dst,
                                                              float* __restrict__ alpha) {
  constexpr int WS = FO * GAT_FIN + 4;
  __shared__ __align__(16) float Wsm[128 * GAT_FIN + 32 * 4];
  __shared__ float att_s[128], att_d[128];
  const int HF = heads * FO;
  for (int idx = threadIdx.x; idx < HF * GAT_FIN; idx += blockDim.x) {
    const int c = idx / GAT_FIN, k = idx % GAT_FIN;
    Wsm[(c / FO) * WS + (c % FO) * GAT_FIN + k] = k < F_in ? W[(size_t)c * F_in + k] : 0.f;
  }
  for (int idx = threadIdx.x; idx < HF; idx += blockDim.x) { att_s[idx] = att_src[idx]; att_d[idx] = att_dst[idx]; }
  __syncthreads();
  const int lane = threadIdx.x & 31, rpw = 32 / heads, g = lane / heads, h = lane % heads;
  const float* Wh = Wsm + h * WS;
  float us[GAT_FIN];                     // W_h^T att_src_h: neighbour logits straight from the input features
#pragma unroll
  for (int k = 0; k < GAT_FIN; ++k) us[k] = 0.f;
  for (int f = 0; f < FO; ++f) {
    const float a = att_s[h * FO + f];
#pragma unroll
    for (int k = 0; k < GAT_FIN; ++k) us[k] = fmaf(a, Wh[f * GAT_FIN + k], us[k]);
  }
  const int nwarp = (gridDim.x * blockDim.x) >> 5;
  for (int rg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; rg * rpw < Z; rg += nwarp) {
    const int i = rg * rpw + g;
    const bool on = i < Z;
    float xi[GAT_FIN];
#pragma unroll
    for (int k = 0; k < GAT_FIN; ++k) xi[k] = (on && k < F_in) ? x[(size_t)i * F_in + k] : 0.f;
    float ps = 0.f, pd = 0.f;
#pragma unroll 1      // keep W in shared memory: unrolled, the 2 * FO weight rows get hoisted out of the row loop (220 registers)
    for (int f4 = 0; f4 < FO / 4; ++f4) {
      float p[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 w0 = *reinterpret_cast<const float4*>(Wh + (f4 * 4 + q) * GAT_FIN);
        const float4 w1 = *reinterpret_cast<const float4*>(Wh + (f4 * 4 + q) * GAT_FIN + 4);
        p[q] = xi[0] * w0.x + xi[1] * w0.y + xi[2] * w0.z + xi[3] * w0.w + xi[4] * w1.x + xi[5] * w1.y + xi[6] * w1.z + xi[7] * w1.w;
        ps = fmaf(p[q], att_s[h * FO + f4 * 4 + q], ps);
        pd = fmaf(p[q], att_d[h * FO + f4 * 4 + q], pd);
      }
      if (on) *reinterpret_cast<float4*>(xw + (size_t)i * HF + h * FO + f4 * 4) = make_float4(p[0], p[1], p[2], p[3]);
    }
    int e0 = 0, deg = 0;
    if (on) {
      a_src[(size_t)i * heads + h] = ps;
      a_dst[(size_t)i * heads + h] = pd;
      e0 = rowptr[i]; deg = rowptr[i + 1] - e0;
    }
    const float ad = pd;
    float m = -INFINITY, ssum = 0.f;
    for (int e = 0; e < deg; ++e) {
      const int j = col[e0 + e];
      float pre = ad;
#pragma unroll
      for (int k = 0; k < GAT_FIN; ++k)
        if (k < F_in) pre = fmaf(us[k], x[(size_t)j * F_in + k], pre);
      const float sc = lrelu(pre, slope), mn = fmaxf(m, sc);
      ssum = ssum * __expf(m - mn) + __expf(sc - mn);
      m = mn;
    }
    const float inv = deg > 0 ? 1.f / ssum : 0.f;
    float agg[GAT_FIN];
#pragma unroll
    for (int k = 0; k < GAT_FIN; ++k) agg[k] = 0.f;
    for (int e = 0; e < deg; ++e) {
      const int j = col[e0 + e];
      float xj[GAT_FIN];
      float pre = ad;
#pragma unroll
      for (int k = 0; k < GAT_FIN; ++k) {
        xj[k] = k < F_in ? x[(size_t)j * F_in + k] : 0.f;
        pre = fmaf(us[k], xj[k], pre);
      }
      const float w = __expf(lrelu(pre, slope) - m) * inv;
      alpha[(size_t)(e0 + e) * heads + h] = w;
#pragma unroll
      for (int k = 0; k < GAT_FIN; ++k) agg[k] = fmaf(w, xj[k], agg[k]);
    }
    const float sc_mean = 1.f / heads;
#pragma unroll 1      // keep W in shared memory: unrolled, the 2 * FO weight rows get hoisted out of the row loop (220 registers)
    for (int f4 = 0; f4 < FO / 4; ++f4) {
      float o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 w0 = *reinterpret_cast<const float4*>(Wh + (f4 * 4 + q) * GAT_FIN);
        const float4 w1 = *reinterpret_cast<const float4*>(Wh + (f4 * 4 + q) * GAT_FIN + 4);
        o[q] = agg[0] * w0.x + agg[1] * w0.y + agg[2] * w0.z + agg[3] * w0.w + agg[4] * w1.x + agg[5] * w1.y + agg[6] * w1.z + agg[7] * w1.w;
      }
      if (concat) {
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias) b = *reinterpret_cast<const float4*>(bias + h * FO + f4 * 4);
        if (on) *reinterpret_cast<float4*>(out + (size_t)i * HF + h * FO + f4 * 4) = make_float4(o[0] + b.x, o[1] + b.y, o[2] + b.z, o[3] + b.w);
      } else {
        for (int w = 1; w < heads; w <<= 1) {       // mean over the heads of the row (adjacent lanes)
#pragma unroll
          for (int q = 0; q < 4; ++q) o[q] += __shfl_xor_sync(0xffffffffu, o[q], w);
        }
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias) b = *reinterpret_cast<const float4*>(bias + f4 * 4);
        if (on && h == 0) *reinterpret_cast<float4*>(out + (size_t)i * FO + f4 * 4) =
            make_float4(o[0] * sc_mean + b.x, o[1] * sc_mean + b.y, o[2] * sc_mean + b.z, o[3] * sc_mean + b.w);
      }
    }
  }
}

template <int FO>
__global__ void __launch_bounds__(256) gat_bwd_dst_rowhead_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, int Z,
                                                                  const float* __restrict__ x, int F_in, const float* __restrict__ W,
                                                                  const float* __restrict__ a_src, const float* __restrict__ a_dst,
                                                                  const float* __restrict__ alpha, const float* __restrict__ gout,
                                                                  int heads, int concat, float slope, float* __restrict__ de,
                                                                  float* __restrict__ d_adst) {
  constexpr int WS = FO * GAT_FIN + 4;
  __shared__ __align__(16) float Wsm[128 * GAT_FIN + 32 * 4];
  const int HF = heads * FO;
  for (int idx = threadIdx.x; idx < HF * GAT_FIN; idx += blockDim.x) {
    const int c = idx / GAT_FIN, k = idx % GAT_FIN;
    Wsm[(c / FO) * WS + (c % FO) * GAT_FIN + k] = k < F_in ? W[(size_t)c * F_in + k] : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, rpw = 32 / heads, g = lane / heads, h = lane % heads;
  const float* Wh = Wsm + h * WS;
  const float sc = concat ? 1.f : 1.f / heads;
  const int nwarp = (gridDim.x * blockDim.x) >> 5;
  for (int rg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; rg * rpw < Z; rg += nwarp) {
    const int i = rg * rpw + g;
    if (i >= Z) continue;
    float qv[GAT_FIN];                    // W_h^T g_ih
#pragma unroll
    for (int k = 0; k < GAT_FIN; ++k) qv[k] = 0.f;
#pragma unroll
    for (int f4 = 0; f4 < FO / 4; ++f4) {
      const float4 gv = *reinterpret_cast<const float4*>(gout + (concat ? (size_t)i * HF + h * FO : (size_t)i * FO) + f4 * 4);
      const float gq[4] = {gv.x * sc, gv.y * sc, gv.z * sc, gv.w * sc};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 w0 = *reinterpret_cast<const float4*>(Wh + (f4 * 4 + q) * GAT_FIN);
        const float4 w1 = *reinterpret_cast<const float4*>(Wh + (f4 * 4 + q) * GAT_FIN + 4);
        qv[0] = fmaf(gq[q], w0.x, qv[0]); qv[1] = fmaf(gq[q], w0.y, qv[1]); qv[2] = fmaf(gq[q], w0.z, qv[2]); qv[3] = fmaf(gq[q], w0.w, qv[3]);
        qv[4] = fmaf(gq[q], w1.x, qv[4]); qv[5] = fmaf(gq[q], w1.y, qv[5]); qv[6] = fmaf(gq[q], w1.z, qv[6]); qv[7] = fmaf(gq[q], w1.w, qv[7]);
      }
    }
    const int e0 = rowptr[i], deg = rowptr[i + 1] - e0;
    const float ad = a_dst[(size_t)i * heads + h];
    float ci = 0.f;
    for (int e = 0; e < deg; ++e) {
      const int j = col[e0 + e];
      float d = 0.f;
#pragma unroll
      for (int k = 0; k < GAT_FIN; ++k)
        if (k < F_in) d = fmaf(qv[k], x[(size_t)j * F_in + k], d);
      ci = fmaf(alpha[(size_t)(e0 + e) * heads + h], d, ci);
    }
    float dad = 0.f;
    for (int e = 0; e < deg; ++e) {
      const int j = col[e0 + e];
      float d = 0.f;
#pragma unroll
      for (int k = 0; k < GAT_FIN; ++k)
        if (k < F_in) d = fmaf(qv[k], x[(size_t)j * F_in + k], d);
      const size_t eh = (size_t)(e0 + e) * heads + h;
      const float pre = a_src[(size_t)j * heads + h] + ad;
      const float dev = alpha[eh] * (d - ci) * (pre > 0.f ? 1.f : slope);
      dad += dev;
      de[eh] = dev;
    }
    d_adst[(size_t)i * heads + h] = dad;
  }
}

static bool gat_rowhead_ok(int heads, int F_out, int F_in) {
  return F_in <= GAT_FIN && heads <= 32 && (heads & (heads - 1)) == 0 && heads * F_out <= 128 &&
         (F_out == 4 || F_out == 8 || F_out == 16 || F_out == 32);
}
static int gat_rowhead_grid(int Z, int heads) {
  const int rows_per_block = 8 * (32 / heads);
  const int need = (Z + rows_per_block - 1) / rows_per_block;
  const int cap = 148 * 8;
  return need < cap ? (need < 1 ? 1 : need) : cap;
}
#define GAT_DISPATCH_FO(FO_, CALL)                  \
  switch (FO_) {                                    \
    case 4: { constexpr int F_ = 4; CALL; } break;   \
    case 8: { constexpr int F_ = 8; CALL; } break;   \
    case 16: { constexpr int F_ = 16; CALL; } break; \
    default: { constexpr int F_ = 32; CALL; } break; \
  }

// dx[z,:] = W^T dxw[z,:] on sub-warps: the row of dxw is read once, coalesced; F_in partial sums are reduced over the
// LPR lanes of the row
template <int LPR>
__global__ void __launch_bounds__(256) gat_bwd_input_sub_kernel(const float* __restrict__ dxw, const float* __restrict__ W, int Z,
                                                                int F_in, float* __restrict__ dx) {
  constexpr int RPW = 32 / LPR, HF = 4 * LPR;
  const int lane = threadIdx.x & 31, g = lane / LPR, sl = lane % LPR, c = 4 * sl;
  float wr[4][GAT_FIN];
#pragma unroll
  for (int k = 0; k < GAT_FIN; ++k)
#pragma unroll
    for (int q = 0; q < 4; ++q) wr[q][k] = k < F_in ? W[(size_t)(c + q) * F_in + k] : 0.f;
  const int nwarp = (gridDim.x * blockDim.x) >> 5;
  for (int rg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; rg * RPW < Z; rg += nwarp) {
    const int z = rg * RPW + g;
    const bool on = z < Z;
    const float4 d = on ? *reinterpret_cast<const float4*>(dxw + (size_t)z * HF + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    float mine = 0.f;
#pragma unroll
    for (int k = 0; k < GAT_FIN; ++k) {
      float v = d.x * wr[0][k] + d.y * wr[1][k] + d.z * wr[2][k] + d.w * wr[3][k];
      for (int o = LPR >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (sl == k) mine = v;
    }
    if (on && sl < F_in) dx[(size_t)z * F_in + sl] = mine;     // LPR >= 2; F_in <= 8 needs LPR >= 8 or the loop below
    if (LPR < GAT_FIN) {
      // fewer lanes than input features (8 or 16 channels): lane 0 of the row recomputes the remaining ones
#pragma unroll
      for (int k = LPR; k < GAT_FIN; ++k) {
        float v = d.x * wr[0][k] + d.y * wr[1][k] + d.z * wr[2][k] + d.w * wr[3][k];
        for (int o = LPR >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (on && sl == 0 && k < F_in) dx[(size_t)z * F_in + k] = v;
      }
    }
  }
}

constexpr int GAT_PARAM_BLOCKS = 148 * 8;
static bool gat_sub_ok(int heads, int F_out, int F_in) {
  const int HF = heads * F_out;
  return F_in <= GAT_FIN && heads <= 32 && (HF == 8 || HF == 16 || HF == 32 || HF == 64 || HF == 128);
}
static int gat_sub_grid(int Z, int lpr) {
  const int rows_per_block = 8 * (32 / lpr);
  const int need = (Z + rows_per_block - 1) / rows_per_block;
  const int cap = 148 * 8;
  return need < cap ? (need < 1 ? 1 : need) : cap;
}
#define GAT_DISPATCH_LPR(HF_, CALL)              \
  switch ((HF_) / 4) {                           \
    case 2: { constexpr int L_ = 2; CALL; } break;   \
    case 4: { constexpr int L_ = 4; CALL; } break;   \
    case 8: { constexpr int L_ = 8; CALL; } break;   \
    case 16: { constexpr int L_ = 16; CALL; } break; \
    default: { constexpr int L_ = 32; CALL; } break; \
  }

static bool gat_shape_ok(int heads, int F_out, int F_in) {
  return heads > 0 && F_out >= 4 && F_out % 4 == 0 && (F_out & (F_out - 1)) == 0 && F_in > 0 && F_in <= 32;
}

int gat_forward(const int* rowptr, const int* col, int Z, int nnz, const float* x, int F_in, const float* W, const float* att_src,
                const float* att_dst, const float* bias, int heads, int F_out, int concat, float slope, float* out, float* xw,
                float* a_src, float* a_dst, float* alpha, cudaStream_t st) {
  (void)nnz;
  if (!gat_shape_ok(heads, F_out, F_in)) return AB200_ERR_UNSUPPORTED;
  if (gat_rowhead_ok(heads, F_out, F_in)) {
    GAT_DISPATCH_FO(F_out, (gat_fwd_rowhead_kernel<F_><<<gat_rowhead_grid(Z, heads), 256, 0, st>>>(
                               rowptr, col, Z, x, F_in, W, att_src, att_dst, bias, heads, concat, slope, out, xw, a_src, a_dst, alpha)));
    return check_launch();
  }
  if (gat_sub_ok(heads, F_out, F_in)) {
    const int HF = heads * F_out;
    GAT_DISPATCH_LPR(HF, (gat_fwd_fused_kernel<L_><<<gat_sub_grid(Z, L_), 256, 0, st>>>(
                             rowptr, col, Z, x, F_in, W, att_src, att_dst, bias, heads, F_out, concat, slope, out, xw, a_src, a_dst, alpha)));
    return check_launch();
  }
  const int wpb = 8, blocks = (Z + wpb - 1) / wpb;
  if (F_out >= 32) {
    cudaMemsetAsync(a_src, 0, sizeof(float) * (size_t)Z * heads, st);
    cudaMemsetAsync(a_dst, 0, sizeof(float) * (size_t)Z * heads, st);
  }
  gat_project_kernel<<<blocks, 256, 0, st>>>(x, W, att_src, att_dst, Z, F_in, heads, F_out, xw, a_src, a_dst);
  int rc = check_launch();
  if (rc) return rc;
  if (!concat) {
    gat_mean_init_kernel<<<148, 256, 0, st>>>(out, bias, Z, F_out);
    if ((rc = check_launch())) return rc;
  }
  gat_aggregate_kernel<<<blocks, 256, 0, st>>>(rowptr, col, Z, xw, a_src, a_dst, bias, heads, F_out, concat, slope, out, alpha);
  return check_launch();
}

size_t gat_backward_workspace(int Z, int nnz, int heads, int F_out) {
  // de [nnz,H], d_adst [Z,H], d_asrc [Z,H], dxw [Z,HF]
  // + per-block partial rows of the tiled parameter reduction
  return align_up(sizeof(float) * (size_t)nnz * heads, 256) + 2 * align_up(sizeof(float) * (size_t)Z * heads, 256) +
         align_up(sizeof(float) * (size_t)Z * heads * F_out, 256) +
         align_up(sizeof(float) * (size_t)GAT_PARAM_BLOCKS * (GAT_FIN + 3) * heads * F_out, 256);
}

int gat_backward(const int* rowptr, const int* col, const int* rowptr_t, const int* col_t, const int* eid_t, int Z, int nnz,
                 const float* x, int F_in, const float* W, const float* att_src, const float* att_dst, int heads, int F_out,
                 int concat, float slope, const float* xw, const float* a_src, const float* a_dst, const float* alpha,
                 const float* gout, float* grad_x, float* grad_W, float* grad_att_src, float* grad_att_dst, float* grad_bias,
                 void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!gat_shape_ok(heads, F_out, F_in)) return AB200_ERR_UNSUPPORTED;
  if (ws_bytes < gat_backward_workspace(Z, nnz, heads, F_out)) return AB200_ERR_WORKSPACE;
  char* p = (char*)ws;
  float* de = (float*)p; p += align_up(sizeof(float) * (size_t)nnz * heads, 256);
  float* d_adst = (float*)p; p += align_up(sizeof(float) * (size_t)Z * heads, 256);
  float* d_asrc = (float*)p; p += align_up(sizeof(float) * (size_t)Z * heads, 256);
  float* dxw = (float*)p; p += align_up(sizeof(float) * (size_t)Z * heads * F_out, 256);
  float* partial = (float*)p;
  const int HF = heads * F_out;
  int rc = 0;
  if (gat_sub_ok(heads, F_out, F_in)) {
    if (gat_rowhead_ok(heads, F_out, F_in)) {
      GAT_DISPATCH_FO(F_out, (gat_bwd_dst_rowhead_kernel<F_><<<gat_rowhead_grid(Z, heads), 256, 0, st>>>(
                                 rowptr, col, Z, x, F_in, W, a_src, a_dst, alpha, gout, heads, concat, slope, de, d_adst)));
    } else {
      GAT_DISPATCH_LPR(HF, (gat_bwd_dst_sub_kernel<L_><<<gat_sub_grid(Z, L_), 256, 0, st>>>(
                               rowptr, col, Z, x, F_in, W, a_src, a_dst, alpha, gout, heads, F_out, concat, slope, de, d_adst)));
    }
    if ((rc = check_launch())) return rc;
    GAT_DISPATCH_LPR(HF, (gat_bwd_src_sub_kernel<L_><<<gat_sub_grid(Z, L_), 256, 0, st>>>(
                             rowptr_t, col_t, eid_t, Z, alpha, de, gout, d_adst, att_src, att_dst, heads, F_out, concat, dxw, d_asrc)));
    if ((rc = check_launch())) return rc;
    const int zl = 256 / HF;
    int nblk = (Z + zl - 1) / zl;
    if (nblk > GAT_PARAM_BLOCKS) nblk = GAT_PARAM_BLOCKS;
    GAT_DISPATCH_LPR(HF, (gat_bwd_param_tiled_kernel<4 * L_><<<nblk, 256, 0, st>>>(x, xw, dxw, d_asrc, d_adst, gout, Z, F_in, heads, F_out,
                                                                                    concat, partial)));
    if ((rc = check_launch())) return rc;
    const int nfin = (GAT_FIN + 3) * HF;
    gat_param_finalize_kernel<<<(nfin + 255) / 256, 256, 0, st>>>(partial, nblk, HF, F_in, F_out, concat, grad_W, grad_att_src,
                                                                 grad_att_dst, grad_bias);
    if ((rc = check_launch())) return rc;
    if (grad_x) {
      GAT_DISPATCH_LPR(HF, (gat_bwd_input_sub_kernel<L_><<<gat_sub_grid(Z, L_), 256, 0, st>>>(dxw, W, Z, F_in, grad_x)));
      rc = check_launch();
    }
    return rc;
  }
  const int wpb = 8, blocks = (Z + wpb - 1) / wpb;
  gat_bwd_dst_kernel<<<blocks, 256, 0, st>>>(rowptr, col, Z, xw, a_src, a_dst, alpha, gout, heads, F_out, concat, slope, de, d_adst);
  rc = check_launch();
  if (rc) return rc;
  gat_bwd_src_kernel<<<blocks, 256, 0, st>>>(rowptr_t, col_t, eid_t, Z, alpha, de, gout, d_adst, att_src, att_dst, heads, F_out,
                                             concat, dxw, d_asrc);
  if ((rc = check_launch())) return rc;
  const size_t smem = sizeof(float) * (size_t)(F_in + 3) * 256;
  gat_bwd_param_kernel<<<HF, 256, smem, st>>>(x, xw, dxw, d_asrc, d_adst, gout, Z, F_in, heads, F_out, concat, grad_W, grad_att_src,
                                              grad_att_dst, grad_bias);
  if ((rc = check_launch())) return rc;
  if (grad_x) {
    gat_bwd_input_kernel<<<148, 256, 0, st>>>(dxw, W, Z, F_in, HF, grad_x);
    rc = check_launch();
  }
  return rc;
}

}  // namespace ab200
