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

static bool gat_shape_ok(int heads, int F_out, int F_in) {
  return heads > 0 && F_out >= 4 && F_out % 4 == 0 && (F_out & (F_out - 1)) == 0 && F_in > 0 && F_in <= 32;
}

int gat_forward(const int* rowptr, const int* col, int Z, int nnz, const float* x, int F_in, const float* W, const float* att_src,
                const float* att_dst, const float* bias, int heads, int F_out, int concat, float slope, float* out, float* xw,
                float* a_src, float* a_dst, float* alpha, cudaStream_t st) {
  (void)nnz;
  if (!gat_shape_ok(heads, F_out, F_in)) return AB200_ERR_UNSUPPORTED;
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
  return align_up(sizeof(float) * (size_t)nnz * heads, 256) + 2 * align_up(sizeof(float) * (size_t)Z * heads, 256) +
         align_up(sizeof(float) * (size_t)Z * heads * F_out, 256);
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
  float* dxw = (float*)p;
  const int HF = heads * F_out;
  const int wpb = 8, blocks = (Z + wpb - 1) / wpb;
  gat_bwd_dst_kernel<<<blocks, 256, 0, st>>>(rowptr, col, Z, xw, a_src, a_dst, alpha, gout, heads, F_out, concat, slope, de, d_adst);
  int rc = check_launch();
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
