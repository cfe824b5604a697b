// Self-test of the tcgen05 plumbing in umma.cuh: D[128 x N] = A[128 x K] * B[N x K]^T (bf16 in, fp32 out) on one
// CTA, with the operand placement / shared-memory layout selectable at run time.  The GPU tests run every variant
// against a CPU product, so a wrong descriptor encoding shows up here and not inside the fused solver kernel.
//   a_mode: 0 = A in TMEM (tcgen05.st, TS form)   1 = A in smem, no swizzle   2 = A in smem, 128B swizzle
//   b_mode:                                        1 = B in smem, no swizzle   2 = B in smem, 128B swizzle
//   a_mode / b_mode 3 = operand in smem MN-major, no swizzle (LBO = K-direction core stride, SBO = MN-direction)
//   a_mode / b_mode 4 = same bytes, LBO / SBO fields exchanged in the descriptor (to pin the field semantics)
#include "common.cuh"
#include "umma.cuh"

namespace ab200 {
using namespace umma;

__global__ void __launch_bounds__(128, 1) umma_probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                            float* __restrict__ D, int N, int K, int a_mode, int b_mode,
                                                            int* __restrict__ status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [A operand: 128*K*2] [B operand: N*K*2], both 1024-aligned
  uint8_t* sbase = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = sbase;
  uint8_t* sB = sbase + ((128 * K * 2 + 1023) / 1024) * 1024;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  if (warp == 0) tmem_alloc<256>(&tmem_base_s);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tD = tmem;            // columns [0, N)
  const uint32_t tA = tmem + 128;      // columns [128, 128 + K/2)
  const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;

  // ---- operand A: thread = row
  const int row = tid;
  const uint32_t a_lbo = 128u * 16u, a_sbo = 128u;     // no-swizzle: 128 rows x 16 B per k-group
  if (a_mode == 0) {
    for (int k0 = 0; k0 < K; k0 += 32) {
      uint32_t r[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = pack_bf16(A[row * K + k0 + 2 * j], A[row * K + k0 + 2 * j + 1]);
      tmem_st16(tA + lane_sel + (uint32_t)(k0 / 2), r);
    }
    tmem_st_wait();
  } else if (a_mode <= 2) {
    for (int k = 0; k < K; ++k) {
      const uint32_t off = (a_mode == 1) ? off_kmajor_noswz(row, k, a_lbo, a_sbo) : off_kmajor_sw128(row, k, 128u * 128u);
      *reinterpret_cast<__nv_bfloat16*>(sA + off) = __float2bfloat16_rn(A[row * K + k]);
    }
  } else {
    for (int k = 0; k < K; ++k)
      *reinterpret_cast<__nv_bfloat16*>(sA + off_mnmajor_noswz(row, k, 128u, (uint32_t)(K / 8) * 128u)) =
          __float2bfloat16_rn(A[row * K + k]);
  }
  // ---- operand B: all threads
  const uint32_t b_lbo = (uint32_t)N * 16u, b_sbo = 128u;
  for (int i = tid; i < N * K; i += blockDim.x) {
    const int n = i / K, k = i % K;
    const uint32_t off = (b_mode == 1) ? off_kmajor_noswz(n, k, b_lbo, b_sbo)
                         : (b_mode == 2) ? off_kmajor_sw128(n, k, (uint32_t)N * 128u)
                                         : off_mnmajor_noswz(n, k, 128u, (uint32_t)(K / 8) * 128u);
    *reinterpret_cast<__nv_bfloat16*>(sB + off) = __float2bfloat16_rn(B[n * K + k]);
  }
  fence_async_smem();      // generic-proxy writes -> visible to the tensor core's async-proxy reads
  tc_fence_before();
  __syncthreads();

  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, N, false, a_mode >= 3, b_mode >= 3);
    const uint32_t mn_lbo = 128u, mn_sbo = (uint32_t)(K / 8) * 128u;
    for (int ks = 0; ks < K / 16; ++ks) {
      uint64_t bdesc;
      if (b_mode == 1) bdesc = make_smem_desc(smem_u32(sB) + (uint32_t)ks * 2u * b_lbo, b_lbo, b_sbo, SWZ_NONE);
      else if (b_mode == 3) bdesc = make_smem_desc(smem_u32(sB) + (uint32_t)ks * 2u * mn_lbo, mn_lbo, mn_sbo, SWZ_NONE);
      else if (b_mode == 4) bdesc = make_smem_desc(smem_u32(sB) + (uint32_t)ks * 2u * mn_lbo, mn_sbo, mn_lbo, SWZ_NONE);
      else bdesc = make_smem_desc(smem_u32(sB) + (uint32_t)(ks >> 2) * (uint32_t)N * 128u + (uint32_t)(ks & 3) * 32u, 16u, 1024u, SWZ_128B);
      if (a_mode == 0) {
        mma_ts(tD, tA + (uint32_t)ks * 8u, bdesc, idesc, ks > 0 ? 1u : 0u);
      } else {
        uint64_t adesc;
        if (a_mode == 1) adesc = make_smem_desc(smem_u32(sA) + (uint32_t)ks * 2u * a_lbo, a_lbo, a_sbo, SWZ_NONE);
        else if (a_mode == 3) adesc = make_smem_desc(smem_u32(sA) + (uint32_t)ks * 2u * mn_lbo, mn_lbo, mn_sbo, SWZ_NONE);
        else if (a_mode == 4) adesc = make_smem_desc(smem_u32(sA) + (uint32_t)ks * 2u * mn_lbo, mn_sbo, mn_lbo, SWZ_NONE);
        else adesc = make_smem_desc(smem_u32(sA) + (uint32_t)(ks >> 2) * 128u * 128u + (uint32_t)(ks & 3) * 32u, 16u, 1024u, SWZ_128B);
        mma_ss(tD, adesc, bdesc, idesc, ks > 0 ? 1u : 0u);
      }
    }
    mma_commit(&bar);
  }
  const bool ok = mbar_wait(&bar, 0, 2000000000LL);
  tc_fence_after();
  if (!ok) {
    if (tid == 0) *status = 1;    // timed out
  } else {
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tD + lane_sel + (uint32_t)c0, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) D[row * N + c0 + j] = __uint_as_float(r[j]);
    }
    if (tid == 0) *status = 0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

int umma_probe(const float* A, const float* B, float* D, int N, int K, int a_mode, int b_mode, int* status, cudaStream_t st) {
  if ((N != 64 && N != 128) || K % 32 != 0 || K < 32 || K > 192) return AB200_ERR_BAD_ARG;
  if (a_mode < 0 || a_mode > 4 || b_mode < 1 || b_mode > 4) return AB200_ERR_BAD_ARG;
  if ((a_mode == 2 || b_mode == 2) && K % 64 != 0) return AB200_ERR_BAD_ARG;
  const size_t smem = 1024 + ((128 * K * 2 + 1023) / 1024) * 1024 + ((size_t)N * K * 2 + 1023) / 1024 * 1024;
  cudaError_t e = cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_cuda_error(e); return AB200_ERR_CUDA; }
  umma_probe_kernel<<<1, 128, smem, st>>>(A, B, D, N, K, a_mode, b_mode, status);
  return check_launch();
}

}  // namespace ab200
