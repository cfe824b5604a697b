// Shared device/host helpers for the ananke_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ananke_b200.h"

namespace ab200 {

// ---- packed drift-net layout ---------------------------------------------------------------------
// The host hands weights over exactly as torch stores them (`w_flat`, see ananke_b200.h).  Kernels want
// every Linear transposed to [K][N] (k-major rows) so that a k-chunk is one contiguous block and a
// thread's output columns are a contiguous 16/32-byte read.  The pack kernel writes this layout into
// the caller-provided workspace:
//   W0   [2P][hid]   rows of w_in belonging to (p, v)
//   WH   [H][hid]    rows of w_in belonging to h          (folded once per trajectory into a per-agent bias)
//   wsin [hid], wcos [hid]                                 (time-feature columns of w_in)
//   b_in [hid]
//   per residual block: WA [hid][hid], bA [hid], WB [hid][hid], bB [hid]
//   WO   [hid][P], bO [P]
struct PackLayout {
  int P, H, HID, NRES;
  __host__ __device__ int in_dim() const { return 2 * P + H + 2; }
  __host__ __device__ int64_t off_W0() const { return 0; }
  __host__ __device__ int64_t off_WH() const { return off_W0() + (int64_t)2 * P * HID; }
  __host__ __device__ int64_t off_wsin() const { return off_WH() + (int64_t)H * HID; }
  __host__ __device__ int64_t off_wcos() const { return off_wsin() + HID; }
  __host__ __device__ int64_t off_bin() const { return off_wcos() + HID; }
  __host__ __device__ int64_t off_res(int r) const { return off_bin() + HID + (int64_t)r * (2 * (int64_t)HID * HID + 2 * HID); }
  __host__ __device__ int64_t off_WA(int r) const { return off_res(r); }
  __host__ __device__ int64_t off_bA(int r) const { return off_res(r) + (int64_t)HID * HID; }
  __host__ __device__ int64_t off_WB(int r) const { return off_bA(r) + HID; }
  __host__ __device__ int64_t off_bB(int r) const { return off_WB(r) + (int64_t)HID * HID; }
  __host__ __device__ int64_t off_WO() const { return off_res(NRES); }
  __host__ __device__ int64_t off_bO() const { return off_WO() + (int64_t)HID * P; }
  __host__ __device__ int64_t total() const { return off_bO() + P; }
};

// Offsets inside `w_flat` (torch order, [out][in] row-major).
struct FlatLayout {
  int P, H, HID, NRES;
  __host__ __device__ int in_dim() const { return 2 * P + H + 2; }
  __host__ __device__ int64_t off_win() const { return 0; }
  __host__ __device__ int64_t off_bin() const { return (int64_t)HID * in_dim(); }
  __host__ __device__ int64_t off_res(int r) const { return off_bin() + HID + (int64_t)r * (2 * (int64_t)HID * HID + 2 * HID); }
  __host__ __device__ int64_t off_wa(int r) const { return off_res(r); }
  __host__ __device__ int64_t off_ba(int r) const { return off_res(r) + (int64_t)HID * HID; }
  __host__ __device__ int64_t off_wb(int r) const { return off_ba(r) + HID; }
  __host__ __device__ int64_t off_bb(int r) const { return off_wb(r) + (int64_t)HID * HID; }
  __host__ __device__ int64_t off_wout() const { return off_res(NRES); }
  __host__ __device__ int64_t off_bout() const { return off_wout() + (int64_t)P * HID; }
  __host__ __device__ int64_t total() const { return off_bout() + P; }
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// thread-local last CUDA error text (ab200_last_cuda_error)
void set_cuda_error(cudaError_t e);
inline int check_launch() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_cuda_error(e); return AB200_ERR_CUDA; }
  return AB200_OK;
}

// ---- small device helpers --------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Un-contracted fp32 ops: the Runge-Kutta stage algebra follows torchdiffeq's association order
// (rk_common.py rk4_alt_step_func) with separate roundings, as the eager PyTorch reference performs them.
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }

// angle of the time features: sin/cos(t * 2 * pi / period), fp32 throughout as the reference computes it
// (mode_sep/architecture/model.py:62-63; latent_ode/architecture/model.py:86)
__device__ __forceinline__ void time_features(float t, float period, float& s, float& c) {
  float ang = __fdiv_rn(fmul(fmul(t, 2.0f), 3.14159274101257324f), period);
  s = sinf(ang);
  c = cosf(ang);
}

}  // namespace ab200
