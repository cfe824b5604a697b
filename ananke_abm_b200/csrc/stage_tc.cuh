// Shared machinery of the tensor-core STAGE kernels (stage_fwd_tc.cu, stage_bwd_tc.cu, wgrad_tc.cu).
//
// A "stage" is one evaluation of the drift net at a Runge-Kutta stage input that is a linear combination of the
// step's base state y0 = [p0, v0, h] and the accelerations a_1..a_s of earlier stages (the drift is second order:
// k_j = (v_in_j, a_j), so every stage input, the step solution, the dense-output rows and the embedded error are
// linear in (p0, v0, a_1..a_s)).  The stage kernels keep NO Runge-Kutta state in registers, which is what lets one
// CTA run TWO 128-agent tiles at once ("slots", 8 warps each): while one slot is in its epilogue (tcgen05.ld ->
// ReLU -> bf16 -> tcgen05.st) the tensor core runs the other slot's layer, sharing one resident copy of the weights.
//
//   shared memory : the six weight matrices as bf16 in the canonical un-swizzled K-major UMMA layout, each with a
//                   16-wide K extension that carries the bias (hi + lo bf16 split) and, for the first layer, the
//                   sin/cos time-feature columns -- bias and time features are applied by the tensor core, the
//                   epilogue is a bare ReLU + pack.  The same image read MN-major is W^T (used by the dgrad GEMMs).
//   tensor memory : per slot 256 columns: ACC fp32 [0,160) | ACT bf16 pairs [160,224) | HB (context h) [224,240)
//                   | TB ("time/bias block": sin, cos hi/lo splits and two ones) [240,248)
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace ab200 {
namespace stc {
using namespace umma;

constexpr int P = 64, H = 32, HID = 128, NRES = 2, D = 2 * P + H;
constexpr int SLOT_THREADS = 256, NSLOT = 2, THREADS = SLOT_THREADS * NSLOT;
constexpr int TM = 128;                        // agents per tile == UMMA M
constexpr int KX = 16;                         // K extension (one UMMA K step)
constexpr int K1 = 2 * P + H + KX;             // 176
constexpr int KH = HID + KX;                   // 144
constexpr long long STAGE_WAIT_CYCLES = 100000000LL;   // ~50 ms: a layer takes microseconds
constexpr int MAX_A = 7;                       // accelerations a stage may combine (dopri5: 6 + FSAL)

// ---- weight image (bytes) ------------------------------------------------------------------------------
constexpr uint32_t SBO = 128u;                                       // MN-adjacent core matrices
__host__ __device__ constexpr uint32_t lbo(int N) { return (uint32_t)N * 16u; }   // K-adjacent core matrices
constexpr uint32_t OFF_W1 = 0, SZ_W1 = (uint32_t)HID * K1 * 2;       // [128][176]
constexpr uint32_t OFF_RES = OFF_W1 + SZ_W1, SZ_HH = (uint32_t)HID * KH * 2;   // 4 x [128][144]: A0, B0, A1, B1
constexpr uint32_t OFF_WO = OFF_RES + 2u * NRES * SZ_HH, SZ_WO = (uint32_t)P * KH * 2;   // [64][144]
constexpr uint32_t W_BYTES = OFF_WO + SZ_WO;                         // 210,944
__host__ __device__ constexpr uint32_t off_hh(int m) { return OFF_RES + (uint32_t)m * SZ_HH; }
// position of the extension columns inside a K extension block (A side holds the matching constants)
//   0: s_hi*ws_hi  1: s_hi*ws_lo  2: s_lo*ws_hi  3: c_hi*wc_hi  4: c_hi*wc_lo  5: c_lo*wc_hi  6: 1*b_hi  7: 1*b_lo

// ---- tensor-memory columns (per slot) ------------------------------------------------------------------
constexpr uint32_t SLOT_COLS = 256, C_ACC = 0, C_ACT = 160, C_HB = 224, C_TB = 240;

// Optional cycle trace (compile with -DAB200_STAGE_TRACE): thread 0 of every slot of CTA 0 appends (tag, clock) pairs.
#ifdef AB200_STAGE_TRACE
static __device__ long long g_stage_trace[8192];     // one copy per translation unit (no -rdc)
static __device__ int g_stage_trace_n[2];
#define STAGE_TRACE(c, tag)                                                                  \
  do {                                                                                        \
    if (blockIdx.x == 0 && (c).stid == 0) {                                                   \
      const int i_ = g_stage_trace_n[(c).slot]++;                                             \
      if (i_ < 2048) { g_stage_trace[((c).slot * 2048 + i_) * 2] = (tag); g_stage_trace[((c).slot * 2048 + i_) * 2 + 1] = clock64(); } \
    }                                                                                         \
  } while (0)
#else
#define STAGE_TRACE(c, tag) do { } while (0)
#endif

__device__ __forceinline__ void slot_sync(int slot) {
  asm volatile("bar.sync %0, %1;" ::"r"(slot + 1), "n"(SLOT_THREADS) : "memory");
}
// d = {hi: bf16(a), lo: bf16(b)} with ReLU applied
__device__ __forceinline__ uint32_t pack_relu_bf16(float lo_v, float hi_v) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_v), "f"(lo_v));
  return d;
}
__device__ __forceinline__ float bf16lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// Operand-format generic versions (HALF = IEEE fp16 operands: 11-bit significand, 8x less rounding noise than bf16 --
// what an adaptive solver's error estimate needs; values of this net stay far inside the fp16 range).
template <bool HALF>
__device__ __forceinline__ uint32_t pack2(float lo_v, float hi_v) {
  uint32_t d;
  if (HALF) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_v), "f"(lo_v));   // saturate, never inf
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_v), "f"(lo_v));
  return d;
}
template <bool HALF>
__device__ __forceinline__ uint32_t pack2_relu(float lo_v, float hi_v) {
  uint32_t d;
  if (HALF) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_v), "f"(lo_v));
  else asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_v), "f"(lo_v));
  return d;
}
template <bool HALF>
__device__ __forceinline__ float un_lo(uint32_t v) {
  if (!HALF) return bf16lo(v);
  float f;
  asm("{.reg .b16 l, h; mov.b32 {l, h}, %1; cvt.f32.f16 %0, l;}" : "=f"(f) : "r"(v));
  return f;
}
template <bool HALF>
__device__ __forceinline__ float un_hi(uint32_t v) {
  if (!HALF) return bf16hi(v);
  float f;
  asm("{.reg .b16 l, h; mov.b32 {l, h}, %1; cvt.f32.f16 %0, h;}" : "=f"(f) : "r"(v));
  return f;
}
template <bool HALF>
__device__ __forceinline__ float round16(float x) { return un_lo<HALF>(pack2<HALF>(x, 0.0f)); }

// Per-thread view of its slot.
struct SlotCtx {
  uint32_t tmem;        // TMEM base of the slot (column 0, lane 0)
  uint32_t lane_sel;    // (32 * (warp % 4)) << 16 : the TMEM lanes this warp may touch
  uint32_t sbase;       // shared-memory address of the weight image
  uint64_t* bar;        // MMA-completion mbarrier of the slot
  int* lock;            // CTA-wide "MMA issue" mutex (shared memory)
  int flags;            // tuning switches (stage_flags()): 1 = issue mutex, 2 = L2 prefetch of the next tile
  int* status;
  uint32_t phase;
  int slot, stid, row, hf;   // thread index inside the slot, agent row (TMEM lane), column half
  bool alive;
};

// One layer:  ACC[128 x N] = A * W^T  (K-major image)  or  A * W  (MN-major view of the same image, `TRANS`).
//   A = NKS K-steps of bf16 pairs starting at TMEM column a_col, optionally followed by the TB block (EXT).
// Called by all threads of the slot; returns when the accumulator is complete.  Everything about the shape is a
// template parameter so that the single issuing thread runs a fully unrolled stream of tcgen05.mma whose descriptors
// differ by a compile-time constant (the tensor core needs a new MMA every 64 cycles; a descriptor rebuilt with
// shifts and masks per K-step does not keep up).
// The MMA stream of one layer, issued by ONE thread.  Deliberately not inlined and rolled: the stage kernels are
// 80-130 KB of SASS (every layer's epilogue is specialised), far beyond the instruction cache, and this code runs on a
// single thread per slot -- one shared copy costs a call, eleven inlined copies cost instruction-cache misses for all
// sixteen warps.  ~10 instructions per MMA keep ahead of the 64-cycle MMA.
static __device__ __noinline__ void issue_mmas(uint32_t acc, uint32_t a0, uint32_t tb, uint64_t d0, uint32_t step16, uint32_t idesc, int nks,
                                        uint64_t* bar) {
#pragma unroll 1
  for (int ks = 0; ks < nks; ++ks) mma_ts(acc, a0 + (uint32_t)ks * 8u, d0 + (uint64_t)((uint32_t)ks * step16), idesc, ks > 0 ? 1u : 0u);
  if (tb != 0xffffffffu) mma_ts(acc, tb, d0 + (uint64_t)((uint32_t)nks * step16), idesc, 1u);
  mma_commit(bar);
}
static __device__ __noinline__ bool wait_mma(uint64_t* bar, uint32_t phase) { return mbar_wait(bar, phase, STAGE_WAIT_CYCLES); }

// SHARED_ISSUE: use the out-of-line issue routine (backward kernel) or an unrolled inline stream (forward kernel, which
// is smaller and more sensitive to issue latency; measured 95 vs 103 us).
// `after_issue` runs on every thread of the slot between the issue of the layer's MMAs and the wait for their completion: work
// put there (reading the A operand back from tensor memory is allowed -- the tensor core only reads it) stays off the slot's
// critical path while it fits into the MMA time.
struct NoHook { __device__ __forceinline__ void operator()() const {} };
template <bool TRANS, int NKS, bool EXT, int N_IMG, int N, bool SHARED_ISSUE = false, bool HALF = false, class Hook = NoHook>
__device__ __forceinline__ void run_layer(SlotCtx& c, uint32_t a_col, uint32_t w_off, Hook after_issue = Hook()) {
  STAGE_TRACE(c, 1);
  tmem_st_wait();
  tc_fence_before();
  STAGE_TRACE(c, 2);
  slot_sync(c.slot);
  STAGE_TRACE(c, 3);
  if (c.stid == 0) {
    tc_fence_after();
    if (c.flags & 1) {      // optional issue mutex (measured: no gain)
      const long long t0 = clock64();
      while (atomicCAS(c.lock, 0, 1) != 0) {
        if (clock64() - t0 > STAGE_WAIT_CYCLES) { *c.status = 7; break; }
      }
    }
    constexpr uint32_t idesc = make_idesc_bf16(TM, N, false, false, TRANS, HALF);
    constexpr uint32_t L = lbo(N_IMG);
    // K-major: K-adjacent cores L apart (LBO), MN-adjacent cores 128 B apart (SBO); a K-step spans two K cores.
    // MN-major view: B'[n' = in][k' = out] = W[out][in]; K-direction cores are the image's 8-row groups (128 B
    // apart), MN-direction cores its 8-column groups (L apart).
    constexpr uint32_t step16 = (TRANS ? 2u * SBO : 2u * L) >> 4;       // descriptor start-address units per K-step
    const uint64_t d0 = TRANS ? make_smem_desc(c.sbase + w_off, SBO, L, SWZ_NONE) : make_smem_desc(c.sbase + w_off, L, SBO, SWZ_NONE);
    if (SHARED_ISSUE) {
      issue_mmas(c.tmem + C_ACC, c.tmem + a_col, EXT ? c.tmem + C_TB : 0xffffffffu, d0, step16, idesc, NKS, c.bar);
    } else {
      const uint32_t acc = c.tmem + C_ACC, a0 = c.tmem + a_col;
  #pragma unroll
      for (int ks = 0; ks < NKS; ++ks) {
        mma_ts(acc, a0 + (uint32_t)ks * 8u, d0 + (uint64_t)(ks * step16), idesc, ks > 0 ? 1u : 0u);
      }
      if (EXT) mma_ts(acc, c.tmem + C_TB, d0 + (uint64_t)(NKS * step16), idesc, 1u);
        mma_commit(c.bar);
      }
    if (c.flags & 1) atomicExch(c.lock, 0);
  }
  STAGE_TRACE(c, 4);
  __syncwarp();
  after_issue();
  if (c.alive && !(SHARED_ISSUE ? wait_mma(c.bar, c.phase) : mbar_wait(c.bar, c.phase, STAGE_WAIT_CYCLES))) { c.alive = false; *c.status = 1; }
  c.phase ^= 1;
  __syncwarp();
  tc_fence_after();
  STAGE_TRACE(c, 5);
}

// ---- hidden-layer epilogues: this thread's 64 columns (hf*64 ..) of its row ------------------------------
// plain: act = relu(acc) -> ACT.  If KEEP, the packed result is also kept in z[32] (residual stream).
template <bool KEEP, bool HALF = false>
__device__ __forceinline__ void epi_relu(const SlotCtx& c, uint32_t (&z)[32]) {
#pragma unroll
  for (int ch = 0; ch < 2; ++ch) {
    uint32_t r[32];
    tmem_ld32(c.tmem + c.lane_sel + C_ACC + (uint32_t)(c.hf * 64 + ch * 32), r);
    tmem_ld_wait();
    uint32_t o[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      o[j] = pack2_relu<HALF>(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
      if (KEEP) z[ch * 16 + j] = o[j];
    }
    tmem_st16(c.tmem + c.lane_sel + C_ACT + (uint32_t)(c.hf * 32 + ch * 16), o);
  }
}
// residual: z <- relu(acc + z) -> ACT
template <bool HALF = false>
__device__ __forceinline__ void epi_residual(const SlotCtx& c, uint32_t (&z)[32]) {
#pragma unroll
  for (int ch = 0; ch < 2; ++ch) {
    uint32_t r[32];
    tmem_ld32(c.tmem + c.lane_sel + C_ACC + (uint32_t)(c.hf * 64 + ch * 32), r);
    tmem_ld_wait();
    uint32_t o[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const uint32_t zz = z[ch * 16 + j];
      o[j] = pack2_relu<HALF>(__uint_as_float(r[2 * j]) + un_lo<HALF>(zz), __uint_as_float(r[2 * j + 1]) + un_hi<HALF>(zz));
      z[ch * 16 + j] = o[j];
    }
    tmem_st16(c.tmem + c.lane_sel + C_ACT + (uint32_t)(c.hf * 32 + ch * 16), o);
  }
}

// TB block of this row: [s_hi, s_hi, s_lo, c_hi, c_hi, c_lo, 1, 1, 0 x 8] as 8 packed columns
template <bool HALF = false>
__device__ __forceinline__ void write_time_block(const SlotCtx& c, float t, float period) {
  if (c.hf != 0) return;
  float s, co;
  time_features(t, period, s, co);
  const float s_hi = round16<HALF>(s), c_hi = round16<HALF>(co);
  uint32_t o[8];
  o[0] = pack2<HALF>(s_hi, s_hi);
  o[1] = pack2<HALF>(s - s_hi, c_hi);
  o[2] = pack2<HALF>(c_hi, co - c_hi);
  o[3] = pack2<HALF>(1.0f, 1.0f);
  o[4] = o[5] = o[6] = o[7] = 0u;
  tmem_st8(c.tmem + c.lane_sel + C_TB, o);
}

// Host-visible description of how a stage input / output is combined from (p0, v0, a_1..a_n).
//   p = p0 + cpv * v0 + sum_j cpa[j] * a_j        v = v0 + sum_j cva[j] * a_j
struct Combo {
  float cpv;
  float cpa[MAX_A + 1];
  float cva[MAX_A + 1];
};

// ---- "blocked" fp32 matrices --------------------------------------------------------------------------------
// Every per-agent buffer the stage kernels touch ([B][160] states and their adjoints, [B][64] accelerations and their
// adjoints) is stored tile-blocked: rows are padded to a multiple of 128; inside a 128-agent tile the float4 holding
// features 4 f4 .. 4 f4 + 3 of agent r sits at float4 index  (tile * F/4 + f4) * 128 + r.  A thread that owns one
// agent row (= one TMEM lane) therefore reads/writes 16 B that are adjacent to its neighbour lanes' 16 B: every
// warp-wide access is one contiguous 512 B segment, where a row-major layout would touch 32 cache lines.
// `ab200_rows_block` / `ab200_rows_unblock` convert from / to the public row-major tensors.
__device__ __forceinline__ float4* blk4(float* base, int tile, int F4, int f4, int row) {
  return reinterpret_cast<float4*>(base) + ((size_t)tile * F4 + f4) * TM + row;
}
__device__ __forceinline__ const float4* blk4(const float* base, int tile, int F4, int f4, int row) {
  return reinterpret_cast<const float4*>(base) + ((size_t)tile * F4 + f4) * TM + row;
}
// read-only 128-bit load (ld.global.nc): lets the compiler hoist input loads above the blob / output stores
__device__ __forceinline__ float4 ldro(const float4* p) { return __ldg(p); }
// ask the L2 for a whole tile of a blocked buffer (contiguous F * 512 bytes); one thread, no registers, no wait
__device__ __forceinline__ void prefetch_tile_l2(const float* base, int tile, int F4) {
  const char* p = reinterpret_cast<const char*>(base) + (size_t)tile * F4 * TM * 16;
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"((uint32_t)(F4 * TM * 16)) : "memory");
}
// per-line L2 prefetch of the float4 groups [f4_begin, f4_begin + n) of this thread's row in the NEXT tile: lanes whose
// row is a multiple of 8 cover one 128-byte line each (8 rows x 16 B)
__device__ __forceinline__ void prefetch_rows_l2(const float* base, int tile, int F4, int f4_begin, int n, int row) {
  if (row & 7) return;
  for (int j = 0; j < n; ++j) {
    const float4* p = reinterpret_cast<const float4*>(base) + ((size_t)tile * F4 + f4_begin + j) * TM + row;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
  }
}
constexpr int YF4 = D / 4, AF4 = P / 4;       // float4 groups per row of a state / an acceleration buffer

// setup shared by the kernels: weights -> smem (one bulk copy engine transfer), TMEM, barriers.
template <uint32_t IMG_BYTES = W_BYTES>
__device__ __forceinline__ SlotCtx stage_setup(uint8_t* smem, const uint8_t* wimg, uint64_t* bars, uint32_t* tmem_base_s,
                                               int* lock, int* status, int flags) {
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc<512>(tmem_base_s);
  if (tid == 0) {
    *lock = 0;
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    mbar_fence_init();
    constexpr uint32_t CH = IMG_BYTES / 4;     // 52,736 for the extended image; multiple of 16
    static_assert(CH * 4 == IMG_BYTES && CH % 16 == 0, "image chunking");
    mbar_arrive_expect_tx(&bars[2], IMG_BYTES);
#pragma unroll
    for (int i = 0; i < 4; ++i) bulk_g2s(smem + i * CH, wimg + i * CH, CH, &bars[2]);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (!mbar_wait(&bars[2], 0, STAGE_WAIT_CYCLES)) *status = 6;
  SlotCtx c;
  c.slot = tid / SLOT_THREADS;
  c.stid = tid % SLOT_THREADS;
  const int sw = c.stid >> 5, lane = tid & 31;
  const int q = warp & 3;
  c.hf = sw >> 2;
  c.row = q * 32 + lane;
  c.tmem = *tmem_base_s + (uint32_t)c.slot * SLOT_COLS;
  c.lane_sel = (uint32_t)(q * 32) << 16;
  c.sbase = smem_u32(smem);
  c.bar = &bars[c.slot];
  c.lock = lock;
  c.flags = flags;
  c.status = status;
  c.phase = 0;
  c.alive = true;
  if ((flags & 8) && c.slot == 1) {      // experiment: start the second slot half a layer late (anti-phase)
    const long long t0 = clock64();
    while (clock64() - t0 < 1200) { }
  }
  return c;
}

// Tile of iteration `it` for this (CTA, slot), or -1 when the slot is done.  Full waves fill both slots of every CTA; the tiles
// of a last, partial wave are spread over the FIRST slots of as many CTAs as possible before any second slot is used (a slot
// that runs alone is faster than two that share the tensor pipe and the SM's load/store path): matters when a rank owns few
// tiles, e.g. 1M agents sharded over 8 GPUs = 977 tiles on 296 slots.  Launch with grid = min(ntiles, #SMs).
__device__ __forceinline__ int slot_tile(int it, int ntiles, int slot) {
  const int per_wave = (int)gridDim.x * NSLOT;
  const int full = ntiles / per_wave * per_wave;
  int tile = it * per_wave + (int)blockIdx.x * NSLOT + slot;
  if (tile < full) return tile;
  if (it * per_wave != full) return -1;
  tile = full + slot * (int)gridDim.x + (int)blockIdx.x;
  return tile < ntiles ? tile : -1;
}

__device__ __forceinline__ void stage_teardown(uint32_t tmem_base) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) tmem_dealloc<512>(tmem_base);
}

}  // namespace stc

// L2 residency hints (kernel instantiations L2POL = true; backward kernel: default, AB200_STAGE_FLAGS bit 4 = off; forward attempt
// kernel: AB200_STAGE_FLAGS bit 128 = on, it measured SLOWER there: 1,151 -> 1,207 us): the accelerations a_j of a forward attempt
// (and the gx tiles of a fused backward launch) are written and re-read by the same thread over ~160 us while the launch streams ~600 MB of blobs through the 126 MB L2 (hit rate
// 22 %).  Tag the a_out stores and the a_j loads of all but their last use evict_last, and the LAST use evict_first (demotes the
// line so that dead lines do not pile up -- the un-demoted variant of round 2 was slower).
__device__ __forceinline__ uint64_t l2_policy_keep() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_drop() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float4 ld_l2hint(const float4* ptr, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr), "l"(pol) : "memory");
  return v;
}
__device__ __forceinline__ void st_l2hint(float4* ptr, const float4& v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;"
               :: "l"(ptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}

// tuning switches, read once from the environment (AB200_STAGE_FLAGS, default 0 = both off)
int stage_flags();

// host entry points (stage_fwd_tc.cu / stage_bwd_tc.cu / wgrad_tc.cu)
size_t stage_tc_image_bytes();
int stage_tc_pack(const float* w_flat, uint8_t* image, cudaStream_t st);
constexpr size_t IMG_STRIDE = (stc::W_BYTES + 255) / 256 * 256;   // bf16 image at 0, fp16 image at IMG_STRIDE
// Third image (stage_fwd2_tc.cu, the split-activation forward kernel): fp16 matrices WITHOUT K extensions followed by an
// fp32 table of the time-feature columns and biases (applied in fp32 by the epilogue); the status word follows it.
constexpr size_t IMG2_OFFSET = 2 * IMG_STRIDE;
constexpr size_t IMG2_BYTES = 192256;                              // stage_fwd2_tc.cu: f2::IMG_BYTES (static_assert there)
constexpr size_t STATUS_OFFSET = IMG2_OFFSET + IMG2_BYTES;         // multiple of 256
inline int* stage_status_ptr(const uint8_t* image) { return reinterpret_cast<int*>(const_cast<uint8_t*>(image) + STATUS_OFFSET); }
int stage_fwd2_pack(const float* w_flat, uint8_t* image2, cudaStream_t st);

}  // namespace ab200
