// Layout of the activation / gradient "blob" buffer that stage_bwd_tc.cu writes and wgrad_tc.cu consumes, and of the
// per-CTA partial weight-gradient buffer.
//
// A blob holds one 128-agent tile of one layer operand as bf16, already in the canonical un-swizzled MN-major UMMA
// operand image with K = agent:   byte(f, agent) = (f / 8) * 2048 + agent * 16 + (f % 8) * 2
// (core matrix = 8 agents x 8 features; K-adjacent cores 128 B apart, feature groups 2048 B apart), so a warp of 32
// agents writes 512 contiguous bytes per feature group and the weight-gradient kernel copies blobs verbatim.
#pragma once
#include <stdint.h>
#include <stddef.h>

namespace ab200 {
namespace wg {

constexpr int TM = 128;
constexpr uint32_t FG_BYTES = TM * 16;                       // one group of 8 features for 128 agents
constexpr uint32_t X1_FEATS = 176, X1_BYTES = X1_FEATS / 8 * FG_BYTES;    // p, v, h, sin, cos, 1, 0-pad
constexpr uint32_t HID_BYTES = 128 / 8 * FG_BYTES;           // 32,768
constexpr uint32_t GO_BYTES = 64 / 8 * FG_BYTES;             // 16,384
constexpr uint32_t BYTES_PER_BLOB_SET = X1_BYTES + 10 * HID_BYTES + GO_BYTES;   // 389,120 per tile-stage

struct SpillLayout {
  int nblobs;
  __host__ __device__ size_t x1(int b) const { return (size_t)b * X1_BYTES; }
  __host__ __device__ size_t act_base() const { return (size_t)nblobs * X1_BYTES; }
  __host__ __device__ size_t act(int i, int b) const { return act_base() + ((size_t)i * nblobs + b) * HID_BYTES; }      // z0,u0,z1,u1,z2
  __host__ __device__ size_t grad_base() const { return act_base() + (size_t)5 * nblobs * HID_BYTES; }
  __host__ __device__ size_t grad(int i, int b) const { return grad_base() + ((size_t)i * nblobs + b) * HID_BYTES; }   // g1,gA0,gB0,gA1,gB1
  __host__ __device__ size_t go_base() const { return grad_base() + (size_t)5 * nblobs * HID_BYTES; }
  __host__ __device__ size_t go(int b) const { return go_base() + (size_t)b * GO_BYTES; }
  __host__ __device__ size_t total() const { return (size_t)nblobs * BYTES_PER_BLOB_SET; }
};

// What a TRAINING forward launch (stage_fwd2_tc.cu) can save per stage for the backward pass, so that the backward kernel does not
// recompute the net and the weight-gradient kernel reads the layer inputs from here:
//   level 1: the stage input X blob;   level 2: + the five hidden activation blobs (z0,u0,z1,u1,z2) + the ReLU masks
// (two words per thread and layer, the bit layout of stage_bwd_tc.cu's bwd_fwd_epi: [layer][hf * 128 + row] uint2).
// A level-1 buffer is the prefix of a level-2 buffer.
constexpr uint32_t MASK_BYTES = 5 * 256 * 8;                 // 10,240 per tile
struct FwdSaveLayout {
  int ntiles;
  __host__ __device__ size_t x1(int tile) const { return (size_t)tile * X1_BYTES; }
  __host__ __device__ size_t act(int i, int tile) const { return (size_t)ntiles * X1_BYTES + ((size_t)i * ntiles + tile) * HID_BYTES; }
  __host__ __device__ size_t mask(int tile) const { return (size_t)ntiles * (X1_BYTES + 5 * HID_BYTES) + (size_t)tile * MASK_BYTES; }
  __host__ __device__ size_t total(int level) const {
    return level >= 2 ? (size_t)ntiles * (X1_BYTES + 5 * HID_BYTES + MASK_BYTES) : (size_t)ntiles * X1_BYTES;
  }
};

// per-CTA fp32 partial weight gradients: pair 0 [128][176] | pairs 1..4 [128][128] + bias block [128][16] | pair 5 [128 in][64 out]
constexpr int PART_W1 = 0, PART_W1_N = 176;
constexpr int PART_HH = 128 * 176, PART_HH_N = 144, PART_HH_SZ = 128 * 144;
constexpr int PART_WO = PART_HH + 4 * PART_HH_SZ, PART_WO_N = 64;
constexpr int PART_TOTAL = PART_WO + 128 * 64;               // 104,448 floats

}  // namespace wg
}  // namespace ab200
