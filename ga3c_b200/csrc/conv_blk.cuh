// Shared by the kernels that stream fp32 frames into the space-to-depth "block matrix" operand
// (conv_fwd.cu, conv_bwd.cu: conv11 forward and conv11 weight gradient on tcgen05).
//
// The zero-padded 88x88x4 bf16 image is cut into 22x22 blocks of 4x4 pixels; block (Y, X) is row Y*22 + X of
// Blk[484, K = 64 = (dy, dx, c)].  Storage is the no-swizzle UMMA layout with ALL rows contiguous: 16-byte k-chunk
// j = dy*2 + (dx>>1) of row r lives at j*BLK_LBO + r*16 (8 elements = (dx&1, c)).  Read K-major it is the A operand of
// the forward GEMMs, read MN-major it is the transposed operand of the weight gradient; in both uses a ROW SHIFT
// (the 2x2 block quadrants of an 8x8 stride-4 window) is just a different descriptor start address.
#pragma once
#include "common.cuh"

namespace ga3c {

constexpr int CH_ROWS = 12, CF_NCHUNK = IMG / CH_ROWS, CH_BYTES = CH_ROWS * IMG * 16;     // 7 chunks of 16,128 B
constexpr int CF_NSLOT = 4;
static_assert(CF_NCHUNK * CH_ROWS == IMG && CH_BYTES % 16 == 0, "chunks must tile the frame");
// block matrix: 22 x 22 blocks (+ slack rows read by the dead part of the last M tile); chunk arrays padded so that
// neighbouring k-chunks start 16 banks apart.  548 rows cover conv_fwd's four 128-row tiles; kernels that read fewer rows
// pass their own plane stride (LBO) to the helpers below.
constexpr int BLK_W = 22, BLK_ROWS = 548, BLK_LBO = BLK_ROWS * 16, BLK_BYTES = 8 * BLK_LBO;              // 8,768 / 70,144

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}
// no-swizzle K-major operand: 16-byte k-chunk j of row r at start + j*LBO + (r/8)*SBO + (r%8)*16
__device__ __forceinline__ uint64_t make_desc_ns(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

// per-lane part of a pixel's destination in Blk: padded pixel px = x + 2 -> block column X = px >> 2, dx = px & 3
template <int LBO = BLK_LBO>
__device__ __forceinline__ void blk_lane_offsets(int lane, uint32_t (&lane_off)[3]) {
#pragma unroll
  for (int it = 0; it < 3; ++it) {
    const int px = lane + 32 * it + 2;
    lane_off[it] = ((px >> 1) & 1) * LBO + (px >> 2) * 16 + (px & 1) * 8;
  }
}

// one 12-row fp32 chunk (staging slot `src`) -> bf16 -> Blk; warp w of AUX_WARPS converts rows w, w + AUX_WARPS, ...
// All 16-byte loads are issued before the first conversion so their latencies overlap.
template <int AUX_WARPS, int LBO = BLK_LBO>
__device__ __forceinline__ void blk_convert_chunk(uint32_t src, uint32_t blk, int c, int warp, int lane,
                                                  const uint32_t (&lane_off)[3]) {
  constexpr int RPW = CH_ROWS / AUX_WARPS;
  static_assert(RPW * AUX_WARPS == CH_ROWS, "image rows of a chunk split evenly over the aux warps");
  uint32_t px4[RPW][3][4];
#pragma unroll
  for (int rr = 0; rr < RPW; ++rr)
#pragma unroll
    for (int it = 0; it < 3; ++it)
      if (lane + 32 * it < IMG) lds128(px4[rr][it], src + ((warp + rr * AUX_WARPS) * IMG + lane + 32 * it) * 16);
#pragma unroll
  for (int rr = 0; rr < RPW; ++rr) {
    const int py = c * CH_ROWS + warp + rr * AUX_WARPS + 2;        // padded row -> block row Y = py >> 2, dy = py & 3
    const uint32_t row_off = blk + (py & 3) * (2 * LBO) + (py >> 2) * (BLK_W * 16);
#pragma unroll
    for (int it = 0; it < 3; ++it) {
      if (lane + 32 * it < IMG) {
        const uint32_t* r = px4[rr][it];
        sts64(row_off + lane_off[it], pack_bf16(__uint_as_float(r[0]), __uint_as_float(r[1])),
              pack_bf16(__uint_as_float(r[2]), __uint_as_float(r[3])));
      }
    }
  }
}

}  // namespace ga3c
