// Shared by the kernels that stream fp32 frames into the space-to-depth "block matrix" operand
// (conv_fwd.cu, conv_bwd_fused.cu: conv11 forward and conv11 weight gradient on tcgen05).
//
// The zero-padded 88x88x4 bf16 image is cut into 22x22 blocks of 4x4 pixels; block (Y, X) is row Y*22 + X of
// Blk[484, K = 64 = (dy, dx, c)].  Storage is the no-swizzle UMMA layout with ALL rows contiguous: 16-byte k-chunk
// j = dy*2 + (dx>>1) of row r lives at j*BLK_LBO + r*16 (8 elements = (dx&1, c)).  Read K-major it is the A operand of
// the forward GEMMs, read MN-major it is the transposed operand of the weight gradient; in both uses a ROW SHIFT
// (the 2x2 block quadrants of an 8x8 stride-4 window) is just a different descriptor start address.
#pragma once
#include "common.cuh"

namespace ga3c {

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

// ---- per-warp streaming pipeline --------------------------------------------------------------------------------
// Every aux warp runs its OWN chunk pipeline: chunk = 4 image rows (5,376 B fp32, one cp.async.bulk), two ring slots per
// warp, no barrier between the warps.  (With all aux warps converting one 12-row chunk in lockstep -- wait, load, convert,
// fence, block barrier, re-arm -- the per-chunk latency chain, not HBM, set the frame rate.)  Chunk c (image rows 4c..4c+3)
// is padded rows 4c+2..4c+5, i.e. (block row c, dy = 2, 3) and (block row c+1, dy = 0, 1); block row Y is complete after
// chunks Y-1 and Y.  Consumers are four position groups i (conv11 tiles forward, k-step groups in the weight gradient),
// each reading Blk rows [~128 i, 128 i + 150]: block rows floor(128 i / 22) .. min((128 i + 150) / 22, 21).
constexpr int PW_ROWS = 4, PW_NCHUNK = IMG / PW_ROWS, PW_BYTES = PW_ROWS * IMG * 16, PW_SLOTS = 2;       // 21 chunks of 5,376 B
static_assert(PW_NCHUNK * PW_ROWS == IMG && PW_BYTES % 16 == 0, "chunks must tile the frame");
// first group that needs chunk c (the group's "rows ready" barrier counts its own chunks: 7, 6, 6, 2) ...
__device__ __forceinline__ int pw_first_consumer(int c) { return c <= 6 ? 0 : c <= 12 ? 1 : c <= 18 ? 2 : 3; }
__host__ __device__ constexpr int pw_group_chunks(int i) { return i == 0 ? 7 : i == 3 ? 2 : 6; }
// ... and last group of the PREVIOUS frame that still reads block rows c, c+1 (they are about to be rewritten)
__device__ __forceinline__ int pw_last_consumer(int c) { return c >= 16 ? 3 : c >= 10 ? 2 : c >= 4 ? 1 : 0; }

// uint8 frames (F2 ingestion: the reference's `image.astype(np.float32) / 128.0 - 1.0`, Environment.py:60, done here): a
// 4-row chunk is 1,344 B, one 32-bit load per pixel; x = k / 128 - 1 is exact in fp32 and in bf16.
constexpr int PW_BYTES_U8 = PW_ROWS * IMG * 4;
__device__ __forceinline__ void u8x4_to_bf16x4(uint32_t k4, uint32_t& lo, uint32_t& hi) {
  // byte -> float through the 2^23 trick (0x4B000000 | k = 8388608 + k), then (k - 128) / 128
  const float f0 = __uint_as_float(0x4B000000u | (k4 & 0xFFu)) - 8388608.f, f1 = __uint_as_float(0x4B000000u | ((k4 >> 8) & 0xFFu)) - 8388608.f,
              f2 = __uint_as_float(0x4B000000u | ((k4 >> 16) & 0xFFu)) - 8388608.f, f3 = __uint_as_float(0x4B000000u | (k4 >> 24)) - 8388608.f;
  lo = pack_bf16(fmaf(f0, 0.0078125f, -1.f), fmaf(f1, 0.0078125f, -1.f));
  hi = pack_bf16(fmaf(f2, 0.0078125f, -1.f), fmaf(f3, 0.0078125f, -1.f));
}

// One warp, one 4-row chunk, in two phases so that the ring slot can be re-armed (and the Blk rows' previous readers
// awaited) BETWEEN them: load + convert into 24 registers of packed bf16 (the slot is free afterwards), then store into Blk.
template <bool U8>
__device__ __forceinline__ void blk_load_rows4(uint32_t src, int lane, uint32_t (&pk)[PW_ROWS][3][2]) {
  if (U8) {
    uint32_t k4[PW_ROWS][3];
#pragma unroll
    for (int rr = 0; rr < PW_ROWS; ++rr)
#pragma unroll
      for (int it = 0; it < 3; ++it)
        if (lane + 32 * it < IMG) asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(k4[rr][it]) : "r"(src + (rr * IMG + lane + 32 * it) * 4));
#pragma unroll
    for (int rr = 0; rr < PW_ROWS; ++rr)
#pragma unroll
      for (int it = 0; it < 3; ++it)
        if (lane + 32 * it < IMG) u8x4_to_bf16x4(k4[rr][it], pk[rr][it][0], pk[rr][it][1]);
  } else {
#pragma unroll
    for (int half = 0; half < 2; ++half) {                          // two rows at a time: 24 registers of loads in flight
      uint32_t px4[2][3][4];
#pragma unroll
      for (int rr = 0; rr < 2; ++rr)
#pragma unroll
        for (int it = 0; it < 3; ++it)
          if (lane + 32 * it < IMG) lds128(px4[rr][it], src + ((2 * half + rr) * IMG + lane + 32 * it) * 16);
#pragma unroll
      for (int rr = 0; rr < 2; ++rr)
#pragma unroll
        for (int it = 0; it < 3; ++it)
          if (lane + 32 * it < IMG) {
            const uint32_t* r = px4[rr][it];
            pk[2 * half + rr][it][0] = pack_bf16(__uint_as_float(r[0]), __uint_as_float(r[1]));
            pk[2 * half + rr][it][1] = pack_bf16(__uint_as_float(r[2]), __uint_as_float(r[3]));
          }
    }
  }
}
template <int LBO>
__device__ __forceinline__ void blk_store_rows4(uint32_t blk, int c, int lane, const uint32_t (&lane_off)[3],
                                                const uint32_t (&pk)[PW_ROWS][3][2]) {
#pragma unroll
  for (int rr = 0; rr < PW_ROWS; ++rr) {
    const int py = c * PW_ROWS + rr + 2;                           // padded row -> block row Y = py >> 2, dy = py & 3
    const uint32_t row_off = blk + (py & 3) * (2 * LBO) + (py >> 2) * (BLK_W * 16);
#pragma unroll
    for (int it = 0; it < 3; ++it)
      if (lane + 32 * it < IMG) sts64(row_off + lane_off[it], pk[rr][it][0], pk[rr][it][1]);
  }
}

}  // namespace ga3c
