// Fused conv11 (8x8 s4, 16) -> conv12 (4x4 s2, 32) forward; one persistent CTA per SM, one frame
// per iteration, the next frame's bytes in flight while the current one is computed.
//
// Reference op: tf.nn.conv2d(..., padding='SAME') + b, relu  (NetworkVP.py:224-226), wired as
// NetworkDNav.py:81-82.  Implicit GEMM on bf16 tensor-core tiles, fp32 accumulate:
//   conv11: M = 441 positions, N = 16, K = 256 = (kh, kw, c)      -- A from the padded smem image
//   conv12: M = 121 positions, N = 32, K = 256 = (kh, kw, ci)     -- A from the padded smem conv11 output
// Both N are too narrow for a tcgen05 tile to be anything but shared-memory-read bound (A-operand
// bytes per MMA are fixed while N shrinks the math), and the im2col duplication (4x) would have to be
// materialised in shared memory for a UMMA descriptor.  Warp-level mma.sync reads the *un-duplicated*
// image straight out of shared memory with conflict-free 8-byte fragment loads, so the frame is
// converted once, stays on chip, and conv11's output never leaves the SM before conv12 consumes it.
//
// Pipeline per frame (HBM traffic: the fp32 frame, 112,896 B / sample, read exactly once):
//   TMA engine : cp.async.bulk of frame i+1 into the fp32 staging buffer, 4 chunks, each re-armed as soon as it is consumed
//   all warps  : wait(frame i) -> fp32 staging -> padded bf16 image -> [issue frame i+1] -> conv11 -> conv12 -> stores
// Measured limits (profiles/): shared-memory wavefronts (74 % of peak) and the legacy HMMA pipe
// (33 % active; its ceiling is ~44 M frames/s, below the 58 M frames/s HBM roofline).
#include "common.cuh"
#include "kernels.h"

namespace ga3c {

constexpr int CF_THREADS = 512, CF_WARPS = CF_THREADS / 32;
constexpr int CF_NCH = 1, CF_CHUNK_PIX = IMG * IMG / CF_NCH;     // staging chunks per frame
constexpr int CF_OFF_STG = 0;
constexpr int CF_OFF_XS = CF_OFF_STG + FRAME_BYTES;         // 112,896
constexpr int CF_OFF_N1P = CF_OFF_XS + XS_BYTES;            // 174,848
constexpr int CF_OFF_W12F = CF_OFF_N1P + N1P_BYTES;         // 193,280
constexpr int CF_OFF_N2S = CF_OFF_W12F + 16 * 2 * 32 * 16;  // 209,664
constexpr int CF_OFF_BIAS = CF_OFF_N2S + N2_POS * 64;       // 217,408
constexpr int CF_OFF_BAR = CF_OFF_BIAS + (C1_OUT + C2_OUT) * 4;   // 217,600
constexpr int CF_SMEM = CF_OFF_BAR + 8 * CF_NCH;        // 217,632 <= 232,448

// one chunk of the fp32 staging buffer (dense NHWC) -> zero-bordered bf16 image, 8 B per pixel
template <int NT>
__device__ __forceinline__ void convert_chunk(uint32_t stg, uint32_t xs, int c, int tid) {
#pragma unroll 2
  for (int k = tid; k < CF_CHUNK_PIX; k += NT) {
    const int i = c * CF_CHUNK_PIX + k;
    uint32_t r[4];
    lds128(r, stg + i * 16);
    const int y = i / IMG, xx = i - y * IMG;
    sts64(xs + (y + 2) * XS_ROW_BYTES + (xx + 2) * 8,
          pack_bf16(__uint_as_float(r[0]), __uint_as_float(r[1])), pack_bf16(__uint_as_float(r[2]), __uint_as_float(r[3])));
  }
}

__global__ void __launch_bounds__(CF_THREADS, 1)
conv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w11, const float* __restrict__ b11,
                const float* __restrict__ w12, const float* __restrict__ b12,
                uint16_t* __restrict__ n1_out, uint16_t* __restrict__ n2_out, int batch) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t stg = sbase + CF_OFF_STG, xs = sbase + CF_OFF_XS, n1p = sbase + CF_OFF_N1P,
                 w12f = sbase + CF_OFF_W12F, n2s = sbase + CF_OFF_N2S, bar = sbase + CF_OFF_BAR;
  float* bias_s = reinterpret_cast<float*>(smem + CF_OFF_BIAS);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;

  if (tid == 0) {
    for (int c = 0; c < CF_NCH; ++c) mbar_init(bar + 8 * c, 1);
    fence_mbar_init();
  }
  __syncthreads();
  int b = blockIdx.x;
  if (tid == 0 && b < batch)                                      // overlaps the weight setup below
    for (int c = 0; c < CF_NCH; ++c) stg_issue_chunk<CF_NCH>(stg, x + (size_t)b * STATE_DIM, c, bar);

  // zero the padded buffers once: the borders are never written again
  for (int i = tid; i < (XS_BYTES + N1P_BYTES) / 16; i += CF_THREADS) sts128(xs + i * 16, make_uint4(0, 0, 0, 0));
  griddep_launch();
  griddep_wait();               // the weights below come from the optimizer kernel that precedes this one in the stream
  // conv12 weights in mma B-fragment order: [kstep = tap][n-tile pair][lane] -> {b0,b1 (tile 2np), b0,b1 (tile 2np+1)}
  // K permutation inside a k16 step (one tap, 16 ci): logical cols (2t,2t+1,2t+8,2t+9) <-> ci (4t..4t+3)
  for (int i = tid; i < 16 * 2 * 32; i += CF_THREADS) {
    const int ks = i >> 6, np = (i >> 5) & 1, ln = i & 31, gg = ln >> 2, tt = ln & 3;
    uint32_t r[4];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int n = 8 * (2 * np + q) + gg;
      const float* w = w12 + (ks * 16 + 4 * tt) * C2_OUT + n;
      r[2 * q] = pack_bf16(w[0], w[C2_OUT]);
      r[2 * q + 1] = pack_bf16(w[2 * C2_OUT], w[3 * C2_OUT]);
    }
    sts128(w12f + i * 16, make_uint4(r[0], r[1], r[2], r[3]));
  }
  if (tid < C1_OUT) bias_s[tid] = b11[tid];
  if (tid < C2_OUT) bias_s[C1_OUT + tid] = b12[tid];
  // conv11 weights as register-resident B fragments: k16 step = (kh, half): pixels kw = 4*half + t, 4 channels
  uint32_t wb[16][2][2];
#pragma unroll
  for (int ks = 0; ks < 16; ++ks) {
    const int kh = ks >> 1, kw = 4 * (ks & 1) + t;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const float* w = w11 + ((kh * 8 + kw) * 4) * C1_OUT + 8 * nt + g;
      wb[ks][nt][0] = pack_bf16(w[0], w[C1_OUT]);
      wb[ks][nt][1] = pack_bf16(w[2 * C1_OUT], w[3 * C1_OUT]);
    }
  }
  __syncthreads();

  uint32_t phase = 0;
  for (; b < batch; b += gridDim.x) {
    const bool more = b + (int)gridDim.x < batch;
#pragma unroll 1
    for (int c = 0; c < CF_NCH; ++c) {
      mbar_wait(bar + 8 * c, phase);          // chunk c of frame b has landed in the staging buffer
      convert_chunk<CF_THREADS>(stg, xs, c, tid);
      __syncthreads();                        // every thread is done with chunk c (after the last one: image complete)
      if (tid == 0 && more) {
        fence_proxy_async();                  // order the generic-proxy reads above before the async-proxy refill
        stg_issue_chunk<CF_NCH>(stg, x + (size_t)(b + gridDim.x) * STATE_DIM, c, bar);
      }
    }
    phase ^= 1;

    // ---------------- conv11: 28 m16 tiles over 16 warps ----------------
    for (int tile = warp; tile < 28; tile += CF_WARPS) {
      const int r0 = tile * 16 + g, r1 = r0 + 8;
      const int p0 = min(r0, N1_POS - 1), p1 = min(r1, N1_POS - 1);
      const int oy0 = p0 / H1, ox0 = p0 - oy0 * H1, oy1 = p1 / H1, ox1 = p1 - oy1 * H1;
      const uint32_t a0 = xs + (4 * oy0) * XS_ROW_BYTES + (4 * ox0 + t) * 8;
      const uint32_t a1 = xs + (4 * oy1) * XS_ROW_BYTES + (4 * ox1 + t) * 8;
      float acc[2][4] = {};
#pragma unroll
      for (int ks = 0; ks < 16; ++ks) {
        const int off = (ks >> 1) * XS_ROW_BYTES + (ks & 1) * 32;
        uint32_t a[4];
        lds64(a[0], a[2], a0 + off);
        lds64(a[1], a[3], a1 + off);
        mma_bf16_16816(acc[0], a, wb[ks][0][0], wb[ks][0][1]);
        mma_bf16_16816(acc[1], a, wb[ks][1][0], wb[ks][1][1]);
      }
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const float bz0 = bias_s[8 * nt + 2 * t], bz1 = bias_s[8 * nt + 2 * t + 1];
        if (r0 < N1_POS)
          sts32(n1p + n1p_off(oy0 + 1, ox0 + 1, nt) + 4 * t,
                pack_bf16(fmaxf(acc[nt][0] + bz0, 0.f), fmaxf(acc[nt][1] + bz1, 0.f)));
        if (r1 < N1_POS)
          sts32(n1p + n1p_off(oy1 + 1, ox1 + 1, nt) + 4 * t,
                pack_bf16(fmaxf(acc[nt][2] + bz0, 0.f), fmaxf(acc[nt][3] + bz1, 0.f)));
      }
    }
    __syncthreads();

    // conv11 output to HBM (training only): coalesced 16-B chunks, un-swizzled
    if (n1_out != nullptr) {
      uint4* dst = reinterpret_cast<uint4*>(n1_out + (size_t)b * N1_POS * C1_OUT);
      for (int i = tid; i < N1_POS * 2; i += CF_THREADS) {
        const int pos = i >> 1, oy = pos / H1, ox = pos - oy * H1;
        uint32_t r[4];
        lds128(r, n1p + n1p_off(oy + 1, ox + 1, i & 1));
        dst[i] = make_uint4(r[0], r[1], r[2], r[3]);
      }
    }

    // ---------------- conv12: 8 m16 tiles x 2 n-halves = 16 warp tasks ----------------
    {
      const int mt = warp >> 1, nh = warp & 1;
      const int r0 = mt * 16 + g, r1 = r0 + 8;
      const int p0 = min(r0, N2_POS - 1), p1 = min(r1, N2_POS - 1);
      const int oy0 = p0 / H2, ox0 = p0 - oy0 * H2, oy1 = p1 / H2, ox1 = p1 - oy1 * H2;
      float acc[2][4] = {};
#pragma unroll
      for (int kh = 0; kh < 4; ++kh) {
#pragma unroll
        for (int kw = 0; kw < 4; ++kw) {
          const int ks = kh * 4 + kw;
          uint32_t a[4], bw[4];
          lds64(a[0], a[2], n1p + n1p_off(2 * oy0 + kh, 2 * ox0 + kw, t >> 1) + (t & 1) * 8);
          lds64(a[1], a[3], n1p + n1p_off(2 * oy1 + kh, 2 * ox1 + kw, t >> 1) + (t & 1) * 8);
          lds128(bw, w12f + ((ks * 2 + nh) * 32 + lane) * 16);
          mma_bf16_16816(acc[0], a, bw[0], bw[1]);
          mma_bf16_16816(acc[1], a, bw[2], bw[3]);
        }
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int nt = 2 * nh + q;
        const float bz0 = bias_s[C1_OUT + 8 * nt + 2 * t], bz1 = bias_s[C1_OUT + 8 * nt + 2 * t + 1];
        if (r0 < N2_POS)
          sts32(n2s + r0 * 64 + (8 * nt + 2 * t) * 2, pack_bf16(fmaxf(acc[q][0] + bz0, 0.f), fmaxf(acc[q][1] + bz1, 0.f)));
        if (r1 < N2_POS)
          sts32(n2s + r1 * 64 + (8 * nt + 2 * t) * 2, pack_bf16(fmaxf(acc[q][2] + bz0, 0.f), fmaxf(acc[q][3] + bz1, 0.f)));
      }
    }
    __syncthreads();
    {
      uint4* dst = reinterpret_cast<uint4*>(n2_out + (size_t)b * FLAT);
      for (int i = tid; i < FLAT * 2 / 16; i += CF_THREADS) {
        uint32_t r[4];
        lds128(r, n2s + i * 16);
        dst[i] = make_uint4(r[0], r[1], r[2], r[3]);
      }
    }
    // no barrier needed here: the next iteration's convert only writes xs (last read two barriers ago),
    // its conv11 writes n1p only after the barrier that follows the convert, and n2s is rewritten only
    // after two more barriers.
  }
}

int configure_conv_fwd() {
  return (int)cudaFuncSetAttribute(conv_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CF_SMEM);
}

int launch_conv_fwd(const float* x, const float* w11, const float* b11, const float* w12, const float* b12,
                    uint16_t* n1_out, uint16_t* n2_out, int batch, int num_sms, cudaStream_t stream) {
  const int grid = min(batch, num_sms);
  return launch_pdl(conv_fwd_kernel, dim3(grid), dim3(CF_THREADS), CF_SMEM, stream, x, w11, b11, w12, b12, n1_out, n2_out, batch);
}

}  // namespace ga3c
