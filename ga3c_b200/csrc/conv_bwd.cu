// Backward of the two convolutions (the autodiff half of opt.minimize, NetworkVP_discrate.py:130,
// for the layers defined at NetworkVP.py:212-228 / NetworkDNav.py:81-82).
//
//   conv12_bwd_kernel  per frame: dn1 = conv12 data-gradient of dn2 (masked by relu'(n1)),
//                      g_w12 += patches(n1)^T dn2, g_b12 += colsum(dn2)
//   conv11_wgrad_kernel per frame: g_w11 += patches(x)^T dn1, g_b11 += colsum(dn1)
//                      (conv11 has no data-gradient: x is the input)
//
// bf16 operands / fp32 accumulate on mma.sync tiles fed from shared memory; weight-gradient
// accumulators stay in registers across all frames a CTA processes and are flushed with one
// atomicAdd per element per CTA.
#include "common.cuh"
#include "kernels.h"

namespace ga3c {

// ------------------------------------------------------------------------------------------------
// conv12 backward.  One persistent CTA per SM, 16 warps, inputs double-buffered with cp.async.
// Every shared-memory layout is AFFINE (padding instead of XOR swizzles) so that tap / k-step offsets
// fold into instruction immediates, and everything that depends only on the thread (tile -> pixel
// maps, staging destinations) is computed once per CTA: the frame loop is loads + HMMA.
//   n1pl  conv11 output, SAME-padded (1 before, 2 after) 24x24 pixels, split by x parity into two planes
//         of 24 x 12 pixels, 48 B per pixel (32 B data + 16 B pad): the stride-2 pixel walks of conv12 become
//         unit-stride walks at 48 B, conflict-free for ldmatrix and for the 4-byte mask reads
//   dn2a  dn2 on a zero-bordered 12x12 grid (index (oy+1, ox+1)), 96 B per position: dgrad A fragments (8-byte loads)
//   dn2b  dn2 as 128 plain positions (121 + zero rows), 80 B per position: wgrad B fragments (ldmatrix.trans)
constexpr int B12_THREADS = 512, B12_WARPS = B12_THREADS / 32;
constexpr int N1PL_PITCH = 48, N1PL_PLANE = 24 * 12 * N1PL_PITCH, N1PL_BYTES = 2 * N1PL_PLANE;   // 13,824 / 27,648
constexpr int DN2A_W = 12, DN2A_PITCH = 96, DN2A_BYTES = DN2A_W * DN2A_W * DN2A_PITCH;             // 13,824
constexpr int DN2B_PITCH = 80, DN2B_BYTES = 128 * DN2B_PITCH;                                      // 10,240
constexpr int B12_IN_BYTES = N1PL_BYTES + DN2A_BYTES + DN2B_BYTES;                                 // 51,712
constexpr int B12_TILES = 29;                                       // dgrad m16 tiles: 7 + 7 + 7 + 8 over the 4 parity classes
constexpr int B12_OFF_IN = 0;
constexpr int B12_OFF_WDF = B12_OFF_IN + 2 * B12_IN_BYTES;          // 103,424
constexpr int B12_OFF_DN1S = B12_OFF_WDF + 4 * 8 * 32 * 16;         // 119,808
constexpr int B12_OFF_LUT = B12_OFF_DN1S + N1_POS * 32;             // 133,920
constexpr int B12_OFF_RED = B12_OFF_LUT + B12_TILES * 16 * 8;       // 137,632
constexpr int B12_SMEM = B12_OFF_RED + B12_THREADS * 4;             // 139,680

__device__ __forceinline__ int n1pl_off(int Y, int X) {              // padded pixel (Y, X) in [0,24)^2
  return (X & 1) * N1PL_PLANE + (Y * 12 + (X >> 1)) * N1PL_PITCH;
}

__global__ void __launch_bounds__(B12_THREADS, 1)
conv12_bwd_kernel(const uint16_t* __restrict__ n1, const uint16_t* __restrict__ dn2, const float* __restrict__ w12,
                  uint16_t* __restrict__ dn1, float* __restrict__ g_w12, float* __restrict__ g_b12, int batch) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t wdf = sbase + B12_OFF_WDF, dn1s = sbase + B12_OFF_DN1S, lut = sbase + B12_OFF_LUT;
  float* red = reinterpret_cast<float*>(smem + B12_OFF_RED);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int j = lane >> 3, rr = lane & 7;
  const int stride = gridDim.x;

  for (int i = tid; i < 2 * B12_IN_BYTES / 16; i += B12_THREADS) sts128(sbase + i * 16, make_uint4(0, 0, 0, 0));
  // data-gradient weights in B-fragment order: [class = par_y*2+par_x][kstep = (a*2+b)*2+half][lane]
  //   K = (a, b, co): taps kh = par_y + 2a, kw = par_x + 2b ; N = ci
  for (int i = tid; i < 4 * 8 * 32; i += B12_THREADS) {
    const int cls = i >> 8, ks = (i >> 5) & 7, ln = i & 31, gg = ln >> 2, tt = ln & 3;
    const int kh = (cls >> 1) + 2 * (ks >> 2), kw = (cls & 1) + 2 * ((ks >> 1) & 1), half = ks & 1;
    uint32_t r[4];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int ci = 8 * q + gg;
      const float* w = w12 + ((kh * 4 + kw) * C1_OUT + ci) * C2_OUT + 16 * half + 4 * tt;
      r[2 * q] = pack_bf16(w[0], w[1]);
      r[2 * q + 1] = pack_bf16(w[2], w[3]);
    }
    sts128(wdf + i * 16, make_uint4(r[0], r[1], r[2], r[3]));
  }
  // dgrad row table: entry (tile, row) = { dn2a byte offset of its (untapped) position, dn1 staging offset or 0xFFFF,
  //                                        n1pl offset of its pixel (relu mask), 0 }
  //   class (pary, parx): pixel y = 2*iy + 1 - pary ; q = i + 2 - par ; tap a reaches dn2 row q - a on the padded grid
  for (int e = tid; e < B12_TILES * 16; e += B12_THREADS) {
    const int tile = e >> 4, r = e & 15;
    const int cls = tile < 7 ? 0 : tile < 14 ? 1 : tile < 21 ? 2 : 3;
    const int pary = cls >> 1, parx = cls & 1, nx = 10 + parx, cnt = (10 + pary) * nx;
    const int rix = (tile - cls * 7) * 16 + r, c = min(rix, cnt - 1);
    const int iy = c / nx, ix = c - iy * nx;
    const int y = 2 * iy + 1 - pary, xx = 2 * ix + 1 - parx;
    const uint32_t a_off = ((iy + 2 - pary) * DN2A_W + (ix + 2 - parx)) * DN2A_PITCH;
    const uint32_t o_off = rix < cnt ? (uint32_t)(y * H1 + xx) * 32 : 0xFFFFu;
    const uint32_t m_off = n1pl_off(y + 1, xx + 1);
    sts64(lut + e * 8, a_off | (o_off << 16), m_off);
  }
  // staging destinations of this thread's 16-B chunks (frame-invariant)
  uint32_t dst_n1[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int i = min(tid + k * B12_THREADS, N1_POS * 2 - 1), pos = i >> 1, y = pos / H1, xx = pos - y * H1;
    dst_n1[k] = n1pl_off(y + 1, xx + 1) + (i & 1) * 16;
  }
  uint32_t dst_a, dst_b;
  {
    const int i = min(tid, N2_POS * 4 - 1), pos = i >> 2, oy = pos / H2, ox = pos - oy * H2;
    dst_a = N1PL_BYTES + ((oy + 1) * DN2A_W + ox + 1) * DN2A_PITCH + (i & 3) * 16;
    dst_b = N1PL_BYTES + DN2A_BYTES + pos * DN2B_PITCH + (i & 3) * 16;
  }
  // wgrad A bases: row (ks, j>>1, rr) -> pixel (2*oy, 2*ox) of the tap-(0,0) window, 16-B chunk j&1
  uint32_t wa_base[8];
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    const int pos = min(ks * 16 + (j >> 1) * 8 + rr, N2_POS - 1), oy = pos / H2, ox = pos - oy * H2;   // rows >= 121 meet zero dn2b rows
    wa_base[ks] = (2 * oy * 12 + ox) * N1PL_PITCH + (j & 1) * 16;
  }
  const int kh = warp >> 2, kw = warp & 3;                            // wgrad: one tap per warp
  const uint32_t tap_off = (kw & 1) * N1PL_PLANE + (kh * 12 + (kw >> 1)) * N1PL_PITCH;
  griddep_launch();
  __syncthreads();
  griddep_wait();               // dn2 comes from the dense1 data-gradient GEMM that precedes this kernel

  auto prefetch = [&](int b, uint32_t in) {
    const uint4* s1 = reinterpret_cast<const uint4*>(n1 + (size_t)b * N1_POS * C1_OUT);
    cp_async16(in + dst_n1[0], s1 + tid, 16);
    if (tid + B12_THREADS < N1_POS * 2) cp_async16(in + dst_n1[1], s1 + tid + B12_THREADS, 16);
    if (tid < N2_POS * 4) {
      const uint4* s2 = reinterpret_cast<const uint4*>(dn2 + (size_t)b * FLAT) + tid;
      cp_async16(in + dst_a, s2, 16);
      cp_async16(in + dst_b, s2, 16);
    }
  };

  float wacc[4][4] = {};      // wgrad: tap = warp, 16 ci x 32 co
  float bacc = 0.f;           // db2 partial: co = tid & 31, part = tid >> 5

  int b = blockIdx.x, buf = 0;
  if (b < batch) prefetch(b, sbase + B12_OFF_IN);
  cp_async_commit();
  for (; b < batch; b += stride, buf ^= 1) {
    const uint32_t in = sbase + B12_OFF_IN + buf * B12_IN_BYTES;
    const uint32_t n1pl = in, dn2a = in + N1PL_BYTES, dn2b = dn2a + DN2A_BYTES;
    cp_async_wait<0>();
    __syncthreads();          // frame b visible to every warp; the other buffer and dn1s are free (barrier at loop end)
    if (b + stride < batch) prefetch(b + stride, sbase + B12_OFF_IN + (buf ^ 1) * B12_IN_BYTES);
    cp_async_commit();

    // ---------------- data gradient: 29 m16 tiles over 4 parity classes ----------------
    for (int tile = warp; tile < B12_TILES; tile += B12_WARPS) {
      const int cls = tile < 7 ? 0 : tile < 14 ? 1 : tile < 21 ? 2 : 3;
      uint32_t e0a, e0m, e1a, e1m;
      lds64(e0a, e0m, lut + (tile * 16 + g) * 8);
      lds64(e1a, e1m, lut + (tile * 16 + g + 8) * 8);
      const uint32_t a0 = dn2a + (e0a & 0xFFFFu) + (t >> 1) * 16 + (t & 1) * 8;
      const uint32_t a1 = dn2a + (e1a & 0xFFFFu) + (t >> 1) * 16 + (t & 1) * 8;
      const uint32_t bsrc = wdf + cls * 4096 + lane * 16;
      float acc[2][4] = {};
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const int off = -((ks >> 2) * DN2A_W + ((ks >> 1) & 1)) * DN2A_PITCH + (ks & 1) * 32;
        uint32_t a[4], bw[4];
        lds64(a[0], a[2], a0 + off);
        lds64(a[1], a[3], a1 + off);
        lds128(bw, bsrc + ks * 512);
        mma_bf16_16816(acc[0], a, bw[0], bw[1]);
        mma_bf16_16816(acc[1], a, bw[2], bw[3]);
      }
      const uint32_t o0 = e0a >> 16, o1 = e1a >> 16;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        if (o0 != 0xFFFFu) {
          uint32_t m;
          asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(m) : "r"(n1pl + e0m + nt * 16 + 4 * t));
          sts32(dn1s + o0 + nt * 16 + 4 * t, pack_bf16((m & 0x7FFFu) ? acc[nt][0] : 0.f, (m & 0x7FFF0000u) ? acc[nt][1] : 0.f));
        }
        if (o1 != 0xFFFFu) {
          uint32_t m;
          asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(m) : "r"(n1pl + e1m + nt * 16 + 4 * t));
          sts32(dn1s + o1 + nt * 16 + 4 * t, pack_bf16((m & 0x7FFFu) ? acc[nt][2] : 0.f, (m & 0x7FFF0000u) ? acc[nt][3] : 0.f));
        }
      }
    }

    // ---------------- weight gradient: M = 16 ci of this warp's tap, N = 32 co, K = 121 positions ----------------
    {
      const uint32_t brow = dn2b + ((j & 1) * 8 + rr) * DN2B_PITCH + (j >> 1) * 16;
      const uint32_t asrc = n1pl + tap_off;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        uint32_t bf0[4], bf1[4], af[4];
        ldsm_x4_t(bf0, brow + ks * 16 * DN2B_PITCH);
        ldsm_x4_t(bf1, brow + ks * 16 * DN2B_PITCH + 32);
        ldsm_x4_t(af, asrc + wa_base[ks]);
        mma_bf16_16816(wacc[0], af, bf0[0], bf0[1]);
        mma_bf16_16816(wacc[1], af, bf0[2], bf0[3]);
        mma_bf16_16816(wacc[2], af, bf1[0], bf1[1]);
        mma_bf16_16816(wacc[3], af, bf1[2], bf1[3]);
      }
    }
    // bias gradient partials
    {
      const uint32_t src = dn2b + (tid & 31) * 2;
      for (int pos = tid >> 5; pos < N2_POS; pos += B12_WARPS) {
        uint16_t v;
        asm volatile("ld.shared.u16 %0, [%1];\n" : "=h"(v) : "r"(src + pos * DN2B_PITCH));
        bacc += __uint_as_float((uint32_t)v << 16);
      }
    }
    __syncthreads();
    {
      uint4* dst = reinterpret_cast<uint4*>(dn1 + (size_t)b * N1_POS * C1_OUT);
      for (int i = tid; i < N1_POS * 2; i += B12_THREADS) {
        uint32_t r[4];
        lds128(r, dn1s + i * 16);
        dst[i] = make_uint4(r[0], r[1], r[2], r[3]);
      }
    }
    __syncthreads();   // dn1s and this input buffer are rewritten by the following iterations
  }

#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    float* o = g_w12 + (warp * C1_OUT) * C2_OUT + 8 * nt + 2 * t;
    atomicAdd(o + g * C2_OUT, wacc[nt][0]);
    atomicAdd(o + g * C2_OUT + 1, wacc[nt][1]);
    atomicAdd(o + (g + 8) * C2_OUT, wacc[nt][2]);
    atomicAdd(o + (g + 8) * C2_OUT + 1, wacc[nt][3]);
  }
  red[tid] = bacc;
  __syncthreads();
  if (tid < C2_OUT) {
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < B12_WARPS; ++p) s += red[p * 32 + tid];
    atomicAdd(g_b12 + tid, s);
  }
}

// ------------------------------------------------------------------------------------------------
// conv11 weight gradient.  One persistent CTA per SM, 16 warps; per frame
//   TMA engine : cp.async.bulk of frame i+1 into the fp32 staging buffer (mbarrier), cp.async of dn1(i+1)
//   all warps  : wait -> fp32 staging -> padded bf16 image -> [issue frame i+1] -> wgrad MMAs
// warp = (kh, part): M tile pair (kh, half 0/1) x N 16, K = the 14 k16 position steps of its part.
constexpr int W11_THREADS = 512;
constexpr int W11_NCH = 2, W11_CHUNK_PIX = IMG * IMG / W11_NCH;   // staging chunks per frame
constexpr int DN1S_ROWS = 448;                                    // 441 padded to 28 k16 steps
constexpr int DN1S_BYTES = DN1S_ROWS * 32;                        // 14,336
constexpr int C11_OFF_STG = 0;
constexpr int C11_OFF_XS = C11_OFF_STG + FRAME_BYTES;         // 112,896
constexpr int C11_OFF_DN1S = C11_OFF_XS + XS_BYTES;               // 174,848 (two buffers)
constexpr int C11_OFF_RED = C11_OFF_DN1S + 2 * DN1S_BYTES;        // 203,520
constexpr int C11_OFF_BAR = C11_OFF_RED + W11_THREADS * 4;        // 205,568
constexpr int C11_SMEM = C11_OFF_BAR + 8 * W11_NCH;            // 205,600

// 16-B chunk h (pixels 2h, 2h+1) of pixel quad q sits at h ^ ((q >> 2) & 1): the transposed ldmatrix
// reads of 8 consecutive positions (32 B apart) then touch all 32 banks once
__device__ __forceinline__ uint32_t xs_chunk_off(int px) {        // byte offset of the 16-B chunk holding padded pixel px (even)
  const int q = px >> 2, h = (px >> 1) & 1;
  return q * 32 + ((h ^ ((q >> 2) & 1)) << 4);
}

__device__ __forceinline__ void w11_prefetch_dn1(const uint16_t* __restrict__ dn1, int b, uint32_t dn1s, int tid) {
  const uint4* s1 = reinterpret_cast<const uint4*>(dn1 + (size_t)b * N1_POS * C1_OUT);
  for (int i = tid; i < N1_POS * 2; i += W11_THREADS) {
    const int pos = i >> 1;
    cp_async16(dn1s + pos * 32 + ((((i & 1) ^ (pos >> 2)) & 1) << 4), s1 + i, 16);
  }
}

__global__ void __launch_bounds__(W11_THREADS, 1)
conv11_wgrad_kernel(const float* __restrict__ x, const uint16_t* __restrict__ dn1, float* __restrict__ g_w11,
                    float* __restrict__ g_b11, int batch) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t stg = sbase + C11_OFF_STG, xs = sbase + C11_OFF_XS, bar = sbase + C11_OFF_BAR;
  float* red = reinterpret_cast<float*>(smem + C11_OFF_RED);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int j = lane >> 3, rr = lane & 7;
  const int kh = warp & 7, part = warp >> 3;
  const int stride = gridDim.x;

  if (tid == 0) {
    for (int c = 0; c < W11_NCH; ++c) mbar_init(bar + 8 * c, 1);
    fence_mbar_init();
  }
  for (int i = tid; i < (XS_BYTES + 2 * DN1S_BYTES) / 16; i += W11_THREADS) sts128(xs + i * 16, make_uint4(0, 0, 0, 0));
  __syncthreads();
  int b = blockIdx.x;
  if (b < batch) {
    if (tid == 0)                                 // x is an input of the step: stream it before the dependency wait
      for (int c = 0; c < W11_NCH; ++c) stg_issue_chunk<W11_NCH>(stg, x + (size_t)b * STATE_DIM, c, bar);
  }
  griddep_launch();
  griddep_wait();               // dn1 comes from conv12_bwd, which precedes this kernel
  if (b < batch) w11_prefetch_dn1(dn1, b, sbase + C11_OFF_DN1S, tid);
  cp_async_commit();

  float wacc[2][2][4] = {};   // m-tiles (kh, half = 0/1) x 2 n-tiles
  float bacc = 0.f;           // db11 partial: co = tid & 15, part = tid >> 4

  uint32_t phase = 0;
  int buf = 0;
  for (; b < batch; b += stride, buf ^= 1) {
    const uint32_t dn1s = sbase + C11_OFF_DN1S + buf * DN1S_BYTES;
    const bool more = b + stride < batch;
    cp_async_wait<0>();                       // dn1(b): visible to all after the first barrier below
#pragma unroll 1
    for (int c = 0; c < W11_NCH; ++c) {    // fp32 staging -> zero-bordered, chunk-swizzled bf16 image
      mbar_wait(bar + 8 * c, phase);
#pragma unroll 2
      for (int k = tid; k < W11_CHUNK_PIX; k += W11_THREADS) {
        const int i = c * W11_CHUNK_PIX + k;
        uint32_t r[4];
        lds128(r, stg + i * 16);
        const int y = i / IMG, px = i - y * IMG + 2;
        sts64(xs + (y + 2) * XS_ROW_BYTES + xs_chunk_off(px & ~1) + (px & 1) * 8,
              pack_bf16(__uint_as_float(r[0]), __uint_as_float(r[1])), pack_bf16(__uint_as_float(r[2]), __uint_as_float(r[3])));
      }
      __syncthreads();                        // chunk c is free (after the last one: image + dn1(b) complete)
      if (tid == 0 && more) {
        fence_proxy_async();
        stg_issue_chunk<W11_NCH>(stg, x + (size_t)(b + stride) * STATE_DIM, c, bar);
      }
    }
    phase ^= 1;
    if (more) w11_prefetch_dn1(dn1, b + stride, sbase + C11_OFF_DN1S + (buf ^ 1) * DN1S_BYTES, tid);
    cp_async_commit();

#pragma unroll 2
    for (int kk = 0; kk < 14; ++kk) {
      const int ks = part * 14 + kk;
      uint32_t bf[4];
      {
        const int pos = ks * 16 + (j & 1) * 8 + rr;                 // rows 441..447 are zero
        ldsm_x4_t(bf, dn1s + pos * 32 + ((((j >> 1) ^ (pos >> 2)) & 1) << 4));
      }
      const int pos = min(ks * 16 + (j >> 1) * 8 + rr, N1_POS - 1);
      const int oy = pos / H1, ox = pos - oy * H1;
      const uint32_t arow = xs + (4 * oy + kh) * XS_ROW_BYTES;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t af[4];
        ldsm_x4_t(af, arow + xs_chunk_off(4 * ox + 4 * half + 2 * (j & 1)));
        mma_bf16_16816(wacc[half][0], af, bf[0], bf[1]);
        mma_bf16_16816(wacc[half][1], af, bf[2], bf[3]);
      }
    }
    {
      const int co = tid & 15;
      for (int pos = tid >> 4; pos < N1_POS; pos += W11_THREADS / 16) {
        uint16_t v;
        asm volatile("ld.shared.u16 %0, [%1];\n" : "=h"(v)
                     : "r"(dn1s + pos * 32 + ((((co >> 3) ^ (pos >> 2)) & 1) << 4) + (co & 7) * 2));
        bacc += __uint_as_float((uint32_t)v << 16);
      }
    }
    __syncthreads();                          // the image is rewritten by the next iteration's convert
  }

  // m_local = kw_local*4 + c within tile (kh, kw = 4*half + kw_local); the two parts of a kh add up in global
#pragma unroll
  for (int half = 0; half < 2; ++half) {
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      float* o = g_w11 + ((kh * 8 + 4 * half) * 4) * C1_OUT + 8 * nt + 2 * t;
      atomicAdd(o + g * C1_OUT, wacc[half][nt][0]);
      atomicAdd(o + g * C1_OUT + 1, wacc[half][nt][1]);
      atomicAdd(o + (g + 8) * C1_OUT, wacc[half][nt][2]);
      atomicAdd(o + (g + 8) * C1_OUT + 1, wacc[half][nt][3]);
    }
  }
  red[tid] = bacc;
  __syncthreads();
  if (tid < C1_OUT) {
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < W11_THREADS / 16; ++p) s += red[p * 16 + tid];
    atomicAdd(g_b11 + tid, s);
  }
}

int configure_conv_bwd() {
  cudaError_t e = cudaFuncSetAttribute(conv12_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, B12_SMEM);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(conv11_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C11_SMEM);
  return (int)e;
}

int launch_conv12_bwd(const uint16_t* n1, const uint16_t* dn2, const float* w12, uint16_t* dn1, float* g_w12,
                      float* g_b12, int batch, int num_sms, cudaStream_t stream) {
  const int grid = min(batch, num_sms);
  return launch_pdl(conv12_bwd_kernel, dim3(grid), dim3(B12_THREADS), B12_SMEM, stream, n1, dn2, w12, dn1, g_w12, g_b12, batch);
}

int launch_conv11_wgrad(const float* x, const uint16_t* dn1, float* g_w11, float* g_b11, int batch, int num_sms,
                        cudaStream_t stream) {
  const int grid = min(batch, num_sms);
  return launch_pdl(conv11_wgrad_kernel, dim3(grid), dim3(W11_THREADS), C11_SMEM, stream, x, dn1, g_w11, g_b11, batch);
}

}  // namespace ga3c
