// Backward of the two convolutions (the autodiff half of opt.minimize, NetworkVP_discrate.py:130,
// for the layers defined at NetworkVP.py:212-228 / NetworkDNav.py:81-82).
//
//   conv12_bwd_kernel  per frame: dn1 = conv12 data-gradient of dn2 (masked by relu'(n1)),
//                      g_w12 += patches(n1)^T dn2, g_b12 += colsum(dn2)
//   conv11_wgrad_kernel per frame: g_w11 += patches(x)^T dn1, g_b11 += colsum(dn1)
//                      (conv11 has no data-gradient: x is the input)
//
// bf16 operands / fp32 accumulate (conv12: mma.sync tiles fed from shared memory, conv11: tcgen05 with TMEM
// accumulators); weight-gradient accumulators stay on chip across all frames a CTA processes and are stored once,
// into the CTA's own slab of the gradient-partial workspace.  grad_reduce (elementwise.cu) adds the slabs in a
// fixed order: contended atomics cost more than the whole frame loop (measured), and the sums become reproducible.
#include "common.cuh"
#include "kernels.h"
#include "tcgen05.cuh"
#include "conv_blk.cuh"

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
                  uint16_t* __restrict__ dn1, float* __restrict__ g_w12, float* __restrict__ g_b12, int64_t gp_stride,
                  int batch) {
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
  griddep_wait(K_CONV12_BWD); // dn2 comes from the dense1 data-gradient GEMM that precedes this kernel

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

  // this CTA's partial sums -> its slab (summed over CTAs by grad_reduce)
  g_w12 += (int64_t)blockIdx.x * gp_stride;
  g_b12 += (int64_t)blockIdx.x * gp_stride;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    float* o = g_w12 + (warp * C1_OUT) * C2_OUT + 8 * nt + 2 * t;
    *reinterpret_cast<float2*>(o + g * C2_OUT) = make_float2(wacc[nt][0], wacc[nt][1]);
    *reinterpret_cast<float2*>(o + (g + 8) * C2_OUT) = make_float2(wacc[nt][2], wacc[nt][3]);
  }
  red[tid] = bacc;
  __syncthreads();
  if (tid < C2_OUT) {
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < B12_WARPS; ++p) s += red[p * 32 + tid];
    g_b12[tid] = s;
  }
  trace_mark(K_CONV12_BWD, 2);
}

// ------------------------------------------------------------------------------------------------
// conv11 weight gradient on tcgen05.  With the space-to-depth block matrix Blk (conv_blk.cuh) the gradient of
// quadrant (a, b) of the 8x8 kernel is one long-K GEMM over output positions m = oy*22 + ox:
//     dW_ab[64 (dy, dx, c), 16 cout] = sum_m Blk[m + 22a + b, :]^T . dn1[m, :]
// A = Blk read MN-major (transposed) at a row-shifted start address, B = dn1 staged MN-major (two 8-channel planes,
// dead column and tail rows zero), M = 64, N = 16, K = 464 positions = 29 UMMAs per quadrant and frame.  The four
// accumulators live in TMEM for the whole kernel (all frames of the CTA) and are flushed once with atomics.
//   warps 0-5   fp32 chunk (TMA ring) -> bf16 -> Blk (as in conv_fwd)
//   warp  6     one thread issues the UMMAs group by group as the Blk rows land
//   warps 8-11  dn1 frame -> shared memory (cp.async, double buffered), bias gradient; final TMEM flush
constexpr int W11_THREADS = 384, W11_AUX_WARPS = 6, W11_AUX_THREADS = 32 * W11_AUX_WARPS, W11_ISSUE_WARP = 6, W11_EPI_WARP0 = 8;
constexpr int W11_KSTEPS = 29;                                   // 464 >= 462 positions (21 rows x 22, column 21 dead)
constexpr int DN1_PLANE = 512 * 16, DN1_BUF = 2 * DN1_PLANE;     // 8 channels x 512 rows per plane, 2 planes per frame
constexpr int W11_OFF_BLK = 0;
constexpr int W11_OFF_RING = W11_OFF_BLK + BLK_BYTES;            //  70,144
constexpr int W11_OFF_DN1 = W11_OFF_RING + CF_NSLOT * CH_BYTES;  // 134,656 (two buffers)
constexpr int W11_OFF_RED = W11_OFF_DN1 + 2 * DN1_BUF;           // 167,424
constexpr int W11_OFF_BAR = W11_OFF_RED + 128 * 4;               // 167,936
constexpr int WB_RING = 0;        // [4] TMA chunk landed
constexpr int WB_BLKRDY = 4;      // [4] Blk rows of position group i converted          (aux -> issuer)
constexpr int WB_GRP = 8;         // [4] UMMAs of group i retired (tcgen05.commit)       (-> aux: Blk rows free)
constexpr int WB_DN1RDY = 12;     // [2] dn1 buffer staged                                (stager -> issuer)
constexpr int WB_DN1FREE = 14;    // [2] every UMMA reading the dn1 buffer retired        (-> stager)
constexpr int WB_DONE = 16;       //     every UMMA of the kernel retired                   (-> final flush)
constexpr int W11_NBAR = 17;
constexpr int W11_OFF_TSLOT = W11_OFF_BAR + W11_NBAR * 8;
constexpr int C11_SMEM = W11_OFF_TSLOT + 16 + 128;

__global__ void __launch_bounds__(W11_THREADS, 1)
conv11_wgrad_kernel(const float* __restrict__ x, const uint16_t* __restrict__ dn1, float* __restrict__ g_w11,
                    float* __restrict__ g_b11, int64_t gp_stride, int batch) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* const smem = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t blk = sbase + W11_OFF_BLK, ring = sbase + W11_OFF_RING, dn1s = sbase + W11_OFF_DN1, bars = sbase + W11_OFF_BAR,
                 tslot = sbase + W11_OFF_TSLOT;
  float* red = reinterpret_cast<float*>(smem + W11_OFF_RED);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int stride = gridDim.x;
  const int n_frames = ((int)blockIdx.x < batch) ? (batch - 1 - (int)blockIdx.x) / stride + 1 : 0;
  const int n_chunks = n_frames * CF_NCHUNK;
  auto frame_of = [&](int k) { return (size_t)(blockIdx.x + k * stride); };
  auto bar = [&](int i) { return bars + i * 8; };
  auto issue_chunk = [&](int q) {                                  // one thread
    const int k = q / CF_NCHUNK, c = q - k * CF_NCHUNK, slot = q % CF_NSLOT;
    mbar_expect_tx(bar(WB_RING + slot), CH_BYTES);
    bulk_load(ring + slot * CH_BYTES, reinterpret_cast<const uint8_t*>(x + frame_of(k) * STATE_DIM) + c * CH_BYTES, CH_BYTES,
              bar(WB_RING + slot));
  };

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar(WB_RING + i), 1);
      mbar_init(bar(WB_BLKRDY + i), 1);
      mbar_init(bar(WB_GRP + i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(WB_DN1RDY + i), 1);
      mbar_init(bar(WB_DN1FREE + i), 1);
    }
    mbar_init(bar(WB_DONE), 1);
    fence_mbar_init();
  }
  if (warp == W11_EPI_WARP0) tmem_alloc<64>(tslot);
  __syncthreads();
  if (tid == 0)                                                    // x is an input of the step: stream it before the dependency wait
    for (int q = 0; q < CF_NSLOT && q < n_chunks; ++q) issue_chunk(q);
  // zero once: image borders / slack rows of Blk; dead-column and tail rows of both dn1 buffers
  for (int i = tid; i < BLK_BYTES / 16; i += W11_THREADS) sts128(blk + i * 16, make_uint4(0, 0, 0, 0));
  for (int i = tid; i < 2 * DN1_BUF / 16; i += W11_THREADS) sts128(dn1s + i * 16, make_uint4(0, 0, 0, 0));
  griddep_launch();
  griddep_wait(K_CONV11_WGRAD); // dn1 comes from conv12_bwd, which precedes this kernel
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(tmem_base) : "r"(tslot));

  if (warp < W11_AUX_WARPS) {
    // =========================== aux: fp32 chunk -> bf16 block matrix, slot re-arm ===========================
    uint32_t lane_off[3];
    blk_lane_offsets(lane, lane_off);
    for (int k = 0; k < n_frames; ++k) {
#pragma unroll 1
      for (int c = 0; c < CF_NCHUNK; ++c) {
        const int q = k * CF_NCHUNK + c, slot = q % CF_NSLOT;
        // block rows 3c..3c+3 are rewritten: the last position group of frame k-1 that reads them must have retired
        if (k > 0) mbar_wait(bar(WB_GRP + (c + 1) / 2), (k - 1) & 1);
        mbar_wait(bar(WB_RING + slot), (q / CF_NSLOT) & 1);
        blk_convert_chunk<W11_AUX_WARPS>(ring + slot * CH_BYTES, blk, c, warp, lane, lane_off);
        fence_proxy_async();
        named_bar_sync(1, W11_AUX_THREADS);
        if (tid == 0) {
          if (q + CF_NSLOT < n_chunks) issue_chunk(q + CF_NSLOT);
          // position group i (m in [128 i, 128 i + 128)) reads block rows up to (128 i + 150) / 22
          if (c == 2) mbar_arrive(bar(WB_BLKRDY + 0));
          if (c == 4) mbar_arrive(bar(WB_BLKRDY + 1));
          if (c == 6) { mbar_arrive(bar(WB_BLKRDY + 2)); mbar_arrive(bar(WB_BLKRDY + 3)); }
        }
      }
    }
  } else if (warp == W11_ISSUE_WARP) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_m(64, C1_OUT, true, true);
      for (int k = 0; k < n_frames; ++k) {
        const uint32_t dbuf = dn1s + (k & 1) * DN1_BUF;
        mbar_wait(bar(WB_DN1RDY + (k & 1)), (k >> 1) & 1);
        for (int gi = 0; gi < 4; ++gi) {
          mbar_wait(bar(WB_BLKRDY + gi), k & 1);
          tc_fence_after();
          const int s1 = gi == 3 ? W11_KSTEPS : 8 * gi + 8;
          for (int s = 8 * gi; s < s1; ++s) {
            // B: dn1 rows [16 s, 16 s + 16): MN-major, 8-channel planes at SBO = DN1_PLANE, 8-row k groups at LBO = 128
            const uint64_t db = make_desc_ns(dbuf + s * 256, 128, DN1_PLANE);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              // A: Blk rows 16 s + 22 a + b ..: MN-major, k-chunks (8 of the 64 block elements) at SBO = BLK_LBO
              tc_mma_bf16(tmem_base + 16 * q, make_desc_ns(blk + (16 * s + BLK_W * (q >> 1) + (q & 1)) * 16, 128, BLK_LBO), db,
                          idesc, (k | s) ? 1u : 0u);
          }
          tc_commit(bar(WB_GRP + gi));
        }
        tc_commit(bar(WB_DN1FREE + (k & 1)));
      }
      tc_commit(bar(WB_DONE));
    }
  } else if (warp >= W11_EPI_WARP0) {
    // =========================== dn1 staging + bias gradient; final flush ===========================
    const int etid = tid - 32 * W11_EPI_WARP0, ew = warp - W11_EPI_WARP0;
    float bacc = 0.f;                                              // db11 partial: channel etid & 15, row phase etid >> 4
    for (int k = 0; k < n_frames; ++k) {
      const uint32_t dbuf = dn1s + (k & 1) * DN1_BUF;
      if (k >= 2) mbar_wait(bar(WB_DN1FREE + (k & 1)), ((k >> 1) - 1) & 1);   // frame k-2 (same buffer) has been consumed
      const uint4* src = reinterpret_cast<const uint4*>(dn1 + frame_of(k) * (N1_POS * C1_OUT));
      for (int i = etid; i < N1_POS * 2; i += 128) {
        const int p = i >> 1, oy = p / H1, ox = p - oy * H1;
        cp_async16(dbuf + (i & 1) * DN1_PLANE + (oy * BLK_W + ox) * 16, src + i, 16);
      }
      cp_async_commit();
      cp_async_wait<0>();
      fence_proxy_async();
      named_bar_sync(2, 128);
      if (etid == 0) mbar_arrive(bar(WB_DN1RDY + (k & 1)));
      {
        const int co = etid & 15;
        for (int m = etid >> 4; m < H1 * BLK_W; m += 8) {          // dead-column rows are zero
          uint16_t v;
          asm volatile("ld.shared.u16 %0, [%1];\n" : "=h"(v) : "r"(dbuf + (co >> 3) * DN1_PLANE + m * 16 + (co & 7) * 2));
          bacc += __uint_as_float((uint32_t)v << 16);
        }
      }
    }
    red[etid] = bacc;
    named_bar_sync(2, 128);
    if (etid < C1_OUT) {
      float sum = 0.f;
#pragma unroll
      for (int ph = 0; ph < 8; ++ph) sum += red[ph * 16 + etid];
      g_b11[(int64_t)blockIdx.x * gp_stride + etid] = sum;
    }
    // flush the four accumulators: M = 64 rows sit on TMEM lanes 32 w + (0..15) (rows 16 w ..), 16 columns per quadrant
    if (n_frames > 0) {
      mbar_wait(bar(WB_DONE), 0);
      tc_fence_after();
      float* const slab = g_w11 + (int64_t)blockIdx.x * gp_stride;  // this CTA's partial sums (summed by grad_reduce)
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(tmem_base + ((uint32_t)(ew * 32) << 16) + 16 * q));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        if (lane < 16) {
          // row = block element j*8 + e: j = dy*2 + (dx>>1), e = (dx&1)*4 + c
          const int row = 16 * ew + lane, j = row >> 3, e = row & 7;
          const int kh = 4 * (q >> 1) + (j >> 1), kw = 4 * (q & 1) + (j & 1) * 2 + (e >> 2), c = e & 3;
          float4* o = reinterpret_cast<float4*>(slab + ((kh * 8 + kw) * 4 + c) * C1_OUT);
#pragma unroll
          for (int n = 0; n < C1_OUT / 4; ++n)
            o[n] = make_float4(__uint_as_float(r[4 * n]), __uint_as_float(r[4 * n + 1]), __uint_as_float(r[4 * n + 2]),
                               __uint_as_float(r[4 * n + 3]));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  trace_mark(K_CONV11_WGRAD, 2);
  if (warp == W11_EPI_WARP0) {
    tc_fence_after();
    tmem_dealloc<64>(tmem_base);
  }
}

GA3C_TRACE_ATTACH(trace_attach_conv_bwd)

int configure_conv_bwd() {
  cudaError_t e = cudaFuncSetAttribute(conv12_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, B12_SMEM);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(conv11_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C11_SMEM);
  return (int)e;
}

int conv_bwd_grid(int batch, int num_sms) { return min(batch, num_sms); }

int launch_conv12_bwd(const uint16_t* n1, const uint16_t* dn2, const float* w12, uint16_t* dn1, float* g_w12,
                      float* g_b12, int64_t gp_stride, int batch, int num_sms, cudaStream_t stream) {
  return launch_pdl(conv12_bwd_kernel, dim3(conv_bwd_grid(batch, num_sms)), dim3(B12_THREADS), B12_SMEM, stream, n1, dn2,
                    w12, dn1, g_w12, g_b12, gp_stride, batch);
}

int launch_conv11_wgrad(const float* x, const uint16_t* dn1, float* g_w11, float* g_b11, int64_t gp_stride, int batch,
                        int num_sms, cudaStream_t stream) {
  return launch_pdl(conv11_wgrad_kernel, dim3(conv_bwd_grid(batch, num_sms)), dim3(W11_THREADS), C11_SMEM, stream, x, dn1,
                    g_w11, g_b11, gp_stride, batch);
}

}  // namespace ga3c
