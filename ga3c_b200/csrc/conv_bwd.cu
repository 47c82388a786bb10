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

constexpr int CB_THREADS = 256;

// padded dn2 in smem: 12x12 positions (index (oy+1, ox+1); row/col 0 are zero), 64 B per position,
// 16-B chunk c stored at c ^ f(pp) so that 8 consecutive positions x one chunk are conflict-free
constexpr int DN2P_W = 12, DN2P_BYTES = DN2P_W * DN2P_W * 64;     // 9216
__device__ __forceinline__ int dn2p_off(int pp, int chunk) {
  const int f = (((pp >> 1) & 1) << 1) | ((pp >> 2) & 1);
  return (pp << 6) + ((chunk ^ f) << 4);
}

// ------------------------------------------------------------------------------------------------
// inputs are double-buffered: frame i+1 streams in (cp.async) while frame i is computed
constexpr int C12_IN_BYTES = N1P_BYTES + DN2P_BYTES;              // 27648 per buffer
constexpr int C12_OFF_IN = 0;
constexpr int C12_OFF_WDF = C12_OFF_IN + 2 * C12_IN_BYTES;        // 55296
constexpr int C12_OFF_DN1S = C12_OFF_WDF + 4 * 8 * 32 * 16;       // 71680
constexpr int C12_OFF_RED = C12_OFF_DN1S + N1_POS * 32;           // 85792
constexpr int C12_SMEM = C12_OFF_RED + CB_THREADS * 4;            // 86816  (2 CTAs / SM)

__device__ __forceinline__ void c12_prefetch(const uint16_t* __restrict__ n1, const uint16_t* __restrict__ dn2, int b,
                                             uint32_t n1p, uint32_t dn2p, int tid) {
  const uint4* s1 = reinterpret_cast<const uint4*>(n1 + (size_t)b * N1_POS * C1_OUT);
  for (int i = tid; i < N1_POS * 2; i += CB_THREADS) {
    const int pos = i >> 1, oy = pos / H1, ox = pos - oy * H1;
    cp_async16(n1p + n1p_off(oy + 1, ox + 1, i & 1), s1 + i, 16);
  }
  const uint4* s2 = reinterpret_cast<const uint4*>(dn2 + (size_t)b * FLAT);
  for (int i = tid; i < N2_POS * 4; i += CB_THREADS) {
    const int pos = i >> 2, oy = pos / H2, ox = pos - oy * H2;
    cp_async16(dn2p + dn2p_off((oy + 1) * DN2P_W + ox + 1, i & 3), s2 + i, 16);
  }
}

__global__ void __launch_bounds__(CB_THREADS, 2)
conv12_bwd_kernel(const uint16_t* __restrict__ n1, const uint16_t* __restrict__ dn2, const float* __restrict__ w12,
                  uint16_t* __restrict__ dn1, float* __restrict__ g_w12, float* __restrict__ g_b12, int batch) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t wdf = sbase + C12_OFF_WDF, dn1s = sbase + C12_OFF_DN1S;
  float* red = reinterpret_cast<float*>(smem + C12_OFF_RED);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int j = lane >> 3, rr = lane & 7;

  for (int i = tid; i < 2 * C12_IN_BYTES / 16; i += CB_THREADS) sts128(sbase + i * 16, make_uint4(0, 0, 0, 0));
  // data-gradient weights in B-fragment order: [class = par_y*2+par_x][kstep = (a*2+b)*2+half][lane]
  //   K = (a, b, co): taps kh = par_y + 2a, kw = par_x + 2b ; N = ci
  for (int i = tid; i < 4 * 8 * 32; i += CB_THREADS) {
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
  __syncthreads();

  float wacc[2][4][4] = {};   // wgrad: taps 2*warp, 2*warp+1 ; 4 n-tiles
  float bacc = 0.f;           // db2 partial: co = tid & 31, part = tid >> 5

  // the zero fill above must land before the first cp.async overwrites the interiors (same threads do
  // not own the same bytes), so the prefetch of the first frame follows the setup barrier
  if ((int)blockIdx.x < batch) c12_prefetch(n1, dn2, blockIdx.x, sbase + C12_OFF_IN, sbase + C12_OFF_IN + N1P_BYTES, tid);
  cp_async_commit();
  int buf = 0;
  for (int b = blockIdx.x; b < batch; b += gridDim.x, buf ^= 1) {
    const uint32_t n1p = sbase + C12_OFF_IN + buf * C12_IN_BYTES, dn2p = n1p + N1P_BYTES;
    cp_async_wait<0>();
    __syncthreads();          // frame b is visible to every warp; the other buffer and dn1s are free (barrier at loop end)
    if (b + (int)gridDim.x < batch)
      c12_prefetch(n1, dn2, b + gridDim.x, sbase + C12_OFF_IN + (buf ^ 1) * C12_IN_BYTES,
                   sbase + C12_OFF_IN + (buf ^ 1) * C12_IN_BYTES + N1P_BYTES, tid);
    cp_async_commit();

    // ---------------- data gradient: 29 m16 tiles over 4 parity classes ----------------
    for (int tile = warp; tile < 29; tile += 8) {
      const int cls = tile < 7 ? 0 : tile < 14 ? 1 : tile < 21 ? 2 : 3;
      const int lt = tile - cls * 7;
      const int pary = cls >> 1, parx = cls & 1;
      const int ny = 10 + pary, nx = 10 + parx, cnt = ny * nx;
      const int r0 = lt * 16 + g, r1 = r0 + 8;
      const int c0 = min(r0, cnt - 1), c1 = min(r1, cnt - 1);
      const int iy0 = c0 / nx, ix0 = c0 - iy0 * nx, iy1 = c1 / nx, ix1 = c1 - iy1 * nx;
      // pixel y = 2*i + 1 - par ; q = i + 1 - par ; tap a -> oy = q - a ; padded row = oy + 1
      const int qy0 = iy0 + 2 - pary, qx0 = ix0 + 2 - parx, qy1 = iy1 + 2 - pary, qx1 = ix1 + 2 - parx;
      float acc[2][4] = {};
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const int a_ = ks >> 2, b_ = (ks >> 1) & 1, half = ks & 1;
        const int pp0 = (qy0 - a_) * DN2P_W + (qx0 - b_), pp1 = (qy1 - a_) * DN2P_W + (qx1 - b_);
        uint32_t a[4], bw[4];
        lds64(a[0], a[2], dn2p + dn2p_off(pp0, 2 * half + (t >> 1)) + (t & 1) * 8);
        lds64(a[1], a[3], dn2p + dn2p_off(pp1, 2 * half + (t >> 1)) + (t & 1) * 8);
        lds128(bw, wdf + ((cls * 8 + ks) * 32 + lane) * 16);
        mma_bf16_16816(acc[0], a, bw[0], bw[1]);
        mma_bf16_16816(acc[1], a, bw[2], bw[3]);
      }
      const int y0 = 2 * iy0 + 1 - pary, x0 = 2 * ix0 + 1 - parx, y1 = 2 * iy1 + 1 - pary, x1 = 2 * ix1 + 1 - parx;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        if (r0 < cnt) {
          uint32_t m;
          asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(m) : "r"(n1p + n1p_off(y0 + 1, x0 + 1, nt) + 4 * t));
          sts32(dn1s + (y0 * H1 + x0) * 32 + nt * 16 + 4 * t,
                pack_bf16((m & 0x7FFFu) ? acc[nt][0] : 0.f, (m & 0x7FFF0000u) ? acc[nt][1] : 0.f));
        }
        if (r1 < cnt) {
          uint32_t m;
          asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(m) : "r"(n1p + n1p_off(y1 + 1, x1 + 1, nt) + 4 * t));
          sts32(dn1s + (y1 * H1 + x1) * 32 + nt * 16 + 4 * t,
                pack_bf16((m & 0x7FFFu) ? acc[nt][2] : 0.f, (m & 0x7FFF0000u) ? acc[nt][3] : 0.f));
        }
      }
    }

    // ---------------- weight gradient: M = 256 (tap, ci), N = 32, K = 121 positions ----------------
#pragma unroll 2
    for (int ks = 0; ks < 8; ++ks) {
      uint32_t bf[2][4];
      {
        const int pos = ks * 16 + (j & 1) * 8 + rr;
        int pp = 0;
        if (pos < N2_POS) { const int oy = pos / H2, ox = pos - oy * H2; pp = (oy + 1) * DN2P_W + ox + 1; }
        ldsm_x4_t(bf[0], dn2p + dn2p_off(pp, (j >> 1)));
        ldsm_x4_t(bf[1], dn2p + dn2p_off(pp, 2 + (j >> 1)));
      }
      const int pos = ks * 16 + (j >> 1) * 8 + rr;
      int oy = 0, ox = 0;
      const bool ok = pos < N2_POS;
      if (ok) { oy = pos / H2; ox = pos - oy * H2; }
#pragma unroll
      for (int ti = 0; ti < 2; ++ti) {
        const int tap = 2 * warp + ti, kh = tap >> 2, kw = tap & 3;
        uint32_t af[4];
        const int py = ok ? 2 * oy + kh : N1P_W - 1, px = ok ? 2 * ox + kw : N1P_W - 1;
        ldsm_x4_t(af, n1p + n1p_off(py, px, j & 1));
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
          mma_bf16_16816(wacc[ti][nt], af, bf[nt >> 1][(nt & 1) * 2], bf[nt >> 1][(nt & 1) * 2 + 1]);
      }
    }
    // bias gradient partials
    {
      const int co = tid & 31;
      for (int pos = tid >> 5; pos < N2_POS; pos += 8) {
        const int oy = pos / H2, ox = pos - oy * H2;
        uint16_t v;
        asm volatile("ld.shared.u16 %0, [%1];\n" : "=h"(v)
                     : "r"(dn2p + dn2p_off((oy + 1) * DN2P_W + ox + 1, co >> 3) + (co & 7) * 2));
        bacc += __uint_as_float((uint32_t)v << 16);
      }
    }
    __syncthreads();
    {
      uint4* dst = reinterpret_cast<uint4*>(dn1 + (size_t)b * N1_POS * C1_OUT);
      for (int i = tid; i < N1_POS * 2; i += CB_THREADS) {
        uint32_t r[4];
        lds128(r, dn1s + i * 16);
        dst[i] = make_uint4(r[0], r[1], r[2], r[3]);
      }
    }
    __syncthreads();   // dn1s / n1p / dn2p are rewritten by the next iteration
  }

#pragma unroll
  for (int ti = 0; ti < 2; ++ti) {
    const int tap = 2 * warp + ti;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      float* o = g_w12 + (tap * C1_OUT) * C2_OUT + 8 * nt + 2 * t;
      atomicAdd(o + g * C2_OUT, wacc[ti][nt][0]);
      atomicAdd(o + g * C2_OUT + 1, wacc[ti][nt][1]);
      atomicAdd(o + (g + 8) * C2_OUT, wacc[ti][nt][2]);
      atomicAdd(o + (g + 8) * C2_OUT + 1, wacc[ti][nt][3]);
    }
  }
  red[tid] = bacc;
  __syncthreads();
  if (tid < C2_OUT) {
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < 8; ++p) s += red[p * 32 + tid];
    atomicAdd(g_b12 + tid, s);
  }
}

// ------------------------------------------------------------------------------------------------
// conv11 weight gradient.  One persistent CTA per SM, 16 warps; per frame
//   TMA engine : cp.async.bulk of frame i+1 into the fp32 staging buffer (mbarrier), cp.async of dn1(i+1)
//   all warps  : wait -> fp32 staging -> padded bf16 image -> [issue frame i+1] -> wgrad MMAs
// warp = (kh, part): M tile pair (kh, half 0/1) x N 16, K = the 14 k16 position steps of its part.
constexpr int W11_THREADS = 512;
constexpr int W11_FRAME_BYTES = STATE_DIM * 4;                    // 112,896
constexpr int W11_CHUNKS = 8, W11_CHUNK_BYTES = W11_FRAME_BYTES / W11_CHUNKS;
constexpr int DN1S_ROWS = 448;                                    // 441 padded to 28 k16 steps
constexpr int DN1S_BYTES = DN1S_ROWS * 32;                        // 14,336
constexpr int C11_OFF_STG = 0;
constexpr int C11_OFF_XS = C11_OFF_STG + W11_FRAME_BYTES;         // 112,896
constexpr int C11_OFF_DN1S = C11_OFF_XS + XS_BYTES;               // 174,848 (two buffers)
constexpr int C11_OFF_RED = C11_OFF_DN1S + 2 * DN1S_BYTES;        // 203,520
constexpr int C11_OFF_BAR = C11_OFF_RED + W11_THREADS * 4;        // 205,568
constexpr int C11_SMEM = C11_OFF_BAR + 16;                        // 205,584

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

__device__ __forceinline__ void w11_issue_frame(uint32_t stg, const float* src, uint32_t bar) {
  mbar_expect_tx(bar, W11_FRAME_BYTES);
#pragma unroll
  for (int c = 0; c < W11_CHUNKS; ++c)
    bulk_load(stg + c * W11_CHUNK_BYTES, reinterpret_cast<const uint8_t*>(src) + c * W11_CHUNK_BYTES, W11_CHUNK_BYTES, bar);
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
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  for (int i = tid; i < (XS_BYTES + 2 * DN1S_BYTES) / 16; i += W11_THREADS) sts128(xs + i * 16, make_uint4(0, 0, 0, 0));
  __syncthreads();
  int b = blockIdx.x;
  if (b < batch) {
    if (tid == 0) w11_issue_frame(stg, x + (size_t)b * STATE_DIM, bar);
    w11_prefetch_dn1(dn1, b, sbase + C11_OFF_DN1S, tid);
  }
  cp_async_commit();

  float wacc[2][2][4] = {};   // m-tiles (kh, half = 0/1) x 2 n-tiles
  float bacc = 0.f;           // db11 partial: co = tid & 15, part = tid >> 4

  uint32_t phase = 0;
  int buf = 0;
  for (; b < batch; b += stride, buf ^= 1) {
    const uint32_t dn1s = sbase + C11_OFF_DN1S + buf * DN1S_BYTES;
    mbar_wait(bar, phase);                    // frame b has landed in the staging buffer
    phase ^= 1;
    {                                         // fp32 staging -> zero-bordered, chunk-swizzled bf16 image
      constexpr int NPIX = IMG * IMG;
#pragma unroll 2
      for (int i = tid; i < NPIX; i += W11_THREADS) {
        uint32_t r[4];
        lds128(r, stg + i * 16);
        const int y = i / IMG, px = i - y * IMG + 2;
        sts64(xs + (y + 2) * XS_ROW_BYTES + xs_chunk_off(px & ~1) + (px & 1) * 8,
              pack_bf16(__uint_as_float(r[0]), __uint_as_float(r[1])), pack_bf16(__uint_as_float(r[2]), __uint_as_float(r[3])));
      }
    }
    cp_async_wait<0>();                       // dn1(b)
    __syncthreads();                          // staging and the other dn1 buffer are free; image + dn1(b) complete
    if (b + stride < batch) {
      if (tid == 0) {
        fence_proxy_async();
        w11_issue_frame(stg, x + (size_t)(b + stride) * STATE_DIM, bar);
      }
      w11_prefetch_dn1(dn1, b + stride, sbase + C11_OFF_DN1S + (buf ^ 1) * DN1S_BYTES, tid);
    }
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
  cudaError_t e = cudaFuncSetAttribute(conv12_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C12_SMEM);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(conv11_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C11_SMEM);
  return (int)e;
}

int launch_conv12_bwd(const uint16_t* n1, const uint16_t* dn2, const float* w12, uint16_t* dn1, float* g_w12,
                      float* g_b12, int batch, int num_sms, cudaStream_t stream) {
  const int grid = min(batch, 2 * num_sms);
  conv12_bwd_kernel<<<grid, CB_THREADS, C12_SMEM, stream>>>(n1, dn2, w12, dn1, g_w12, g_b12, batch);
  return (int)cudaGetLastError();
}

int launch_conv11_wgrad(const float* x, const uint16_t* dn1, float* g_w11, float* g_b11, int batch, int num_sms,
                        cudaStream_t stream) {
  const int grid = min(batch, num_sms);
  conv11_wgrad_kernel<<<grid, W11_THREADS, C11_SMEM, stream>>>(x, dn1, g_w11, g_b11, batch);
  return (int)cudaGetLastError();
}

}  // namespace ga3c
