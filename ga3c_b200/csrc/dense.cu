// dense1 GEMMs: forward, data-gradient, weight-gradient  (NetworkDNav.py:90, dense_layer :256-269).
// bf16 operands, fp32 accumulate.  64x64x32 CTA tile, 4 warps (2x2), 4-stage cp.async pipeline,
// ldmatrix fragment loads from XOR-swizzled shared memory (conflict-free).
//
//   fwd  : C[M=B,   N=256 ] = A[M,K=3872] (n2, K-major)   x  B[K,N] (w1, N-major)   + bias, relu -> fp32
//   dgrad: C[M=B,   N=3872] = A[M,K=256 ] (dd1, K-major)  x  B[N,K] (w1, K-major)   masked by n2>0 -> bf16
//   wgrad: C[M=3872,N=256 ] = A[K=B,M]    (n2, M-major)   x  B[K,N] (dd1, N-major)  -> fp32
#include "common.cuh"
#include "kernels.h"

namespace ga3c {

constexpr int GB_M = 64, GB_N = 64, GB_K = 32, G_STAGES = 4, G_THREADS = 128;
constexpr int G_TILE_BYTES = 64 * 32 * 2;                 // 4096 for either operand
constexpr int G_SMEM = G_STAGES * 2 * G_TILE_BYTES;       // 32768

enum { LAY_KMAJOR = 0, LAY_MNMAJOR = 1 };

// Loads one operand tile.  `mn0` is the tile origin along M (or N), `k0` along K.
//   KMAJOR : gmem [MN][K], smem [64 mn][32 k], 64-B rows, 4 chunks, chunk ^= (row>>1)&3
//   MNMAJOR: gmem [K][MN], smem [32 k][64 mn], 128-B rows, 8 chunks, chunk ^= row&7
template <int LAY>
__device__ __forceinline__ void load_tile(uint32_t sdst, const uint16_t* __restrict__ src, int ld, int mn0, int k0,
                                          int MN, int K, int tid) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = tid + i * G_THREADS;
    if (LAY == LAY_KMAJOR) {
      const int row = c >> 2, ch = c & 3;
      const int mn = mn0 + row, k = k0 + ch * 8;
      const bool ok = (mn < MN) && (k < K);
      const uint16_t* p = ok ? src + (size_t)mn * ld + k : src;
      cp_async16(sdst + row * 64 + ((ch ^ ((row >> 1) & 3)) << 4), p, ok ? 16 : 0);
    } else {
      const int row = c >> 3, ch = c & 7;
      const int k = k0 + row, mn = mn0 + ch * 8;
      const bool ok = (k < K) && (mn < MN);
      const uint16_t* p = ok ? src + (size_t)k * ld + mn : src;
      cp_async16(sdst + row * 128 + ((ch ^ (row & 7)) << 4), p, ok ? 16 : 0);
    }
  }
}

struct EpiBiasReluF32 {
  float* out; const float* bias; int ldc;
  __device__ __forceinline__ void operator()(int m, int n, float v0, float v1) const {
    float2 r = make_float2(fmaxf(v0 + bias[n], 0.f), fmaxf(v1 + bias[n + 1], 0.f));
    *reinterpret_cast<float2*>(out + (size_t)m * ldc + n) = r;
  }
};
struct EpiStoreF32 {
  float* out; int ldc;
  __device__ __forceinline__ void operator()(int m, int n, float v0, float v1) const {
    *reinterpret_cast<float2*>(out + (size_t)m * ldc + n) = make_float2(v0, v1);
  }
};
struct EpiReluMaskBf16 {
  uint16_t* out; const uint16_t* act; int ldc;
  __device__ __forceinline__ void operator()(int m, int n, float v0, float v1) const {
    const uint32_t a = *reinterpret_cast<const uint32_t*>(act + (size_t)m * ldc + n);
    const float r0 = (a & 0x7FFFu) != 0 && !(a & 0x8000u) ? v0 : 0.f;                 // act > 0 (post-ReLU: never negative)
    const float r1 = ((a >> 16) & 0x7FFFu) != 0 && !(a & 0x80000000u) ? v1 : 0.f;
    *reinterpret_cast<uint32_t*>(out + (size_t)m * ldc + n) = pack_bf16(r0, r1);
  }
};

template <int ALAY, int BLAY, class Epi>
__global__ void __launch_bounds__(G_THREADS)
gemm_bf16_kernel(const uint16_t* __restrict__ A, int lda, const uint16_t* __restrict__ Bm, int ldb, int M, int N, int K,
                 Epi epi) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sA = smem_u32(smem), sB = sA + G_STAGES * G_TILE_BYTES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp >> 1, wn = warp & 1;
  const int m0 = blockIdx.y * GB_M, n0 = blockIdx.x * GB_N;
  const int KT = (K + GB_K - 1) / GB_K;

  float acc[2][4][4] = {};

#pragma unroll
  for (int s = 0; s < G_STAGES - 1; ++s) {
    if (s < KT) {
      load_tile<ALAY>(sA + s * G_TILE_BYTES, A, lda, m0, s * GB_K, M, K, tid);
      load_tile<BLAY>(sB + s * G_TILE_BYTES, Bm, ldb, n0, s * GB_K, N, K, tid);
    }
    cp_async_commit();
  }

  const int j = lane >> 3, rr = lane & 7;
  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<G_STAGES - 2>();
    __syncthreads();
    {
      const int nx = kt + G_STAGES - 1;
      if (nx < KT) {
        const int s = nx % G_STAGES;
        load_tile<ALAY>(sA + s * G_TILE_BYTES, A, lda, m0, nx * GB_K, M, K, tid);
        load_tile<BLAY>(sB + s * G_TILE_BYTES, Bm, ldb, n0, nx * GB_K, N, K, tid);
      }
      cp_async_commit();
    }
    const uint32_t a_s = sA + (kt % G_STAGES) * G_TILE_BYTES, b_s = sB + (kt % G_STAGES) * G_TILE_BYTES;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      uint32_t af[2][4], bf[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        if (ALAY == LAY_KMAJOR) {
          const int row = wm * 32 + mt * 16 + (j & 1) * 8 + rr, ch = kk * 2 + (j >> 1);
          ldsm_x4(af[mt], a_s + row * 64 + ((ch ^ ((row >> 1) & 3)) << 4));
        } else {
          const int krow = kk * 16 + (j >> 1) * 8 + rr, ch = (wm * 32 + mt * 16) / 8 + (j & 1);
          ldsm_x4_t(af[mt], a_s + krow * 128 + ((ch ^ (krow & 7)) << 4));
        }
      }
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        if (BLAY == LAY_MNMAJOR) {
          const int krow = kk * 16 + (j & 1) * 8 + rr, ch = (wn * 32 + np * 16) / 8 + (j >> 1);
          ldsm_x4_t(bf[np], b_s + krow * 128 + ((ch ^ (krow & 7)) << 4));
        } else {
          const int row = wn * 32 + np * 16 + (j >> 1) * 8 + rr, ch = kk * 2 + (j & 1);
          ldsm_x4(bf[np], b_s + row * 64 + ((ch ^ ((row >> 1) & 3)) << 4));
        }
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
          mma_bf16_16816(acc[mt][nt], af[mt], bf[nt >> 1][(nt & 1) * 2], bf[nt >> 1][(nt & 1) * 2 + 1]);
    }
  }
  cp_async_wait<0>();

  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int n = n0 + wn * 32 + nt * 8 + 2 * t;
      const int ma = m0 + wm * 32 + mt * 16 + g, mb = ma + 8;
      if (n < N) {
        if (ma < M) epi(ma, n, acc[mt][nt][0], acc[mt][nt][1]);
        if (mb < M) epi(mb, n, acc[mt][nt][2], acc[mt][nt][3]);
      }
    }
  }
}

int configure_dense() { return 0; }   // 32 KB dynamic smem: below the 48 KB opt-in threshold

int launch_dense_fwd(const uint16_t* n2, const uint16_t* w1bf, const float* b1, float* d1, int batch, cudaStream_t stream) {
  dim3 grid(FC / GB_N, (batch + GB_M - 1) / GB_M);
  gemm_bf16_kernel<LAY_KMAJOR, LAY_MNMAJOR, EpiBiasReluF32><<<grid, G_THREADS, G_SMEM, stream>>>(
      n2, FLAT, w1bf, FC, batch, FC, FLAT, EpiBiasReluF32{d1, b1, FC});
  return (int)cudaGetLastError();
}

int launch_dense_dgrad(const uint16_t* dd1, const uint16_t* w1bf, const uint16_t* n2, uint16_t* dn2, int batch,
                       cudaStream_t stream) {
  dim3 grid((FLAT + GB_N - 1) / GB_N, (batch + GB_M - 1) / GB_M);
  gemm_bf16_kernel<LAY_KMAJOR, LAY_KMAJOR, EpiReluMaskBf16><<<grid, G_THREADS, G_SMEM, stream>>>(
      dd1, FC, w1bf, FC, batch, FLAT, FC, EpiReluMaskBf16{dn2, n2, FLAT});
  return (int)cudaGetLastError();
}

int launch_dense_wgrad(const uint16_t* n2, const uint16_t* dd1, float* g_w1, int batch, cudaStream_t stream) {
  dim3 grid(FC / GB_N, (FLAT + GB_M - 1) / GB_M);
  gemm_bf16_kernel<LAY_MNMAJOR, LAY_MNMAJOR, EpiStoreF32><<<grid, G_THREADS, G_SMEM, stream>>>(
      n2, FLAT, dd1, FC, FLAT, FC, batch, EpiStoreF32{g_w1, FC});
  return (int)cudaGetLastError();
}

}  // namespace ga3c
