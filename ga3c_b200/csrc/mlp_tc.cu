// The wide layers of the low-dimensional MLP networks (BASELINE configs[3]; fork NetworkVP.py:79-105: 256 -> 256 and 256 -> 100 hold
// 92 % of the flops) on the 5th-generation tensor cores, in fp32-equivalent precision: tcgen05.mma.kind::tf32 with the 3xTF32
// operand split.  Every fp32 operand x is split into  hi = tf32(x)  (round to nearest, 10 mantissa bits)  and  lo = x - hi  (exact
// in fp32; the tensor core reads its leading 11 bits), and  a.b ~ hi_a.hi_b + hi_a.lo_b + lo_a.hi_b  accumulates in fp32 in TMEM:
// the dropped terms are 2^-22 relative, the error of the products is that of an fp32 FMA chain (experiments/tf32_mlp_error_study.py).
//
//   fwd  : out_l [B, n]   = act(in_l [B, k] x W_l [k, n] + b_l)           (A K-major,  B MN-major)
//   dgrad: dz_{l-1} [B,k] = (dz_l [B, n] x W_l [k, n]^T) . act'(out_{l-1})  (A K-major,  B K-major)
//   wgrad: dW_l [k, n]    = in_l [B, k]^T x dz_l [B, n]                    (A MN-major, B MN-major), split over the batch into
//                                                                           the partial arenas mlp_reduce sums in a fixed order
// Operands are the fp32 row-major matrices the fused kernel (mlp.cu) already keeps in HBM for training; nothing is transposed
// or converted in HBM.  One 128 x BN output tile per CTA, 448 threads:
//   warp 0     TMA producer: fp32 boxes of 32 floats (one 128-byte swizzle row) x rows, SWIZZLE_128B
//   warp 1     TMEM allocator + UMMA issuer: per 32-float k-block 4 k-steps (K = 8) x 3 products
//   warps 2-5  epilogue: tcgen05.ld 32 lanes x 32 columns -> bias / activation / act' -> global
//   warps 6-13 splitters: rewrite a landed stage in place as hi and write lo into the stage's twin buffer (the operand layout
//              is whatever TMA produced: the split is element-wise), fence.proxy.async, arrive on the stage's "split" barrier
#include "common.cuh"
#include "kernels.h"
#include "mlp.cuh"
#include "tcgen05.cuh"

namespace ga3c {

namespace {

constexpr int T3_BK = 32;                          // floats of K per stage: one 128-byte swizzle row
constexpr int T3_THREADS = 448, T3_SPLIT_WARP0 = 6, T3_SPLIT_WARPS = 8;
constexpr int T3_A_BYTES = TC_BM * T3_BK * 4;      // 16 KB

template <int BN>
constexpr int t3_half() { return T3_A_BYTES + BN * T3_BK * 4; }      // hi (or lo) operands of one stage
template <int BN>
constexpr int t3_stages() { return BN >= 256 ? 2 : 3; }
template <int BN>
constexpr int t3_smem() { return t3_stages<BN>() * 2 * t3_half<BN>() + (3 * t3_stages<BN>() + 1) * 8 + 16 + 1024; }

// instruction descriptor, kind::tf32: D = f32, A = B = tf32
__host__ __device__ constexpr uint32_t make_idesc_tf32(int m, int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}
// Shared-memory descriptor of an fp32 operand.  K-major: SWIZZLE_128B, rows of 128 B (32 floats of K), 8-row atoms: SBO = 1024.
// MN-major: 32-bit operands can only be transposed from the SWIZZLE_128B_BASE32B layout (layout type 1: 32-byte chunks swizzled
// within 128 B, pattern period 4 rows; TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): atom = 4 k-rows x 128 B (32 floats of M / N),
// SBO = 512 between k-atoms, LBO = 4096 between the 32-wide atoms (one TMA box each).
__device__ __forceinline__ uint64_t make_desc_f32(uint32_t saddr, bool mn_major) {
  const uint64_t lbo = mn_major ? (4096u >> 4) : 1u, sbo = mn_major ? (512u >> 4) : (1024u >> 4), type = mn_major ? 1ull : 2ull;
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (type << 61);
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}

// ---- epilogues: one thread owns output row m and 32 consecutive columns n .. n + 31 (n, N multiples of 4) -------------------
struct Epi3Fwd {         // out = act(acc + b)
  const float* bias; float* out; int N, act;
  __device__ __forceinline__ void operator()(int, int m, int n, const uint32_t (&r)[32], bool zero) const {
    float* dst = out + (size_t)m * N + n;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (n + 4 * i >= N) break;
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias + n) + i);
      float v[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[e] += zero ? 0.f : __uint_as_float(r[4 * i + e]);
        if (act == MLP_ACT_SIGMOID) v[e] = 1.f / (1.f + expf(-v[e]));
      }
      reinterpret_cast<float4*>(dst)[i] = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
};
struct Epi3Dgrad {       // dz_prev = acc * act'(out_prev)
  const float* out_prev; float* dz; int N, act;
  __device__ __forceinline__ void operator()(int, int m, int n, const uint32_t (&r)[32], bool zero) const {
    float* dst = dz + (size_t)m * N + n;
    const float4* op = reinterpret_cast<const float4*>(out_prev + (size_t)m * N + n);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (n + 4 * i >= N) break;
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = zero ? 0.f : __uint_as_float(r[4 * i + e]);
      if (act == MLP_ACT_SIGMOID) {
        const float4 o = op[i];
        v[0] *= o.x * (1.f - o.x); v[1] *= o.y * (1.f - o.y); v[2] *= o.z * (1.f - o.z); v[3] *= o.w * (1.f - o.w);
      }
      reinterpret_cast<float4*>(dst)[i] = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
};
struct Epi3Wgrad {       // raw tile into partial arena `split` (part points at the layer's weights in arena 0)
  float* part; int64_t part_stride; int N;
  __device__ __forceinline__ void operator()(int split, int m, int n, const uint32_t (&r)[32], bool zero) const {
    float* dst = part + (size_t)split * part_stride + (size_t)m * N + n;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (n + 4 * i >= N) break;
      reinterpret_cast<float4*>(dst)[i] =
          zero ? make_float4(0.f, 0.f, 0.f, 0.f)
               : make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                             __uint_as_float(r[4 * i + 3]));
    }
  }
};

template <int BN, bool A_MN, bool B_MN, class Epi>
__global__ void __launch_bounds__(T3_THREADS, 1)
gemm3_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, int M, int N, int k_blocks,
             int k_blocks_per_split, const Epi epi) {
  constexpr int STAGES = t3_stages<BN>(), HALF = t3_half<BN>(), STAGE = 2 * HALF;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;     // SWIZZLE_128B atoms need 1024-byte alignment
  uint8_t* const smem = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bars = sbase + STAGES * STAGE;     // full[STAGES] | split[STAGES] | empty[STAGES] | tmem_full
  const uint32_t tmem_slot = bars + (3 * STAGES + 1) * 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * TC_BM, n0 = blockIdx.y * BN, split = blockIdx.z;
  const int kb0 = split * k_blocks_per_split;
  const int kb1 = min(kb0 + k_blocks_per_split, k_blocks);
  auto bar_full = [&](int s) { return bars + s * 8; };
  auto bar_split = [&](int s) { return bars + (STAGES + s) * 8; };
  auto bar_empty = [&](int s) { return bars + (2 * STAGES + s) * 8; };
  const uint32_t bar_done = bars + 3 * STAGES * 8;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tm_a));
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tm_b));
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_split(s), T3_SPLIT_WARPS); mbar_init(bar_empty(s), 1); }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) tmem_alloc<BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_launch();
  griddep_wait(K_MLP_TC);           // operands are produced by the preceding kernels
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    for (int kb = kb0; kb < kb1; ++kb) {
      const int it = kb - kb0, s = it % STAGES;
      mbar_wait(bar_empty(s), ((it / STAGES) & 1) ^ 1);
      if (elect_one()) {
        const uint32_t sa = sbase + s * STAGE, sb = sa + T3_A_BYTES;
        mbar_expect_tx(bar_full(s), HALF);
        if (A_MN) {
#pragma unroll
          for (int j = 0; j < TC_BM / 32; ++j) tma_load_2d(sa + j * 4096, &tm_a, m0 + 32 * j, kb * T3_BK, bar_full(s));
        } else {
          tma_load_2d(sa, &tm_a, kb * T3_BK, m0, bar_full(s));
        }
        if (B_MN) {
#pragma unroll
          for (int j = 0; j < BN / 32; ++j) tma_load_2d(sb + j * 4096, &tm_b, n0 + 32 * j, kb * T3_BK, bar_full(s));
        } else {
          tma_load_2d(sb, &tm_b, kb * T3_BK, n0, bar_full(s));
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_tf32(TC_BM, BN, A_MN, B_MN);
    for (int kb = kb0; kb < kb1; ++kb) {
      const int it = kb - kb0, s = it % STAGES;
      mbar_wait(bar_split(s), (it / STAGES) & 1);          // hi / lo of the stage written (and visible to the async proxy)
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = sbase + s * STAGE, sb = sa + T3_A_BYTES;
        const uint64_t a_hi = make_desc_f32(sa, A_MN), b_hi = make_desc_f32(sb, B_MN);
        const uint64_t a_lo = make_desc_f32(sa + HALF, A_MN), b_lo = make_desc_f32(sb + HALF, B_MN);
#pragma unroll
        for (int k = 0; k < T3_BK / 8; ++k) {
          // advance K by 8: +32 B inside the 128-B swizzle row (K-major), +1 k-atom = 1024 B (MN-major)
          const uint64_t ka = (uint64_t)((A_MN ? 1024u : 32u) * k >> 4), kbv = (uint64_t)((B_MN ? 1024u : 32u) * k >> 4);
          tc_mma_tf32(tmem_base, a_lo + ka, b_hi + kbv, idesc, (it > 0 || k > 0) ? 1u : 0u);      // small terms first
          tc_mma_tf32(tmem_base, a_hi + ka, b_lo + kbv, idesc, 1u);
          tc_mma_tf32(tmem_base, a_hi + ka, b_hi + kbv, idesc, 1u);
        }
        tc_commit(bar_empty(s));
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit(bar_done);
    __syncwarp();
  } else if (warp < T3_SPLIT_WARP0) {
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int m = m0 + q * 32 + lane;
    const bool zero = kb1 <= kb0;                 // an empty split still writes its (zero) tile
    if (!zero) {
      mbar_wait(bar_done, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      if (!zero) tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, r);
      const int n = n0 + c * 32;
      if (m < M && n < N) epi(split, m, n, r, zero);
    }
  } else {
    const int t = threadIdx.x - 32 * T3_SPLIT_WARP0;      // 0 .. 255
    for (int kb = kb0; kb < kb1; ++kb) {
      const int it = kb - kb0, s = it % STAGES;
      mbar_wait(bar_full(s), (it / STAGES) & 1);           // TMA bytes landed
      float4* hi = reinterpret_cast<float4*>(smem + s * STAGE);
      float4* lo = reinterpret_cast<float4*>(smem + s * STAGE + HALF);
#pragma unroll
      for (int i = 0; i < HALF / 16 / (32 * T3_SPLIT_WARPS); ++i) {
        const int idx = t + 32 * T3_SPLIT_WARPS * i;
        const float4 x = hi[idx];
        float4 h;
        h.x = __uint_as_float((__float_as_uint(x.x) + 0x1000u) & 0xFFFFE000u);
        h.y = __uint_as_float((__float_as_uint(x.y) + 0x1000u) & 0xFFFFE000u);
        h.z = __uint_as_float((__float_as_uint(x.z) + 0x1000u) & 0xFFFFE000u);
        h.w = __uint_as_float((__float_as_uint(x.w) + 0x1000u) & 0xFFFFE000u);
        hi[idx] = h;
        lo[idx] = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
      }
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_split(s));
    }
  }
  tc_fence_before();
  __syncthreads();
  trace_mark(K_MLP_TC, 2);
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<BN>(tmem_base);
  }
}
static_assert(t3_half<256>() % (16 * 32 * T3_SPLIT_WARPS) == 0 && t3_half<128>() % (16 * 32 * T3_SPLIT_WARPS) == 0,
              "the splitters cover a stage in whole rounds");

// Weight gradient of a layer with fan-in K <= 4 (the state and the 4-wide first layer of the fork NetworkVP; K = 0: none) and the
// bias gradient of any layer, per batch split: dW[k][n] = sum_r in[r][k] dz[r][n], db[n] = sum_r dz[r][n].  Pure streaming --
// dz is read once, coalesced -- where a 64 x 64 tile kernel would spend its time on zero padding.  1024 threads = 16 row groups
// x 64 columns (the loop is bound by memory latency: rows in flight are what counts); the row groups are added in a fixed order.
constexpr int SK_RG = 16;
template <int K>
__global__ void __launch_bounds__(64 * SK_RG) mlp_skinny_wgrad_kernel(const float* __restrict__ in, const float* __restrict__ dz,
                                                                      int batch, int N, int rows_per_split, float* part_w,
                                                                      float* part_b, int64_t part_stride) {
  __shared__ float red[SK_RG][K + 1][64];
  griddep_launch();
  griddep_wait(K_MLP_TC);
  const int rg = threadIdx.x >> 6, c = threadIdx.x & 63, n = blockIdx.y * 64 + c;
  const int r0 = blockIdx.x * rows_per_split, r1 = min(batch, r0 + rows_per_split);
  float acc[K + 1];
#pragma unroll
  for (int k = 0; k <= K; ++k) acc[k] = 0.f;
  if (n < N) {
    constexpr int U = 4;
    int r = r0 + rg;
    for (; r + SK_RG * (U - 1) < r1; r += SK_RG * U) {
      float d[U], x[U][K > 0 ? K : 1];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        d[u] = __ldcs(dz + (size_t)(r + SK_RG * u) * N + n);
#pragma unroll
        for (int k = 0; k < K; ++k) x[u][k] = __ldg(in + (size_t)(r + SK_RG * u) * K + k);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = fmaf(x[u][k], d[u], acc[k]);
        acc[K] += d[u];
      }
    }
    for (; r < r1; r += SK_RG) {
      const float d = dz[(size_t)r * N + n];
#pragma unroll
      for (int k = 0; k < K; ++k) acc[k] = fmaf(__ldg(in + (size_t)r * K + k), d, acc[k]);
      acc[K] += d;
    }
  }
#pragma unroll
  for (int k = 0; k <= K; ++k) red[rg][k][c] = acc[k];
  __syncthreads();
  if (rg == 0 && n < N) {
    float* pw = part_w + (size_t)blockIdx.x * part_stride;
#pragma unroll
    for (int k = 0; k <= K; ++k) {
      float t = 0.f;
#pragma unroll
      for (int g = 0; g < SK_RG; ++g) t += red[g][k][c];
      if (k < K) pw[(size_t)k * N + n] = t;
      else part_b[(size_t)blockIdx.x * part_stride + n] = t;
    }
  }
}

// 2-D fp32 row-major matrix [rows][cols] (ld floats between rows), box = 32 inner x box_rows, 128-B swizzle (of 32-byte chunks
// for an operand that is consumed MN-major); out-of-bounds
// elements read as zero, so ragged M / N / K tails need no special casing in the kernel
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}
int make_tmap_f32(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, bool mn_major) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return (int)cudaErrorNotSupported;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 4};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

template <int BN, bool A_MN, bool B_MN, class Epi>
int launch_gemm3(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int k_blocks, int per, int splits, const Epi& epi,
                 cudaStream_t stream) {
  static const int configured = (int)cudaFuncSetAttribute(gemm3_kernel<BN, A_MN, B_MN, Epi>,
                                                          cudaFuncAttributeMaxDynamicSharedMemorySize, t3_smem<BN>());
  if (configured) return configured;
  const dim3 grid((M + TC_BM - 1) / TC_BM, (N + BN - 1) / BN, splits);
  return launch_pdl(gemm3_kernel<BN, A_MN, B_MN, Epi>, grid, dim3(T3_THREADS), (size_t)t3_smem<BN>(), stream, ta, tb, M, N,
                    k_blocks, per, epi);
}

}  // namespace

GA3C_TRACE_ATTACH(trace_attach_mlp_tc)

bool mlp_tc_layer_ok(const MlpLayerDesc& L) { return L.k >= 64 && L.n >= 64 && L.k % 4 == 0 && L.n % 4 == 0 && L.n <= 256 && L.k <= 256; }

// out = act(in [B, k] x W [k, n] + bias)
int launch_mlp_tc_fwd(const float* W, const float* bias, int k, int n, int act, const float* in, float* out, int batch,
                      cudaStream_t stream) {
  CUtensorMap ta, tb;
  if (make_tmap_f32(&ta, in, batch, k, k, TC_BM, false)) return (int)cudaErrorInvalidValue;            // A: [B][k], K inner
  if (make_tmap_f32(&tb, W, k, n, n, T3_BK, true)) return (int)cudaErrorInvalidValue;                 // B: [k][n], N inner
  const int kblocks = (k + T3_BK - 1) / T3_BK;
  const Epi3Fwd epi{bias, out, n, act};
  if (n > 128) return launch_gemm3<256, false, true>(ta, tb, batch, n, kblocks, kblocks, 1, epi, stream);
  return launch_gemm3<128, false, true>(ta, tb, batch, n, kblocks, kblocks, 1, epi, stream);
}

// dz_prev [B, k] = (dz [B, n] x W [k, n]^T) * act'(out_prev);  prev_act: activation of the layer that produced out_prev
int launch_mlp_tc_dgrad(const float* W, int k, int n, const float* dz, const float* out_prev, int prev_act, float* dz_prev,
                        int batch, cudaStream_t stream) {
  CUtensorMap ta, tb;
  if (make_tmap_f32(&ta, dz, batch, n, n, TC_BM, false)) return (int)cudaErrorInvalidValue;            // A: [B][n], reduction n inner
  const int bn = k > 128 ? 256 : 128;
  if (make_tmap_f32(&tb, W, k, n, n, bn, false)) return (int)cudaErrorInvalidValue;                    // B: [k][n] = [N_out][K_red]
  const int kblocks = (n + T3_BK - 1) / T3_BK;
  const Epi3Dgrad epi{out_prev, dz_prev, k, prev_act};
  if (bn == 256) return launch_gemm3<256, false, false>(ta, tb, batch, k, kblocks, kblocks, 1, epi, stream);
  return launch_gemm3<128, false, false>(ta, tb, batch, k, kblocks, kblocks, 1, epi, stream);
}

// dW [k, n] = in [B, k]^T x dz [B, n]: batch split s into part_w + s * part_stride (every split writes its whole tile, zeros
// if it has no rows); the bias gradient comes from launch_mlp_skinny_wgrad with k = 0
int launch_mlp_tc_wgrad(int k, int n, const float* in, const float* dz, int batch, int splits, int rows_per_split, float* part_w,
                        float* part_b, int64_t part_stride, cudaStream_t stream) {
  CUtensorMap ta, tb;
  if (make_tmap_f32(&ta, in, batch, k, k, T3_BK, true)) return (int)cudaErrorInvalidValue;            // A: [K=B][M=k], M inner
  if (make_tmap_f32(&tb, dz, batch, n, n, T3_BK, true)) return (int)cudaErrorInvalidValue;            // B: [K=B][N=n], N inner
  const int kblocks = (batch + T3_BK - 1) / T3_BK;
  const int per = rows_per_split / T3_BK;
  const Epi3Wgrad epi{part_w, part_stride, n};
  int r;
  if (n > 128) r = launch_gemm3<256, true, true>(ta, tb, k, n, kblocks, per, splits, epi, stream);
  else r = launch_gemm3<128, true, true>(ta, tb, k, n, kblocks, per, splits, epi, stream);
  (void)part_b;
  return r;
}

int launch_mlp_skinny_wgrad(int k, int n, const float* in, const float* dz, int batch, int splits, int rows_per_split,
                            float* part_w, float* part_b, int64_t part_stride, cudaStream_t stream) {
  const dim3 grid(splits, (n + 63) / 64), block(64 * SK_RG);
  switch (k) {
    case 0: return launch_pdl(mlp_skinny_wgrad_kernel<0>, grid, block, 0, stream, in, dz, batch, n, rows_per_split, part_w, part_b, part_stride);
    case 1: return launch_pdl(mlp_skinny_wgrad_kernel<1>, grid, block, 0, stream, in, dz, batch, n, rows_per_split, part_w, part_b, part_stride);
    case 2: return launch_pdl(mlp_skinny_wgrad_kernel<2>, grid, block, 0, stream, in, dz, batch, n, rows_per_split, part_w, part_b, part_stride);
    case 3: return launch_pdl(mlp_skinny_wgrad_kernel<3>, grid, block, 0, stream, in, dz, batch, n, rows_per_split, part_w, part_b, part_stride);
    case 4: return launch_pdl(mlp_skinny_wgrad_kernel<4>, grid, block, 0, stream, in, dz, batch, n, rows_per_split, part_w, part_b, part_stride);
  }
  return (int)cudaErrorInvalidValue;
}

}  // namespace ga3c
