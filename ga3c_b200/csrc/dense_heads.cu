// dense1 forward + value / policy heads + A3C loss and its backward in ONE launch.
//
// Reference ops: dense1 = relu(flat @ w + b) (NetworkDNav.py:90, :256-269), then NetworkVP_discrate.py:60-85.
//
// The dense1 GEMM ([B,3872] x [3872,256]) is far too small for its K: split-K is what fills the SMs.  The first version left
// `splits` fp32 partial tiles in L2 (9.4 MB at B = 1024) for a second kernel to sum; here the K-slices of one 128-row tile are
// the CTAs of ONE thread-block cluster, each CTA parks its 128 x 256 fp32 partial tile in its own shared memory, and every
// CTA then finishes 16 of the tile's rows: it sums their eight partial rows over distributed shared memory (fixed rank order:
// deterministic), adds the bias, applies the ReLU, and runs the heads, the loss and its backward on them (heads_core.cuh) --
// the partial sums never reach L2, and the step has one launch less.
//
//   grid (ceil(B / 128), 1, 8), cluster (1, 1, 8): blockIdx.z = cluster rank = K-slice (8 k-blocks of 64; the last one 5)
//   warp 0     TMA producer (A: n2 rows, K-major; B: the bf16 shadow of dense1/w, MN-major; SWIZZLE_128B, 3-stage ring)
//   warp 1     TMEM allocator + MMA issuer (tcgen05.mma M = 128, N = 256)
//   warps 4-7  TMEM -> partial tile in shared memory (over the retired stage buffers)
//   all 8      after the cluster barrier: warp = sample (two rounds of 8), lane = 8 features, as in heads_kernel
#include "common.cuh"
#include "kernels.h"
#include "tcgen05.cuh"
#include "heads_core.cuh"

namespace ga3c {

constexpr int DH_CLUSTER = 8, DH_BN = FC, DH_BK = 64, DH_STAGES = 3;
constexpr int DH_A_STAGE = TC_BM * DH_BK * 2, DH_B_STAGE = DH_BN * DH_BK * 2, DH_STAGE = DH_A_STAGE + DH_B_STAGE;   // 16 + 32 KB
constexpr int DH_ROWS = TC_BM / DH_CLUSTER;                        // 16 rows finished per CTA
constexpr int DH_PITCH = (FC + 4) * 4;                             // partial-tile row pitch: 260 floats, conflict-free float4 rows
static_assert(TC_BM * DH_PITCH <= DH_STAGES * DH_STAGE, "the partial tile overlays the stage buffers");
static_assert(DH_ROWS == 2 * HD_CHUNK && HD_THREADS == 256, "two chunks of 8 samples, one warp per sample");
constexpr int DH_SMEM = DH_STAGES * DH_STAGE + (2 * DH_STAGES + 1) * 8 + 16 + 1024;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

template <int A>
__global__ void __launch_bounds__(HD_THREADS, 1)
dense_heads_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, HeadsArgs p, int k_blocks) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ HeadsSmem<A> hs;
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;    // SWIZZLE_128B atoms need 1024-byte alignment
  const uint32_t bars = sbase + DH_STAGES * DH_STAGE;              // full[S], empty[S], tmem_full
  const uint32_t tmem_slot = bars + (2 * DH_STAGES + 1) * 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  EvtLog evt_i = evt_open();                   // debug event log of CTA 0 (ga3c_evt_*): one predicated load when detached
  const int m0 = blockIdx.x * TC_BM;
  const int per = (k_blocks + DH_CLUSTER - 1) / DH_CLUSTER;
  const int kb0 = (int)rank * per, kb1 = min(kb0 + per, k_blocks);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tm_a));
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tm_b));
    for (int s = 0; s < DH_STAGES; ++s) { mbar_init(bars + s * 8, 1); mbar_init(bars + (DH_STAGES + s) * 8, 1); }
    mbar_init(bars + 2 * DH_STAGES * 8, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<DH_BN>(tmem_slot);
  if (p.preload) heads_load_weights<A>(p, hs, tid, HD_THREADS);    // see heads_kernel for when the early read is safe
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_launch();
  evt_mark(evt_i, 80, 0);
  griddep_wait(K_DENSE_FWD);          // n2 comes from the conv forward, the weights from the previous step's optimizer
  evt_mark(evt_i, 81, 0);
  if (!p.preload) heads_load_weights<A>(p, hs, tid, HD_THREADS);
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    for (int kb = kb0; kb < kb1; ++kb) {
      const int it = kb - kb0, s = it % DH_STAGES;
      mbar_wait(bars + (DH_STAGES + s) * 8, ((it / DH_STAGES) & 1) ^ 1);      // slot free
      if (elect_one()) {
        const uint32_t full = bars + s * 8, sa = sbase + s * DH_STAGE, sb = sa + DH_A_STAGE;
        mbar_expect_tx(full, DH_STAGE);
        tma_load_2d(sa, &tm_a, kb * DH_BK, m0, full);
#pragma unroll
        for (int j = 0; j < DH_BN / 64; ++j) tma_load_2d(sb + j * 8192, &tm_b, 64 * j, kb * DH_BK, full);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(DH_BN, false, true);
    for (int kb = kb0; kb < kb1; ++kb) {
      const int it = kb - kb0, s = it % DH_STAGES;
      mbar_wait(bars + s * 8, (it / DH_STAGES) & 1);                          // TMA bytes landed
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = sbase + s * DH_STAGE, sb = sa + DH_A_STAGE;
        const uint64_t da = make_desc(sa, false), db = make_desc(sb, true);
#pragma unroll
        for (int k = 0; k < DH_BK / 16; ++k)      // advance K by 16: +32 B inside the swizzle row (A, K-major), +2 k-atoms = 2048 B (B, MN-major)
          tc_mma_bf16(tmem_base, da + (uint64_t)(32u * k >> 4), db + (uint64_t)(2048u * k >> 4), idesc, (it > 0 || k > 0) ? 1u : 0u);
        tc_commit(bars + (DH_STAGES + s) * 8);
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit(bars + 2 * DH_STAGES * 8);                     // accumulator complete
    __syncwarp();
    evt_mark(evt_i, 83, 0);
  } else if (warp >= 4) {
    // TMEM -> this CTA's partial tile in shared memory, over the stage buffers (every TMA load has landed and every UMMA has
    // read its operands once the accumulator barrier fires)
    const int q = warp & 3, row = q * 32 + lane;
    if (kb1 > kb0) {
      mbar_wait(bars + 2 * DH_STAGES * 8, 0);
      tc_fence_after();
    }
    evt_mark(evt_i, 84, 0);
#pragma unroll 1
    for (int c = 0; c < DH_BN / 32; ++c) {
      uint32_t r[32];
      if (kb1 > kb0) {
        tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, r);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = 0u;                               // a K-slice beyond the last k-block contributes zeros
      }
      const uint32_t dst = sbase + row * DH_PITCH + c * 128;
#pragma unroll
      for (int i = 0; i < 8; ++i) sts128(dst + i * 16, make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]));
    }
  }
  evt_mark(evt_i, 85, 0);
  tc_fence_before();
  cluster_sync_all();                 // every CTA's partial tile is complete and visible in the cluster
  evt_mark(evt_i, 86, 0);

  // ---- the 16 rows this CTA finishes: sum of the eight partial rows (rank order), + bias, ReLU, heads ----
  HeadsAcc<A> ac;
  ac.clear();
  const float inv_mix = 1.f / (1.f + p.min_policy * (float)A);
#pragma unroll 1
  for (int c = 0; c < DH_ROWS / HD_CHUNK; ++c) {
    const int sl = warp, trow = (int)rank * DH_ROWS + c * HD_CHUNK + sl, b = m0 + trow;
    if (b < p.batch) {
      float4 fa = *reinterpret_cast<const float4*>(&hs.b1s[4 * lane]);
      float4 fb = *reinterpret_cast<const float4*>(&hs.b1s[128 + 4 * lane]);
      float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sb = sa;
      const uint32_t local = sbase + trow * DH_PITCH + lane * 16;
      float4 qa[DH_CLUSTER], qb[DH_CLUSTER];
#pragma unroll
      for (int q = 0; q < DH_CLUSTER; ++q) {
        const uint32_t remote = map_to_rank(local, (uint32_t)q);
        qa[q] = ld_dsmem_f4(remote);
        qb[q] = ld_dsmem_f4(remote + 512);
      }
#pragma unroll
      for (int q = 0; q < DH_CLUSTER; ++q) {
        sa.x += qa[q].x; sa.y += qa[q].y; sa.z += qa[q].z; sa.w += qa[q].w;
        sb.x += qb[q].x; sb.y += qb[q].y; sb.z += qb[q].z; sb.w += qb[q].w;
      }
      fa.x = fmaxf(fa.x + sa.x, 0.f); fa.y = fmaxf(fa.y + sa.y, 0.f); fa.z = fmaxf(fa.z + sa.z, 0.f); fa.w = fmaxf(fa.w + sa.w, 0.f);
      fb.x = fmaxf(fb.x + sb.x, 0.f); fb.y = fmaxf(fb.y + sb.y, 0.f); fb.z = fmaxf(fb.z + sb.z, 0.f); fb.w = fmaxf(fb.w + sb.w, 0.f);
      heads_sample<A>(p, hs, ac, b, sl, lane, fa, fb, inv_mix);
    } else if (p.train) {
      heads_pad_sample<A>(hs, sl, lane);
    }
    evt_mark(evt_i, 87, c);
    if (p.train) {
      __syncthreads();
      heads_accumulate<A>(p, hs, ac, m0 + (int)rank * DH_ROWS + c * HD_CHUNK, tid);
      __syncthreads();
    }
    evt_mark(evt_i, 88, c);
  }
  cluster_sync_all();                 // nobody leaves while a peer may still read its partial tile
  evt_mark(evt_i, 89, 0);
  if (p.train) heads_store_slab<A>(p, hs, ac, (int)(blockIdx.x * DH_CLUSTER + rank), tid, warp, lane);
  trace_mark(K_DENSE_FWD, 2);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<DH_BN>(tmem_base);
  }
}

GA3C_TRACE_ATTACH(trace_attach_dense_heads)
GA3C_EVT_ATTACH(evt_attach_dense_heads)

// ---- host side ----------------------------------------------------------------------------------
int make_tmap_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);   // dense_tc.cu

int dense_heads_ctas(int batch) { return ((batch + TC_BM - 1) / TC_BM) * DH_CLUSTER; }

template <int A>
static int configure_dense_heads_t() {
  return (int)cudaFuncSetAttribute(dense_heads_kernel<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, DH_SMEM);
}
int configure_dense_heads() {          // per device: called by ga3c_create
  int r = 0;
#define GA3C_CASE(N) if (!r) r = configure_dense_heads_t<N>();
  GA3C_CASE(1) GA3C_CASE(2) GA3C_CASE(3) GA3C_CASE(4) GA3C_CASE(5) GA3C_CASE(6) GA3C_CASE(7) GA3C_CASE(8) GA3C_CASE(9)
  GA3C_CASE(10) GA3C_CASE(11) GA3C_CASE(12) GA3C_CASE(13) GA3C_CASE(14) GA3C_CASE(15) GA3C_CASE(16) GA3C_CASE(17) GA3C_CASE(18)
#undef GA3C_CASE
  return r;
}

template <int A>
static int launch_dense_heads_t(const CUtensorMap& ta, const CUtensorMap& tb, const HeadsArgs& args, int k_blocks, cudaStream_t stream) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((args.batch + TC_BM - 1) / TC_BM, 1, DH_CLUSTER);
  cfg.blockDim = dim3(HD_THREADS);
  cfg.dynamicSmemBytes = DH_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = 1; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = DH_CLUSTER;
  cfg.attrs = attr; cfg.numAttrs = 2;
  return (int)cudaLaunchKernelEx(&cfg, dense_heads_kernel<A>, ta, tb, args, k_blocks);
}

int launch_dense_heads(const uint16_t* n2, const uint16_t* w1bf, const HeadsArgs& args, cudaStream_t stream) {
  CUtensorMap ta, tb;
  if (make_tmap_bf16(&ta, n2, args.batch, FLAT, FLAT, TC_BM)) return (int)cudaErrorInvalidValue;      // A: [B][3872], K inner
  if (make_tmap_bf16(&tb, w1bf, FLAT, FC, FC, 64)) return (int)cudaErrorInvalidValue;                 // B: [3872][256], N inner
  const int k_blocks = (FLAT + DH_BK - 1) / DH_BK;
  switch (args.num_actions) {
#define GA3C_CASE(N) case N: return launch_dense_heads_t<N>(ta, tb, args, k_blocks, stream);
    GA3C_CASE(1) GA3C_CASE(2) GA3C_CASE(3) GA3C_CASE(4) GA3C_CASE(5) GA3C_CASE(6) GA3C_CASE(7) GA3C_CASE(8)
    GA3C_CASE(9) GA3C_CASE(10) GA3C_CASE(11) GA3C_CASE(12) GA3C_CASE(13) GA3C_CASE(14) GA3C_CASE(15)
    GA3C_CASE(16) GA3C_CASE(17) GA3C_CASE(18)
#undef GA3C_CASE
    default: return (int)cudaErrorInvalidValue;
  }
}

}  // namespace ga3c
