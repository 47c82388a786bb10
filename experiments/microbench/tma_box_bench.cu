// Microbenchmark: what does ONE cp.async.bulk.tensor box cost an SM when the tensor sits in L2?
// The dense1 GEMMs (dense_tc.cu) read L2-resident operands (n2 7.9 MB, dd1 0.5 MB, the bf16 shadow of dense1/w 2 MB) in 8 KB and
// 16 KB SWIZZLE_128B boxes and run at a third of ncu's L2 throughput: is the limit bytes, boxes, or bytes in flight?
// Every CTA's producer lane keeps `depth` boxes in flight (one mbarrier per slot, re-issued as soon as it lands), boxes walk a
// bf16 [4096][3872] matrix (31.7 MB: L2-resident after the warm-up pass) at CTA-dependent coordinates.
// Variants: 2-D boxes {64 x rows} (a K-major operand tile: rows of 128 B, 7,744 B apart) and 3-D boxes {64 x rows x halves}
// (an MN-major operand tile: `halves` 64-column blocks of the same rows in ONE instruction).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_box_bench tma_box_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
constexpr int ROWS = 4096, COLS = 3872;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma2d(uint32_t dst, const CUtensorMap* m, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
               ::"r"(dst), "l"(m), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma3d(uint32_t dst, const CUtensorMap* m, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
               ::"r"(dst), "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

// lane 0 of every warp of the CTA is a producer with its own ring: `n_boxes` boxes of `box_bytes` each, `depth` in flight
__global__ void box_kernel(const __grid_constant__ CUtensorMap map, int three_d, int box_rows, int halves, int box_bytes, int depth,
                           int n_boxes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const uint32_t base = ((smem_u32(smem) + 1023u) & ~1023u) + warp * depth * box_bytes;
  const uint32_t bars = ((smem_u32(smem) + 1023u) & ~1023u) + warps * depth * box_bytes + warp * depth * 8;
  if ((threadIdx.x & 31) == 0) {
    for (int i = 0; i < depth; ++i) mbar_init(bars + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    const int row_tiles = ROWS / box_rows, col_blocks = COLS / 64 / halves;      // 60 column blocks of 64 (the ragged 61st is left out)
    auto issue = [&](int q, int slot) {
      const unsigned t = (unsigned)((blockIdx.x * 4 + warp) * 131 + q * 17);
      const int r0 = (int)(t % row_tiles) * box_rows, c0 = (int)((t / row_tiles) % col_blocks) * halves;
      mbar_expect_tx(bars + 8 * slot, box_bytes);
      if (three_d) tma3d(base + slot * box_bytes, &map, 0, r0, c0, bars + 8 * slot);
      else tma2d(base + slot * box_bytes, &map, c0 * 64, r0, bars + 8 * slot);
    };
    for (int j = 0; j < depth && j < n_boxes; ++j) issue(j, j);
    for (int q = 0; q < n_boxes; ++q) {
      const int slot = q % depth;
      mbar_wait(bars + 8 * slot, (q / depth) & 1);
      if (q + depth < n_boxes) issue(q + depth, slot);
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  EncodeTiledFn encode = (EncodeTiledFn)fnp;
  uint16_t* x;
  CK(cudaMalloc(&x, (size_t)ROWS * COLS * 2));
  CK(cudaMemset(x, 0, (size_t)ROWS * COLS * 2));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaFuncSetAttribute(box_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  struct V { int three_d, rows, halves, depth, ctas_per_sm, warps; };
  const V vs[] = {
      {0, 64, 1, 2, 1},  {0, 64, 1, 4, 1},  {0, 64, 1, 8, 1},  {0, 64, 1, 16, 1},     // 8 KB boxes
      {0, 128, 1, 2, 1}, {0, 128, 1, 4, 1}, {0, 128, 1, 8, 1},                        // 16 KB
      {0, 256, 1, 2, 1}, {0, 256, 1, 4, 1},                                           // 32 KB
      {1, 64, 2, 2, 1},  {1, 64, 2, 4, 1},  {1, 64, 2, 8, 1},                         // 16 KB as {64 x 64 x 2}
      {1, 64, 4, 2, 1},  {1, 64, 4, 4, 1},                                            // 32 KB as {64 x 64 x 4}
      {0, 64, 1, 3, 3},  {0, 64, 1, 8, 3},  {0, 128, 1, 2, 3}, {0, 128, 1, 4, 3},     // three CTAs per SM (dense_bwd's residency)
      {1, 64, 2, 4, 3},  {1, 64, 4, 2, 3},
      {0, 64, 1, 4, 1, 2}, {0, 64, 1, 4, 1, 3}, {0, 64, 1, 4, 1, 4},                  // several producer warps in ONE CTA
      {0, 128, 1, 2, 1, 2}, {0, 128, 1, 2, 1, 3}, {1, 64, 2, 2, 1, 2}, {1, 64, 2, 2, 1, 3},
      {0, 64, 1, 3, 3, 2}, {1, 64, 2, 2, 3, 2},                                       // ... and in each of three CTAs per SM
  };
  printf("# %d SMs; bf16 [%d][%d] (%.1f MB) in L2; producer = lane 0 of a warp\n", sms, ROWS, COLS, ROWS * COLS * 2 / 1e6);
  for (const V& v : vs) {
    CUtensorMap map;
    const int box_bytes = 128 * v.rows * v.halves;
    CUresult r;
    if (v.three_d) {
      cuuint64_t dims[3] = {64, (cuuint64_t)ROWS, (cuuint64_t)(COLS / 64)};
      cuuint64_t strides[2] = {(cuuint64_t)COLS * 2, 128};
      cuuint32_t box[3] = {64, (cuuint32_t)v.rows, (cuuint32_t)v.halves}, es[3] = {1, 1, 1};
      r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      cuuint64_t dims[2] = {(cuuint64_t)COLS, (cuuint64_t)ROWS};
      cuuint64_t strides[1] = {(cuuint64_t)COLS * 2};
      cuuint32_t box[2] = {64, (cuuint32_t)v.rows}, es[2] = {1, 1};
      r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) { printf("encode failed (%d) for 3d=%d rows=%d halves=%d\n", (int)r, v.three_d, v.rows, v.halves); continue; }
    const int warps = v.warps ? v.warps : 1;
    const int smem = warps * (v.depth * box_bytes + 8 * v.depth) + 2048;
    if (smem * v.ctas_per_sm > 220 * 1024) { printf("skip (smem)\n"); continue; }
    const int n_boxes = (4 << 20) / box_bytes / v.ctas_per_sm / warps;        // 4 MB per SM and launch
    const int grid = sms * v.ctas_per_sm;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      CK(cudaEventRecord(e0));
      box_kernel<<<grid, 32 * warps, smem>>>(map, v.three_d, v.rows, v.halves, box_bytes, v.depth, n_boxes);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaGetLastError());
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep > 0 && ms < best) best = ms;
    }
    const double bytes = (double)grid * warps * n_boxes * box_bytes;
    printf("%s box {64 x %3d%s} %5d B, depth %2d, %d CTA/SM x %d producer(s) (%3d KB in flight per SM): %7.1f us  %7.1f GB/s  %6.1f ns per box and SM  %5.1f B/clk/SM\n",
           v.three_d ? "3-D" : "2-D", v.rows, v.three_d ? (v.halves == 2 ? " x 2" : " x 4") : "    ", box_bytes, v.depth, v.ctas_per_sm, warps,
           warps * v.depth * box_bytes * v.ctas_per_sm / 1024, best * 1e3, bytes / best / 1e6,
           best * 1e6 / (n_boxes * v.ctas_per_sm * warps), (double)box_bytes * n_boxes * v.ctas_per_sm * warps / (best * 1e-3 * 1.965e9));
  }
  return 0;
}
