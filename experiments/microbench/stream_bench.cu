// Microbenchmark: how fast can 148 persistent CTAs stream fp32 frames (112,896 B each, one frame per CTA at a time, like
// conv_fwd / conv_bwd) from HBM into shared memory?  Variants: cp.async.bulk ring (chunk bytes x slots x warps) and plain
// vectorised loads.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_bench stream_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
constexpr int FRAME = 112896;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// each of `warps` warps streams chunks q = w, w + warps, ... of the CTA's chunk stream through `slots` private slots
__global__ void bulk_kernel(const uint8_t* x, int n_frames_total, int chunk, int slots, int warps, int touch, float* sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_frame = FRAME / chunk;
  const int n_frames = ((int)blockIdx.x < n_frames_total) ? (n_frames_total - 1 - (int)blockIdx.x) / gridDim.x + 1 : 0;
  const int n_chunks = n_frames * per_frame;
  const uint32_t ring = smem_u32(smem), bars = ring + warps * slots * chunk;
  if (threadIdx.x == 0) { for (int i = 0; i < warps * slots; ++i) mbar_init(bars + 8 * i, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  __syncthreads();
  auto issue = [&](int q, int slot) {
    const int k = q / per_frame, c = q - k * per_frame;
    const size_t f = blockIdx.x + (size_t)k * gridDim.x;
    mbar_expect_tx(bars + 8 * slot, chunk);
    bulk_load(ring + slot * chunk, x + f * FRAME + (size_t)c * chunk, chunk, bars + 8 * slot);
  };
  float acc = 0.f;
  if (warp < warps) {
    if (lane == 0) for (int j = 0; j < slots; ++j) if (warp + j * warps < n_chunks) issue(warp + j * warps, warp * slots + j);
    int j = 0;
    for (int q = warp; q < n_chunks; q += warps, ++j) {
      const int slot = warp * slots + (j % slots);
      mbar_wait(bars + 8 * slot, (j / slots) & 1);
      if (touch) {   // read the chunk like the converter does (one 16-byte load per lane per 512 B)
        for (int o = lane * 16; o < chunk; o += 512) {
          float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ring + slot * chunk + o));
          acc += v.x + v.y + v.z + v.w;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      __syncwarp();
      if (lane == 0 && q + slots * warps < n_chunks) issue(q + slots * warps, slot);
    }
  }
  if (acc == 12345.678f) sink[0] = acc;
}

// plain loads: every thread keeps `unroll` 16-byte loads in flight
template <int UNROLL>
__global__ void ldg_kernel(const float4* x, size_t n4, float* sink) {
  float acc = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (UNROLL - 1) * stride < n4; i += UNROLL * stride) {
    float4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(x + i + u * stride));
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  if (acc == 12345.678f) sink[0] = acc;
}

int main() {
  const int B = 4096;                       // 462 MB > L2
  uint8_t* x; float* sink;
  CK(cudaMalloc(&x, (size_t)B * FRAME)); CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(x, 1, (size_t)B * FRAME));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  auto report = [&](const char* name, float ms) { printf("%-60s %8.1f us  %7.1f GB/s\n", name, ms * 1e3, (double)B * FRAME / ms / 1e6); };
  struct Cfg { int chunk, slots, warps, touch, ctas; } cfgs[] = {
      {5376, 2, 6, 1, 148},  {5376, 2, 6, 0, 148},  {5376, 4, 6, 1, 148},  {5376, 6, 6, 1, 148}, {5376, 3, 12, 1, 148},
      {16128, 1, 4, 1, 148}, {16128, 1, 8, 1, 148}, {16128, 1, 12, 1, 148}, {28224, 1, 4, 1, 148}, {28224, 1, 8, 1, 148},
      {5376, 2, 6, 1, 296},  {5376, 2, 8, 1, 296}, {16128, 1, 6, 1, 296}, {2688, 4, 6, 1, 148}, {2688, 8, 6, 1, 148}, {2688, 12, 6, 1, 148}};
  for (auto c : cfgs) {
    const int smem = c.chunk * c.slots * c.warps + 8 * c.slots * c.warps + 128;
    if (smem > 227 * 1024 / (c.ctas > 148 ? 2 : 1)) { printf("skip\n"); continue; }
    for (int it = 0; it < 3; ++it) {
      CK(cudaEventRecord(e0));
      bulk_kernel<<<c.ctas, 32 * c.warps, smem>>>(x, B, c.chunk, c.slots, c.warps, c.touch, sink);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    }
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    char name[128]; snprintf(name, sizeof name, "bulk chunk %5d x %2d slots x %2d warps (%3d KB in flight) touch %d ctas %d", c.chunk, c.slots, c.warps, c.chunk * c.slots * c.warps / 1024, c.touch, c.ctas);
    report(name, ms);
  }
  const size_t n4 = (size_t)B * FRAME / 16;
  for (int blocks : {148 * 2, 148 * 4, 148 * 8}) {
    for (int it = 0; it < 3; ++it) { CK(cudaEventRecord(e0)); ldg_kernel<8><<<blocks, 512>>>((const float4*)x, n4, sink); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); }
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    char name[128]; snprintf(name, sizeof name, "ld.global.nc v4 x8 unroll, %d blocks x 512", blocks); report(name, ms);
  }
  return 0;
}
