// Microbenchmark: issue-to-retire cost of small tcgen05.mma (kind::f16, bf16, cta_group::1) by shape and operand layout,
// operands in shared memory (SS mode).  One CTA; `reps` UMMAs round-robin over `nacc` accumulators, then one commit.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bench umma_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mma(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %6, 0;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\n"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n}\n" ::"r"(d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
}
struct Cfg { int M, N, a_mn, b_mn, swz, nacc, a_lbo, a_sbo, b_lbo, b_sbo, a_step, b_step, nrot; };

__global__ void __launch_bounds__(128, 1) bench(Cfg c, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) unsigned long long bar_s;
  __shared__ uint32_t tslot;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u, bar = smem_u32(&bar_s);
  for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tslot;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)c.a_mn << 15) | ((uint32_t)c.b_mn << 16) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24);
  const uint32_t lay = c.swz ? (2u << 29) : 0u;       // layout type in bits 61..63 of the descriptor
  const uint32_t a_base = base, b_base = base + 128 * 1024;
  const uint32_t a_lo0 = ((a_base & 0x3FFFFu) >> 4) | ((uint32_t)(c.a_lbo >> 4) << 16), a_hi = (uint32_t)(c.a_sbo >> 4) | (1u << 14) | lay;
  const uint32_t b_lo0 = ((b_base & 0x3FFFFu) >> 4) | ((uint32_t)(c.b_lbo >> 4) << 16), b_hi = (uint32_t)(c.b_sbo >> 4) | (1u << 14) | lay;
  long long t0 = 0, t1 = 0;
  if (threadIdx.x < 32) {
    for (int round = 0; round < 2; ++round) {          // round 0 warms up
      t0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < reps; ++i) {
          const int r = i % c.nrot;
          mma(tmem + (i % c.nacc) * c.N, a_lo0 + r * (c.a_step >> 4), a_hi, b_lo0 + r * (c.b_step >> 4), b_hi, idesc, 1u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
      }
      __syncwarp();
      mbar_wait(bar, round & 1);
      t1 = clock64();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem) : "memory");
}

int main() {
  long long* out; CK(cudaMalloc(&out, 8 * 148));
  CK(cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024 + 1024));
  const int reps = 2048;
  struct Named { const char* name; Cfg c; } tests[] = {
      // no-swizzle planes: MN-major LBO = 128 (8 k-rows), SBO = plane; K-major LBO = plane, SBO = 128
      {"wgrad11  M64  N16  A MN-ns  B MN-ns  4 acc", {64, 16, 1, 1, 0, 4, 128, 7872, 128, 7424, 256, 256, 16}},
      {"wgrad11  M64  N16  A MN-ns  B MN-ns  1 acc", {64, 16, 1, 1, 0, 1, 128, 7872, 128, 7424, 256, 256, 16}},
      {"wgrad12  M64  N32  A MN-ns  B MN-ns  4 acc", {64, 32, 1, 1, 0, 4, 128, 2560, 128, 2816, 256, 256, 8}},
      {"merged   M64  N64  A MN-ns  B MN-ns  1 acc", {64, 64, 1, 1, 0, 1, 128, 7872, 128, 7424, 256, 256, 16}},
      {"merged   M64  N128 A MN-ns  B MN-ns  1 acc", {64, 128, 1, 1, 0, 1, 128, 7872, 128, 4096, 256, 256, 8}},
      {"         M128 N16  A MN-ns  B MN-ns  4 acc", {128, 16, 1, 1, 0, 4, 128, 7872, 128, 7424, 256, 256, 16}},
      {"         M128 N64  A MN-ns  B MN-ns  1 acc", {128, 64, 1, 1, 0, 1, 128, 7872, 128, 7424, 256, 256, 16}},
      {"dgrad    M128 N64  A K-ns   B K-ns   2 acc", {128, 64, 0, 0, 0, 2, 2816, 128, 1024, 128, 16, 0, 8}},
      {"conv11f  M128 N32  A K-ns   B K-ns   4 acc", {128, 32, 0, 0, 0, 4, 8768, 128, 512, 128, 16, 0, 8}},
      {"         M128 N16  A K-ns   B K-ns   4 acc", {128, 16, 0, 0, 0, 4, 8768, 128, 512, 128, 16, 0, 8}},
      {"         M64  N16  A K-ns   B K-ns   4 acc", {64, 16, 0, 0, 0, 4, 8768, 128, 512, 128, 16, 0, 8}},
      {"conv12f  M128 N32  A K-sw128 B K-sw128 1 acc", {128, 32, 0, 0, 1, 1, 16, 1024, 16, 1024, 32, 32, 4}},
      {"         M128 N64  A K-sw128 B K-sw128 1 acc", {128, 64, 0, 0, 1, 1, 16, 1024, 16, 1024, 32, 32, 4}},
      {"         M128 N128 A K-sw128 B K-sw128 1 acc", {128, 128, 0, 0, 1, 1, 16, 1024, 16, 1024, 32, 32, 4}},
      {"         M128 N256 A K-sw128 B K-sw128 1 acc", {128, 256, 0, 0, 1, 1, 16, 1024, 16, 1024, 32, 32, 4}},
      {"         M64  N32  A K-sw128 B K-sw128 1 acc", {64, 32, 0, 0, 1, 1, 16, 1024, 16, 1024, 32, 32, 4}},
      {"         M64  N64  A MN-sw128 B MN-sw128 1 acc", {64, 64, 1, 1, 1, 1, 8192, 1024, 8192, 1024, 2048, 2048, 4}},
      {"         M64  N16  A MN-sw128 B MN-sw128 4 acc", {64, 16, 1, 1, 1, 4, 8192, 1024, 8192, 1024, 2048, 2048, 4}},
  };
  for (auto& t : tests) {
    bench<<<1, 128, 201 * 1024 + 1024>>>(t.c, reps, out);
    CK(cudaDeviceSynchronize());
    long long cyc; CK(cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost));
    const double per = (double)cyc / reps;
    const double bytes = (t.c.M * 16 + t.c.N * 16) * 2.0;
    printf("%-52s %7.1f cycles/UMMA   operands %5.0f B -> %5.1f B/clk   %6.1f%% of dense bf16 peak\n", t.name, per, bytes, bytes / per,
           100.0 * (2.0 * t.c.M * t.c.N * 16 / per) / 8192.0);
  }
  return 0;
}
