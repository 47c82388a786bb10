// Native predictor batcher: the body of ThreadPredictor.run (ThreadPredictor.py:45-66) as a thread inside the library, over the
// shared-memory slab transport of ga3c_b200.transport.
//
// The Python predictor thread spends its time in the interpreter: scanning the pending bytes, gathering 128 rows of 28 KB into a
// batch with numpy, calling into the library, scattering 128 replies and releasing 128 semaphores, all under the GIL that the
// trainer thread needs too (76 k predictions/s per GPU against 31.7 M/s of kernel capacity).  Here the loop never
// touches the interpreter, and the batch is never gathered on the host at all: the agents' state slab is page-locked and mapped
// (cudaHostRegister), and one kernel copies the pending rows straight from host memory into the device batch buffer -- 128 rows
// of 28 KB cross PCIe once, at PCIe speed, with no staging copy.  Replies go back through the slab's reply rows and the agents'
// own semaphores (the `wait_q` of ProcessAgent.py:64), exactly as ga3c_b200.transport.SlabPredictionQueue.reply_batch does.
//
// Serialisation with the trainer: a handle is not re-entrant (one activation workspace); ga3c_lock / ga3c_unlock is the mutex the
// Python Network holds around train() and predict_p_and_v() too.
#include <semaphore.h>
#include <time.h>

#include <atomic>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ga3c_b200.h"
#include "host_util.h"

namespace {

// rows[ids[i]] -> out[i], rows of `row16` 16-byte chunks; `rows` is host memory mapped into the device address space
__global__ void __launch_bounds__(256) gather_rows_kernel(const uint4* __restrict__ rows, const int32_t* __restrict__ ids,
                                                          uint4* __restrict__ out, int row16) {
  const uint4* src = rows + (size_t)ids[blockIdx.x] * row16;
  uint4* dst = out + (size_t)blockIdx.x * row16;
  constexpr int U = 8;                       // 8 x 16 B per thread in flight: PCIe reads are latency-bound
  for (int i0 = threadIdx.x; i0 < row16; i0 += U * 256) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i0 + u * 256 < row16) v[u] = src[i0 + u * 256];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i0 + u * 256 < row16) dst[i0 + u * 256] = v[u];
  }
}

}  // namespace

struct ga3c_batcher {
  ga3c_net* net = nullptr;
  ga3c_batcher_config cfg{};
  std::vector<sem_t*> wake;
  int device = 0;
  cudaStream_t stream = nullptr;
  void* slab_dev = nullptr;                  // device alias of cfg.states
  bool registered = false;
  uint8_t* x_dev = nullptr;                  // [max_batch][state_bytes]
  float *p_dev = nullptr, *v_dev = nullptr;
  int32_t* ids_host = nullptr;               // pinned + mapped
  int32_t* ids_dev = nullptr;
  float *p_host = nullptr, *v_host = nullptr;   // pinned
  std::thread th;
  std::atomic<bool> stop{false};
  std::atomic<long long> batches{0}, rows{0};
  std::atomic<int> error{0};
  std::string error_msg;
  int cursor = 0;
};

static void batcher_loop(ga3c_batcher* b) {
  const ga3c_batcher_config& c = b->cfg;
  cudaSetDevice(b->device);
  sem_t* work = static_cast<sem_t*>(c.work_sem);
  const int n_agents = c.num_agents, A = c.num_actions;
  while (!b->stop.load(std::memory_order_relaxed)) {
    // the semaphore is a wake-up hint (one permit per posted request); the pending bytes are the truth
    timespec ts;
    clock_gettime(CLOCK_REALTIME, &ts);
    ts.tv_nsec += 20 * 1000 * 1000;
    if (ts.tv_nsec >= 1000000000L) { ts.tv_nsec -= 1000000000L; ts.tv_sec += 1; }
    const bool got = sem_timedwait(work, &ts) == 0;
    int n = 0;
    for (int k = 0; k < n_agents && n < c.max_batch; ++k) {       // round robin from the cursor: nobody is starved
      int a = b->cursor + k;
      if (a >= n_agents) a -= n_agents;
      if (reinterpret_cast<volatile uint8_t*>(c.pending)[a]) b->ids_host[n++] = a;
    }
    if (n == 0) continue;
    b->cursor = (b->ids_host[n - 1] + 1) % n_agents;
    for (int k = got ? 1 : 0; k < n; ++k) sem_trywait(work);      // one permit per row taken (best effort)
    std::atomic_thread_fence(std::memory_order_acquire);          // the rows were written before their pending bytes
    int rc = ga3c_lock(b->net);
    if (rc == 0) {
      gather_rows_kernel<<<n, 256, 0, b->stream>>>(static_cast<const uint4*>(b->slab_dev), b->ids_dev,
                                                    reinterpret_cast<uint4*>(b->x_dev), c.state_bytes / 16);
      rc = (int)cudaGetLastError();
      if (rc == 0)
        rc = c.x_u8 ? ga3c_predict_u8(b->net, b->x_dev, n, b->p_dev, b->v_dev, b->stream)
                    : ga3c_predict(b->net, reinterpret_cast<const float*>(b->x_dev), n, b->p_dev, b->v_dev, b->stream);
      if (rc == 0) rc = (int)cudaMemcpyAsync(b->p_host, b->p_dev, (size_t)n * A * 4, cudaMemcpyDeviceToHost, b->stream);
      if (rc == 0) rc = (int)cudaMemcpyAsync(b->v_host, b->v_dev, (size_t)n * 4, cudaMemcpyDeviceToHost, b->stream);
      if (rc == 0) rc = (int)cudaStreamSynchronize(b->stream);
      ga3c_unlock(b->net);
    }
    if (rc != 0) {
      b->error_msg = std::string("native predictor batcher: ") + (ga3c_last_error() ? ga3c_last_error() : "") + " (rc " +
                     std::to_string(rc) + ")";
      b->error.store(rc);
      return;                                                     // the agents time out on their wait_q, the host reads the error
    }
    for (int i = 0; i < n; ++i) {
      const int a = b->ids_host[i];
      memcpy(c.reply_p + (size_t)a * A, b->p_host + (size_t)i * A, (size_t)A * 4);
      c.reply_v[a] = b->v_host[i];
      c.pending[a] = 0;
    }
    std::atomic_thread_fence(std::memory_order_release);          // replies before the wake-ups
    for (int i = 0; i < n; ++i) sem_post(b->wake[b->ids_host[i]]);
    b->batches.fetch_add(1, std::memory_order_relaxed);
    b->rows.fetch_add(n, std::memory_order_relaxed);
  }
}

extern "C" int ga3c_batcher_create(ga3c_net* net, const ga3c_batcher_config* cfg, ga3c_batcher** out) {
  if (!net || !cfg || !out) return ga3c::set_error("ga3c_batcher_create: null argument");
  *out = nullptr;
  if (cfg->num_agents < 1 || cfg->max_batch < 1 || cfg->state_bytes < 16 || cfg->state_bytes % 16 != 0)
    return ga3c::set_error("ga3c_batcher_create: need num_agents, max_batch >= 1 and state rows of a multiple of 16 bytes");
  if (!cfg->states || !cfg->pending || !cfg->reply_p || !cfg->reply_v || !cfg->work_sem || !cfg->wake_sems)
    return ga3c::set_error("ga3c_batcher_create: null slab pointer");
  ga3c_batcher* b = new ga3c_batcher();
  b->net = net;
  b->cfg = *cfg;
  b->wake.assign(reinterpret_cast<sem_t* const*>(cfg->wake_sems), reinterpret_cast<sem_t* const*>(cfg->wake_sems) + cfg->num_agents);
  b->cfg.wake_sems = nullptr;
  b->device = cfg->device;
#define GA3C_B(call)                                                                          \
  do {                                                                                        \
    cudaError_t _e = (call);                                                                  \
    if (_e != cudaSuccess) { ga3c_batcher_destroy(b); return ga3c::fail_cuda(#call, _e); }    \
  } while (0)
  GA3C_B(cudaSetDevice(b->device));
  GA3C_B(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
  const size_t slab_bytes = (size_t)cfg->num_agents * cfg->state_bytes;
  GA3C_B(cudaHostRegister(cfg->states, slab_bytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
  b->registered = true;
  GA3C_B(cudaHostGetDevicePointer(&b->slab_dev, cfg->states, 0));
  GA3C_B(cudaMalloc((void**)&b->x_dev, (size_t)cfg->max_batch * cfg->state_bytes));
  GA3C_B(cudaMalloc((void**)&b->p_dev, (size_t)cfg->max_batch * cfg->num_actions * 4));
  GA3C_B(cudaMalloc((void**)&b->v_dev, (size_t)cfg->max_batch * 4));
  GA3C_B(cudaHostAlloc((void**)&b->ids_host, (size_t)cfg->max_batch * 4, cudaHostAllocMapped));
  GA3C_B(cudaHostGetDevicePointer((void**)&b->ids_dev, b->ids_host, 0));
  GA3C_B(cudaHostAlloc((void**)&b->p_host, (size_t)cfg->max_batch * cfg->num_actions * 4, cudaHostAllocDefault));
  GA3C_B(cudaHostAlloc((void**)&b->v_host, (size_t)cfg->max_batch * 4, cudaHostAllocDefault));
#undef GA3C_B
  if (int r = ga3c_reserve(net, cfg->max_batch)) { ga3c_batcher_destroy(b); return r; }
  *out = b;
  return 0;
}

extern "C" int ga3c_batcher_start(ga3c_batcher* b) {
  if (!b) return ga3c::set_error("ga3c_batcher_start: null handle");
  if (b->th.joinable()) return ga3c::set_error("ga3c_batcher_start: already running");
  b->stop.store(false);
  b->error.store(0);
  b->th = std::thread(batcher_loop, b);
  return 0;
}

extern "C" int ga3c_batcher_stop(ga3c_batcher* b) {
  if (!b) return 0;
  b->stop.store(true);
  if (b->th.joinable()) b->th.join();
  return 0;
}

extern "C" int ga3c_batcher_stats(ga3c_batcher* b, int64_t* batches, int64_t* rows, int32_t* error) {
  if (!b) return ga3c::set_error("ga3c_batcher_stats: null handle");
  if (batches) *batches = b->batches.load();
  if (rows) *rows = b->rows.load();
  if (error) *error = b->error.load();
  if (b->error.load() != 0) return ga3c::set_error(b->error_msg);
  return 0;
}

extern "C" int ga3c_batcher_destroy(ga3c_batcher* b) {
  if (!b) return 0;
  ga3c_batcher_stop(b);
  cudaSetDevice(b->device);
  if (b->stream) { cudaStreamSynchronize(b->stream); cudaStreamDestroy(b->stream); }
  if (b->registered) cudaHostUnregister(b->cfg.states);
  cudaFree(b->x_dev); cudaFree(b->p_dev); cudaFree(b->v_dev);
  if (b->ids_host) cudaFreeHost(b->ids_host);
  if (b->p_host) cudaFreeHost(b->p_host);
  if (b->v_host) cudaFreeHost(b->v_host);
  delete b;
  return 0;
}
