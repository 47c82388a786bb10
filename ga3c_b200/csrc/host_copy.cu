// Host staging of pageable caller arrays (ga3c_stage_h2d).  Host code only: no kernels in this file.
//
// The reference's ThreadTrainer hands Network.train ordinary numpy arrays -- np.concatenate output, ThreadTrainer.py:54-58 -- and
// a cudaMemcpyAsync from pageable memory runs at a fraction of the PCIe rate (the driver stages it through its own bounce
// buffers with one thread).  Here a small pool of worker threads copies the array into the caller's page-locked staging buffer
// in 512 KB pieces, claimed in address order, with non-temporal stores (the destination is read next by the DMA engine, not by a
// core: no read-for-ownership, no cache pollution), and the calling thread enqueues the DMA of every chunk as soon as the
// chunk's pieces are in place -- the copy engine works on chunk k while the cores fill chunk k + 1.  The interpreter is not
// involved (ctypes releases the GIL for the call), so the second trainer thread's step runs underneath.
#include <cuda_runtime.h>
#include <emmintrin.h>
#include <sched.h>

#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/ga3c_b200.h"
#include "host_util.h"

namespace {

constexpr int64_t PIECE = 512 << 10;
constexpr int MAX_WORKERS = 32;

// GA3C_COPY_NT=0: plain memcpy per piece (on a host whose last-level cache holds a whole batch -- and whose DMA reads are served
// from that cache -- ordinary stores can beat streaming ones); read per call so that a bench can compare both in one process
bool use_streaming_stores() {
  const char* e = getenv("GA3C_COPY_NT");
  return !(e && atoi(e) == 0);
}

// dst, src any alignment; streaming (non-temporal) 16-byte stores for the aligned body
void stream_copy(uint8_t* d, const uint8_t* s, size_t n, bool nt) {
  if (!nt) { std::memcpy(d, s, n); return; }
  size_t head = (64 - (reinterpret_cast<uintptr_t>(d) & 63)) & 63;
  if (head > n) head = n;
  if (head) { std::memcpy(d, s, head); d += head; s += head; n -= head; }
  size_t blocks = n / 64;
  for (size_t i = 0; i < blocks; ++i) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s));
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 16));
    const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 32));
    const __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 48));
    _mm_stream_si128(reinterpret_cast<__m128i*>(d), a);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + 16), b);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + 32), c);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + 48), e);
    s += 64; d += 64;
  }
  _mm_sfence();                                  // the pieces' stores are globally visible before the piece is reported done
  if (n & 63) std::memcpy(d, s, n & 63);
}

struct Job {
  uint8_t* dst = nullptr;
  const uint8_t* src = nullptr;
  int64_t bytes = 0, n_pieces = 0, pieces_per_chunk = 1;
  int workers = 0;
  bool nt = true;
  std::atomic<int64_t> next{0};
  std::atomic<int> active{0};
  std::unique_ptr<std::atomic<int>[]> done;      // pieces finished, per chunk
};

class Pool {
 public:
  // never destroyed: a trainer thread may still be inside a job when the process exits (daemon threads), and a static destructor
  // joining the workers underneath it would race with that call; the OS reaps the workers with the process
  static Pool& get() { static Pool* p = new Pool; return *p; }

  // copies the whole job; `on_chunk(c)` is called by the calling thread, in order, as soon as chunk c is complete
  template <class F>
  void run(uint8_t* dst, const uint8_t* src, int64_t bytes, int64_t chunk_bytes, int threads, F on_chunk) {
    std::lock_guard<std::mutex> one_job(job_mutex_);
    const int64_t ppc = chunk_bytes / PIECE > 0 ? chunk_bytes / PIECE : 1;
    const int64_t n_pieces = (bytes + PIECE - 1) / PIECE, n_chunks = (n_pieces + ppc - 1) / ppc;
    if (threads > MAX_WORKERS) threads = MAX_WORKERS;
    if ((int64_t)threads > n_pieces) threads = (int)n_pieces;
    ensure_workers(threads);
    job_.dst = dst; job_.src = src; job_.bytes = bytes; job_.n_pieces = n_pieces; job_.pieces_per_chunk = ppc;
    job_.workers = threads;
    job_.nt = use_streaming_stores();
    job_.next.store(0);
    job_.done.reset(new std::atomic<int>[n_chunks]);
    for (int64_t c = 0; c < n_chunks; ++c) job_.done[c].store(0);
    job_.active.store(threads);
    {
      std::lock_guard<std::mutex> l(m_);
      ++generation_;
    }
    cv_.notify_all();
    for (int64_t c = 0; c < n_chunks; ++c) {
      const int64_t in_chunk = (c + 1 < n_chunks) ? ppc : n_pieces - c * ppc;
      int spins = 0;
      while (job_.done[c].load(std::memory_order_acquire) < in_chunk) {
        if (++spins < 2000) _mm_pause();
        else { sched_yield(); spins = 0; }
      }
      on_chunk(c);
    }
    while (job_.active.load(std::memory_order_acquire) != 0) _mm_pause();     // no worker still looks at this job
  }

 private:
  Pool() = default;

  void ensure_workers(int n) {
    while ((int)threads_.size() < n) {
      const int id = (int)threads_.size();
      uint64_t seen;
      {
        std::lock_guard<std::mutex> l(m_);
        seen = generation_;
      }
      threads_.emplace_back([this, id, seen]() mutable { worker(id, seen); });
    }
  }

  void worker(int id, uint64_t seen) {
    for (;;) {
      {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [&] { return stop_ || generation_ != seen; });
        if (stop_) return;
        seen = generation_;
      }
      if (id >= job_.workers) continue;
      for (;;) {
        const int64_t p = job_.next.fetch_add(1, std::memory_order_relaxed);
        if (p >= job_.n_pieces) break;
        const int64_t off = p * PIECE, n = (off + PIECE <= job_.bytes) ? PIECE : job_.bytes - off;
        stream_copy(job_.dst + off, job_.src + off, (size_t)n, job_.nt);
        job_.done[p / job_.pieces_per_chunk].fetch_add(1, std::memory_order_release);
      }
      job_.active.fetch_sub(1, std::memory_order_release);
    }
  }

  std::mutex job_mutex_, m_;
  std::condition_variable cv_;
  std::vector<std::thread> threads_;
  uint64_t generation_ = 0;
  bool stop_ = false;
  Job job_;
};

}  // namespace

extern "C" int ga3c_stage_h2d(void* dst_dev, void* staging, const void* src, int64_t bytes, int64_t chunk_bytes, int32_t threads,
                              void* stream) {
  if (bytes < 0 || chunk_bytes < 0 || threads < 0) return ga3c::set_error("ga3c_stage_h2d: negative size");
  if (bytes == 0) return 0;
  if (!staging || !src) return ga3c::set_error("ga3c_stage_h2d: null buffer");
  if (chunk_bytes == 0) chunk_bytes = 8 << 20;
  chunk_bytes = (chunk_bytes + PIECE - 1) / PIECE * PIECE;
  uint8_t* st = static_cast<uint8_t*>(staging);
  const uint8_t* s = static_cast<const uint8_t*>(src);
  uint8_t* d = static_cast<uint8_t*>(dst_dev);
  cudaError_t err = cudaSuccess;
  auto dma = [&](int64_t c) {
    if (!d || err != cudaSuccess) return;
    const int64_t off = c * chunk_bytes, n = (off + chunk_bytes <= bytes) ? chunk_bytes : bytes - off;
    err = cudaMemcpyAsync(d + off, st + off, (size_t)n, cudaMemcpyHostToDevice, (cudaStream_t)stream);
  };
  if (threads <= 1 || bytes <= PIECE) {
    const int64_t n_chunks = (bytes + chunk_bytes - 1) / chunk_bytes;
    for (int64_t c = 0; c < n_chunks; ++c) {
      const int64_t off = c * chunk_bytes, n = (off + chunk_bytes <= bytes) ? chunk_bytes : bytes - off;
      stream_copy(st + off, s + off, (size_t)n, use_streaming_stores());
      dma(c);
    }
  } else {
    Pool::get().run(st, s, bytes, chunk_bytes, threads, dma);
  }
  if (err != cudaSuccess) return ga3c::fail_cuda("ga3c_stage_h2d: cudaMemcpyAsync", err);
  return 0;
}
