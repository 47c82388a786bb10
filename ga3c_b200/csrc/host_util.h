// Host-side helpers shared by the C-ABI layers (net.cu, mlp_net.cu): error reporting behind ga3c_last_error and the launch
// bookkeeping (launch counter + optional per-kernel CUDA-event brackets, ga3c_timing_* / ga3c_mlp_timing_*).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

namespace ga3c {

int set_error(const std::string& m);                 // net.cu: thread-local message, returns -1
inline int fail_cuda(const char* where, cudaError_t e) {
  set_error(std::string(where) + ": " + cudaGetErrorString(e));
  return (int)e ? (int)e : -1;
}

#define CK(call)                                                  \
  do {                                                            \
    cudaError_t _e = (call);                                      \
    if (_e != cudaSuccess) return ::ga3c::fail_cuda(#call, _e);   \
  } while (0)
#define CKL(call)                                                         \
  do {                                                                    \
    int _r = (call);                                                      \
    if (_r != 0) return ::ga3c::fail_cuda(#call, (cudaError_t)_r);        \
  } while (0)

// per-handle launch bookkeeping: record r uses events 2r (before) and 2r+1 (after) on the launch stream
struct LaunchLog {
  long long launches = 0;
  std::vector<cudaEvent_t> tev;
  std::vector<int> tkid;
  int tcursor = 0;

  void clear() {
    for (cudaEvent_t e : tev) cudaEventDestroy(e);
    tev.clear(); tkid.clear(); tcursor = 0;
  }
  int enable(int max_records) {                      // 0 switches timing off; caller has synchronised the device
    clear();
    tev.resize((size_t)2 * max_records);
    tkid.assign((size_t)max_records, 0);
    for (auto& e : tev) CK(cudaEventCreate(&e));
    return 0;
  }
  // sums the durations per kernel id and rewinds the cursor; caller has synchronised the device
  int collect(double* total_ms, long long* counts, int n_kernels) {
    for (int k = 0; k < n_kernels; ++k) { total_ms[k] = 0.0; counts[k] = 0; }
    for (int r = 0; r < tcursor; ++r) {
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, tev[2 * r], tev[2 * r + 1]));
      total_ms[tkid[r]] += ms;
      counts[tkid[r]] += 1;
    }
    tcursor = 0;
    return 0;
  }
};

// launch one kernel of the path; when timing is enabled bracket it with events on the same stream
#define LAUNCH(net, kid, st, call)                                                                          \
  do {                                                                                                      \
    ::ga3c::LaunchLog& _l = (net)->log;                                                                     \
    const bool _t = !_l.tev.empty() && (size_t)(2 * _l.tcursor + 1) < _l.tev.size();                        \
    if (_t) CK(cudaEventRecord(_l.tev[2 * _l.tcursor], (st)));                                              \
    CKL(call);                                                                                              \
    if (_t) {                                                                                               \
      CK(cudaEventRecord(_l.tev[2 * _l.tcursor + 1], (st)));                                                \
      _l.tkid[_l.tcursor++] = (kid);                                                                        \
    }                                                                                                       \
    _l.launches++;                                                                                          \
  } while (0)

}  // namespace ga3c
