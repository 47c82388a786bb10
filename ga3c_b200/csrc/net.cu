// C-ABI host layer: handle, flat arenas, workspace, and the launch sequences of the hot path.
// See include/ga3c_b200.h for the contract and the reference call sites each entry replaces.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/ga3c_b200.h"
#include "common.cuh"
#include "kernels.h"
#include "host_util.h"
#include "mlp.cuh"
#include "dp_exchange.cuh"

using namespace ga3c;

namespace {

thread_local std::string g_err;

int fail(const char* where, cudaError_t e) { return fail_cuda(where, e); }
int fail_msg(const std::string& m) { g_err = m; return -1; }

struct ParamDesc {
  std::string name;
  int64_t offset;
  int32_t ndim;
  int64_t shape[4];
  int64_t count;
};

constexpr int64_t ALIGN_FLOATS = 64;
int64_t align_up(int64_t v) { return (v + ALIGN_FLOATS - 1) / ALIGN_FLOATS * ALIGN_FLOATS; }

}  // namespace

struct ga3c_net {
  ga3c_config cfg;
  int num_sms = 148;
  std::vector<ParamDesc> params;   // TF creation order
  int64_t arena_floats = 0;
  int64_t small_floats = 0;        // prefix holding every tensor except dense1/w (grads zeroed per step)
  // one slab holds [params | grads | ms | mom | bf16 shadow of dense1/w | LL receive buffers | comm flags]: a single CUDA IPC handle
  // exposes everything a data-parallel peer needs (ga3c_dp_*)
  uint8_t* slab = nullptr;
  size_t slab_bytes = 0;
  float *w = nullptr, *g = nullptr, *ms = nullptr, *mom = nullptr;
  uint16_t* w1_shadow = nullptr;   // bf16 [3872,256]
  // data parallel over peer memory (single node, <= 8 ranks)
  int dp_rank = 0, dp_world = 1;
  uint8_t* dp_peer[DP_MAX_WORLD] = {};   // slab base of every rank as mapped in this process (own slab at dp_rank)
  bool dp_ipc[DP_MAX_WORLD] = {};        // mapped with cudaIpcOpenMemHandle (to be closed on detach)
  uint64_t dp_step = 0;
  int dp_exch = 20;                // overlapped exchange: exchange CTAs appended to the conv backward launch (GA3C_DP_EXCH_CTAS)
  int dp_exchange = 4;             // GA3C_DP_EXCHANGE: 4 "side" (default: dense1/w exchanged by a small-footprint kernel on a side stream
                                   // that is resident next to the conv backward of this step and the conv forward of the next;
                                   // small tensors in dp_small), 0 "tail" (one launch on every SM at the end of the step), 3 "warps" (the
                                   // dense1/w exchange on the optimizer warps of every conv backward CTA, small tensors + completion
                                   // in dp_small), 1 "overlap" (exchange CTAs inside the conv backward launch + dp_small), 2
                                   // "single" (the first version: one kernel that also broadcasts the fp32 weights).  Measured at 2
                                   // GPUs, B = 1024 per GPU (profiles/r2l_*): tail 0.1196 ms, warps 0.1205, overlap 0.1403
  int fuse_heads = 0;              // GA3C_FUSE_HEADS=1: dense1 forward + heads as ONE cluster launch (dense_heads.cu: split-K over a
                                   // thread-block cluster, partial tiles reduced over distributed shared memory).  Parity-green, but
                                   // measured slower at B = 1024 (step 0.1015 -> 0.1077 ms, profiles/r2r_*): 64 CTAs run the heads of
                                   // 16 samples each in two serial rounds where the separate kernel runs 128 CTAs of one round
  int fuse_opt = 0;                // GA3C_FUSE_OPT=1: single GPU, dense1/w updated by the optimizer warps of the conv backward CTAs
                                   // instead of the launch at the end.  Measured: conv_bwd 26.3 -> 30.6 us, rmsprop 8.9 -> 5.0 us,
                                   // step 0.1015 -> 0.1032 ms: the conv backward is HBM-bound, the 22 MB cost what they cost alone
  int cur_exch = 0;                // ... of the step being enqueued (0 outside the overlapped data-parallel step)
  int64_t xbuf_off = 0, comm_off = 0;    // byte offsets in the slab: LL receive buffers [2][8][small prefix * 8 B], comm block
  int64_t bigrecv_off = 0;               // ... and the LL receive buffers of the pushed dense1/w gradient slices [2][(n4 + 8) * 32 B]
  int dp_push = 0;                       // GA3C_DP_PUSH=1 (with GA3C_DP_EXCHANGE=tail): slices pushed by the dense1 wgrad epilogues in the
                                         // LL format instead of pulled by their owners.  Measured at 2 GPUs: 0.1463 vs 0.1181 ms per
                                         // step -- scattered 16-byte stores across NVLink from the GEMM epilogue triple its time
  cudaStream_t dp_stream = nullptr;      // GA3C_DP_EXCHANGE=side: the dense1/w exchange runs here, next to the conv kernels
  cudaEvent_t dp_big_done = nullptr;     // ... and whatever reads the dense1/w shadow next waits for this
  bool dp_big_pending = false;
  cudaEvent_t dp_grad_ready = nullptr;   // recorded behind dense_bwd: the side stream waits for it
  const DpBigArgs* dp_side = nullptr;    // set while a side-mode step is enqueued: fb_tail_impl launches the exchange behind dense_bwd
  // workspace
  uint16_t *n2 = nullptr, *dd1 = nullptr, *dn1 = nullptr;
  // training-only activations, stored in the UMMA operand layouts the conv backward consumes (common.cuh); their zero borders /
  // slack rows are never written, so they are cleared once, at allocation
  uint8_t *n1 = nullptr, *dn2 = nullptr, *xblk = nullptr;
  float* d1 = nullptr;
  float* d1_part = nullptr;        // [splits][B,256] raw split-K partials of dense1 (dense_tc.cu)
  // gradient partials: one slab per CTA of the heads / conv backward kernels, laid out like the small-tensor prefix
  // of the gradient arena followed by the 4 loss sums; summed by grad_reduce at the end of ga3c_fb_tail
  float* gpart = nullptr;
  int64_t gp_stride = 0;
  int gp_heads_grid = 0;           // slabs the heads kernel of the current step wrote
  float* loss_out = nullptr;       // caller's loss buffer of the current step (may be null)
  float *clip_ss = nullptr, *clip_scale = nullptr;   // Config.USE_GRAD_CLIP scratch: chunk sums of squares, per-tensor scale
  float *clip_ss2 = nullptr, *clip_scale2 = nullptr; // ... of the second optimizer's gradient (DUAL_RMSPROP + USE_GRAD_CLIP)
  float *g2 = nullptr, *ms2 = nullptr, *mom2 = nullptr;   // Config.DUAL_RMSPROP: gradient of cost_v and the second optimizer's slots
  bool keep_dn1 = false;           // tests: also store dn1 (which otherwise never leaves the SM) to the workspace
  int64_t global_step = 0;
  std::mutex mtx;                  // ga3c_lock / ga3c_unlock
  LaunchLog log;                   // launch counter + per-kernel CUDA-event timing (ga3c_timing_*)
  int last_batch = 0;
  unsigned long long* evt = nullptr;       // pipeline event log of CTA 0 (ga3c_evt_*), device memory
  unsigned long long* trace = nullptr;     // [K_COUNT][TRACE_SLOTS] globaltimer stamps (ga3c_trace_*), device memory

  int64_t off(int i) const { return params[i].offset; }
};


static const char* const kKernelNames[K_COUNT] = {"conv_fwd", "dense_fwd", "heads", "dense_wgrad", "dense_bwd",
                                                  "conv_bwd", "conv11_wgrad", "rmsprop", "grad_reduce",
                                                  "mlp_fused", "mlp_wgrad", "mlp_reduce", "dp_big", "mlp_tc"};

enum { P_C11W = 0, P_C11B, P_C12W, P_C12B, P_D1W, P_D1B, P_VW, P_VB, P_PW, P_PB, P_COUNT };

static int alloc_workspace(ga3c_net* n, int max_batch);
static int trace_attach_all(unsigned long long* buf);
static int dp_attach_finish(ga3c_net* n, int32_t rank, int32_t world);

namespace ga3c {
int set_error(const std::string& m) { g_err = m; return -1; }     // for the other host files (mlp_net.cu)
const char* kernel_name(int kid) { return (kid >= 0 && kid < K_COUNT) ? kKernelNames[kid] : nullptr; }
}
extern "C" const char* ga3c_last_error(void) { return g_err.c_str(); }
extern "C" int ga3c_abi_version(void) { return 6; }

extern "C" int ga3c_create(const ga3c_config* cfg, ga3c_net** out) {
  if (!cfg || !out) return fail_msg("ga3c_create: null argument");
  *out = nullptr;
  if (cfg->num_actions < 1 || cfg->num_actions > MAX_ACTIONS) return fail_msg("ga3c_create: num_actions must be 1..18");
  if (cfg->max_batch < 1) return fail_msg("ga3c_create: max_batch must be >= 1");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return fail_msg("ga3c_create: no CUDA device (there is no CPU fallback)");
  if (cfg->device < 0 || cfg->device >= ndev) return fail_msg("ga3c_create: device ordinal out of range");
  CK(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) return fail_msg("ga3c_create: kernels are built for sm_100a only; device is sm_" +
                                        std::to_string(prop.major) + std::to_string(prop.minor));
  ga3c_net* n = new ga3c_net();
  n->cfg = *cfg;
  n->num_sms = prop.multiProcessorCount;
  const int A = cfg->num_actions;
  // TF creation order (NetworkVP.py:284-285 would list them so); offsets pack the small tensors first
  struct { const char* name; int nd; int64_t s[4]; } spec[P_COUNT] = {
      {"conv11/w:0", 4, {8, 8, 4, 16}}, {"conv11/b:0", 1, {16, 0, 0, 0}},
      {"conv12/w:0", 4, {4, 4, 16, 32}}, {"conv12/b:0", 1, {32, 0, 0, 0}},
      {"dense1/w:0", 2, {FLAT, FC, 0, 0}}, {"dense1/b:0", 1, {FC, 0, 0, 0}},
      {"logits_v/w:0", 2, {FC, 1, 0, 0}}, {"logits_v/b:0", 1, {1, 0, 0, 0}},
      {"logits_p/w:0", 2, {FC, A, 0, 0}}, {"logits_p/b:0", 1, {A, 0, 0, 0}}};
  n->params.resize(P_COUNT);
  for (int i = 0; i < P_COUNT; ++i) {
    ParamDesc& d = n->params[i];
    d.name = spec[i].name;
    d.ndim = spec[i].nd;
    d.count = 1;
    for (int k = 0; k < 4; ++k) { d.shape[k] = spec[i].s[k]; if (k < d.ndim) d.count *= d.shape[k]; }
  }
  int64_t cur = 0;
  const int order[P_COUNT] = {P_C11W, P_C11B, P_C12W, P_C12B, P_D1B, P_VW, P_VB, P_PW, P_PB, P_D1W};
  for (int k = 0; k < P_COUNT; ++k) {
    if (order[k] == P_D1W) n->small_floats = cur;
    n->params[order[k]].offset = cur;
    cur = align_up(cur + n->params[order[k]].count);
  }
  n->arena_floats = cur;

  const size_t ab = (size_t)n->arena_floats * sizeof(float);
#define GA3C_ALLOC(ptr, bytes)                                                   \
  do {                                                                           \
    cudaError_t _e = cudaMalloc((void**)&(ptr), (bytes));                        \
    if (_e != cudaSuccess) { ga3c_destroy(n); return fail("cudaMalloc", _e); }   \
  } while (0)
  const size_t shadow_bytes = (size_t)FLAT * FC * 2;
  n->xbuf_off = (int64_t)(4 * ab + shadow_bytes);
  n->comm_off = n->xbuf_off + 2 * (int64_t)DP_MAX_WORLD * n->small_floats * 8;
  n->bigrecv_off = (n->comm_off + DP_COMM_BYTES + 127) / 128 * 128;
  n->slab_bytes = (size_t)n->bigrecv_off + 2 * ((size_t)FLAT * FC / 4 + DP_MAX_WORLD) * 32;
  GA3C_ALLOC(n->slab, n->slab_bytes);
#undef GA3C_ALLOC
  n->w = reinterpret_cast<float*>(n->slab);
  n->g = reinterpret_cast<float*>(n->slab + ab);
  n->ms = reinterpret_cast<float*>(n->slab + 2 * ab);
  n->mom = reinterpret_cast<float*>(n->slab + 3 * ab);
  n->w1_shadow = reinterpret_cast<uint16_t*>(n->slab + 4 * ab);
  cudaMemset(n->slab + n->xbuf_off, 0, n->slab_bytes - (size_t)n->xbuf_off);
  n->gp_stride = n->small_floats + 64;
  e = cudaMalloc((void**)&n->gpart, (size_t)2 * n->num_sms * n->gp_stride * sizeof(float));
  if (e != cudaSuccess) { ga3c_destroy(n); return fail("cudaMalloc", e); }
  cudaMemset(n->gpart, 0, (size_t)2 * n->num_sms * n->gp_stride * sizeof(float));
  if (cfg->dual_rmsprop) {
    float** extra[3] = {&n->g2, &n->ms2, &n->mom2};
    for (float** a : extra) {
      e = cudaMalloc((void**)a, ab);
      if (e != cudaSuccess) { ga3c_destroy(n); return fail("cudaMalloc", e); }
      cudaMemset(*a, 0, ab);
    }
    std::vector<float> ones((size_t)n->arena_floats, 1.0f);
    cudaMemcpy(n->ms2, ones.data(), ab, cudaMemcpyHostToDevice);
  }
  if (cfg->use_grad_clip) {
    e = cudaMalloc((void**)&n->clip_ss, (size_t)P_COUNT * clip_chunks((int64_t)FLAT * FC) * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&n->clip_scale, P_COUNT * sizeof(float));
    if (e == cudaSuccess && cfg->dual_rmsprop) {
      e = cudaMalloc((void**)&n->clip_ss2, (size_t)P_COUNT * clip_chunks((int64_t)FLAT * FC) * sizeof(float));
      if (e == cudaSuccess) e = cudaMalloc((void**)&n->clip_scale2, P_COUNT * sizeof(float));
    }
    if (e != cudaSuccess) { ga3c_destroy(n); return fail("cudaMalloc", e); }
  }
  if (int r = alloc_workspace(n, cfg->max_batch)) { ga3c_destroy(n); return r; }
  if (const char* f = getenv("GA3C_FUSE_OPT")) n->fuse_opt = atoi(f) != 0;
  if (const char* f = getenv("GA3C_FUSE_HEADS")) n->fuse_heads = atoi(f) != 0;
  cudaMemset(n->w, 0, ab); cudaMemset(n->g, 0, ab); cudaMemset(n->mom, 0, ab);
  cudaMemset(n->w1_shadow, 0, (size_t)FLAT * FC * 2);
  {  // ms slot starts at 1.0 [TF-SEMANTICS]
    std::vector<float> ones((size_t)n->arena_floats, 1.0f);
    cudaMemcpy(n->ms, ones.data(), ab, cudaMemcpyHostToDevice);
  }
  int r;
  if ((r = configure_conv_fwd()) || (r = configure_conv_bwd_fused()) || (r = configure_dense_tc()) || (r = configure_dense_heads())) {
    ga3c_destroy(n);
    return fail("cudaFuncSetAttribute", (cudaError_t)r);
  }
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { ga3c_destroy(n); return fail("ga3c_create sync", e); }
  *out = n;
  return 0;
}

static void free_workspace(ga3c_net* n) {
  cudaFree(n->n1); cudaFree(n->n2); cudaFree(n->d1); cudaFree(n->dd1); cudaFree(n->dn2); cudaFree(n->dn1); cudaFree(n->xblk);
  cudaFree(n->d1_part);
  n->n2 = n->dd1 = n->dn1 = nullptr; n->n1 = n->dn2 = n->xblk = nullptr; n->d1 = n->d1_part = nullptr;
}

static int alloc_workspace(ga3c_net* n, int max_batch) {
  const size_t mb = (size_t)max_batch;
  CK(cudaMalloc((void**)&n->n1, mb * B2_BYTES)); CK(cudaMalloc((void**)&n->n2, mb * FLAT * 2));
  CK(cudaMalloc((void**)&n->d1, mb * FC * 4)); CK(cudaMalloc((void**)&n->dd1, mb * FC * 2));
  CK(cudaMalloc((void**)&n->dn2, mb * G_BYTES)); CK(cudaMalloc((void**)&n->dn1, mb * N1_POS * C1_OUT * 2));
  CK(cudaMalloc((void**)&n->xblk, mb * XB_FRAME_BYTES));
  CK(cudaMemset(n->n1, 0, mb * B2_BYTES)); CK(cudaMemset(n->dn2, 0, mb * G_BYTES)); CK(cudaMemset(n->xblk, 0, mb * XB_FRAME_BYTES));
  // splits * batch <= max(batch, 64 * num_sms) rows for every batch (dense_fwd_splits)
  const size_t part_rows = mb > (size_t)64 * n->num_sms ? mb : (size_t)64 * n->num_sms;
  CK(cudaMalloc((void**)&n->d1_part, part_rows * FC * 4));
  n->cfg.max_batch = max_batch;
  return 0;
}

extern "C" int ga3c_reserve(ga3c_net* n, int32_t max_batch) {
  if (!n) return fail_msg("ga3c_reserve: null handle");
  if (max_batch <= n->cfg.max_batch) return 0;
  CK(cudaSetDevice(n->cfg.device));
  CK(cudaDeviceSynchronize());
  free_workspace(n);
  if (int r = alloc_workspace(n, max_batch)) { n->cfg.max_batch = 0; return r; }
  return 0;
}

extern "C" int ga3c_destroy(ga3c_net* n) {
  if (!n) return 0;
  ga3c_dp_detach(n);
  cudaFree(n->slab);
  cudaFree(n->gpart);
  cudaFree(n->clip_ss); cudaFree(n->clip_scale); cudaFree(n->clip_ss2); cudaFree(n->clip_scale2);
  cudaFree(n->g2); cudaFree(n->ms2); cudaFree(n->mom2);
  if (n->trace) { trace_attach_all(nullptr); cudaFree(n->trace); }
  free_workspace(n);
  n->log.clear();
  delete n;
  return 0;
}

extern "C" int ga3c_lock(ga3c_net* n) {
  if (!n) return fail_msg("ga3c_lock: null handle");
  n->mtx.lock();
  return 0;
}
extern "C" int ga3c_unlock(ga3c_net* n) {
  if (!n) return fail_msg("ga3c_unlock: null handle");
  n->mtx.unlock();
  return 0;
}

extern "C" int ga3c_param_count(const ga3c_net* n) { return n ? (int)n->params.size() : 0; }

extern "C" int ga3c_param_info(const ga3c_net* n, int i, const char** name, int64_t* offset, int32_t* ndim,
                               int64_t shape[4]) {
  if (!n || i < 0 || i >= (int)n->params.size()) return fail_msg("ga3c_param_info: bad index");
  const ParamDesc& d = n->params[i];
  if (name) *name = d.name.c_str();
  if (offset) *offset = d.offset;
  if (ndim) *ndim = d.ndim;
  if (shape) for (int k = 0; k < 4; ++k) shape[k] = d.shape[k];
  return 0;
}

extern "C" int64_t ga3c_arena_floats(const ga3c_net* n) { return n ? n->arena_floats : 0; }

extern "C" int ga3c_arena_ptrs(ga3c_net* n, float** p, float** g, float** ms, float** mom) {
  if (!n) return fail_msg("ga3c_arena_ptrs: null handle");
  if (p) *p = n->w;
  if (g) *g = n->g;
  if (ms) *ms = n->ms;
  if (mom) *mom = n->mom;
  return 0;
}

static float* arena_of(ga3c_net* n, int which) {
  switch (which) {
    case 0: return n->w; case 1: return n->g; case 2: return n->ms; case 3: return n->mom;
    case 4: return n->g2; case 5: return n->ms2; case 6: return n->mom2;      // null unless dual_rmsprop
  }
  return nullptr;
}

extern "C" int ga3c_arena_ptr(ga3c_net* n, int which, float** ptr_dev) {
  if (!n || !ptr_dev) return fail_msg("ga3c_arena_ptr: null argument");
  *ptr_dev = arena_of(n, which);
  if (!*ptr_dev) return fail_msg("ga3c_arena_ptr: bad arena id");
  return 0;
}

extern "C" int ga3c_arena_upload(ga3c_net* n, int which, const float* host, int64_t nf) {
  if (!n || !host) return fail_msg("ga3c_arena_upload: null argument");
  float* dst = arena_of(n, which);
  if (!dst || nf != n->arena_floats) return fail_msg("ga3c_arena_upload: bad arena id or size");
  CK(cudaSetDevice(n->cfg.device));
  CK(cudaMemcpy(dst, host, (size_t)nf * 4, cudaMemcpyHostToDevice));
  if (which == 0) {
    CKL(launch_f32_to_bf16(n->w + n->off(P_D1W), n->w1_shadow, (int64_t)FLAT * FC, 0));
    n->log.launches++;
    CK(cudaDeviceSynchronize());
  }
  return 0;
}

extern "C" int ga3c_arena_download(ga3c_net* n, int which, float* host, int64_t nf) {
  if (!n || !host) return fail_msg("ga3c_arena_download: null argument");
  float* src = arena_of(n, which);
  if (!src || nf != n->arena_floats) return fail_msg("ga3c_arena_download: bad arena id or size");
  CK(cudaSetDevice(n->cfg.device));
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(host, src, (size_t)nf * 4, cudaMemcpyDeviceToHost));
  if (n->dp_world > 1 && (which == 0 || which == 2 || which == 3)) {
    // overlapped exchange: the fp32 master copy and the RMSProp slots of a dense1/w slice live on its owner (only the bf16
    // shadow travels).  A peer's slice is stable here: it changes only inside a step this rank takes part in.
    const int64_t n4 = (int64_t)FLAT * FC / 4, per = (n4 + n->dp_world - 1) / n->dp_world;
    for (int q = 0; q < n->dp_world; ++q) {
      if (q == n->dp_rank) continue;
      const int64_t lo = per * q, hi = lo + per < n4 ? lo + per : n4;
      if (hi <= lo) continue;
      const int64_t off = n->off(P_D1W) + lo * 4;
      CK(cudaMemcpy(host + off, n->dp_peer[q] + (size_t)which * nf * 4 + (size_t)off * 4, (size_t)(hi - lo) * 16,
                    cudaMemcpyDeviceToHost));
    }
  }
  return 0;
}

extern "C" int64_t ga3c_global_step(const ga3c_net* n) { return n ? n->global_step : -1; }
extern "C" int ga3c_set_global_step(ga3c_net* n, int64_t s) { if (!n) return -1; n->global_step = s; return 0; }

static int check_batch(ga3c_net* n, int batch, const char* who) {
  if (!n) return fail_msg(std::string(who) + ": null handle");
  if (batch < 1 || batch > n->cfg.max_batch)
    return fail_msg(std::string(who) + ": batch " + std::to_string(batch) + " outside 1.." + std::to_string(n->cfg.max_batch));
  return 0;
}

static HeadsArgs heads_args(ga3c_net* n, int batch, int splits) {
  HeadsArgs h{};
  h.d1 = n->d1; h.d1_part = n->d1_part; h.n_split = splits; h.b1 = n->w + n->off(P_D1B);
  h.wp = n->w + n->off(P_PW); h.bp = n->w + n->off(P_PB);
  h.wv = n->w + n->off(P_VW); h.bv = n->w + n->off(P_VB);
  h.batch = batch; h.num_actions = n->cfg.num_actions;
  h.log_eps = n->cfg.log_epsilon; h.min_policy = n->cfg.min_policy; h.log_softmax = n->cfg.use_log_softmax != 0;
  h.preload = batch >= n->num_sms;
  return h;
}

// data parallel, side-stream exchange: the bf16 shadow of dense1/w is complete when the exchange kernel of the previous step is
static int wait_dp_big(ga3c_net* n, cudaStream_t st) {
  // (not cleared: predict and train may come on different streams, and each of them has to see the exchange complete)
  if (n->dp_big_pending) CK(cudaStreamWaitEvent(st, n->dp_big_done, 0));
  return 0;
}
// ... or, by default (GA3C_DP_GATE=0 brings the event wait back), when dense_fwd's TMA producer has seen every rank's "slice
// landed everywhere" flag of the last exchanged step in this rank's comm block (DpGate, dense_tc.cu): no stream operation between
// conv_fwd and dense_fwd, so the launch chain stays programmatic
static bool dp_gated(const ga3c_net* n) {
  static const bool on = [] { const char* e = getenv("GA3C_DP_GATE"); return !e || atoi(e) != 0; }();
  return on && n->dp_big_pending && n->dp_world > 1;
}
static DpGate dp_gate(const ga3c_net* n) {
  DpGate g{};
  if (dp_gated(n)) {
    uint8_t* comm = n->dp_peer[n->dp_rank] + n->comm_off;
    g.flags = comm + DPC_BIGDONE; g.err = comm + DPC_ERR; g.step = n->dp_step; g.world = n->dp_world;
  }
  return g;
}

// dense1 forward + heads in one cluster launch (dense_heads.cu) while its slabs fit the gradient-partial workspace
static bool fused_heads(const ga3c_net* n, int batch) {
  return n->fuse_heads && !n->cfg.dual_rmsprop && dense_heads_ctas(batch) <= 2 * n->num_sms;
}

static int predict_impl(ga3c_net* n, const void* x, bool x_u8, int32_t batch, float* p_out, float* v_out, void* stream) {
  if (int r = check_batch(n, batch, "ga3c_predict")) return r;
  if (!x || !p_out || !v_out) return fail_msg("ga3c_predict: null buffer");
  CK(cudaSetDevice(n->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const float* w = n->w;
  LAUNCH(n, K_CONV_FWD, st, launch_conv_fwd(x, x_u8, w + n->off(P_C11W), w + n->off(P_C11B), w + n->off(P_C12W),
                                            w + n->off(P_C12B), nullptr, nullptr, n->n2, batch, n->num_sms, st));
  const bool fused_p = fused_heads(n, batch);
  const DpGate gate = fused_p ? DpGate{} : dp_gate(n);
  if (gate.flags == nullptr)
    if (int r = wait_dp_big(n, st)) return r;
  const int splits = dense_fwd_splits(batch, n->num_sms);
  HeadsArgs h = heads_args(n, batch, splits);
  h.p_out = p_out; h.v_out = v_out; h.train = 0;
  if (fused_p) {
    LAUNCH(n, K_DENSE_FWD, st, launch_dense_heads(n->n2, n->w1_shadow, h, st));
  } else {
    LAUNCH(n, K_DENSE_FWD, st, launch_dense_fwd_tc(n->n2, n->w1_shadow, n->d1_part, batch, splits, st, &gate));
    LAUNCH(n, K_HEADS, st, launch_heads(h, n->num_sms, st));
  }
  n->last_batch = batch;
  return 0;
}

extern "C" int ga3c_predict(ga3c_net* n, const float* x, int32_t batch, float* p_out, float* v_out, void* stream) {
  return predict_impl(n, x, false, batch, p_out, v_out, stream);
}
extern "C" int ga3c_predict_u8(ga3c_net* n, const uint8_t* x, int32_t batch, float* p_out, float* v_out, void* stream) {
  return predict_impl(n, x, true, batch, p_out, v_out, stream);
}

// forward, fused loss forward/backward, and the dense1 weight gradient: on return (in stream order) the
// gradients of dense1/w, dense1/b, logits_v/*, logits_p/* are final, so their allreduce can start while
// ga3c_fb_tail computes the conv gradients.
static int fb_head_impl(ga3c_net* n, const void* x, bool x_u8, const float* yr, const float* a, int32_t batch, float beta,
                        float* loss, void* stream, bool with_wgrad, int part = 0, bool skip_forward = false) {
  if (int r = check_batch(n, batch, "ga3c_fb_head")) return r;
  if (!x || !yr || !a) return fail_msg("ga3c_fb_head: null buffer");
  CK(cudaSetDevice(n->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const float* w = n->w;
  float* g = n->g;
  float* gp = n->gpart;       // small-tensor gradients: per-CTA partial slabs, summed at the end of ga3c_fb_tail
  const int splits = dense_fwd_splits(batch, n->num_sms);
  const bool fused = fused_heads(n, batch);
  if (!skip_forward) {        // the second DUAL_RMSPROP pass reuses n1 / n2 / the dense1 partials of the first
    LAUNCH(n, K_CONV_FWD, st, launch_conv_fwd(x, x_u8, w + n->off(P_C11W), w + n->off(P_C11B), w + n->off(P_C12W),
                                              w + n->off(P_C12B), n->n1, n->xblk, n->n2, batch, n->num_sms, st));
    const DpGate gate = fused ? DpGate{} : dp_gate(n);
    if (gate.flags == nullptr)
      if (int r = wait_dp_big(n, st)) return r;
    if (!fused) LAUNCH(n, K_DENSE_FWD, st, launch_dense_fwd_tc(n->n2, n->w1_shadow, n->d1_part, batch, splits, st, &gate));
  }
  HeadsArgs h = heads_args(n, batch, splits);
  h.yr = yr; h.a = a; h.beta = beta; h.train = 1; h.dd1 = n->dd1; h.part = part;
  h.g_wp = gp + n->off(P_PW); h.g_bp = gp + n->off(P_PB); h.g_wv = gp + n->off(P_VW); h.g_bv = gp + n->off(P_VB);
  h.g_b1 = gp + n->off(P_D1B); h.loss = gp + n->small_floats; h.gp_stride = n->gp_stride;
  n->gp_heads_grid = fused ? dense_heads_ctas(batch) : heads_grid(batch, n->num_sms);
  n->loss_out = loss;
  if (fused) LAUNCH(n, K_DENSE_FWD, st, launch_dense_heads(n->n2, n->w1_shadow, h, st));
  else LAUNCH(n, K_HEADS, st, launch_heads(h, n->num_sms, st));
  (void)g;
  if (with_wgrad) LAUNCH(n, K_DENSE_WGRAD, st, launch_dense_wgrad_tc(n->n2, n->dd1, n->g + n->off(P_D1W), batch, st));
  n->last_batch = batch;
  return 0;
}

extern "C" int ga3c_fb_head(ga3c_net* n, const float* x, const float* yr, const float* a, int32_t batch, float beta,
                            float* loss, void* stream) {
  return fb_head_impl(n, x, false, yr, a, batch, beta, loss, stream, true);
}

// sum of the per-CTA slabs: conv tensors (the first four of the arena) over the conv grids, head tensors and the loss
// sums over the heads grid
static GradReduceArgs reduce_args(ga3c_net* n, int batch, float* g_dst = nullptr) {
  GradReduceArgs r{};
  r.part = n->gpart; r.stride = n->gp_stride; r.out = g_dst ? g_dst : n->g; r.out_tail = n->loss_out;
  r.out_floats = (int)n->small_floats; r.n_floats = (int)n->small_floats + 4;
  for (int s = 0; s < GR_MAX_SEG; ++s) { r.seg_end[s] = r.n_floats; r.seg_count[s] = n->gp_heads_grid; }
  r.seg_end[0] = (int)n->off(P_D1B); r.seg_count[0] = conv_bwd_grid(batch, n->num_sms, n->cur_exch);
  return r;
}

// dense1 data gradient and the two conv backward kernels; with `reduce` the slabs are summed into the gradient arena
// (conv11/*, conv12/*, dense1/b, heads, loss sums), otherwise the caller does it (fused with RMSProp).
static WgradPush wgrad_push(const ga3c_net* n, uint64_t step) {
  WgradPush p{};
  for (int r = 0; r < n->dp_world; ++r) p.peer[r] = n->dp_peer[r];
  p.rank = n->dp_rank; p.world = n->dp_world;
  p.per4 = ((long long)FLAT * FC / 4 + n->dp_world - 1) / n->dp_world;
  p.recv_off = n->bigrecv_off;
  p.flag = (unsigned int)step;
  return p;
}

static int fb_tail_impl(ga3c_net* n, const void* x, bool x_u8, int32_t batch, void* stream, bool with_wgrad, bool reduce,
                        const DpBigArgs* dp = nullptr, float* g_dst = nullptr, const ConvBwdOpt* opt = nullptr) {
  if (int r = check_batch(n, batch, "ga3c_fb_tail")) return r;
  if (!x) return fail_msg("ga3c_fb_tail: null buffer");
  if (batch != n->last_batch) return fail_msg("ga3c_fb_tail: batch differs from the preceding ga3c_fb_head");
  CK(cudaSetDevice(n->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const float* w = n->w;
  float* gp = n->gpart;
  if (with_wgrad) {   // dgrad + wgrad tiles of dense1 in one grid
    WgradPush push{};
    const bool pushed = dp != nullptr && dp->pushed;
    if (pushed) push = wgrad_push(n, dp->step);
    LAUNCH(n, K_DENSE_DGRAD, st, launch_dense_bwd_tc(n->dd1, n->w1_shadow, n->n2, n->dn2, (g_dst ? g_dst : n->g) + n->off(P_D1W), batch,
                                                     pushed ? &push : nullptr, st));
  } else
    LAUNCH(n, K_DENSE_DGRAD, st, launch_dense_dgrad_tc(n->dd1, n->w1_shadow, n->n2, n->dn2, batch, st));
  if (n->dp_side != nullptr) {
    // side-stream exchange of dense1/w: its gradient is final behind this point of the stream
    CK(cudaEventRecord(n->dp_grad_ready, st));
    CK(cudaStreamWaitEvent(n->dp_stream, n->dp_grad_ready, 0));
    CKL(launch_dp_big_side(*n->dp_side, true, n->num_sms, n->dp_stream));
    n->log.launches++;
    CK(cudaEventRecord(n->dp_big_done, n->dp_stream));
    n->dp_big_pending = true;
  }
  LAUNCH(n, K_CONV12_BWD, st, launch_conv_bwd(n->xblk, n->n1, n->dn2, w + n->off(P_C12W), n->keep_dn1 ? n->dn1 : nullptr,
                                              gp + n->off(P_C11W), gp + n->off(P_C11B), gp + n->off(P_C12W),
                                              gp + n->off(P_C12B), n->gp_stride, batch, n->num_sms, x_u8, opt ? *opt : ConvBwdOpt{}, dp, st));
  if (reduce) LAUNCH(n, K_GRAD_REDUCE, st, launch_grad_reduce(reduce_args(n, batch, g_dst), st));
  return 0;
}

extern "C" int ga3c_fb_tail(ga3c_net* n, const float* x, int32_t batch, void* stream) {
  return fb_tail_impl(n, x, false, batch, stream, false, true);
}

static int forward_backward_impl(ga3c_net* n, const void* x, bool x_u8, const float* yr, const float* a, int32_t batch,
                                 float beta, float* loss, void* stream) {
  if (int r = fb_head_impl(n, x, x_u8, yr, a, batch, beta, loss, stream, false)) return r;
  return fb_tail_impl(n, x, x_u8, batch, stream, true, true);
}
extern "C" int ga3c_forward_backward(ga3c_net* n, const float* x, const float* yr, const float* a, int32_t batch,
                                     float beta, float* loss, void* stream) {
  return forward_backward_impl(n, x, false, yr, a, batch, beta, loss, stream);
}
extern "C" int ga3c_forward_backward_u8(ga3c_net* n, const uint8_t* x, const float* yr, const float* a, int32_t batch,
                                        float beta, float* loss, void* stream) {
  return forward_backward_impl(n, x, true, yr, a, batch, beta, loss, stream);
}
extern "C" int ga3c_fb_head_u8(ga3c_net* n, const uint8_t* x, const float* yr, const float* a, int32_t batch, float beta,
                               float* loss, void* stream) {
  return fb_head_impl(n, x, true, yr, a, batch, beta, loss, stream, true);
}
extern "C" int ga3c_fb_tail_u8(ga3c_net* n, const uint8_t* x, int32_t batch, void* stream) {
  return fb_tail_impl(n, x, true, batch, stream, false, true);
}

static RmsPropArgs rmsprop_args(ga3c_net* n, float lr) {
  RmsPropArgs a{};
  a.w = n->w; a.ms = n->ms; a.mom = n->mom; a.g = n->g; a.w1_shadow = n->w1_shadow;
  a.n_floats = n->arena_floats; a.w1_offset = n->off(P_D1W); a.w1_count = (int64_t)FLAT * FC;
  a.lr = lr; a.decay = n->cfg.rmsprop_decay; a.momentum = n->cfg.rmsprop_momentum; a.eps = n->cfg.rmsprop_epsilon;
  return a;
}

static ClipArgs clip_args(ga3c_net* n, int which = 0) {        // which: 0 the (first) optimizer's gradient, 1 the second one's
  ClipArgs c{};
  c.g = which ? n->g2 : n->g; c.n_tensors = P_COUNT; c.max_chunks = clip_chunks((int64_t)FLAT * FC);
  for (int i = 0; i < P_COUNT; ++i) { c.offset[i] = n->params[i].offset; c.count[i] = n->params[i].count; }
  c.clip = n->cfg.grad_clip_norm; c.chunk_ss = which ? n->clip_ss2 : n->clip_ss; c.scale = which ? n->clip_scale2 : n->clip_scale;
  return c;
}

static int apply_rmsprop_impl(ga3c_net* n, float lr, void* stream, const GradReduceArgs* red) {
  if (!n) return fail_msg("ga3c_apply_rmsprop: null handle");
  CK(cudaSetDevice(n->cfg.device));
  RmsPropArgs a = rmsprop_args(n, lr);
  if (n->cfg.dual_rmsprop) return fail_msg("ga3c_apply_rmsprop: with DUAL_RMSPROP use ga3c_train_step (two backward passes)");
  if (n->cfg.use_grad_clip) {
    // clip_by_average_norm needs the whole (reduced) gradient of a variable: local arena only (single GPU, or after the
    // host's NCCL allreduce in dp_mode 'nccl')
    if (n->dp_world > 1) return fail_msg("ga3c_apply_rmsprop: USE_GRAD_CLIP is not available with the peer-memory exchange");
    LAUNCH(n, K_RMSPROP, (cudaStream_t)stream, launch_rmsprop_clipped(a, clip_args(n), (cudaStream_t)stream));
    n->log.launches += 2;
    return 0;          // NetworkVP_discrate.py:121: apply_gradients without global_step -- the step counter stays put
  }
  if (n->dp_world > 1) {
    // fused reduce-scatter(grads) -> RMSProp on this rank's slice -> all-gather(weights) over peer memory
    RmsPropDpArgs d{};
    d.base = a;
    for (int r = 0; r < n->dp_world; ++r) d.peer[r] = n->dp_peer[r];
    d.rank = n->dp_rank; d.world = n->dp_world; d.step = ++n->dp_step;
    d.arena_bytes = (int64_t)n->arena_floats * 4;
    d.comm_offset = n->comm_off;
    d.has_red = red != nullptr;
    if (red) d.red = *red;
    LAUNCH(n, K_RMSPROP, (cudaStream_t)stream, launch_rmsprop_dp(d, n->num_sms, (cudaStream_t)stream));
  } else {
    LAUNCH(n, K_RMSPROP, (cudaStream_t)stream, launch_rmsprop(a, (cudaStream_t)stream));
  }
  n->global_step += 1;   // opt.minimize(..., global_step=self.global_step), NetworkVP_discrate.py:130
  return 0;
}

extern "C" int ga3c_apply_rmsprop(ga3c_net* n, float lr, void* stream) { return apply_rmsprop_impl(n, lr, stream, nullptr); }

// ---- data parallel over CUDA IPC peer memory ---------------------------------------------------------
extern "C" int ga3c_dp_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

extern "C" int ga3c_dp_export(ga3c_net* n, void* handle_out) {
  if (!n || !handle_out) return fail_msg("ga3c_dp_export: null argument");
  CK(cudaSetDevice(n->cfg.device));
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, n->slab));
  memcpy(handle_out, &h, sizeof(h));
  return 0;
}

extern "C" int ga3c_dp_detach(ga3c_net* n) {
  if (!n) return 0;
  if (n->dp_stream) {
    cudaStreamSynchronize(n->dp_stream);
    cudaStreamDestroy(n->dp_stream);
    cudaEventDestroy(n->dp_big_done);
    cudaEventDestroy(n->dp_grad_ready);
    n->dp_stream = nullptr; n->dp_big_done = nullptr; n->dp_big_pending = false;
  }
  for (int r = 0; r < DP_MAX_WORLD; ++r)
    if (n->dp_peer[r] && n->dp_ipc[r]) cudaIpcCloseMemHandle(n->dp_peer[r]);
  for (int r = 0; r < DP_MAX_WORLD; ++r) { n->dp_peer[r] = nullptr; n->dp_ipc[r] = false; }
  n->dp_world = 1; n->dp_rank = 0;
  return 0;
}

extern "C" int ga3c_dp_attach(ga3c_net* n, int32_t rank, int32_t world, const void* handles) {
  if (!n || !handles) return fail_msg("ga3c_dp_attach: null argument");
  if (world < 1 || world > DP_MAX_WORLD || rank < 0 || rank >= world)
    return fail_msg("ga3c_dp_attach: world must be 1.." + std::to_string(DP_MAX_WORLD) + " and 0 <= rank < world");
  CK(cudaSetDevice(n->cfg.device));
  CK(cudaDeviceSynchronize());
  ga3c_dp_detach(n);
  const cudaIpcMemHandle_t* hs = static_cast<const cudaIpcMemHandle_t*>(handles);
  for (int r = 0; r < world; ++r) {
    if (r == rank) { n->dp_peer[r] = n->slab; continue; }
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, hs[r], cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      ga3c_dp_detach(n);
      return fail("cudaIpcOpenMemHandle (peer-to-peer access between the ranks' GPUs is required)", e);
    }
    n->dp_peer[r] = static_cast<uint8_t*>(p);
    n->dp_ipc[r] = true;
  }
  return dp_attach_finish(n, rank, world);
}

// ranks that live in ONE process (tests; a host that drives several GPUs from one process): the peers' slabs are ordinary
// device pointers, no IPC handles involved.  peers[r] is rank r's handle (peers[rank] == net).
extern "C" int ga3c_dp_attach_local(ga3c_net* n, int32_t rank, int32_t world, ga3c_net* const* peers) {
  if (!n || !peers) return fail_msg("ga3c_dp_attach_local: null argument");
  if (world < 1 || world > DP_MAX_WORLD || rank < 0 || rank >= world || peers[rank] != n)
    return fail_msg("ga3c_dp_attach_local: need 0 <= rank < world <= 8 and peers[rank] == net");
  CK(cudaSetDevice(n->cfg.device));
  CK(cudaDeviceSynchronize());
  ga3c_dp_detach(n);
  for (int r = 0; r < world; ++r) {
    if (!peers[r] || peers[r]->arena_floats != n->arena_floats) { ga3c_dp_detach(n); return fail_msg("ga3c_dp_attach_local: peers must be handles of the same network"); }
    if (peers[r]->cfg.device != n->cfg.device) {
      int can = 0;
      cudaDeviceCanAccessPeer(&can, n->cfg.device, peers[r]->cfg.device);
      if (!can) { ga3c_dp_detach(n); return fail_msg("ga3c_dp_attach_local: no peer access between the devices"); }
      cudaError_t e = cudaDeviceEnablePeerAccess(peers[r]->cfg.device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { ga3c_dp_detach(n); return fail("cudaDeviceEnablePeerAccess", e); }
      cudaGetLastError();
    }
    n->dp_peer[r] = peers[r]->slab;
  }
  return dp_attach_finish(n, rank, world);
}

static int dp_attach_finish(ga3c_net* n, int32_t rank, int32_t world) {
  n->dp_rank = rank; n->dp_world = world; n->dp_step = 0;
  n->dp_exchange = 4;
  if (const char* e = getenv("GA3C_DP_EXCHANGE")) {
    const std::string m(e);
    if (m == "side") n->dp_exchange = 4;
    else if (m == "warps") n->dp_exchange = 3;
    else if (m == "tail") n->dp_exchange = 0;
    else if (m == "overlap") n->dp_exchange = 1;
    else if (m == "single") n->dp_exchange = 2;
    else return fail_msg("GA3C_DP_EXCHANGE must be side, warps, tail, overlap or single");
  }
  if (const char* e = getenv("GA3C_DP_PUSH")) n->dp_push = atoi(e) != 0;
  if (const char* e = getenv("GA3C_DP_EXCH_CTAS")) n->dp_exch = atoi(e);
  if (n->dp_exch < 1 || n->dp_exch > n->num_sms / 2) n->dp_exch = 20;
  CK(cudaMemset(n->slab + n->xbuf_off, 0, n->slab_bytes - (size_t)n->xbuf_off));
  CKL(configure_dp());
  if (!n->dp_stream) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    CK(cudaStreamCreateWithPriority(&n->dp_stream, cudaStreamNonBlocking, hi));
    CK(cudaEventCreateWithFlags(&n->dp_big_done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&n->dp_grad_ready, cudaEventDisableTiming));
  }
  n->dp_big_pending = false;
  return 0;
}

// 0: every cross-rank wait of the exchange kernels completed; 1: one gave up (a rank died, or the ranks' train calls fell out
// of step) and the weights can no longer be trusted.  Synchronises the device.
extern "C" int ga3c_dp_error(ga3c_net* n, int32_t* error_out) {
  if (!n || !error_out) return fail_msg("ga3c_dp_error: null argument");
  *error_out = 0;
  if (n->dp_world <= 1) return 0;
  CK(cudaSetDevice(n->cfg.device));
  CK(cudaDeviceSynchronize());
  unsigned int e = 0;
  CK(cudaMemcpy(&e, n->slab + n->comm_off + DPC_ERR, 4, cudaMemcpyDeviceToHost));
  *error_out = (int32_t)e;
  return 0;
}

// Config.DUAL_RMSPROP: one forward, two backward passes (cost_p into g, cost_v into g2) ...
static int dual_forward_backward_impl(ga3c_net* n, const void* x, bool x_u8, const float* yr, const float* a, int32_t batch,
                                      float beta, float* loss, void* stream) {
  if (!n || !n->cfg.dual_rmsprop) return fail_msg("ga3c_dual_forward_backward: the handle was not created with dual_rmsprop");
  if (int r = fb_head_impl(n, x, x_u8, yr, a, batch, beta, loss, stream, false, 1)) return r;
  if (int r = fb_tail_impl(n, x, x_u8, batch, stream, true, true)) return r;
  if (int r = fb_head_impl(n, x, x_u8, yr, a, batch, beta, nullptr, stream, false, 2, true)) return r;
  return fb_tail_impl(n, x, x_u8, batch, stream, true, true, nullptr, n->g2);
}
// ... and one update that subtracts both optimizers' steps, each taken at the weights the call started with
static int dual_apply_impl(ga3c_net* n, float lr, void* stream) {
  if (!n || !n->cfg.dual_rmsprop) return fail_msg("ga3c_dual_apply: the handle was not created with dual_rmsprop");
  CK(cudaSetDevice(n->cfg.device));
  RmsPropDualArgs d{};
  d.a = rmsprop_args(n, lr);
  d.g2 = n->g2; d.ms2 = n->ms2; d.mom2 = n->mom2;
  const int skip[4] = {P_VW, P_VB, P_PW, P_PB};          // cost_p does not reach logits_v (stop_gradient), cost_v not logits_p
  for (int i = 0; i < 4; ++i) { d.skip_lo[i] = n->off(skip[i]); d.skip_hi[i] = n->off(skip[i]) + n->params[skip[i]].count; }
  if (n->cfg.use_grad_clip) {
    // NetworkVP_discrate.py:107-117: tf.clip_by_norm per variable and optimizer, then apply_gradients WITHOUT global_step
    ClipArgs c1 = clip_args(n, 0), c2 = clip_args(n, 1);
    c1.by_norm = c2.by_norm = 1;
    LAUNCH(n, K_RMSPROP, (cudaStream_t)stream, launch_rmsprop_dual_clipped(d, c1, c2, (cudaStream_t)stream));
    n->log.launches += 4;
    return 0;
  }
  LAUNCH(n, K_RMSPROP, (cudaStream_t)stream, launch_rmsprop_dual(d, (cudaStream_t)stream));
  n->global_step += 2;      // both minimize calls advance it (NetworkVP_discrate.py:125-126)
  return 0;
}
extern "C" int ga3c_dual_forward_backward(ga3c_net* n, const float* x, const float* yr, const float* a, int32_t batch, float beta,
                                          float* loss, void* stream) {
  return dual_forward_backward_impl(n, x, false, yr, a, batch, beta, loss, stream);
}
extern "C" int ga3c_dual_forward_backward_u8(ga3c_net* n, const uint8_t* x, const float* yr, const float* a, int32_t batch,
                                             float beta, float* loss, void* stream) {
  return dual_forward_backward_impl(n, x, true, yr, a, batch, beta, loss, stream);
}
extern "C" int ga3c_dual_apply(ga3c_net* n, float lr, void* stream) { return dual_apply_impl(n, lr, stream); }

static DpBigArgs dp_big_args(ga3c_net* n, float lr, int n_exch) {
  DpBigArgs b{};
  for (int r = 0; r < n->dp_world; ++r) b.peer[r] = n->dp_peer[r];
  b.rank = n->dp_rank; b.world = n->dp_world; b.n_exch = n_exch; b.step = ++n->dp_step;
  b.arena_bytes = (int64_t)n->arena_floats * 4; b.shadow_off = 4 * b.arena_bytes; b.comm_offset = n->comm_off;
  b.w1_offset = n->off(P_D1W); b.w1_count = (int64_t)FLAT * FC;
  b.lr = lr; b.decay = n->cfg.rmsprop_decay; b.momentum = n->cfg.rmsprop_momentum; b.eps = n->cfg.rmsprop_epsilon;
  b.pushed = n->dp_exchange == 0 && n->dp_push;       // exchange at the end of the step: slices pushed by the wgrad epilogues
  b.bigrecv_off = n->bigrecv_off;
  return b;
}
static RmsPropDpArgs dp_small_args(ga3c_net* n, float lr, int batch, const DpBigArgs& b) {
  RmsPropDpArgs d{};
  d.base = rmsprop_args(n, lr);
  d.base.preload = batch >= n->num_sms;         // see rmsprop_reduce_kernel
  for (int q = 0; q < n->dp_world; ++q) d.peer[q] = n->dp_peer[q];
  d.rank = n->dp_rank; d.world = n->dp_world; d.step = b.step;
  d.arena_bytes = b.arena_bytes; d.comm_offset = n->comm_off;
  d.has_red = 1; d.red = reduce_args(n, batch);
  return d;
}

// A rank of a data-parallel job that has no experiences this round still takes part in the exchange (every rank must enter
// every step: the asynchronous trainer loop of the reference, ThreadTrainer.py:42-62, gives no such guarantee, so the host
// ticks all ranks in lock step and idle ranks contribute a zero gradient).  No forward / backward kernels run.
static int train_step_empty(ga3c_net* n, float lr, float* loss, void* stream) {
  CK(cudaSetDevice(n->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  if (n->cfg.dual_rmsprop || n->cfg.use_grad_clip)
    return fail_msg("ga3c_train_step: batch 0 is only defined for the peer-memory data-parallel step");
  n->gp_heads_grid = 0;
  n->loss_out = loss;
  n->last_batch = 0;
  if (n->dp_exchange == 4) {          // side-stream exchange: it publishes "gradient final" itself; dp_small with zero slabs
    CK(cudaMemsetAsync(n->g + n->off(P_D1W), 0, (size_t)FLAT * FC * 4, st));
    DpBigArgs b = dp_big_args(n, lr, 0);
    b.pushed = 0;
    n->cur_exch = 0;
    CK(cudaEventRecord(n->dp_grad_ready, st));                    // the zeroed gradient is in place
    CK(cudaStreamWaitEvent(n->dp_stream, n->dp_grad_ready, 0));
    CKL(launch_dp_big_side(b, true, n->num_sms, n->dp_stream));
    n->log.launches++;
    CK(cudaEventRecord(n->dp_big_done, n->dp_stream));
    n->dp_big_pending = true;
    const RmsPropDpArgs d = dp_small_args(n, lr, 0, b);
    LAUNCH(n, K_RMSPROP, st, launch_dp_small(d, n->xbuf_off, st, false));
    n->global_step += 1;
    return 0;
  }
  if (n->dp_exchange == 0 || n->dp_exchange == 3) {   // nothing but the exchange launch (dp_tail: both instalments); it publishes "gradient final"
    CK(cudaMemsetAsync(n->g + n->off(P_D1W), 0, (size_t)FLAT * FC * 4, st));
    const DpBigArgs b = dp_big_args(n, lr, 0);
    n->cur_exch = 0;
    if (b.pushed) {
      CKL(launch_dp_push_zero(wgrad_push(n, b.step), (long long)FLAT * FC / 4, st));
      n->log.launches++;
    }
    const RmsPropDpArgs d = dp_small_args(n, lr, 0, b);    // zero slabs in every segment: the sums are 0
    LAUNCH(n, K_RMSPROP, st, launch_dp_tail(d, b, n->xbuf_off, n->num_sms, st));
    n->global_step += 1;
    return 0;
  }
  if (n->dp_exchange == 1) {
    CK(cudaMemsetAsync(n->g + n->off(P_D1W), 0, (size_t)FLAT * FC * 4, st));
    const DpBigArgs b = dp_big_args(n, lr, n->dp_exch);
    n->cur_exch = b.n_exch;
    float* gp = n->gpart;
    LAUNCH(n, K_CONV12_BWD, st, launch_conv_bwd(n->xblk, n->n1, n->dn2, n->w + n->off(P_C12W), nullptr,
                                                gp + n->off(P_C11W), gp + n->off(P_C11B), gp + n->off(P_C12W),
                                                gp + n->off(P_C12B), n->gp_stride, 0, n->num_sms, false, ConvBwdOpt{}, &b, st));
    const RmsPropDpArgs d = dp_small_args(n, lr, 0, b);
    n->cur_exch = 0;
    LAUNCH(n, K_RMSPROP, st, launch_dp_small(d, n->xbuf_off, st, true));
    n->global_step += 1;
    return 0;
  }
  CK(cudaMemsetAsync(n->g, 0, (size_t)n->arena_floats * 4, st));
  if (loss) CK(cudaMemsetAsync(loss, 0, 16, st));
  return apply_rmsprop_impl(n, lr, stream, nullptr);
}

static int train_step_impl(ga3c_net* n, const void* x, bool x_u8, const float* yr, const float* a, int32_t batch, float lr,
                           float beta, float* loss, void* stream) {
  if (n && batch == 0 && n->dp_world > 1) return train_step_empty(n, lr, loss, stream);
  if (n->cfg.dual_rmsprop) {
    if (n->dp_world > 1) return fail_msg("ga3c_train_step: DUAL_RMSPROP is not available with the peer-memory exchange (dp_mode 'nccl' is)");
    if (int r = dual_forward_backward_impl(n, x, x_u8, yr, a, batch, beta, loss, stream)) return r;
    return dual_apply_impl(n, lr, stream);
  }
  if (int r = fb_head_impl(n, x, x_u8, yr, a, batch, beta, loss, stream, false)) return r;
  if (n->cfg.use_grad_clip) {
    if (int r = fb_tail_impl(n, x, x_u8, batch, stream, true, true)) return r;
    return apply_rmsprop_impl(n, lr, stream, nullptr);
  }
  if (n->dp_world > 1 && n->dp_exchange == 4) {
    // default: the conv backward keeps every SM; the exchange of dense1/w is launched behind dense_bwd on the side stream and
    // runs next to it (and next to the conv forward of the following step); dp_small follows the conv
    // backward on this stream with the small tensors.  dense_fwd of the next call waits for the exchange's completion event.
    DpBigArgs b = dp_big_args(n, lr, 0);
    b.pushed = 0;
    n->cur_exch = 0;
    // (launched behind an event after dense_bwd, NOT on a flag the conv backward would publish: at batch >= 148 a spinning side
    //  grid that got its blocks resident first kept conv_bwd CTA 0 -- the publisher -- off its SM, and the step timed out)
    n->dp_side = &b;
    const int r = fb_tail_impl(n, x, x_u8, batch, stream, true, false);
    n->dp_side = nullptr;
    if (r) return r;
    RmsPropDpArgs d = dp_small_args(n, lr, batch, b);
    LAUNCH(n, K_RMSPROP, (cudaStream_t)stream, launch_dp_small(d, n->xbuf_off, (cudaStream_t)stream, false));
    n->global_step += 1;
    return 0;
  }
  if (n->dp_world > 1 && n->dp_exchange == 3) {
    // default: the dense1/w exchange rides on the optimizer warps of the conv backward CTAs (all SMs keep computing conv
    // gradients); dp_small follows with the small tensors and holds the step open until every rank's slice has landed
    const DpBigArgs b = dp_big_args(n, lr, 0);
    n->cur_exch = 0;
    ConvBwdOpt opt{};
    opt.mode = 2;
    if (int r = fb_tail_impl(n, x, x_u8, batch, stream, true, false, &b, nullptr, &opt)) return r;
    const RmsPropDpArgs d = dp_small_args(n, lr, batch, b);
    LAUNCH(n, K_RMSPROP, (cudaStream_t)stream, launch_dp_small(d, n->xbuf_off, (cudaStream_t)stream, true));
    n->global_step += 1;
    return 0;
  }
  if (n->dp_world > 1 && n->dp_exchange == 0) {
    // exchange at the end of the step (dp_tail_kernel): the conv backward keeps every SM; its CTA 0 publishes "dense1/w gradient
    // final" to the peers as soon as dense_bwd is complete, so that no rank waits for another rank's gradient at the tail
    const DpBigArgs b = dp_big_args(n, lr, 0);
    n->cur_exch = 0;
    if (int r = fb_tail_impl(n, x, x_u8, batch, stream, true, false, &b)) return r;
    const RmsPropDpArgs d = dp_small_args(n, lr, batch, b);
    LAUNCH(n, K_RMSPROP, (cudaStream_t)stream, launch_dp_tail(d, b, n->xbuf_off, n->num_sms, (cudaStream_t)stream));
    n->global_step += 1;
    return 0;
  }
  if (n->dp_world > 1 && n->dp_exchange == 1) {
    // overlapped exchange (dp_exchange.cuh): dense1/w moves between the ranks on exchange CTAs of the conv backward
    // launch; the small tensors follow in dp_small, which also holds the step open until every slice has landed
    const DpBigArgs b = dp_big_args(n, lr, n->dp_exch);
    n->cur_exch = b.n_exch;
    int r = fb_tail_impl(n, x, x_u8, batch, stream, true, false, &b);
    const RmsPropDpArgs d = dp_small_args(n, lr, batch, b);
    n->cur_exch = 0;
    if (r) return r;
    LAUNCH(n, K_RMSPROP, (cudaStream_t)stream, launch_dp_small(d, n->xbuf_off, (cudaStream_t)stream, true));
    n->global_step += 1;
    return 0;
  }
  if (n->dp_world == 1 && n->fuse_opt) {
    // single GPU: dense1/w is updated by the optimizer warps of the conv backward CTAs; the launch at the end of the step sums the
    // slabs and updates the small tensors
    ConvBwdOpt opt{};
    opt.mode = 1;
    opt.g = n->g + n->off(P_D1W); opt.w = n->w + n->off(P_D1W); opt.ms = n->ms + n->off(P_D1W); opt.mom = n->mom + n->off(P_D1W);
    opt.shadow = n->w1_shadow; opt.n4 = (long long)FLAT * FC / 4;
    opt.lr = lr; opt.decay = n->cfg.rmsprop_decay; opt.momentum = n->cfg.rmsprop_momentum; opt.eps = n->cfg.rmsprop_epsilon;
    if (int r = fb_tail_impl(n, x, x_u8, batch, stream, true, false, nullptr, nullptr, &opt)) return r;
    RmsPropArgs ra = rmsprop_args(n, lr);
    ra.n_floats = n->small_floats;                  // dense1/w is done
    ra.preload = batch >= n->num_sms;
    LAUNCH(n, K_RMSPROP, (cudaStream_t)stream, launch_rmsprop_reduce(ra, reduce_args(n, batch), (cudaStream_t)stream));
    n->global_step += 1;
    return 0;
  }
  if (int r = fb_tail_impl(n, x, x_u8, batch, stream, true, false)) return r;
  const GradReduceArgs red = reduce_args(n, batch);
  if (n->dp_world > 1)            // peers read this rank's gradient arena: the exchange kernel first sums the slabs into it
    return apply_rmsprop_impl(n, lr, stream, &red);
  // single GPU: the slab reduction rides in the optimizer launch
  RmsPropArgs ra = rmsprop_args(n, lr);
  ra.preload = batch >= n->num_sms;               // see rmsprop_reduce_kernel
  LAUNCH(n, K_RMSPROP, (cudaStream_t)stream, launch_rmsprop_reduce(ra, red, (cudaStream_t)stream));
  n->global_step += 1;   // opt.minimize(..., global_step=self.global_step), NetworkVP_discrate.py:130
  return 0;
}
extern "C" int ga3c_train_step(ga3c_net* n, const float* x, const float* yr, const float* a, int32_t batch, float lr,
                               float beta, float* loss, void* stream) {
  return train_step_impl(n, x, false, yr, a, batch, lr, beta, loss, stream);
}
extern "C" int ga3c_train_step_u8(ga3c_net* n, const uint8_t* x, const float* yr, const float* a, int32_t batch, float lr,
                                  float beta, float* loss, void* stream) {
  return train_step_impl(n, x, true, yr, a, batch, lr, beta, loss, stream);
}

extern "C" int ga3c_returns(const double* rewards, const int64_t* seg, int32_t n_segments, const double* terminal,
                            double discount, int32_t flags, double rmin, double rmax, double* out, void* stream) {
  if (n_segments < 0) return fail_msg("ga3c_returns: negative segment count");
  if (n_segments == 0) return 0;
  if (!rewards || !seg || !terminal || !out) return fail_msg("ga3c_returns: null buffer");
  CKL(launch_returns(rewards, seg, n_segments, terminal, discount, flags, rmin, rmax, out, (cudaStream_t)stream));
  return 0;
}

extern "C" int ga3c_select_actions(const float* p, const double* u, int32_t batch, int32_t na, int32_t* action,
                                   void* stream) {
  if (batch < 0) return fail_msg("ga3c_select_actions: negative batch");
  if (batch == 0) return 0;
  if (!p || !u || !action) return fail_msg("ga3c_select_actions: null buffer");
  if (na < 1 || na > MAX_ACTIONS) return fail_msg("ga3c_select_actions: num_actions must be 1..18");
  CKL(launch_select_actions(p, u, batch, na, action, (cudaStream_t)stream));
  return 0;
}

extern "C" int ga3c_workspace_ptr(ga3c_net* n, int which, void** ptr, int64_t* bytes) {
  if (!n || !ptr) return fail_msg("ga3c_workspace_ptr: null argument");
  const int64_t b = n->last_batch;
  switch (which) {
    case 0: *ptr = n->n1; if (bytes) *bytes = b * B2_BYTES; break;
    case 1: *ptr = n->n2; if (bytes) *bytes = b * FLAT * 2; break;
    case 2: *ptr = n->d1; if (bytes) *bytes = b * FC * 4; break;
    case 3: *ptr = n->dd1; if (bytes) *bytes = b * FC * 2; break;
    case 4: *ptr = n->dn2; if (bytes) *bytes = b * G_BYTES; break;
    case 5: *ptr = n->dn1; if (bytes) *bytes = b * N1_POS * C1_OUT * 2; break;
    case 6: *ptr = n->w1_shadow; if (bytes) *bytes = (int64_t)FLAT * FC * 2; break;
    case 7: *ptr = n->xblk; if (bytes) *bytes = b * XB_FRAME_BYTES; break;
    default: return fail_msg("ga3c_workspace_ptr: bad id");
  }
  return 0;
}

extern "C" int ga3c_keep_dn1(ga3c_net* n, int32_t on) {
  if (!n) return fail_msg("ga3c_keep_dn1: null handle");
  n->keep_dn1 = on != 0;
  return 0;
}

extern "C" int64_t ga3c_launch_count(const ga3c_net* n) { return n ? n->log.launches : 0; }

extern "C" int ga3c_kernel_count(void) { return K_COUNT; }
extern "C" const char* ga3c_kernel_name(int kid) { return (kid >= 0 && kid < K_COUNT) ? kKernelNames[kid] : nullptr; }

// ---- step timeline trace --------------------------------------------------------------------------
static int trace_attach_all(unsigned long long* buf) {
  int r;
  if ((r = trace_attach_conv_fwd(buf)) || (r = trace_attach_conv_bwd_fused(buf)) || (r = trace_attach_dense_tc(buf)) ||
      (r = trace_attach_heads(buf)) || (r = trace_attach_dense_heads(buf)) || (r = trace_attach_elementwise(buf)) || (r = trace_attach_mlp(buf)) || (r = trace_attach_mlp_tc(buf)) || (r = trace_attach_mlp_stream(buf)))
    return r;
  return 0;
}

extern "C" int ga3c_trace_begin(ga3c_net* n, void* stream) {
  if (!n) return fail_msg("ga3c_trace_begin: null handle");
  CK(cudaSetDevice(n->cfg.device));
  const size_t words = (size_t)K_COUNT * TRACE_SLOTS;
  if (!n->trace) CK(cudaMalloc((void**)&n->trace, words * 8));
  std::vector<unsigned long long> init(words);
  for (size_t i = 0; i < words; ++i) init[i] = (i & 1) ? 0ull : ~0ull;      // even slots take a min, odd slots a max
  CK(cudaStreamSynchronize((cudaStream_t)stream));
  CK(cudaMemcpy(n->trace, init.data(), words * 8, cudaMemcpyHostToDevice));
  CKL(trace_attach_all(n->trace));
  return 0;
}

extern "C" int ga3c_trace_end(ga3c_net* n, uint64_t* stamps, int32_t n_kernels) {
  if (!n || !stamps || n_kernels < K_COUNT) return fail_msg("ga3c_trace_end: bad argument");
  if (!n->trace) return fail_msg("ga3c_trace_end: no trace in progress");
  CK(cudaSetDevice(n->cfg.device));
  CK(cudaDeviceSynchronize());
  CKL(trace_attach_all(nullptr));
  CK(cudaMemcpy(stamps, n->trace, (size_t)K_COUNT * TRACE_SLOTS * 8, cudaMemcpyDeviceToHost));
  return 0;
}

// ---- pipeline event log (debug) ----------------------------------------------------------------------
extern "C" int ga3c_evt_begin(ga3c_net* n) {
  if (!n) return fail_msg("ga3c_evt_begin: null handle");
  CK(cudaSetDevice(n->cfg.device));
  CK(cudaDeviceSynchronize());
  if (!n->evt) CK(cudaMalloc((void**)&n->evt, (size_t)2 * 16384 * 8));
  CK(cudaMemset(n->evt, 0, (size_t)2 * 16384 * 8));
  // the kernels share the 16 warp regions of one buffer: GA3C_EVT_DENSE=1 (tools/evt_dense.py, a -DGA3C_DENSE_EVT build) leaves them
  // to the dense1 GEMMs
  if (getenv("GA3C_EVT_DENSE") == nullptr) {
    CKL(evt_attach_conv_fwd(n->evt));
    CKL(evt_attach_conv_bwd(n->evt));
    CKL(evt_attach_elementwise(n->evt));
    CKL(evt_attach_dense_heads(n->evt));
  } else {
    CKL(evt_attach_dense_tc(n->evt));
  }
  return 0;
}

extern "C" int ga3c_evt_end(ga3c_net* n, uint64_t* records, int32_t cap, int32_t* count) {
  if (!n || !records || !count || !n->evt || cap < 16384) return fail_msg("ga3c_evt_end: bad argument");
  CK(cudaSetDevice(n->cfg.device));
  CK(cudaDeviceSynchronize());
  CKL(evt_attach_conv_fwd(nullptr));
  CKL(evt_attach_conv_bwd(nullptr));
  CKL(evt_attach_elementwise(nullptr));
  CKL(evt_attach_dense_heads(nullptr));
  CKL(evt_attach_dense_tc(nullptr));
  std::vector<uint64_t> all((size_t)2 * 16384);
  CK(cudaMemcpy(all.data(), n->evt, all.size() * 8, cudaMemcpyDeviceToHost));
  int32_t c = 0;
  for (int i = 0; i < 16384; ++i)
    if (all[2 * i] != 0) { records[2 * c] = all[2 * i]; records[2 * c + 1] = all[2 * i + 1]; ++c; }
  *count = c;
  return 0;
}

extern "C" int ga3c_timing_enable(ga3c_net* n, int32_t max_records) {
  if (!n || max_records < 0) return fail_msg("ga3c_timing_enable: bad argument");
  CK(cudaSetDevice(n->cfg.device));
  CK(cudaDeviceSynchronize());
  return n->log.enable(max_records);
}

extern "C" int ga3c_timing_collect(ga3c_net* n, double* total_ms, int64_t* counts, int32_t n_kernels) {
  if (!n || !total_ms || !counts || n_kernels < K_COUNT) return fail_msg("ga3c_timing_collect: bad argument");
  CK(cudaSetDevice(n->cfg.device));
  CK(cudaDeviceSynchronize());
  return n->log.collect(total_ms, reinterpret_cast<long long*>(counts), n_kernels);
}
