// C-ABI host layer of the low-dimensional MLP networks (ga3c_mlp_*, include/ga3c_b200.h): handle, arenas, workspace and
// the launch sequences.  Kernels: mlp.cu (+ the arena RMSProp of elementwise.cu).
#include <string>
#include <algorithm>
#include <vector>

#include "../../include/ga3c_b200.h"
#include "common.cuh"
#include "kernels.h"
#include "host_util.h"
#include "mlp.cuh"

using namespace ga3c;

namespace {

int fail(const char* where, cudaError_t e) { return fail_cuda(where, e); }

struct MlpParam {
  std::string name;
  int64_t offset, count;
  int32_t ndim;
  int64_t shape[4];
  int32_t live;
};

constexpr int64_t ALIGN_FLOATS = 64;
int64_t align_up(int64_t v) { return (v + ALIGN_FLOATS - 1) / ALIGN_FLOATS * ALIGN_FLOATS; }

}  // namespace

struct ga3c_mlp {
  ga3c_mlp_config cfg;
  int num_sms = 148;
  MlpNet net{};
  std::vector<MlpParam> params;          // TF creation order
  int64_t arena_floats = 0, live_floats = 0;   // live tensors are packed first; [live_floats, arena_floats) is never updated
  float *w = nullptr, *g = nullptr, *ms = nullptr, *mom = nullptr;
  float* part = nullptr;                 // [MLP_MAX_SPLITS][live_floats] weight-gradient partial arenas
  float *clip_ss = nullptr, *clip_scale = nullptr;   // Config.USE_GRAD_CLIP scratch
  float *clip_ss2 = nullptr, *clip_scale2 = nullptr; // ... of the second optimizer's gradient (DUAL_RMSPROP + USE_GRAD_CLIP)
  float *g2 = nullptr, *ms2 = nullptr, *mom2 = nullptr;   // Config.DUAL_RMSPROP: gradient of cost_v, second optimizer's slots
  int64_t skip_lo[2] = {}, skip_hi[2] = {};          // arena ranges of logits_v/* and of the policy head
  // workspace for max_batch rows
  float* act[MLP_MAX_LAYERS] = {};
  float* dz[MLP_MAX_LAYERS] = {};
  float* dlogits = nullptr;
  float* loss_part = nullptr;
  int64_t global_step = 0;
  // tensor-core mode (mlp_tc.cu): the run of wide layers [tc_lo, tc_hi) as 3xTF32 GEMMs for training batches >= tc_min_batch
  int tc_lo = 0, tc_hi = 0, tc_min_batch = 4096;
  bool stream_ok = false, stream_ok_pending = false;                // ... with the narrow ends as streaming passes (mlp_stream.cu; GA3C_MLP_STREAM=0: fused-kernel phases)
  LaunchLog log;                         // launch counter + per-kernel CUDA-event timing (ga3c_mlp_timing_*)
};

static void free_workspace(ga3c_mlp* n) {
  for (int l = 0; l < MLP_MAX_LAYERS; ++l) { cudaFree(n->act[l]); cudaFree(n->dz[l]); n->act[l] = n->dz[l] = nullptr; }
  cudaFree(n->dlogits); cudaFree(n->loss_part);
  n->dlogits = n->loss_part = nullptr;
}

static int alloc_workspace(ga3c_mlp* n, int max_batch) {
  const size_t mb = (size_t)max_batch;
  for (int l = 0; l < n->net.n_layers; ++l) {
    CK(cudaMalloc((void**)&n->act[l], mb * n->net.L[l].n * 4));
    CK(cudaMalloc((void**)&n->dz[l], mb * n->net.L[l].n * 4));
  }
  CK(cudaMalloc((void**)&n->dlogits, mb * n->net.n_out_ld * 4));
  // one row per batch tile (tiles of >= 16 rows), or one per block of the streaming heads kernel (mlp_heads_loss_rows <= 8 per SM)
  const size_t loss_rows = std::max((mb + 15) / 16, (size_t)n->num_sms * 64);
  CK(cudaMalloc((void**)&n->loss_part, loss_rows * 4 * 4));
  n->cfg.max_batch = max_batch;
  return 0;
}

extern "C" int ga3c_mlp_destroy(ga3c_mlp* n) {
  if (!n) return 0;
  cudaFree(n->w); cudaFree(n->g); cudaFree(n->ms); cudaFree(n->mom); cudaFree(n->part);
  cudaFree(n->clip_ss); cudaFree(n->clip_scale); cudaFree(n->clip_ss2); cudaFree(n->clip_scale2);
  cudaFree(n->g2); cudaFree(n->ms2); cudaFree(n->mom2);
  free_workspace(n);
  n->log.clear();
  delete n;
  return 0;
}

extern "C" int ga3c_mlp_create(const ga3c_mlp_config* cfg, ga3c_mlp** out) {
  if (!cfg || !out) return set_error("ga3c_mlp_create: null argument");
  *out = nullptr;
  if (cfg->kind != GA3C_MLP_FORK_VP && cfg->kind != GA3C_MLP_DISCRATE) return set_error("ga3c_mlp_create: unknown kind");
  if (cfg->num_actions < 1 || cfg->num_actions > MAX_ACTIONS) return set_error("ga3c_mlp_create: num_actions must be 1..18");
  if (cfg->state_dim < 1 || cfg->state_dim > MLP_MAX_WIDTH) return set_error("ga3c_mlp_create: state_dim must be 1..256");
  if (cfg->max_batch < 1) return set_error("ga3c_mlp_create: max_batch must be >= 1");
  if (cfg->kind == GA3C_MLP_DISCRATE) {
    if (cfg->n_dense < 1 || cfg->n_dense > MLP_MAX_LAYERS) return set_error("ga3c_mlp_create: n_dense must be 1..8");
    for (int i = 0; i < cfg->n_dense; ++i)
      if (cfg->dense_width[i] < 1 || cfg->dense_width[i] > MLP_MAX_WIDTH)
        return set_error("ga3c_mlp_create: dense widths must be 1..256");
    if (cfg->dense_width[cfg->n_dense - 1] > MLP_MAX_HID) return set_error("ga3c_mlp_create: the last dense width must be <= 128");
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return set_error("ga3c_mlp_create: no CUDA device (there is no CPU fallback)");
  if (cfg->device < 0 || cfg->device >= ndev) return set_error("ga3c_mlp_create: device ordinal out of range");
  CK(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) return set_error("ga3c_mlp_create: kernels are built for sm_100a only");

  ga3c_mlp* n = new ga3c_mlp();
  n->cfg = *cfg;
  n->num_sms = prop.multiProcessorCount;
  const int S = cfg->state_dim, A = cfg->num_actions;
  MlpNet& net = n->net;
  net.kind = cfg->kind; net.state_dim = S; net.num_actions = A;
  net.log_eps = cfg->log_epsilon; net.min_policy = cfg->min_policy; net.log_softmax = cfg->use_log_softmax != 0;

  auto add = [&](const std::string& scope, int k, int width, int live) {    // dense_layer: 'w' [k, width] then 'b' [width]
    MlpParam w{scope + "/w:0", 0, (int64_t)k * width, 2, {k, width, 0, 0}, live};
    MlpParam b{scope + "/b:0", 0, (int64_t)width, 1, {width, 0, 0, 0}, live};
    n->params.push_back(w); n->params.push_back(b);
    return (int)n->params.size() - 2;
  };
  std::vector<int> layer_param;          // index of each live hidden layer's 'w' in params
  if (cfg->kind == GA3C_MLP_FORK_VP) {   // NetworkVP.py:79-85
    const char* names[5] = {"dense11_p", "dense12_p", "dense13_p", "dense14_p", "dense1"};
    const int widths[5] = {4, 256, 256, 100, 64};
    const int acts[5] = {MLP_ACT_LINEAR, MLP_ACT_LINEAR, MLP_ACT_LINEAR, MLP_ACT_SIGMOID, MLP_ACT_SIGMOID};
    int fan = S;
    net.n_layers = 5;
    for (int l = 0; l < 5; ++l) {
      layer_param.push_back(add(names[l], fan, widths[l], 1));
      net.L[l].k = fan; net.L[l].n = widths[l]; net.L[l].act = acts[l];
      fan = widths[l];
    }
  } else {                               // NetworkVP_discrate.py:52-56: every layer reads x; the last one is self.denselayer
    for (int i = 0; i < cfg->n_dense; ++i) {
      const int live = i == cfg->n_dense - 1;
      const int idx = add("dense1_" + std::to_string(i + 1) + "_p", S, cfg->dense_width[i], live);
      if (live) layer_param.push_back(idx);
    }
    net.n_layers = 1;
    net.L[0].k = S; net.L[0].n = cfg->dense_width[cfg->n_dense - 1]; net.L[0].act = MLP_ACT_SIGMOID;
  }
  net.hid = net.L[net.n_layers - 1].n;
  {  // the first run of consecutive layers wide enough for 128 x BN tensor-core tiles (fork NetworkVP: 256 -> 256 -> 100 -> 64)
    int lo = 0;
    while (lo < net.n_layers && !mlp_tc_layer_ok(net.L[lo])) ++lo;
    int hi = lo;
    while (hi < net.n_layers && mlp_tc_layer_ok(net.L[hi])) ++hi;
    if (lo >= 1 && hi > lo) { n->tc_lo = lo; n->tc_hi = hi; }       // (layer 0 reads x: never wide here)
    if (const char* e = getenv("GA3C_MLP_TC")) {                     // 0: never; N > 0: from training batches of N rows
      const int v = atoi(e);
      if (v <= 0) n->tc_lo = n->tc_hi = 0; else n->tc_min_batch = v;
    }
    const char* es = getenv("GA3C_MLP_STREAM");
    n->stream_ok_pending = !(es && atoi(es) == 0);
  }
  const int pv = add("logits_v", net.hid, 1, 1);
  int px, py = -1;
  if (cfg->kind == GA3C_MLP_FORK_VP) {
    px = add("logits_p/out_x", net.hid, A, 1);
    py = add("logits_p/out_y", net.hid, A, 1);
    net.n_out = 1 + 2 * A;
  } else {
    px = add("logits_p", net.hid, A, 1);
    net.n_out = 1 + A;
  }
  net.n_out_ld = (net.n_out + 3) & ~3;
  int64_t cur = 0;
  for (int pass = 1; pass >= 0; --pass) {          // live tensors first, then the gradient-less ones
    for (MlpParam& p : n->params)
      if (p.live == pass) { p.offset = cur; cur = align_up(cur + p.count); }
    if (pass == 1) n->live_floats = cur;
  }
  n->arena_floats = cur;
  for (int l = 0; l < net.n_layers; ++l) {
    net.L[l].w_off = (int)n->params[layer_param[l]].offset;
    net.L[l].b_off = (int)n->params[layer_param[l] + 1].offset;
  }
  net.wv_off = (int)n->params[pv].offset; net.bv_off = (int)n->params[pv + 1].offset;
  net.wp_off = (int)n->params[px].offset; net.bp_off = (int)n->params[px + 1].offset;
  if (py >= 0) { net.wy_off = (int)n->params[py].offset; net.by_off = (int)n->params[py + 1].offset; }
  n->skip_lo[0] = n->params[pv].offset; n->skip_hi[0] = n->params[pv + 1].offset + n->params[pv + 1].count;
  const int plast = py >= 0 ? py + 1 : px + 1;      // the policy head's variables are consecutive in the arena
  n->skip_lo[1] = n->params[px].offset; n->skip_hi[1] = n->params[plast].offset + n->params[plast].count;
  n->stream_ok = n->stream_ok_pending && n->tc_hi > n->tc_lo && mlp_stream_ok(net, n->tc_lo, n->tc_hi);

  const size_t ab = (size_t)n->arena_floats * 4;
  float** arenas[4] = {&n->w, &n->g, &n->ms, &n->mom};
  for (float** a : arenas) {
    e = cudaMalloc((void**)a, ab);
    if (e != cudaSuccess) { ga3c_mlp_destroy(n); return fail("cudaMalloc", e); }
    cudaMemset(*a, 0, ab);
  }
  {  // ms slot starts at 1.0 [TF-SEMANTICS]
    std::vector<float> ones((size_t)n->arena_floats, 1.0f);
    cudaMemcpy(n->ms, ones.data(), ab, cudaMemcpyHostToDevice);
  }
  e = cudaMalloc((void**)&n->part, (size_t)MLP_MAX_SPLITS * n->live_floats * 4);
  if (e != cudaSuccess) { ga3c_mlp_destroy(n); return fail("cudaMalloc", e); }
  cudaMemset(n->part, 0, (size_t)MLP_MAX_SPLITS * n->live_floats * 4);
  if (cfg->dual_rmsprop) {
    float** extra[3] = {&n->g2, &n->ms2, &n->mom2};
    for (float** a : extra) {
      e = cudaMalloc((void**)a, ab);
      if (e != cudaSuccess) { ga3c_mlp_destroy(n); return fail("cudaMalloc", e); }
      cudaMemset(*a, 0, ab);
    }
    std::vector<float> ones((size_t)n->arena_floats, 1.0f);
    cudaMemcpy(n->ms2, ones.data(), ab, cudaMemcpyHostToDevice);
  }
  if (cfg->use_grad_clip) {
    for (const MlpParam& p : n->params)
      if (!p.live && !cfg->dual_rmsprop) {     // opt.compute_gradients yields (None, var) and tf.clip_by_average_norm(None, ..) raises;
                                               // the DUAL_RMSPROP branch filters `if not g is None` (NetworkVP_discrate.py:110,:115)
        ga3c_mlp_destroy(n);
        return set_error("ga3c_mlp_create: USE_GRAD_CLIP with gradient-less variables (NetworkVP_discrate.py:55 builds every "
                         "DENSE_LAYERS entry from x): the reference graph cannot be built either");
      }
    if ((int)n->params.size() > CLIP_MAX_TENSORS) { ga3c_mlp_destroy(n); return set_error("ga3c_mlp_create: too many variables for USE_GRAD_CLIP"); }
    int64_t mx = 0;
    for (const MlpParam& p : n->params) mx = p.count > mx ? p.count : mx;
    e = cudaMalloc((void**)&n->clip_ss, n->params.size() * clip_chunks(mx) * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&n->clip_scale, n->params.size() * sizeof(float));
    if (e == cudaSuccess && cfg->dual_rmsprop) {
      e = cudaMalloc((void**)&n->clip_ss2, n->params.size() * clip_chunks(mx) * sizeof(float));
      if (e == cudaSuccess) e = cudaMalloc((void**)&n->clip_scale2, n->params.size() * sizeof(float));
    }
    if (e != cudaSuccess) { ga3c_mlp_destroy(n); return fail("cudaMalloc", e); }
  }
  if (int r = alloc_workspace(n, cfg->max_batch)) { ga3c_mlp_destroy(n); return r; }
  if (int r = configure_mlp()) { ga3c_mlp_destroy(n); return fail("cudaFuncSetAttribute", (cudaError_t)r); }
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { ga3c_mlp_destroy(n); return fail("ga3c_mlp_create sync", e); }
  *out = n;
  return 0;
}

extern "C" int ga3c_mlp_reserve(ga3c_mlp* n, int32_t max_batch) {
  if (!n) return set_error("ga3c_mlp_reserve: null handle");
  if (max_batch <= n->cfg.max_batch) return 0;
  CK(cudaSetDevice(n->cfg.device));
  CK(cudaDeviceSynchronize());
  free_workspace(n);
  if (int r = alloc_workspace(n, max_batch)) { n->cfg.max_batch = 0; return r; }
  return 0;
}

extern "C" int ga3c_mlp_param_count(const ga3c_mlp* n) { return n ? (int)n->params.size() : 0; }

extern "C" int ga3c_mlp_param_info(const ga3c_mlp* n, int i, const char** name, int64_t* offset, int32_t* ndim,
                                   int64_t shape[4], int32_t* live) {
  if (!n || i < 0 || i >= (int)n->params.size()) return set_error("ga3c_mlp_param_info: bad index");
  const MlpParam& d = n->params[i];
  if (name) *name = d.name.c_str();
  if (offset) *offset = d.offset;
  if (ndim) *ndim = d.ndim;
  if (shape) for (int k = 0; k < 4; ++k) shape[k] = d.shape[k];
  if (live) *live = d.live;
  return 0;
}

extern "C" int64_t ga3c_mlp_arena_floats(const ga3c_mlp* n) { return n ? n->arena_floats : 0; }

static float* arena_of(ga3c_mlp* n, int which) {
  switch (which) {
    case 0: return n->w; case 1: return n->g; case 2: return n->ms; case 3: return n->mom;
    case 4: return n->g2; case 5: return n->ms2; case 6: return n->mom2;      // null unless dual_rmsprop
  }
  return nullptr;
}

extern "C" int ga3c_mlp_arena_upload(ga3c_mlp* n, int which, const float* host, int64_t nf) {
  if (!n || !host) return set_error("ga3c_mlp_arena_upload: null argument");
  float* dst = arena_of(n, which);
  if (!dst || nf != n->arena_floats) return set_error("ga3c_mlp_arena_upload: bad arena id or size");
  CK(cudaSetDevice(n->cfg.device));
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(dst, host, (size_t)nf * 4, cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int ga3c_mlp_arena_download(ga3c_mlp* n, int which, float* host, int64_t nf) {
  if (!n || !host) return set_error("ga3c_mlp_arena_download: null argument");
  float* src = arena_of(n, which);
  if (!src || nf != n->arena_floats) return set_error("ga3c_mlp_arena_download: bad arena id or size");
  CK(cudaSetDevice(n->cfg.device));
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(host, src, (size_t)nf * 4, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int64_t ga3c_mlp_global_step(const ga3c_mlp* n) { return n ? n->global_step : -1; }
extern "C" int ga3c_mlp_set_global_step(ga3c_mlp* n, int64_t s) { if (!n) return -1; n->global_step = s; return 0; }
extern "C" int64_t ga3c_mlp_launch_count(const ga3c_mlp* n) { return n ? n->log.launches : 0; }

// kernel-level test entry for the 3xTF32 tensor-core GEMMs of mlp_tc.cu (all pointers device, fp32 row-major):
//   mode 0  out [m, n] = act(a [m, k] x b [k, n] + aux [n])                       (forward)
//   mode 1  out [m, k] = (a [m, n] x b [k, n]^T) * act'(aux [m, k])               (data gradient)
//   mode 2  out [splits][k, n] = a [m, k]^T x b [m, n] per batch split of rows_per_split rows (weight gradient)
extern "C" int ga3c_debug_tf32x3_gemm(int32_t mode, const float* a, const float* b, const float* aux, float* out, int32_t m,
                                      int32_t k, int32_t nn, int32_t act, int32_t splits, int32_t rows_per_split, void* stream) {
  if (!a || !b || !out || m < 1 || k < 4 || nn < 4 || k > 256 || nn > 256 || (k & 3) || (nn & 3))
    return set_error("ga3c_debug_tf32x3_gemm: bad argument (k, n multiples of 4 up to 256)");
  cudaStream_t st = (cudaStream_t)stream;
  int r = 0;
  if (mode == 0) r = launch_mlp_tc_fwd(b, aux, k, nn, act, a, out, m, st);
  else if (mode == 1) r = launch_mlp_tc_dgrad(b, k, nn, a, aux, act, out, m, st);
  else if (mode == 2) {
    if (splits < 1 || rows_per_split < 32 || rows_per_split % 32) return set_error("ga3c_debug_tf32x3_gemm: rows_per_split must be a multiple of 32");
    r = launch_mlp_tc_wgrad(k, nn, a, b, m, splits, rows_per_split, out, nullptr, (int64_t)k * nn, st);
  } else return set_error("ga3c_debug_tf32x3_gemm: mode must be 0, 1 or 2");
  if (r) return fail("ga3c_debug_tf32x3_gemm", (cudaError_t)r);
  return 0;
}
extern "C" int ga3c_mlp_workspace_ptr(ga3c_mlp* n, int32_t which, int32_t layer, void** ptr, int32_t* width) {
  if (!n || !ptr || !width) return set_error("ga3c_mlp_workspace_ptr: null argument");
  if (which < 0 || which > 1 || layer < 0 || layer >= n->net.n_layers) return set_error("ga3c_mlp_workspace_ptr: bad matrix id");
  *ptr = which == 0 ? n->act[layer] : n->dz[layer];
  *width = n->net.L[layer].n;
  return 0;
}

static int check_batch(ga3c_mlp* n, int batch, const char* who) {
  if (!n) return set_error(std::string(who) + ": null handle");
  if (batch < 1 || batch > n->cfg.max_batch)
    return set_error(std::string(who) + ": batch " + std::to_string(batch) + " outside 1.." + std::to_string(n->cfg.max_batch));
  return 0;
}

static MlpStepArgs step_args(ga3c_mlp* n, const float* x, int batch) {
  MlpStepArgs s{};
  s.w = n->w; s.x = x; s.batch = batch;
  for (int l = 0; l < n->net.n_layers; ++l) { s.act[l] = n->act[l]; s.dz[l] = n->dz[l]; }
  s.dlogits = n->dlogits; s.loss_part = n->loss_part;
  return s;
}

extern "C" int ga3c_mlp_predict(ga3c_mlp* n, const float* x, int32_t batch, float* p_out, float* v_out, void* stream) {
  if (int r = check_batch(n, batch, "ga3c_mlp_predict")) return r;
  if (!x || !p_out || !v_out) return set_error("ga3c_mlp_predict: null buffer");
  CK(cudaSetDevice(n->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  MlpStepArgs s = step_args(n, x, batch);
  s.p_out = p_out; s.v_out = v_out; s.train = 0;
  if (n->stream_ok && batch >= n->tc_min_batch) {
    // large batches: front as a streaming pass, the wide layers as 3xTF32 GEMMs, heads as a streaming pass (the layer outputs go
    // through the training workspace; the handle mutex of the host layer serialises predict and train)
    const MlpNet& net = n->net;
    LAUNCH(n, K_MLP_FUSED, st, launch_mlp_front_fwd(net, s, n->num_sms, st));
    for (int l = n->tc_lo; l < n->tc_hi; ++l)
      LAUNCH(n, K_MLP_TC, st, launch_mlp_tc_fwd(n->w + net.L[l].w_off, n->w + net.L[l].b_off, net.L[l].k, net.L[l].n, net.L[l].act,
                                                n->act[l - 1], n->act[l], batch, st));
    LAUNCH(n, K_MLP_FUSED, st, launch_mlp_heads(net, s, n->num_sms, st));
    return 0;
  }
  LAUNCH(n, K_MLP_FUSED, st, launch_mlp_fused(n->net, s, n->num_sms, st));
  return 0;
}

static int mlp_fb_impl(ga3c_mlp* n, const float* x, const float* yr, const float* a, int32_t batch, float beta, float* loss,
                       void* stream, int part, float* g_dst) {
  if (int r = check_batch(n, batch, "ga3c_mlp_forward_backward")) return r;
  if (!x || !yr || !a) return set_error("ga3c_mlp_forward_backward: null buffer");
  CK(cudaSetDevice(n->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  MlpStepArgs s = step_args(n, x, batch);
  s.yr = yr; s.a = a; s.beta = beta; s.train = 1; s.part = part;
  const int tm = mlp_tile_rows(batch, n->num_sms);
  if (n->tc_hi > n->tc_lo && batch >= n->tc_min_batch) {
    // the wide layers as 3xTF32 tensor-core GEMMs over the activations the fused kernel keeps in HBM anyway
    const MlpNet& net = n->net;
    const int lo = n->tc_lo, hi = n->tc_hi;
    s.tc_lo = lo; s.tc_hi = hi;
    const bool streamed = n->stream_ok;       // narrow ends as streaming passes (mlp_stream.cu) instead of fused-kernel phases
    s.phase = 1;
    if (streamed) LAUNCH(n, K_MLP_FUSED, st, launch_mlp_front_fwd(net, s, n->num_sms, st));
    else LAUNCH(n, K_MLP_FUSED, st, launch_mlp_fused(net, s, n->num_sms, st));
    for (int l = lo; l < hi; ++l)
      LAUNCH(n, K_MLP_TC, st, launch_mlp_tc_fwd(n->w + net.L[l].w_off, n->w + net.L[l].b_off, net.L[l].k, net.L[l].n, net.L[l].act,
                                                n->act[l - 1], n->act[l], batch, st));
    s.phase = 2;
    if (streamed) LAUNCH(n, K_MLP_FUSED, st, launch_mlp_heads(net, s, n->num_sms, st));
    else LAUNCH(n, K_MLP_FUSED, st, launch_mlp_fused(net, s, n->num_sms, st));
    for (int l = hi - 1; l >= lo; --l)
      LAUNCH(n, K_MLP_TC, st, launch_mlp_tc_dgrad(n->w + net.L[l].w_off, net.L[l].k, net.L[l].n, n->dz[l], n->act[l - 1],
                                                  net.L[l - 1].act, n->dz[l - 1], batch, st));
    if (lo >= 2) {
      s.phase = 3;
      if (streamed) LAUNCH(n, K_MLP_FUSED, st, launch_mlp_front_bwd(net, s, n->num_sms, st));
      else LAUNCH(n, K_MLP_FUSED, st, launch_mlp_fused(net, s, n->num_sms, st));
    }
    const int splits = MLP_MAX_SPLITS;
    int rows = (batch + splits - 1) / splits;
    rows = (rows + 31) / 32 * 32;
    // weight gradients: tensor-core layers by GEMM (+ a streaming pass for the bias), fan-in <= 4 layers by the streaming kernel
    // alone, whatever is left (the head matrix) by the tile kernel
    for (int l = 0; l < net.n_layers; ++l)
      if ((l >= lo && l < hi) || net.L[l].k <= 4) s.wgrad_skip |= 1u << l;
    if (mlp_heads_wgrad_ok(net)) {
      s.wgrad_skip |= 1u << net.n_layers;
      LAUNCH(n, K_MLP_WGRAD, st, launch_mlp_heads_wgrad(net, s, n->part, n->live_floats, splits, rows, st));
    }
    if (s.wgrad_skip != (2u << net.n_layers) - 1u)
      LAUNCH(n, K_MLP_WGRAD, st, launch_mlp_wgrad(net, s, n->part, n->live_floats, splits, st, rows));
    for (int l = 0; l < net.n_layers; ++l) {
      if (!((s.wgrad_skip >> l) & 1u)) continue;
      const bool tc = l >= lo && l < hi;
      const float* in = l == 0 ? x : n->act[l - 1];
      if (tc)
        LAUNCH(n, K_MLP_TC, st, launch_mlp_tc_wgrad(net.L[l].k, net.L[l].n, in, n->dz[l], batch, splits, rows,
                                                    n->part + net.L[l].w_off, n->part + net.L[l].b_off, n->live_floats, st));
      LAUNCH(n, K_MLP_WGRAD, st, launch_mlp_skinny_wgrad(tc ? 0 : net.L[l].k, net.L[l].n, in, n->dz[l], batch, splits, rows,
                                                         n->part + net.L[l].w_off, n->part + net.L[l].b_off, n->live_floats, st));
    }
    LAUNCH(n, K_MLP_REDUCE, st, launch_mlp_reduce(n->part, n->live_floats, splits, g_dst, (int)n->live_floats, n->loss_part,
                                                  streamed ? mlp_heads_loss_rows(batch, n->num_sms) : (batch + tm - 1) / tm, loss, st));
    return 0;
  }
  LAUNCH(n, K_MLP_FUSED, st, launch_mlp_fused(n->net, s, n->num_sms, st));
  const int splits = mlp_wgrad_splits(n->net, batch, n->num_sms);
  LAUNCH(n, K_MLP_WGRAD, st, launch_mlp_wgrad(n->net, s, n->part, n->live_floats, splits, st));
  LAUNCH(n, K_MLP_REDUCE, st, launch_mlp_reduce(n->part, n->live_floats, splits, g_dst, (int)n->live_floats, n->loss_part,
                                                (batch + tm - 1) / tm, loss, st));
  return 0;
}

extern "C" int ga3c_mlp_forward_backward(ga3c_mlp* n, const float* x, const float* yr, const float* a, int32_t batch,
                                         float beta, float* loss, void* stream) {
  if (!n) return set_error("ga3c_mlp_forward_backward: null handle");
  return mlp_fb_impl(n, x, yr, a, batch, beta, loss, stream, 0, n->g);
}

extern "C" int ga3c_mlp_apply_rmsprop(ga3c_mlp* n, float lr, void* stream) {
  if (!n) return set_error("ga3c_mlp_apply_rmsprop: null handle");
  CK(cudaSetDevice(n->cfg.device));
  RmsPropArgs a{};
  a.w = n->w; a.ms = n->ms; a.mom = n->mom; a.g = n->g; a.w1_shadow = nullptr;
  a.n_floats = n->live_floats;           // the gradient-less variables behind the live prefix are never touched
  a.w1_offset = n->live_floats; a.w1_count = 0;
  a.lr = lr; a.decay = n->cfg.rmsprop_decay; a.momentum = n->cfg.rmsprop_momentum; a.eps = n->cfg.rmsprop_epsilon;
  if (n->cfg.dual_rmsprop) return set_error("ga3c_mlp_apply_rmsprop: with DUAL_RMSPROP use ga3c_mlp_train_step (two backward passes)");
  if (n->cfg.use_grad_clip) {            // tf.clip_by_average_norm per variable (NetworkVP.py:138-141, NetworkVP_discrate.py:118-121)
    ClipArgs c{};
    c.g = n->g; c.n_tensors = (int)n->params.size();
    int64_t mx = 0;
    for (int i = 0; i < c.n_tensors; ++i) {
      c.offset[i] = n->params[i].offset; c.count[i] = n->params[i].count;
      mx = c.count[i] > mx ? c.count[i] : mx;
    }
    c.max_chunks = clip_chunks(mx); c.clip = n->cfg.grad_clip_norm; c.chunk_ss = n->clip_ss; c.scale = n->clip_scale;
    LAUNCH(n, K_RMSPROP, (cudaStream_t)stream, launch_rmsprop_clipped(a, c, (cudaStream_t)stream));
    n->log.launches += 2;
    if (n->cfg.kind == GA3C_MLP_FORK_VP) n->global_step += 1;     // the fork's NetworkVP passes global_step, _discrate does not
    return 0;
  }
  LAUNCH(n, K_RMSPROP, (cudaStream_t)stream, launch_rmsprop(a, (cudaStream_t)stream));
  n->global_step += 1;                   // opt.minimize(..., global_step=self.global_step)
  return 0;
}

extern "C" int ga3c_mlp_train_step(ga3c_mlp* n, const float* x, const float* yr, const float* a, int32_t batch, float lr,
                                   float beta, float* loss, void* stream) {
  if (n && n->cfg.dual_rmsprop) {
    // Config.DUAL_RMSPROP (NetworkVP.py:107-118, :143-147): cost_p into g, cost_v into g2, both steps from the same weights
    if (int r = mlp_fb_impl(n, x, yr, a, batch, beta, loss, stream, 1, n->g)) return r;
    if (int r = mlp_fb_impl(n, x, yr, a, batch, beta, nullptr, stream, 2, n->g2)) return r;
    RmsPropDualArgs d{};
    d.a.w = n->w; d.a.ms = n->ms; d.a.mom = n->mom; d.a.g = n->g; d.a.w1_shadow = nullptr;
    d.a.n_floats = n->live_floats; d.a.w1_offset = n->live_floats; d.a.w1_count = 0;
    d.a.lr = lr; d.a.decay = n->cfg.rmsprop_decay; d.a.momentum = n->cfg.rmsprop_momentum; d.a.eps = n->cfg.rmsprop_epsilon;
    d.g2 = n->g2; d.ms2 = n->ms2; d.mom2 = n->mom2;
    d.skip_lo[0] = n->skip_lo[0]; d.skip_hi[0] = n->skip_hi[0];        // cost_p: not logits_v
    d.skip_lo[2] = n->skip_lo[1]; d.skip_hi[2] = n->skip_hi[1];        // cost_v: not the policy head
    if (n->cfg.use_grad_clip) {
      // tf.clip_by_norm per variable and optimizer (NetworkVP.py:127-137, NetworkVP_discrate.py:107-117); the fork's NetworkVP
      // hands global_step to both apply_gradients calls, NetworkVP_discrate to neither
      ClipArgs c[2] = {};
      int64_t mx = 0;
      for (int w = 0; w < 2; ++w) {
        c[w].g = w ? n->g2 : n->g;
        c[w].n_tensors = 0;
        for (const MlpParam& p : n->params) {
          if (!p.live) continue;                                       // gradient is None: filtered by the reference
          c[w].offset[c[w].n_tensors] = p.offset; c[w].count[c[w].n_tensors] = p.count; ++c[w].n_tensors;
          mx = p.count > mx ? p.count : mx;
        }
        c[w].clip = n->cfg.grad_clip_norm; c[w].by_norm = 1;
        c[w].chunk_ss = w ? n->clip_ss2 : n->clip_ss; c[w].scale = w ? n->clip_scale2 : n->clip_scale;
      }
      c[0].max_chunks = c[1].max_chunks = clip_chunks(mx);
      LAUNCH(n, K_RMSPROP, (cudaStream_t)stream, launch_rmsprop_dual_clipped(d, c[0], c[1], (cudaStream_t)stream));
      n->log.launches += 4;
      if (n->cfg.kind == GA3C_MLP_FORK_VP) n->global_step += 2;
      return 0;
    }
    LAUNCH(n, K_RMSPROP, (cudaStream_t)stream, launch_rmsprop_dual(d, (cudaStream_t)stream));
    n->global_step += 2;
    return 0;
  }
  if (int r = ga3c_mlp_forward_backward(n, x, yr, a, batch, beta, loss, stream)) return r;
  return ga3c_mlp_apply_rmsprop(n, lr, stream);
}

extern "C" int ga3c_mlp_timing_enable(ga3c_mlp* n, int32_t max_records) {
  if (!n || max_records < 0) return set_error("ga3c_mlp_timing_enable: bad argument");
  CK(cudaSetDevice(n->cfg.device));
  CK(cudaDeviceSynchronize());
  return n->log.enable(max_records);
}

extern "C" int ga3c_mlp_timing_collect(ga3c_mlp* n, double* total_ms, int64_t* counts, int32_t n_kernels) {
  if (!n || !total_ms || !counts || n_kernels < K_COUNT) return set_error("ga3c_mlp_timing_collect: bad argument");
  CK(cudaSetDevice(n->cfg.device));
  CK(cudaDeviceSynchronize());
  return n->log.collect(total_ms, reinterpret_cast<long long*>(counts), n_kernels);
}
