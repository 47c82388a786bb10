// Internal launch interface between the C-ABI host layer (net.cu) and the kernel files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

namespace ga3c {

// step timeline trace: attach (or detach with nullptr) the [K_COUNT][TRACE_SLOTS] uint64 buffer, one call per kernel file
int trace_attach_conv_fwd(unsigned long long* buf);
int trace_attach_conv_bwd_fused(unsigned long long* buf);
int trace_attach_dense_tc(unsigned long long* buf);
int trace_attach_heads(unsigned long long* buf);
int trace_attach_elementwise(unsigned long long* buf);
int trace_attach_dense_heads(unsigned long long* buf);

// pipeline event log of CTA 0 (debug): attach a zeroed [16 warps][1024][2] uint64 buffer (or nullptr)
int evt_attach_conv_fwd(unsigned long long* buf);
int evt_attach_conv_bwd(unsigned long long* buf);
int evt_attach_elementwise(unsigned long long* buf);
int evt_attach_dense_heads(unsigned long long* buf);
int evt_attach_dense_tc(unsigned long long* buf);      // records only in a -DGA3C_DENSE_EVT build

// L2 residency hints on the conv kernels' loads and stores (common.cuh); GA3C_L2_HINTS=0 switches them off (evict_normal)
// bit 0: frames loaded evict_first, bit 1: activations stored evict_last, bit 2: the conv backward loads evict_first.
// Measured at B = 1024 / 4096 (profiles/r2g_l2_hint_sweep.txt): predict gains 3 % from bit 0; the fp32 train step is fastest
// with none of them (the L2 already keeps what fits), the uint8 train step with all three (-2.5 %).  GA3C_L2_HINTS overrides.
inline int l2_hints(bool train, bool x_u8) {
  static const int env = [] { const char* e = getenv("GA3C_L2_HINTS"); return e ? atoi(e) : -1; }();
  if (env >= 0) return env;
  return train ? (x_u8 ? 7 : 0) : 1;
}

// one-time per-process function-attribute setup (dynamic smem opt-in); returns cudaError_t as int
int configure_conv_fwd();
int configure_conv_bwd_fused();
int configure_dense_tc();

// conv_fwd.cu -- x fp32 (or uint8, x_u8: k/128 - 1 applied on the fly) [B,28224] -> n2 bf16 [B,3872]; when training also n1 (bf16, in
// the Blk2 operand layout) and xblk (the bf16 block matrix of the frame): both are read back by the conv backward (common.cuh)
int launch_conv_fwd(const void* x, bool x_u8, const float* w11, const float* b11, const float* w12, const float* b12,
                    uint8_t* n1_out, uint8_t* xblk_out, uint16_t* n2_out, int batch, int num_sms, cudaStream_t stream);

// dense_tc.cu -- the three dense1 GEMMs (NetworkDNav.py:90 and its autodiff) on tcgen05 / TMEM / TMA.  fwd leaves `splits` raw fp32 partial tiles
// in d1_part[splits][B][256]; the heads kernel sums them, adds the bias and applies the ReLU.
int dense_fwd_splits(int batch, int num_sms);
// data parallel, side-stream exchange: flags = this rank's comm block + DPC_BIGDONE ([world] x 64 B, u64 step numbers), see
// gemm_tc_body (dense_tc.cu).  flags == nullptr: no gate.
struct DpGate {
  const uint8_t* flags;
  void* err;                        // this rank's error word (bit 16 on a wait that gave up)
  unsigned long long step;          // every flag must have reached this step
  int world;
};
int launch_dense_fwd_tc(const uint16_t* n2, const uint16_t* w1bf, float* d1_part, int batch, int splits, cudaStream_t stream,
                        const DpGate* gate = nullptr);
int launch_dense_dgrad_tc(const uint16_t* dd1, const uint16_t* w1bf, const uint16_t* n2, uint8_t* dn2, int batch,
                          cudaStream_t stream);       // dn2 in the G operand layout (common.cuh), borders untouched
int launch_dense_wgrad_tc(const uint16_t* n2, const uint16_t* dd1, float* g_w1, int batch, cudaStream_t stream);
// dgrad + wgrad tiles in one grid (both only need dd1): the single-GPU / fused-DP step uses this one
// data parallel, exchange at the end of the step: where the dense1/w gradient tiles push the slices this rank does not own
// (dense_tc.cu EpiWgradPush; receive buffers [2 parities][world][per4 float4 in LL format] at recv_off of every slab)
struct WgradPush {
  uint8_t* peer[8];
  int rank, world;
  long long per4, recv_off;
  unsigned int flag;         // the step number; parity = flag & 1
};
int launch_dense_bwd_tc(const uint16_t* dd1, const uint16_t* w1bf, const uint16_t* n2, uint8_t* dn2, float* g_w1, int batch,
                        const WgradPush* push, cudaStream_t stream);

// heads.cu -- value / policy heads, softmax, A3C loss and its backward (NetworkVP_discrate.py:60-85)
struct HeadsArgs {
  const float* d1_part; // [n_split][B,256] raw split-K partial sums of dense1 (no bias, no ReLU)
  int n_split;
  const float* b1;      // dense1 bias
  float* d1;            // [B,256] out: relu(sum of partials + b1)
  const float *wp, *bp, *wv, *bv;
  const float *yr, *a;  // train only
  int batch, num_actions;
  float beta, log_eps, min_policy;
  float *p_out, *v_out;        // may be null in train mode
  uint16_t* dd1;               // [B,256] bf16, train only
  // train only: this CTA's partial sums go to slab blockIdx.x of the gradient-partial workspace (pointers are
  // into slab 0, consecutive slabs gp_stride floats apart); launch_grad_reduce sums the slabs
  float *g_wp, *g_bp, *g_wv, *g_bv, *g_b1;
  float* loss;                 // [4] (cost_p_1, cost_p_2, cost_v, 0) partial sums, same slab scheme
  int64_t gp_stride;
  int train;
  int log_softmax;             // Config.USE_LOG_SOFTMAX
  int part;                    // Config.DUAL_RMSPROP passes: 0 gradient of cost_all, 1 of cost_p alone, 2 of cost_v alone
  int preload;                 // the head weights may be read before the dependency wait (heads.cu)
};
int heads_grid(int batch, int num_sms);          // CTAs (= slabs written) of a training launch
int launch_heads(const HeadsArgs& args, int num_sms, cudaStream_t stream);

// dense_heads.cu -- dense1 forward (cluster split-K, partial tiles reduced over distributed shared memory) with the heads, the
// loss and its backward as its epilogue: replaces launch_dense_fwd_tc + launch_heads (args.d1_part / n_split are not used).
// In training it writes dense_heads_ctas(batch) slabs.
int configure_dense_heads();
int dense_heads_ctas(int batch);
int launch_dense_heads(const uint16_t* n2, const uint16_t* w1bf, const HeadsArgs& args, cudaStream_t stream);

// conv_bwd_fused.cu
// Weight / bias gradients are per-CTA partial sums: CTA i stores into slab i (g_* point into slab 0, slabs are
// gp_stride floats apart) and launch_grad_reduce adds the slabs in a fixed order.  No contended atomics, and the
// step is bit-reproducible.
struct DpBigArgs;                                // dp_exchange.cuh
int conv_bwd_grid(int batch, int num_sms, int n_exch = 0);   // conv CTAs (= slabs written); n_exch SMs are left to the exchange CTAs
// conv_bwd_fused.cu -- conv12 data gradient (dn1, kept on chip; dn1_out: optional copy for tests), conv12 and conv11
// weight / bias gradients in one kernel on tcgen05
// what warps 12-15 of every conv backward CTA do with dense1/w, whose gradient is final when that kernel starts:
//   mode 0 nothing; mode 1 RMSProp + bf16 shadow refresh over the whole tensor (single GPU; the pointers are the dense1/w ranges
//   of the arenas, n4 its size in float4); mode 2 the data-parallel exchange described by the DpBigArgs of the launch
struct ConvBwdOpt {
  int mode;
  const float* g;
  float *w, *ms, *mom;
  uint16_t* shadow;
  long long n4;
  float lr, decay, momentum, eps;
};
int launch_conv_bwd(const uint8_t* xblk, const uint8_t* n1b2, const uint8_t* dn2g, const float* w12, uint16_t* dn1_out,
                    float* g_w11, float* g_b11, float* g_w12, float* g_b12, int64_t gp_stride, int batch, int num_sms, bool x_u8,
                    const ConvBwdOpt& opt, const DpBigArgs* dp, cudaStream_t stream);   // dp != null: data parallel (+ dp->n_exch exchange CTAs)

// elementwise.cu
// out[j] = sum over slabs i < count(j) of part[i * stride + j], j in [0, n_floats): the per-CTA gradient partials of
// the heads / conv12_bwd / conv11_wgrad kernels -> the small-tensor prefix of the gradient arena (+ the loss sums).
constexpr int GR_MAX_SEG = 4;
struct GradReduceArgs {
  const float* part;
  int64_t stride;
  float* out;                       // floats [0, out_floats) go here ...
  float* out_tail;                  // ... and floats [out_floats, n_floats) here (the loss sums; may be null)
  int out_floats, n_floats;         // multiples of 4
  int seg_end[GR_MAX_SEG];          // segment s covers [seg_end[s-1], seg_end[s]) and sums seg_count[s] slabs
  int seg_count[GR_MAX_SEG];
};
int launch_grad_reduce(const GradReduceArgs& a, cudaStream_t stream);
struct RmsPropArgs {
  float *w, *ms, *mom;
  const float* g;
  uint16_t* w1_shadow;     // bf16 shadow of dense1/w
  int64_t n_floats;        // arena size (multiple of 4)
  int64_t w1_offset, w1_count;
  float lr, decay, momentum, eps;
  int preload;             // the optimizer kernels may load w / ms / mom before their dependency wait (rmsprop_reduce_kernel)
};
int launch_rmsprop(const RmsPropArgs& a, cudaStream_t stream);
// Config.USE_GRAD_CLIP: tf.clip_by_average_norm per variable, then RMSProp (3 launches: chunk sums of squares, per-tensor
// scale, update).  offset / count: the variables' element ranges in the arena; chunk_ss: [n_tensors][max_chunks] scratch,
// max_chunks = clip_chunks(largest count); scale: [n_tensors] scratch.
constexpr int CLIP_MAX_TENSORS = 24;
struct ClipArgs {
  const float* g;
  int n_tensors, max_chunks;
  int64_t offset[CLIP_MAX_TENSORS], count[CLIP_MAX_TENSORS];
  float clip;
  float *chunk_ss, *scale;
  int by_norm;                      // 0: tf.clip_by_average_norm (||g|| / n against clip), 1: tf.clip_by_norm (||g|| against clip)
};
int clip_chunks(int64_t max_count);
int launch_rmsprop_clipped(const RmsPropArgs& a, const ClipArgs& c, cudaStream_t stream);
// Config.DUAL_RMSPROP: two optimizers with their own slots, both steps taken from the weights the call started with:
// w <- w - step(g, ms, mom) - step(g2, ms2, mom2).  Elements inside [skip_lo[i], skip_hi[i]) are left alone by optimizer i / 2
// (i = 0, 1: first optimizer; 2, 3: second): the variables its cost does not reach.
struct RmsPropDualArgs {
  RmsPropArgs a;                   // w, g, ms, mom + scalars (+ bf16 shadow range)
  const float* g2;
  float *ms2, *mom2;
  int64_t skip_lo[4], skip_hi[4];
  // DUAL_RMSPROP + USE_GRAD_CLIP (NetworkVP_discrate.py:107-117): per-variable tf.clip_by_norm scales of g (scale1) and g2
  // (scale2), indexed like the tensor table below; null = unclipped
  const float *scale1, *scale2;
  int n_tensors;
  int64_t t_offset[CLIP_MAX_TENSORS], t_count[CLIP_MAX_TENSORS];
};
int launch_rmsprop_dual(const RmsPropDualArgs& d, cudaStream_t stream);
// the same with both gradients clipped first: c1 describes g (d.a.g), c2 describes g2; both by_norm.  5 launches.
int launch_rmsprop_dual_clipped(RmsPropDualArgs d, const ClipArgs& c1, const ClipArgs& c2, cudaStream_t stream);
// grad_reduce + RMSProp in one launch (single-GPU step; r.out must be a.g, r.out_floats the small-tensor prefix)
int launch_rmsprop_reduce(const RmsPropArgs& a, const GradReduceArgs& r, cudaStream_t stream);
// data-parallel RMSProp over peer memory: rank r owns arena slice r; it sums that slice of every rank's gradient
// arena (fixed rank order), applies RMSProp, and stores the new weights (+ bf16 shadow) into every rank's slab.
constexpr int DP_MAX_WORLD = 8;
struct RmsPropDpArgs {
  RmsPropArgs base;                 // this rank's own arenas
  uint8_t* peer[DP_MAX_WORLD];      // slab bases: [params | grads | ms | mom | shadow | comm]
  int rank, world;
  uint64_t step;                    // 1, 2, ... : value the ready / done flags reach in this step
  int64_t arena_bytes, comm_offset;
  GradReduceArgs red;               // has_red: first sum this rank's gradient-partial slabs into its gradient arena
  int has_red;                      // (saves the separate grad_reduce launch in front of the exchange)
};
int launch_rmsprop_dp(const RmsPropDpArgs& a, int num_sms, cudaStream_t stream);
// second instalment of the overlapped exchange (dp_exchange.cuh): slab reduction of the small tensors, LL push into every
// rank's receive buffer, identical RMSProp on every rank; returns only when every rank's dense1/w slice has landed.
// a.red is required; recv_offset: byte offset in the slab of the receive buffers [2][DP_MAX_WORLD][small prefix * 8 B].
// wait_big: block 0 also holds the launch open until every rank's dense1/w slice has landed (the overlapped modes)
int launch_dp_small(const RmsPropDpArgs& a, int64_t recv_offset, cudaStream_t stream, bool wait_big);
// the whole exchange in one launch at the end of the step (default): small tensors as launch_dp_small, dense1/w on every SM
int launch_dp_tail(const RmsPropDpArgs& a, const DpBigArgs& big, int64_t recv_offset, int num_sms, cudaStream_t stream);
// the dense1/w exchange alone, on a side stream, resident next to the conv kernels (elementwise.cu dp_big_side_kernel)
int launch_dp_big_side(const DpBigArgs& big, bool push_ready, int num_sms, cudaStream_t side_stream);
int launch_dp_push_zero(const WgradPush& push, long long n4, cudaStream_t stream);   // a rank without rows: zero slices to the owners
int configure_dp();     // load the exchange kernels now (not lazily at their first launch)
int launch_f32_to_bf16(const float* src, uint16_t* dst, int64_t n, cudaStream_t stream);
int launch_returns(const double* rewards, const int64_t* seg_offsets, int n_segments, const double* terminal,
                   double discount, int flags, double rmin, double rmax, double* out, cudaStream_t stream);
int launch_select_actions(const float* p, const double* u, int batch, int num_actions, int32_t* action,
                          cudaStream_t stream);

}  // namespace ga3c
