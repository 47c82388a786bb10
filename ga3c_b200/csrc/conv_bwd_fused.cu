// Backward of both convolutions in ONE persistent, warp-specialised kernel on the 5th-generation tensor cores
// (the autodiff half of opt.minimize, NetworkVP_discrate.py:130, for the layers at NetworkVP.py:212-228 wired as
// NetworkDNav.py:81-82):
//     dn1  = conv12 data-gradient of dn2, masked by relu'(n1)          (never leaves the SM)
//     g_w12 = patches(n1)^T dn2 , g_b12 = colsum(dn2)
//     g_w11 = patches(x)^T dn1  , g_b11 = colsum(dn1)                  (conv11 has no data-gradient: x is the input)
// bf16 operands, fp32 accumulation in TMEM.  Per frame the kernel reads x (112,896 B fp32), n1 (14,112 B) and dn2
// (7,744 B) from HBM exactly once and writes nothing; the weight-gradient accumulators stay in TMEM for all frames of
// the CTA and are stored once into the CTA's slab of the gradient-partial workspace (grad_reduce / rmsprop_reduce add
// the slabs in a fixed order).
//
// All three products are GEMMs over space-to-depth "block matrices" kept in the no-swizzle UMMA layout with every row
// contiguous (16-byte chunk j of row r at j*LBO + r*16), so that a spatial shift is a different descriptor start address:
//   G    dn2 on a zero-bordered 13x13 grid, row (oy+1)*13 + (ox+1), K = 32 co in 4 chunk planes
//   Blk2 n1, SAME-padded (1 before, 2 after) to 24x24 and cut into 12x12 blocks of 2x2 pixels, row Yb*13 + Xb (column 12
//        dead), 64 elements (dy, dx, ci) in 8 chunk planes
//   Blk  x, zero-padded to 88x88 and cut into 22x22 blocks of 4x4 pixels (conv_blk.cuh)
// conv12 data gradient.  Output pixel (2Yh+py, 2Xh+px) of the padded image gets taps kh = py + 2a, kw = px + 2b from
//   dn2[Yh - a, Xh - b].  With m = Yh*13 + Xh:   D[m, (py, px, ci)] = sum_{a,b} G[m + 14 - 13a - b, :] . Wd_ab
//   where Wd_ab[co, (py, px, ci)] = w12[py + 2a, px + 2b, ci, co]: four row-shifted GEMMs, M = 156 (two 128-row tiles, the
//   second one starting at row 28), N = 64, K = 32 each -> 16 UMMAs.  Row m of D is the whole 2x2 block (Yh, Xh) of dn1.
// conv12 weight gradient.  kh = 2a + dy, kw = 2b + dx: quadrant (a, b) is  dW_ab[(dy, dx, ci), co] = sum_m Blk2[m + 13a + b]^T
//   G[m + 14]  over positions m = oy*13 + ox (dead columns hit the zero border of G): A = Blk2 read MN-major, B = G read
//   MN-major, M = 64, N = 32, K = 144 -> 4 x 9 UMMAs.
// conv11 weight gradient.  Quadrant (a, b) is dW_ab = sum_m Blk[m + 22a + b]^T dn1[m] over positions m = oy*22 + ox.  A = Blk
//   read MN-major at row m' + 22a; the b-shift moves to the B side (m' = m + b): dn1 is written twice by the data-gradient
//   epilogue, planes (0, half) at row m and planes (1, half) at row m + 1, so ONE UMMA per (a, k-step) with N = 32 = (b, cout)
//   covers both b: 2 x 29 UMMAs (M = 64, N = 32) per frame instead of 4 x 29 -- the frame loop is bound by UMMA count and
//   operand bytes (every small UMMA costs ~40-80 cycles, microbenchmark in profiles/), not by HBM.
//
//   warps 0-5    fp32 chunks of x (TMA ring, one independent pipeline per warp) -> bf16 -> Blk
//   warp  6      one thread issues every UMMA: conv12 of frame k+1 between the conv11 position groups of frame k
//   warp  7      one thread streams n1 / dn2 of the next frame into a raw staging buffer (cp.async.bulk)
//   warps 8-11   data-gradient epilogue: TMEM -> relu' mask (n1 from Blk2) -> bf16 -> dn1 operand, bias gradient of conv11;
//                final store of the TMEM weight-gradient accumulators
//   warps 12-15  raw staging -> G / Blk2 layouts, bias gradient of conv12
#include "common.cuh"
#include "kernels.h"
#include "tcgen05.cuh"
#include "conv_blk.cuh"
#include "dp_exchange.cuh"

namespace ga3c {

constexpr int FB_THREADS = 512, FB_AUX_WARPS = 6, FB_ISSUE_WARP = 6, FB_TMA_WARP = 7,
              FB_EPI_WARP0 = 8, FB_RE_WARP0 = 12;
static_assert(FB_EPI_WARP0 % 4 == 0, "epilogue warp e must own TMEM lane quarter e");
constexpr int FBLK_ROWS = 492, FBLK_LBO = FBLK_ROWS * 16, FBLK_BYTES = 8 * FBLK_LBO;     // rows read: <= 16*28 + 23 + 15 = 486
constexpr int W11_KSTEPS = 29;                                   // 464 >= 462 positions (21 rows x 22, column 21 dead)
// dn1 operand: 4 planes (b, half) of 8 channels; planes (1, .) hold the SAME rows shifted down by one, so that one UMMA
// with N = 32 computes both column quadrants b = 0, 1 of the conv11 weight gradient (A = Blk is read once for two)
constexpr int DN1_ROWS = 16 * W11_KSTEPS, DN1_PLANE = DN1_ROWS * 16, DN1_BUF = 4 * DN1_PLANE;
constexpr int G_W = 13, G_ROWS = 176, G_LBO = G_ROWS * 16, G_BYTES = 4 * G_LBO;          // rows read: <= 28 + 127 + 14 = 169
constexpr int B2_ROWS = 160, B2_LBO = B2_ROWS * 16, B2_BYTES = 8 * B2_LBO;               // rows read: <= 143 + 14 = 157
constexpr int DG_TILE1 = 28;                                     // first row of the second data-gradient tile
constexpr int DG_ROWS = 12 * G_W;                                // 156 rows of D (12 x 12 blocks + dead column)
constexpr int W12_KSTEPS = 9;                                    // 144 >= 11 rows x 13 positions
constexpr int RAW_DN2 = N2_POS * C2_OUT * 2, RAW_N1 = N1_POS * C1_OUT * 2, RAW_BYTES = RAW_DN2 + RAW_N1;   // 7,744 + 14,112
constexpr int W12D_BYTES = 4 * 4 * 64 * 16;                      // 4 taps x [4 k-chunks (8 co)][64 rows (py, px, ci)][16 B]

constexpr int FB_OFF_BLK = 0;
constexpr int FB_OFF_RING = FB_OFF_BLK + FBLK_BYTES;             //  62,976
constexpr int FB_RING_BYTES = FB_AUX_WARPS * PW_SLOTS * PW_BYTES; //  64,512: two 4-row slots per aux warp
constexpr int FB_OFF_DN1 = FB_OFF_RING + FB_RING_BYTES;          // 127,488 (one buffer of four planes)
constexpr int FB_OFF_G = FB_OFF_DN1 + DN1_BUF;                   // 157,184
constexpr int FB_OFF_B2 = FB_OFF_G + G_BYTES;                    // 168,448
constexpr int FB_OFF_W12D = FB_OFF_B2 + B2_BYTES;                // 188,928
constexpr int FB_OFF_RAW = FB_OFF_W12D + W12D_BYTES;             // 205,312
constexpr int FB_OFF_RED = FB_OFF_RAW + RAW_BYTES;               // 227,168: [5][16] conv11 + [4][32] conv12 bias partials
constexpr int RED_B12 = 5 * C1_OUT;
constexpr int FB_OFF_BAR = FB_OFF_RED + (RED_B12 + 4 * C2_OUT) * 4;
constexpr int FB_RING = 0;        // [12] TMA chunk of x landed (slot = aux warp * 2 + parity)
constexpr int FB_BLKRDY = 12;     // [4] Blk rows of conv11 position group i converted, one arrival per chunk (aux -> issuer)
constexpr int FB_GRP = 16;        // [4] conv11 UMMAs of group i retired (tcgen05.commit)           (-> aux: Blk rows free)
constexpr int FB_DN1RDY = 20;     //     dn1 operand written, 5 arrivals                            (epilogue -> issuer)
constexpr int FB_DN1FREE = 22;    //     every conv11 UMMA of the frame retired                     (-> epilogue: operand free)
constexpr int FB_RAWFULL = 24;    //     n1 / dn2 of a frame landed in the raw buffer               (TMA -> re-layout)
constexpr int FB_RAWFREE = 25;    //     raw buffer consumed                                        (re-layout -> TMA thread)
constexpr int FB_C12RDY = 26;     //     G / Blk2 hold the frame                                    (re-layout -> issuer)
constexpr int FB_MMA12 = 27;      //     conv12 UMMAs of the frame retired (tcgen05.commit)         (-> epilogue)
constexpr int FB_EPI12 = 28;      //     D drained and Blk2 mask reads done, 5 arrivals             (epilogue -> issuer, re-layout)
constexpr int FB_DONE = 29;       //     every UMMA of the kernel retired                           (-> final store)
constexpr int FB_NBAR = 30;
constexpr int FB_OFF_TSLOT = FB_OFF_BAR + FB_NBAR * 8;
constexpr int FB_SMEM = FB_OFF_TSLOT + 16 + 128;                 // incl. slack to align the base to 128 B
static_assert(FB_SMEM <= 232448, "shared memory budget of one CTA per SM");
constexpr int FB_TMEM_COLS = 512, TM_W11 = 0, TM_W12 = 64, TM_D12 = 192;   // 2x32 | 4x32 | 2x64 columns

template <bool U8, bool DP>   // U8: frames are uint8 [B,28224] (x = k/128 - 1 applied on the fly), else fp32; DP: the grid carries exchange CTAs
__global__ void __launch_bounds__(FB_THREADS, 1)
conv_bwd_kernel(const void* __restrict__ x, const uint16_t* __restrict__ n1, const uint16_t* __restrict__ dn2,
                const float* __restrict__ w12, uint16_t* __restrict__ dn1_out, float* __restrict__ g_w11,
                float* __restrict__ g_b11, float* __restrict__ g_w12, float* __restrict__ g_b12, int64_t gp_stride, int batch,
                int n_conv, const DpBigArgs dp) {
  // data parallel: CTAs [n_conv, gridDim.x) are exchange CTAs -- they move dense1/w (final since dense_bwd, the launch
  // this one depends on) between the ranks while CTAs [0, n_conv) compute the conv gradients (dp_exchange.cuh)
  if (DP && (int)blockIdx.x >= n_conv) {
    griddep_launch();
    griddep_wait(K_DP_BIG);
    dp_big_exchange(dp, (int)blockIdx.x - n_conv, (int)gridDim.x - n_conv);
    trace_mark(K_DP_BIG, 2);
    return;
  }
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* const smem = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t blk = sbase + FB_OFF_BLK, ring = sbase + FB_OFF_RING, dn1s = sbase + FB_OFF_DN1, gg = sbase + FB_OFF_G,
                 b2 = sbase + FB_OFF_B2, w12d = sbase + FB_OFF_W12D, raw = sbase + FB_OFF_RAW, bars = sbase + FB_OFF_BAR,
                 tslot = sbase + FB_OFF_TSLOT;
  float* red = reinterpret_cast<float*>(smem + FB_OFF_RED);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  EvtLog evt_i = evt_open();                                      // pipeline event log of CTA 0 (ga3c_evt_*), off unless attached
  const int stride = DP ? (int)gridDim.x - dp.n_exch : (int)gridDim.x;
  const int n_frames = ((int)blockIdx.x < batch) ? (batch - 1 - (int)blockIdx.x) / stride + 1 : 0;
  const int n_chunks = n_frames * PW_NCHUNK;                       // chunk stream of this CTA: q = k * 21 + c, warp q % 6
  auto frame_of = [&](int k) { return (size_t)(blockIdx.x + k * stride); };
  auto bar = [&](int i) { return bars + i * 8; };
  auto issue_chunk = [&](int q, int slot) {                        // one thread
    const int k = q / PW_NCHUNK, c = q - k * PW_NCHUNK;
    constexpr uint32_t bytes = U8 ? PW_BYTES_U8 : PW_BYTES;
    mbar_expect_tx(bar(FB_RING + slot), bytes);
    bulk_load(ring + slot * PW_BYTES, static_cast<const uint8_t*>(x) + (frame_of(k) * PW_NCHUNK + c) * bytes, bytes, bar(FB_RING + slot));
  };

  // ---------------- prologue ----------------
  if (tid == 0) {
    for (int i = 0; i < FB_AUX_WARPS * PW_SLOTS; ++i) mbar_init(bar(FB_RING + i), 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar(FB_BLKRDY + i), pw_group_chunks(i));
      mbar_init(bar(FB_GRP + i), 1);
    }
    mbar_init(bar(FB_DN1RDY), 5);
    mbar_init(bar(FB_DN1FREE), 1);
    mbar_init(bar(FB_RAWFULL), 1);
    mbar_init(bar(FB_RAWFREE), 1);
    mbar_init(bar(FB_C12RDY), 1);
    mbar_init(bar(FB_MMA12), 1);
    mbar_init(bar(FB_EPI12), 5);
    mbar_init(bar(FB_DONE), 1);
    fence_mbar_init();
  }
  if (warp == FB_EPI_WARP0) tmem_alloc<FB_TMEM_COLS>(tslot);
  __syncthreads();
  if (warp < FB_AUX_WARPS && lane == 0)                            // x is an input of the step: stream it before the dependency wait
    for (int j = 0; j < PW_SLOTS; ++j)
      if (warp + j * FB_AUX_WARPS < n_chunks) issue_chunk(warp + j * FB_AUX_WARPS, warp * PW_SLOTS + j);
  // zero once: image borders / slack rows of Blk, dead rows of both dn1 buffers, borders of G and Blk2
  for (int i = tid; i < FBLK_BYTES / 16; i += FB_THREADS) sts128(blk + i * 16, make_uint4(0, 0, 0, 0));
  for (int i = tid; i < (FB_OFF_W12D - FB_OFF_DN1) / 16; i += FB_THREADS) sts128(dn1s + i * 16, make_uint4(0, 0, 0, 0));
  griddep_launch();
  evt_mark(evt_i, 50, 0);
  griddep_wait(K_CONV12_BWD);   // dn2 comes from the dense1 data-gradient GEMM that precedes this kernel
  if (DP && blockIdx.x == 0 && tid < dp.world) {
    // data parallel: dense_bwd of this rank is complete, i.e. its dense1/w gradient is final -- tell every rank now, a whole
    // conv backward before the exchange at the end of the step needs it (dp_tail_kernel)
    __threadfence_system();
    dp_st_flag(dp.peer[tid] + dp.comm_offset + DPC_BIGREADY + 64 * dp.rank, dp.step);
  }
  // data-gradient weights as the UMMA B operand (K-major, no swizzle): tap (a, b), k-chunk j (8 co), row n = (py, px, ci)
  for (int i = tid; i < 4 * 4 * 64; i += FB_THREADS) {
    const int tap = i >> 8, j = (i >> 6) & 3, n = i & 63, a = tap >> 1, b = tap & 1, py = n >> 5, px = (n >> 4) & 1, ci = n & 15;
    const float* w = w12 + (((py + 2 * a) * 4 + px + 2 * b) * C1_OUT + ci) * C2_OUT + 8 * j;
    sts128(w12d + tap * 4096 + j * 1024 + n * 16,
           make_uint4(pack_bf16(w[0], w[1]), pack_bf16(w[2], w[3]), pack_bf16(w[4], w[5]), pack_bf16(w[6], w[7])));
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(tmem_base) : "r"(tslot));
  evt_mark(evt_i, 51, 0);

  // Data-gradient epilogue of one 128-row tile for TMEM lane quarter `qt`: rows of the first tile go to warps 8-11, the 28
  // live rows of the second tile (lanes 100..127) to warp 15, which shares lane quarter 3 with warp 11 and is otherwise idle
  // once the frame's operands are laid out.
  auto dgrad_epilogue = [&](int k, int t, int qt, float (&bacc)[C1_OUT]) {
    const uint32_t tlane = tmem_base + ((uint32_t)(qt * 32) << 16);
    mbar_wait(bar(FB_MMA12), k & 1);
    evt_mark(evt_i, 40, k);
    tc_fence_after();
    uint16_t* dn1_dst = dn1_out ? dn1_out + frame_of(k) * (N1_POS * C1_OUT) : nullptr;
    // Phase 1: TMEM -> relu' mask -> packed bf16 in registers (one 2x2 block of dn1 per row).  Nothing is written yet: the
    // conv11 UMMAs of frame k-1 may still be reading the dn1 operand.
    uint32_t o[4][8];
    int p22[4];
    const int m = (t ? DG_TILE1 : 0) + 32 * qt + lane;
    const int Yh = m / G_W, Xh = m - Yh * G_W;
    const bool row_ok = (t == 0 || m >= 128) && m < DG_ROWS && Xh < 12;
#pragma unroll
    for (int cls = 0; cls < 4; ++cls) {
      uint32_t r[16];
      tc_ld16(tlane + TM_D12 + 64 * t + 16 * cls, r);
      const int y = 2 * Yh + (cls >> 1) - 1, xx = 2 * Xh + (cls & 1) - 1;
      p22[cls] = -1;
      if (row_ok && y >= 0 && y < H1 && xx >= 0 && xx < H1) {
        uint32_t mk[8];                                            // n1 of this pixel: relu'(n1) = (n1 > 0); post-ReLU values are >= 0
        lds128(*reinterpret_cast<uint32_t(*)[4]>(&mk[0]), b2 + (2 * cls) * B2_LBO + m * 16);
        lds128(*reinterpret_cast<uint32_t(*)[4]>(&mk[4]), b2 + (2 * cls + 1) * B2_LBO + m * 16);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const float lo = (mk[jj] & 0x7FFFu) ? __uint_as_float(r[2 * jj]) : 0.f;
          const float hi = (mk[jj] & 0x7FFF0000u) ? __uint_as_float(r[2 * jj + 1]) : 0.f;
          o[cls][jj] = pack_bf16(lo, hi);
          bacc[2 * jj] += bf16_lo(o[cls][jj]);
          bacc[2 * jj + 1] += bf16_hi(o[cls][jj]);
        }
        p22[cls] = y * BLK_W + xx;
        if (dn1_dst) {
          uint4* d = reinterpret_cast<uint4*>(dn1_dst + (y * H1 + xx) * C1_OUT);
          d[0] = make_uint4(o[cls][0], o[cls][1], o[cls][2], o[cls][3]);
          d[1] = make_uint4(o[cls][4], o[cls][5], o[cls][6], o[cls][7]);
        }
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar(FB_EPI12));                     // D is drained and Blk2 has been read: conv12 of the next frame may start
    // Phase 2: the operand buffer is free once every conv11 UMMA of frame k-1 has retired
    if (k > 0) mbar_wait(bar(FB_DN1FREE), (k - 1) & 1);
    evt_mark(evt_i, 41, k);
#pragma unroll
    for (int cls = 0; cls < 4; ++cls) {
      if (p22[cls] >= 0) {
        const uint4 lo4 = make_uint4(o[cls][0], o[cls][1], o[cls][2], o[cls][3]), hi4 = make_uint4(o[cls][4], o[cls][5], o[cls][6], o[cls][7]);
        const uint32_t dst = dn1s + p22[cls] * 16;
        sts128(dst, lo4);                                          // planes (b = 0, half): row p
        sts128(dst + DN1_PLANE, hi4);
        sts128(dst + 2 * DN1_PLANE + 16, lo4);                     // planes (b = 1, half): row p + 1
        sts128(dst + 3 * DN1_PLANE + 16, hi4);
      }
    }
    fence_proxy_async();                                           // the dn1 operand is read by the tensor core
    __syncwarp();
    if (lane == 0) mbar_arrive(bar(FB_DN1RDY));
    evt_mark(evt_i, 42, k);
  };
  auto b11_partial = [&](int slot, const float (&bacc)[C1_OUT]) {   // warp-reduce the 16 channel sums of this warp
#pragma unroll
    for (int c = 0; c < C1_OUT; ++c) {
      float v = bacc[c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[slot * C1_OUT + c] = v;
    }
  };

  if (warp < FB_AUX_WARPS) {
    // =========================== aux: one chunk pipeline per warp (conv_blk.cuh) ===========================
    uint32_t lane_off[3];
    blk_lane_offsets<FBLK_LBO>(lane, lane_off);
    int j = 0;
#pragma unroll 1
    for (int q = warp; q < n_chunks; q += FB_AUX_WARPS, ++j) {
      const int k = q / PW_NCHUNK, c = q - k * PW_NCHUNK, slot = warp * PW_SLOTS + (j & 1);
      uint32_t pk[PW_ROWS][3][2];
      evt_mark(evt_i, 1, q);
      mbar_wait(bar(FB_RING + slot), (j >> 1) & 1);                 // the chunk has landed
      evt_mark(evt_i, 3, q);
      blk_load_rows4<U8>(ring + slot * PW_BYTES, lane, pk);
      __syncwarp();                                                  // every lane has read its part: the slot is free
      if (lane == 0 && q + PW_SLOTS * FB_AUX_WARPS < n_chunks) issue_chunk(q + PW_SLOTS * FB_AUX_WARPS, slot);
      // block rows c, c+1 are rewritten: the last consumer group of frame k-1 that reads them must have retired
      evt_mark(evt_i, 5, q);
      if (k > 0) mbar_wait(bar(FB_GRP + pw_last_consumer(c)), (k - 1) & 1);
      evt_mark(evt_i, 2, q);
      blk_store_rows4<FBLK_LBO>(blk, c, lane, lane_off, pk);
      fence_proxy_async();                                           // Blk is read by the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(FB_BLKRDY + pw_first_consumer(c)));
      evt_mark(evt_i, 4, q);
    }
  } else if (warp == FB_ISSUE_WARP) {
    // =========================== MMA issuer ===========================
    // warp-uniform: all 32 lanes run the loops and wait on the barriers, elect_one() issues
    constexpr uint32_t idesc_dg = make_idesc_m(128, 64, false, false), idesc_w12 = make_idesc_m(64, C2_OUT, true, true),
                       idesc_w11 = make_idesc_m(64, 2 * C1_OUT, true, true);
    // descriptor low words of row 0 / k-chunk 0 of every operand; a shift of r rows is + r, a k-chunk plane is + LBO/16
    const uint32_t g_k = desc_ns_lo(gg, G_LBO), wd_k = desc_ns_lo(w12d, 1024);                 // K-major: LBO = plane, SBO = 128
    const uint32_t g_mn = desc_ns_lo(gg, 128), b2_mn = desc_ns_lo(b2, 128), blk_mn = desc_ns_lo(blk, 128),
                   dn1_mn = desc_ns_lo(dn1s, 128);                                             // MN-major: LBO = 128, SBO = plane
    constexpr uint32_t hi_k = desc_ns_hi(128), hi_g = desc_ns_hi(G_LBO), hi_b2 = desc_ns_hi(B2_LBO), hi_blk = desc_ns_hi(FBLK_LBO),
                       hi_dn1 = desc_ns_hi(DN1_PLANE);
    auto conv12_mmas = [&](int k) {
      evt_mark(evt_i, 10, k);
      mbar_wait(bar(FB_C12RDY), k & 1);                            // G / Blk2 hold frame k
      evt_mark(evt_i, 11, k);
      if (k > 0) mbar_wait(bar(FB_EPI12), (k - 1) & 1);            // D of frame k-1 has been drained
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
          for (int tap = 0; tap < 4; ++tap)
#pragma unroll
            for (int kk = 0; kk < 2; ++kk)
              tc_mma_bf16_w(tmem_base + TM_D12 + 64 * t,
                            g_k + ((t ? DG_TILE1 : 0) + G_W + 1 - G_W * (tap >> 1) - (tap & 1)) + 2 * kk * (G_LBO / 16), hi_k,
                            wd_k + (tap * 4096 + 2 * kk * 1024) / 16, hi_k, idesc_dg, (tap | kk) ? 1u : 0u);
        const uint32_t acc = k ? 1u : 0u;
#pragma unroll
        for (int s = 0; s < W12_KSTEPS; ++s)
#pragma unroll
          for (int q = 0; q < 4; ++q)
            // A: Blk2 rows 16 s + 13 a + b .., B: G rows 16 s + 14 ..; both MN-major
            tc_mma_bf16_w(tmem_base + TM_W12 + 32 * q, b2_mn + 16 * s + G_W * (q >> 1) + (q & 1), hi_b2, g_mn + 16 * s + G_W + 1, hi_g,
                          idesc_w12, s ? 1u : acc);
        tc_commit(bar(FB_MMA12));
      }
      __syncwarp();
      evt_mark(evt_i, 12, k);
    };
    if (n_frames > 0) conv12_mmas(0);
    for (int k = 0; k < n_frames; ++k) {
      evt_mark(evt_i, 13, k);
      mbar_wait(bar(FB_DN1RDY), k & 1);
      evt_mark(evt_i, 14, k);
#pragma unroll
      for (int gi = 0; gi < 4; ++gi) {
        if (gi == 1 && k + 1 < n_frames) conv12_mmas(k + 1);       // one frame ahead of the conv11 gradient
        mbar_wait(bar(FB_BLKRDY + gi), k & 1);
        evt_mark(evt_i, 15, k * 4 + gi);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t acc = k ? 1u : 0u;
#pragma unroll
          for (int s = 8 * gi; s < (gi == 3 ? W11_KSTEPS : 8 * gi + 8); ++s)
#pragma unroll
            for (int a = 0; a < 2; ++a)
              // A: Blk rows 16 s + 22 a .., B: dn1 rows 16 s .. in the four (b, half) planes; both MN-major
              tc_mma_bf16_w(tmem_base + TM_W11 + 32 * a, blk_mn + 16 * s + BLK_W * a, hi_blk, dn1_mn + 16 * s, hi_dn1, idesc_w11,
                            s ? 1u : acc);
          tc_commit(bar(FB_GRP + gi));
          if (gi == 3) tc_commit(bar(FB_DN1FREE));
        }
        __syncwarp();
        evt_mark(evt_i, 16, k * 4 + gi);
      }
    }
    if (elect_one()) tc_commit(bar(FB_DONE));
    __syncwarp();
  } else if (warp == FB_TMA_WARP) {
    // =========================== n1 / dn2 of the next frame -> raw staging buffer ===========================
    if (lane == 0) {
      for (int k = 0; k < n_frames; ++k) {
        if (k > 0) mbar_wait(bar(FB_RAWFREE), (k - 1) & 1);
        evt_mark(evt_i, 20, k);
        mbar_expect_tx(bar(FB_RAWFULL), RAW_BYTES);
        bulk_load(raw, dn2 + frame_of(k) * FLAT, RAW_DN2, bar(FB_RAWFULL));
        bulk_load(raw + RAW_DN2, n1 + frame_of(k) * (N1_POS * C1_OUT), RAW_N1, bar(FB_RAWFULL));
      }
    }
  } else if (warp >= FB_RE_WARP0) {
    // =========================== raw -> G / Blk2, conv12 bias gradient ===========================
    const int rtid = tid - 32 * FB_RE_WARP0;
    float bacc = 0.f;                                              // db12 partial: channel rtid & 31, position phase rtid >> 5
    float bacc11[C1_OUT];                                          // warp 15: db11 partials of the second data-gradient tile
#pragma unroll
    for (int c = 0; c < C1_OUT; ++c) bacc11[c] = 0.f;
    for (int k = 0; k < n_frames; ++k) {
      mbar_wait(bar(FB_RAWFULL), k & 1);
      evt_mark(evt_i, 30, k);
      if (k > 0) mbar_wait(bar(FB_EPI12), (k - 1) & 1);            // conv12 UMMAs and mask reads of frame k-1 are done
      evt_mark(evt_i, 31, k);
      for (int i = rtid; i < N2_POS * 4; i += 128) {               // dn2: 4 chunks of 8 co per position
        const int pos = i >> 2, j = i & 3, oy = pos / H2, ox = pos - oy * H2;
        uint32_t r[4];
        lds128(r, raw + i * 16);
        sts128(gg + j * G_LBO + ((oy + 1) * G_W + ox + 1) * 16, make_uint4(r[0], r[1], r[2], r[3]));
      }
      for (int i = rtid; i < N1_POS * 2; i += 128) {               // n1: 2 chunks of 8 ci per pixel
        const int p = i >> 1, h = i & 1, y = p / H1, Y = y + 1, X = p - y * H1 + 1;
        uint32_t r[4];
        lds128(r, raw + RAW_DN2 + i * 16);
        sts128(b2 + ((((Y & 1) * 2 + (X & 1)) * 2 + h) * B2_LBO) + ((Y >> 1) * G_W + (X >> 1)) * 16, make_uint4(r[0], r[1], r[2], r[3]));
      }
      fence_proxy_async();
      named_bar_sync(3, 128);
      if (rtid == 0) mbar_arrive(bar(FB_C12RDY));                  // the UMMAs can go; the bias sums below only read raw
      evt_mark(evt_i, 32, k);
      {
        const int co = rtid & 31;
        for (int pos = rtid >> 5; pos < N2_POS; pos += 4) {
          uint16_t v;
          asm volatile("ld.shared.u16 %0, [%1];\n" : "=h"(v) : "r"(raw + pos * 64 + co * 2));
          bacc += __uint_as_float((uint32_t)v << 16);
        }
      }
      named_bar_sync(3, 128);
      if (rtid == 0) mbar_arrive(bar(FB_RAWFREE));
      if (warp == FB_RE_WARP0 + 3) dgrad_epilogue(k, 1, 3, bacc11);
    }
    red[RED_B12 + rtid] = bacc;
    if (warp == FB_RE_WARP0 + 3) b11_partial(4, bacc11);
    named_bar_sync(3, 128);
    if (rtid < C2_OUT)
      g_b12[(int64_t)blockIdx.x * gp_stride + rtid] =
          red[RED_B12 + rtid] + red[RED_B12 + 32 + rtid] + red[RED_B12 + 64 + rtid] + red[RED_B12 + 96 + rtid];
  } else {
    // =========================== data-gradient epilogue; final store ===========================
    const int ew = warp - FB_EPI_WARP0;                            // TMEM lane quarter
    const uint32_t tlane = tmem_base + ((uint32_t)(ew * 32) << 16);
    float bacc[C1_OUT];                                            // db11 partials of this thread's pixels
#pragma unroll
    for (int c = 0; c < C1_OUT; ++c) bacc[c] = 0.f;
    for (int k = 0; k < n_frames; ++k) dgrad_epilogue(k, 0, ew, bacc);
    b11_partial(ew, bacc);                                         // summed with warp 15's after the final block barrier
    // weight-gradient accumulators: M = 64 rows sit on TMEM lanes 32 w + (0..15) (rows 16 w ..); row = chunk plane * 8 + e
    if (n_frames > 0) {
      mbar_wait(bar(FB_DONE), 0);
      tc_fence_after();
      float* const s11 = g_w11 + (int64_t)blockIdx.x * gp_stride;
      float* const s12 = g_w12 + (int64_t)blockIdx.x * gp_stride;
      const int row = 16 * ew + lane, j = row >> 3, e = row & 7;
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {                                // q = a*2 + b: accumulator a, columns 16 b ..
        uint32_t r[16];
        tc_ld16(tlane + TM_W11 + 32 * (q >> 1) + 16 * (q & 1), r);
        if (lane < 16) {
          // conv11 block element: j = dy*2 + (dx>>1), e = (dx&1)*4 + c
          const int kh = 4 * (q >> 1) + (j >> 1), kw = 4 * (q & 1) + (j & 1) * 2 + (e >> 2), c = e & 3;
          float4* o = reinterpret_cast<float4*>(s11 + ((kh * 8 + kw) * 4 + c) * C1_OUT);
#pragma unroll
          for (int n = 0; n < C1_OUT / 4; ++n)
            o[n] = make_float4(__uint_as_float(r[4 * n]), __uint_as_float(r[4 * n + 1]), __uint_as_float(r[4 * n + 2]),
                               __uint_as_float(r[4 * n + 3]));
        }
      }
#pragma unroll 1
      for (int q = 0; q < 4; ++q)
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t r[16];
          tc_ld16(tlane + TM_W12 + 32 * q + 16 * hh, r);
          if (lane < 16) {
            // conv12 block element: j = (dy*2 + dx)*2 + (ci>>3), e = ci & 7
            const int kh = 2 * (q >> 1) + (j >> 2), kw = 2 * (q & 1) + ((j >> 1) & 1), ci = (j & 1) * 8 + e;
            float4* o = reinterpret_cast<float4*>(s12 + ((kh * 4 + kw) * C1_OUT + ci) * C2_OUT + 16 * hh);
#pragma unroll
            for (int n = 0; n < 4; ++n)
              o[n] = make_float4(__uint_as_float(r[4 * n]), __uint_as_float(r[4 * n + 1]), __uint_as_float(r[4 * n + 2]),
                                 __uint_as_float(r[4 * n + 3]));
          }
        }
    }
  }

  evt_mark(evt_i, 52, 0);
  tc_fence_before();
  __syncthreads();
  if (tid < C1_OUT)                                                // conv11 bias gradient: the five warp partials in a fixed order
    g_b11[(int64_t)blockIdx.x * gp_stride + tid] =
        red[tid] + red[C1_OUT + tid] + red[2 * C1_OUT + tid] + red[3 * C1_OUT + tid] + red[4 * C1_OUT + tid];
  trace_mark(K_CONV12_BWD, 2);
  if (warp == FB_EPI_WARP0) {
    tc_fence_after();
    tmem_dealloc<FB_TMEM_COLS>(tmem_base);
  }
}

GA3C_TRACE_ATTACH(trace_attach_conv_bwd_fused)
GA3C_EVT_ATTACH(evt_attach_conv_bwd)

int conv_bwd_grid(int batch, int num_sms, int n_exch) { return min(batch, num_sms - n_exch); }

int configure_conv_bwd_fused() {
  cudaError_t e = cudaFuncSetAttribute(conv_bwd_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(conv_bwd_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(conv_bwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaFuncSetAttribute(conv_bwd_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM);
}

int launch_conv_bwd(const void* x, bool x_u8, const uint16_t* n1, const uint16_t* dn2, const float* w12, uint16_t* dn1_out,
                    float* g_w11, float* g_b11, float* g_w12, float* g_b12, int64_t gp_stride, int batch, int num_sms,
                    const DpBigArgs* dp, cudaStream_t stream) {
  DpBigArgs d{};
  if (dp != nullptr) d = *dp;
  const int n_conv = conv_bwd_grid(batch, num_sms, d.n_exch);
  const dim3 grid(n_conv + d.n_exch);
  auto kernel = dp != nullptr ? (x_u8 ? conv_bwd_kernel<true, true> : conv_bwd_kernel<false, true>)
                              : (x_u8 ? conv_bwd_kernel<true, false> : conv_bwd_kernel<false, false>);
  return launch_pdl(kernel, grid, dim3(FB_THREADS), FB_SMEM, stream, x, n1, dn2, w12, dn1_out, g_w11, g_b11, g_w12, g_b12,
                    gp_stride, batch, n_conv, d);
}

}  // namespace ga3c
