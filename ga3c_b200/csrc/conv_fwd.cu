// Fused conv11 (8x8 s4, 16) -> conv12 (4x4 s2, 32) forward: one persistent, warp-specialised CTA per SM,
// BOTH convolutions on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
//
// Reference op: tf.nn.conv2d(..., padding='SAME') + b, relu  (NetworkVP.py:224-226), wired as
// NetworkDNav.py:81-82.  bf16 operands, fp32 accumulate.
//
// conv11 as 4 shifted GEMMs (space-to-depth).  The zero-padded 88x88x4 image is cut into 22x22 blocks of 4x4 pixels;
//   a block is one row (K = 64 = (dy, dx, c)) of the "block matrix" Blk[484, 64].  An 8x8 stride-4 window is exactly 2x2
//   blocks, so   out[oy, ox, :] = sum_{a,b in {0,1}} Blk[(oy+a)*22 + (ox+b), :] . W_ab        (W_ab = w11[4a.., 4b.., :, :])
//   With output rows indexed m = oy*22 + ox (column 21 is a dead column) quadrant (a,b) reads Blk rows m + 22a + b: the same
//   matrix at a ROW-SHIFTED start address.  Blk is kept in the no-swizzle K-major UMMA layout with all rows contiguous
//   (16-byte k-chunk j of row r at j*LBO + r*16), so any row shift is a legal descriptor start address and nothing is
//   duplicated.  The a-shift (+22 rows) is done by the tensor core (two accumulating UMMAs per k-step); the b-shift (+1 row)
//   is folded into N: D[m, b*16 + c] = sum_a Blk[m + 22a] . W_ab[:, c]  (N = 32), and the epilogue adds
//   out[m] = D[m, 0:16] + D[m+1, 16:32] with one warp shuffle.  M tiles start every 127 rows (row 127 of a tile only feeds
//   row 126): 4 tiles x 2 shifts x 4 k16 steps = 32 UMMAs (M=128, N=32) per frame -- the UMMA time is set by the A-operand
//   shared-memory reads, so halving the instruction count halves it.
// conv12 as one im2col GEMM.  M = 121 (one 128-row tile), N = 32, K = 256 = (kh, kw, ci): the conv11 epilogue scatters every
//   output pixel into the (up to four) im2col slots it feeds, directly in the K-major SWIZZLE_128B operand layout:
//   16 UMMAs per frame.
// The mma.sync predecessors of this kernel (experiments/) were bound by instruction issue: ~17k warp instructions per
// frame; here the SM only converts fp32 -> bf16 and runs the two epilogues.
//
// Streaming pipeline per frame (112,896 B of fp32 input read from HBM exactly once; only n2 [+ n1 when training] leave):
//   TMA engine        cp.async.bulk of 4-row chunks (5,376 B) of the fp32 frame, two ring slots per aux warp (mbarrier per slot)
//   warps 0-5  (aux)  one independent pipeline per warp: chunk -> bf16 -> Blk, re-arm the slot with the warp's chunk after next;
//                     they start BEFORE the dependency wait (frames are a step input), the other warps wait and lay out the weights
//   warp  6           one thread issues the UMMAs: conv11 tile i as soon as its Blk rows are converted, conv12 after the
//                     conv11 epilogues; tcgen05.commit signals TMEM-full / operand-free mbarriers
//   warp  7           training only: copies every finished quarter of Blk (128 block rows, 8 planes) to HBM with cp.async.bulk --
//                     the bf16 block matrix the conv backward reads back instead of the fp32 frame (common.cuh: xblk)
//   warps 8-15        epilogues, two sets of four warps (one TMEM lane quarter each): conv11 tile -> +bias, ReLU -> im2col
//                     scatter (+ n1 to HBM in the Blk2 operand layout when training); conv12 tile -> +bias, ReLU -> n2 to HBM
#include "common.cuh"
#include "kernels.h"
#include "tcgen05.cuh"
#include "conv_blk.cuh"

namespace ga3c {

constexpr int CF_THREADS = 512, CF_AUX_WARPS = 6, CF_ISSUE_WARP = 6, CF_STORE_WARP = 7, CF_EPI_WARP0 = 8;     // two epilogue sets: warps 8-11, 12-15
static_assert(CF_EPI_WARP0 % 4 == 0, "epilogue warp e must own TMEM lane quarter e");
constexpr int C11_TILES = 4;                     // 462 output rows (21 x 22, column 21 dead) in 4 x 128
constexpr int C11_TSTRIDE = 127;                                     // output rows per tile (tile row 127 only feeds row 126)
static_assert((C11_TILES - 1) * C11_TSTRIDE + 128 + BLK_W <= BLK_ROWS && C11_TILES * C11_TSTRIDE >= H1 * BLK_W, "tiles must cover the outputs and stay inside the buffer");

constexpr int CF_OFF_A2 = 0;                                         // conv12 A: 4 k-blocks (kh) x [128 rows x 128 B], SW128
constexpr int CF_OFF_B2 = CF_OFF_A2 + 4 * 128 * 128;                 //  65,536: conv12 B: 4 k-blocks x [32 rows x 128 B], SW128
constexpr int CF_OFF_BLK = CF_OFF_B2 + 4 * 32 * 128;                 //  81,920
constexpr int CF_OFF_RING = CF_OFF_BLK + BLK_BYTES;                  // 152,064
constexpr int CF_OFF_WQ = CF_OFF_RING + CF_AUX_WARPS * PW_SLOTS * PW_BYTES;   // 216,576 (ring: two 4-row slots per aux warp): conv11 B: 2 row shifts a x [8 k-chunks][32 rows (b, cout) x 16 B]
constexpr int CF_OFF_LUT = CF_OFF_WQ + 4 * 2048;                     // 224,768
constexpr int CF_OFF_BIAS = CF_OFF_LUT + 3584;                       // 228,352 (441 x 8 B rounded up)
constexpr int CF_OFF_BAR = CF_OFF_BIAS + (C1_OUT + C2_OUT) * 4;      // 228,544
// mbarriers (8 B each)
constexpr int BAR_RING = 0;          // [12] TMA chunk landed (slot = aux warp * 2 + parity)
constexpr int BAR_BLKRDY = 12;       // [4]  Blk rows of conv11 tile i converted, one arrival per chunk (aux -> issuer)
constexpr int BAR_C11 = 16;          // [4]  conv11 tile i accumulated (tcgen05.commit)     (-> epilogue; -> aux: its Blk rows are free)
constexpr int BAR_T1FREE = 20;       // [4]  conv11 TMEM tile i drained, 4 arrivals         (epilogue -> issuer)
constexpr int BAR_A2RDY = 24;        //      im2col operand of the frame complete, 4 arrivals (epilogue -> issuer)
constexpr int BAR_MMA2 = 25;         //      conv12 accumulated (tcgen05.commit)            (-> epilogue; im2col operand free)
constexpr int BAR_T2FREE = 26;       //      conv12 TMEM tile drained, 4 arrivals           (epilogue -> issuer)
constexpr int BAR_STORED = 27;       // [4]  quarter q of Blk has been read by the bulk stores to HBM (store warp -> aux: rows free)
constexpr int CF_NBAR = 31;
constexpr int CF_OFF_TSLOT = CF_OFF_BAR + CF_NBAR * 8;               // 228,696
constexpr int CF_OFF_XCH = CF_OFF_TSLOT + 16;                        // [2 sets][2][4 warps][16 floats] b-shift exchange across warps
constexpr int CF_SMEM = CF_OFF_XCH + 1024 + 1024;                    // incl. slack to align the base to 1024 B
static_assert(CF_SMEM <= 232448, "shared memory budget of one CTA per SM");
constexpr int CF_TMEM_COLS = 256, TMEM_C12 = 128;                    // conv11 tiles at columns 0,32,64,96; conv12 at 128..159

// U8: frames are uint8 [B,28224] (x = k/128 - 1 applied on the fly), else fp32.  EVT: the instantiation with the pipeline event log
// of CTA 0 (ga3c_evt_*, tools/evt_conv_fwd_prologue.py); the production instantiations carry none of it.
template <bool U8, bool EVT>
__global__ void __launch_bounds__(CF_THREADS, 1)
conv_fwd_kernel(const void* __restrict__ x, const float* __restrict__ w11, const float* __restrict__ b11,
                const float* __restrict__ w12, const float* __restrict__ b12,
                uint8_t* __restrict__ n1_out, uint8_t* __restrict__ xblk_out, uint16_t* __restrict__ n2_out, int batch, int hints) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const smem = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t sa2 = sbase + CF_OFF_A2, sb2 = sbase + CF_OFF_B2, blk = sbase + CF_OFF_BLK, ring = sbase + CF_OFF_RING,
                 wq = sbase + CF_OFF_WQ, lut = sbase + CF_OFF_LUT, bars = sbase + CF_OFF_BAR, tslot = sbase + CF_OFF_TSLOT;
  float* bias_s = reinterpret_cast<float*>(smem + CF_OFF_BIAS);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  EvtLog evt_i = EVT ? evt_open() : EvtLog{nullptr, 0};          // pipeline event log of CTA 0 (ga3c_evt_*): EVT instantiation only
  auto mark = [&](int id, int arg) { if (EVT) evt_mark(evt_i, id, arg); };
  const int stride = gridDim.x;
  const int n_frames = ((int)blockIdx.x < batch) ? (batch - 1 - (int)blockIdx.x) / stride + 1 : 0;
  const int n_chunks = n_frames * PW_NCHUNK;                         // chunk stream of this CTA: q = k * 21 + c, warp q % 6
  auto frame_of = [&](int k) { return (size_t)(blockIdx.x + k * stride); };
  auto bar = [&](int i) { return bars + i * 8; };
  // L2 residency (common.cuh): frames are streamed once, the activations written here are read back later in the step
  const uint64_t pol_stream = (hints & 1) ? l2_policy_evict_first() : l2_policy_normal();
  const uint64_t pol_keep = (hints & 2) ? l2_policy_evict_last() : l2_policy_normal();
  auto issue_chunk = [&](int q, int slot) {                          // one thread
    const int k = q / PW_NCHUNK, c = q - k * PW_NCHUNK;
    constexpr uint32_t bytes = U8 ? PW_BYTES_U8 : PW_BYTES;
    mbar_expect_tx(bar(BAR_RING + slot), bytes);
    bulk_load_hint(ring + slot * PW_BYTES, static_cast<const uint8_t*>(x) + (frame_of(k) * PW_NCHUNK + c) * bytes, bytes,
                   bar(BAR_RING + slot), pol_stream);
  };

  // ---------------- prologue ----------------
  mark(60, 0);
  if (tid == 0) {
    for (int i = 0; i < CF_AUX_WARPS * PW_SLOTS; ++i) mbar_init(bar(BAR_RING + i), 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar(BAR_BLKRDY + i), pw_group_chunks(i));
      mbar_init(bar(BAR_C11 + i), 1);
      mbar_init(bar(BAR_T1FREE + i), 4);
    }
    mbar_init(bar(BAR_A2RDY), 8);
    mbar_init(bar(BAR_MMA2), 1);
    mbar_init(bar(BAR_T2FREE), 4);
    for (int i = 0; i < 4; ++i) mbar_init(bar(BAR_STORED + i), 1);
    fence_mbar_init();
  }
  if (warp == CF_EPI_WARP0) tmem_alloc<CF_TMEM_COLS>(tslot);
  __syncthreads();
  mark(61, 0);
  if (warp < CF_AUX_WARPS && lane == 0)                              // x is an input of the step: stream it before the dependency wait
    for (int j = 0; j < PW_SLOTS; ++j)
      if (warp + j * CF_AUX_WARPS < n_chunks) issue_chunk(warp + j * CF_AUX_WARPS, warp * PW_SLOTS + j);
  // zero the operands once: image borders / slack rows of Blk, padding taps and rows 121..127 of the im2col operand
  for (int i = tid; i < (CF_OFF_RING - CF_OFF_A2) / 16; i += CF_THREADS)
    if (i < CF_OFF_B2 / 16 || i >= CF_OFF_BLK / 16) sts128(sbase + i * 16, make_uint4(0, 0, 0, 0));
  // scatter table: conv11 output pixel p = (y, x) feeds conv12 position (oy, ox) through tap (kh, kw) when
  // y + 1 = 2 oy + kh and x + 1 = 2 ox + kw (SAME padding 1 before / 2 after).  Up to 4 (kh, kw) per pixel:
  // kh = ((y+1)&1) + 2 ia, kw = ((x+1)&1) + 2 ib.  Entry = offset, in 16-byte units, of the pixel's channels 0..7 in the
  // SW128 operand: row * 8 + ((2 kw) ^ (row & 7)) with row = kh * 128 + oy * 11 + ox; channels 8..15 sit at entry ^ 1.
  // 0xFFFF = no such tap.
  for (int p = tid; p < N1_POS; p += CF_THREADS) {
    const int y = p / H1, xx = p - y * H1;
    uint32_t e[4];
#pragma unroll
    for (int idx = 0; idx < 4; ++idx) {
      const int kh = ((y + 1) & 1) + 2 * (idx >> 1), kw = ((xx + 1) & 1) + 2 * (idx & 1);
      const int oy2 = y + 1 - kh, ox2 = xx + 1 - kw;
      const bool ok = oy2 >= 0 && ox2 >= 0 && (oy2 >> 1) < H2 && (ox2 >> 1) < H2;
      const int row = kh * 128 + (oy2 >> 1) * H2 + (ox2 >> 1);
      e[idx] = ok ? (uint32_t)(row * 8 + ((2 * kw) ^ (row & 7))) : 0xFFFFu;
    }
    sts64(lut + p * 8, e[0] | (e[1] << 16), e[2] | (e[3] << 16));
  }
  fence_proxy_async();          // the zeroed operand regions are read by the tensor core (async proxy)
  __syncthreads();              // Blk is zeroed before the aux warps write frame 0 into it; the scatter table is complete
  mark(62, 0);
  griddep_launch();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(tmem_base) : "r"(tslot));
  if (warp >= CF_AUX_WARPS) {
    // Warps 6-15 wait for the preceding kernels and lay the weights out as UMMA operands.  The aux warps do NOT wait: they touch
    // the frames (a step input, streamed since the top of the kernel) and shared memory only, so frame 0 is converted while the
    // CTA sits in front of the dependency -- and while the weights are laid out, 2.5 us that used to precede the first chunk.
    trace_mark_by(K_CONV_FWD, 0, tid == CF_AUX_WARPS * 32);
    asm volatile("griddepcontrol.wait;\n" ::: "memory");       // the weights below come from the optimizer kernel that precedes this one in the stream
    trace_mark_by(K_CONV_FWD, 1, tid == CF_AUX_WARPS * 32);
    mark(63, 0);
    const int t = tid - CF_AUX_WARPS * 32;                        // 0 .. 319; every load of a thread is in flight before its first store
    float4 q12[8];
    float2 q11[8];
    const bool do12 = t < 256, do11 = t >= 64;
    // conv12 weights as the UMMA B operand: k-block kh, row n (cout), 128 B = (kw, ci) K-major, 16-B chunks XOR (n & 7).
    // Item = (kh, chunk c = (kw, ci half), four couts): 8 float4 (one per ci) -> 4 chunks
    const int kh = t >> 6, c12 = (t >> 3) & 7, n12 = (t & 7) * 4;
    if (do12) {
      const float* w = w12 + ((kh * 4 + (c12 >> 1)) * C1_OUT + (c12 & 1) * 8) * C2_OUT + n12;
#pragma unroll
      for (int e = 0; e < 8; ++e) q12[e] = *reinterpret_cast<const float4*>(w + e * C2_OUT);
    }
    // conv11 weights for row shift a: B operand [32 rows n2 = b*16 + cout][K = 64], no-swizzle K-major: k-chunk j = dy*2 + (dx>>1)
    // holds (dx&1, c) -> 8 elements; chunk j of row n2 at a*4096 + j*512 + n2*16.  Item = (a, j, b, two couts): 8 float2 -> 2 chunks
    const int i11 = t - 64, a11 = i11 >> 7, j11 = (i11 >> 4) & 7, b11i = (i11 >> 3) & 1, n11 = (i11 & 7) * 2;
    if (do11) {
      const float* w = w11 + (((4 * a11 + (j11 >> 1)) * 8 + 4 * b11i + (j11 & 1) * 2) * 4) * C1_OUT + n11;
#pragma unroll
      for (int e = 0; e < 8; ++e) q11[e] = *reinterpret_cast<const float2*>(w + e * C1_OUT);
    }
    if (t < C1_OUT) bias_s[t] = b11[t];
    if (t >= 32 && t < 32 + C2_OUT) bias_s[C1_OUT + t - 32] = b12[t - 32];
    if (do12) {
#pragma unroll
      for (int qn = 0; qn < 4; ++qn) {
        const int n = n12 + qn;
        auto f = [&](const float4& v) { return qn == 0 ? v.x : qn == 1 ? v.y : qn == 2 ? v.z : v.w; };
        sts128(sb2 + kh * 4096 + n * 128 + ((c12 ^ (n & 7)) << 4),
               make_uint4(pack_bf16(f(q12[0]), f(q12[1])), pack_bf16(f(q12[2]), f(q12[3])), pack_bf16(f(q12[4]), f(q12[5])),
                          pack_bf16(f(q12[6]), f(q12[7]))));
      }
    }
    if (do11) {
      const uint32_t dst = wq + a11 * 4096 + j11 * 512 + (b11i * 16 + n11) * 16;
      sts128(dst, make_uint4(pack_bf16(q11[0].x, q11[1].x), pack_bf16(q11[2].x, q11[3].x), pack_bf16(q11[4].x, q11[5].x),
                             pack_bf16(q11[6].x, q11[7].x)));
      sts128(dst + 16, make_uint4(pack_bf16(q11[0].y, q11[1].y), pack_bf16(q11[2].y, q11[3].y), pack_bf16(q11[4].y, q11[5].y),
                                  pack_bf16(q11[6].y, q11[7].y)));
    }
    fence_proxy_async();          // operands written with generic-proxy stores are read by the tensor core (async proxy)
    tc_fence_before();
    named_bar_sync(4, CF_THREADS - CF_AUX_WARPS * 32);
    tc_fence_after();
    mark(64, 0);
  }

  if (warp < CF_AUX_WARPS) {
    // =========================== aux: one chunk pipeline per warp (conv_blk.cuh) ===========================
    uint32_t lane_off[3];
    blk_lane_offsets(lane, lane_off);
    int j = 0;
#pragma unroll 1
    for (int q = warp; q < n_chunks; q += CF_AUX_WARPS, ++j) {
      const int k = q / PW_NCHUNK, c = q - k * PW_NCHUNK, slot = warp * PW_SLOTS + (j & 1);
      uint32_t pk[PW_ROWS][3][2];
      mbar_wait(bar(BAR_RING + slot), (j >> 1) & 1);                 // the chunk has landed
      blk_load_rows4<U8>(ring + slot * PW_BYTES, lane, pk);
      __syncwarp();                                                  // every lane has read its part: the slot is free
      if (lane == 0 && q + PW_SLOTS * CF_AUX_WARPS < n_chunks) issue_chunk(q + PW_SLOTS * CF_AUX_WARPS, slot);
      // block rows c, c+1 are rewritten: the last consumer group of frame k-1 that reads them must have retired
      if (k > 0) {
        mbar_wait(bar(BAR_C11 + pw_last_consumer(c)), (k - 1) & 1);
        // ... and the copy of these rows to HBM must have read them (block rows c, c+1 end in quarter pw_last_consumer(c))
        if (xblk_out != nullptr) mbar_wait(bar(BAR_STORED + pw_last_consumer(c)), (k - 1) & 1);
      }
      blk_store_rows4<BLK_LBO>(blk, c, lane, lane_off, pk);
      fence_proxy_async();                                           // Blk is read by the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(BAR_BLKRDY + pw_first_consumer(c)));
    }
  } else if (warp == CF_ISSUE_WARP) {
    // =========================== MMA issuer ===========================
    // warp-uniform: all 32 lanes run the loops and wait on the barriers, elect_one() issues (tcgen05.cuh)
    constexpr uint32_t idesc11 = make_idesc(2 * C1_OUT, false, false), idesc12 = make_idesc(C2_OUT, false, false);
    const uint32_t blk_k = desc_ns_lo(blk, BLK_LBO), wq_k = desc_ns_lo(wq, 512);     // K-major, no swizzle: LBO = k-chunk plane, SBO = 128
    constexpr uint32_t hi_k = desc_ns_hi(128);
    for (int k = 0; k < n_frames; ++k) {
#pragma unroll
      for (int i = 0; i < C11_TILES; ++i) {
        mbar_wait(bar(BAR_BLKRDY + i), k & 1);                       // the Blk rows this tile reads hold frame k
        mark(15, k * 4 + i);
        if (k > 0) mbar_wait(bar(BAR_T1FREE + i), (k - 1) & 1);      // its accumulator of frame k-1 has been drained
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              tc_mma_bf16_w(tmem_base + 32 * i, blk_k + C11_TSTRIDE * i + BLK_W * a + 2 * kk * (BLK_LBO / 16), hi_k,
                            wq_k + (a * 4096 + 2 * kk * 512) / 16, hi_k, idesc11, (a | kk) ? 1u : 0u);
          tc_commit(bar(BAR_C11 + i));
        }
        __syncwarp();
        mark(16, k * 4 + i);
      }
      mbar_wait(bar(BAR_A2RDY), k & 1);                              // every conv11 output of frame k sits in the im2col operand
      mark(17, k);
      if (k > 0) mbar_wait(bar(BAR_T2FREE), (k - 1) & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          const uint64_t da = make_desc(sa2 + kb * 16384, false), db = make_desc(sb2 + kb * 4096, false);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            tc_mma_bf16(tmem_base + TMEM_C12, da + (uint64_t)(kk * 32 >> 4), db + (uint64_t)(kk * 32 >> 4), idesc12, (kb | kk) ? 1u : 0u);
        }
        tc_commit(bar(BAR_MMA2));
      }
      __syncwarp();
      mark(18, k);
    }
  } else if (warp == CF_STORE_WARP) {
    // =========================== Blk -> HBM (training) ===========================
    // Quarter q (block rows 128 q ..) is complete when conv11 tile q's rows are (BAR_BLKRDY + q: block rows up to 6 / 12 / 18 /
    // 21 cover rows 153 / 285 / 417 / 483).  Lane j copies chunk plane j: 2 KB contiguous in shared memory and in HBM
    // ([frame][quarter][plane][128 rows][16 B]); the last quarter holds 100 live rows, the rest of it stays zero in HBM.
    if (xblk_out != nullptr) {
      for (int k = 0; k < n_frames; ++k) {
        uint8_t* dst = xblk_out + frame_of(k) * XB_FRAME_BYTES;
#pragma unroll 1
        for (int q = 0; q < XB_QUARTERS; ++q) {
          mbar_wait(bar(BAR_BLKRDY + q), k & 1);
          if (lane < 8) {
            const uint32_t bytes = q < 3 ? XB_PLANE_BYTES : (XB_LIVE_ROWS - 3 * XB_QROWS) * 16;
            bulk_store_hint(dst + q * XB_QBYTES + lane * XB_PLANE_BYTES, blk + lane * BLK_LBO + q * XB_PLANE_BYTES, bytes, pol_keep);
            bulk_commit();
            bulk_wait_read0();                                     // shared memory has been read: the rows may be overwritten
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(BAR_STORED + q));
        }
      }
      if (lane < 8) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");   // the writes are complete before the kernel ends
    }
  } else if (warp >= CF_EPI_WARP0) {
    // =========================== epilogues ===========================
    // Two sets of four warps (one TMEM lane quarter each): set 0 drains conv11 tiles 0 and 2, set 1 the conv12 tile of the
    // previous frame and conv11 tiles 1 and 3.  One set alone is the longest serial chain of the frame loop.
    const int ew = warp & 3, eset = (warp - CF_EPI_WARP0) >> 2;      // TMEM lane quarter, epilogue set
    const uint32_t tlane = tmem_base + ((uint32_t)(ew * 32) << 16);
    auto conv12_epilogue = [&](int k) {                              // TMEM -> +bias, ReLU, bf16 -> n2[frame k]
      mbar_wait(bar(BAR_MMA2), k & 1);
      mark(43, k);
      tc_fence_after();
      uint32_t r[32];
      tc_ld32(tlane + TMEM_C12, r);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(BAR_T2FREE));
      const int pos = ew * 32 + lane;
      if (pos < N2_POS) {
        uint4* dst = reinterpret_cast<uint4*>(n2_out + frame_of(k) * FLAT + pos * C2_OUT);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint32_t o[4];
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int n = 8 * i + 2 * jj;
            o[jj] = pack_bf16(fmaxf(__uint_as_float(r[n]) + bias_s[C1_OUT + n], 0.f),
                              fmaxf(__uint_as_float(r[n + 1]) + bias_s[C1_OUT + n + 1], 0.f));
          }
          stg128_hint(dst + i, make_uint4(o[0], o[1], o[2], o[3]), pol_keep);
        }
      }
    };
    for (int k = 0; k < n_frames; ++k) {
      if (k > 0) {                                                   // the im2col operand of frame k-1 is no longer read
        if (eset == 1) conv12_epilogue(k - 1);
        else mbar_wait(bar(BAR_MMA2), (k - 1) & 1);
      }
      mark(44, k);
      uint8_t* n1_dst = n1_out ? n1_out + frame_of(k) * B2_BYTES : nullptr;      // Blk2 operand layout (common.cuh)
#pragma unroll 1
      for (int i = eset; i < C11_TILES; i += 2) {
        mbar_wait(bar(BAR_C11 + i), k & 1);
        mark(40, k * 4 + i);
        tc_fence_after();
        uint32_t r[32];                                              // [0,16): b = 0 part of row m ; [16,32): b = 1 part, owed to row m-1
        tc_ld32(tlane + 32 * i, r);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(BAR_T1FREE + i));
        // out[m] = D[m, 0:16] + D[m+1, 16:32]: the neighbour row is the next lane; lane 31 takes it from the next warp's lane 0
        float* xch = reinterpret_cast<float*>(smem + CF_OFF_XCH) + eset * 128 + ((i >> 1) & 1) * 64;
        if (lane == 0) {
#pragma unroll
          for (int c = 0; c < 16; ++c) xch[ew * 16 + c] = __uint_as_float(r[16 + c]);
        }
        named_bar_sync(2 + eset, 128);
        float up[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) up[c] = __shfl_down_sync(0xffffffffu, __uint_as_float(r[16 + c]), 1);
        if (lane == 31 && ew < 3) {
#pragma unroll
          for (int c = 0; c < 16; ++c) up[c] = xch[(ew + 1) * 16 + c];
        }
        const int mt = ew * 32 + lane, m = C11_TSTRIDE * i + mt, oy = m / BLK_W, ox = m - oy * BLK_W;
        if (mt < C11_TSTRIDE && oy < H1 && ox < H1) {
          uint32_t o[8];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj)
            o[jj] = pack_bf16(fmaxf(__uint_as_float(r[2 * jj]) + up[2 * jj] + bias_s[2 * jj], 0.f),
                              fmaxf(__uint_as_float(r[2 * jj + 1]) + up[2 * jj + 1] + bias_s[2 * jj + 1], 0.f));
          const uint4 lo = make_uint4(o[0], o[1], o[2], o[3]), hi = make_uint4(o[4], o[5], o[6], o[7]);
          const int p = oy * H1 + ox;
          uint32_t e01, e23;
          lds64(e01, e23, lut + p * 8);
          const uint32_t e[4] = {e01 & 0xFFFFu, e01 >> 16, e23 & 0xFFFFu, e23 >> 16};
#pragma unroll
          for (int idx = 0; idx < 4; ++idx) {
            if (e[idx] != 0xFFFFu) {
              sts128(sa2 + (e[idx] << 4), lo);
              sts128(sa2 + ((e[idx] ^ 1u) << 4), hi);
            }
          }
          if (n1_dst) {
            uint8_t* d = n1_dst + b2_pixel_offset(oy, ox, 0);
            stg128_hint(d, lo, pol_keep);
            stg128_hint(d + B2_LBO, hi, pol_keep);
          }
        }
      }
      fence_proxy_async();                                           // the scatter is read by the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(BAR_A2RDY));
      mark(42, k);
    }
    if (n_frames > 0 && eset == 1) conv12_epilogue(n_frames - 1);
  }

  tc_fence_before();
  __syncthreads();
  trace_mark(K_CONV_FWD, 2);
  if (warp == CF_EPI_WARP0) {
    tc_fence_after();
    tmem_dealloc<CF_TMEM_COLS>(tmem_base);
  }
}

GA3C_TRACE_ATTACH(trace_attach_conv_fwd)
static bool g_evt_attached_cf = false;     // host side: the event-log instantiation is launched only while a log is attached
int evt_attach_conv_fwd(unsigned long long* buf) {
  g_evt_attached_cf = buf != nullptr;
  return (int)cudaMemcpyToSymbol(g_evt, &buf, sizeof(buf));
}

int configure_conv_fwd() {
  cudaError_t e = cudaFuncSetAttribute(conv_fwd_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF_SMEM);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(conv_fwd_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF_SMEM);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(conv_fwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF_SMEM);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaFuncSetAttribute(conv_fwd_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF_SMEM);
}

int launch_conv_fwd(const void* x, bool x_u8, const float* w11, const float* b11, const float* w12, const float* b12,
                    uint8_t* n1_out, uint8_t* xblk_out, uint16_t* n2_out, int batch, int num_sms, cudaStream_t stream) {
  const int grid = min(batch, num_sms);
  const int hints = l2_hints(n1_out != nullptr, x_u8);
  auto kernel = x_u8 ? (g_evt_attached_cf ? conv_fwd_kernel<true, true> : conv_fwd_kernel<true, false>)
                     : (g_evt_attached_cf ? conv_fwd_kernel<false, true> : conv_fwd_kernel<false, false>);
  return launch_pdl(kernel, dim3(grid), dim3(CF_THREADS), CF_SMEM, stream, x, w11, b11, w12, b12, n1_out, xblk_out, n2_out, batch,
                    hints);
}

}  // namespace ga3c
