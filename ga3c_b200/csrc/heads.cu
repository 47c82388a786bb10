// Value / policy heads, softmax, A3C loss and its backward, fused.  All fp32.
//
// Reference graph (NetworkVP_discrate.py:60-85, :100):
//   v = d1.Wv + bv ; z = d1.Wp + bp ; s = softmax(z) ; p = (s + MIN_POLICY) / (1 + MIN_POLICY*A)
//   cost_p_1 = log(max(sum(p*a), eps)) * (R - stop_gradient(v)) ; cost_p_2 = -beta * sum(log(max(p, eps)) * p)
//   cost_v = 0.5 * (R - v)^2 ; every batch reduction is a SUM.
// The kernel is also the tail of dense1 (NetworkDNav.py:90): it sums the split-K partial tiles the
// tcgen05 GEMM left in d1_part (fixed order => deterministic), adds the bias and applies the ReLU.
// Phase 1: one warp per sample, warp-shuffle reductions for the A+1 dot products, then the whole
//          softmax / loss / dlogits chain in registers; writes d1, p, v, dd1 (bf16) and stages (dz, dv).
// Phase 2: one thread per dense1 feature accumulates dWp[j,:], dWv[j], db1[j] over the chunk.
// Gradient / loss accumulators live in registers across chunks; each CTA stores its partial sums into its own slab
// of the gradient-partial workspace (summed in a fixed order by grad_reduce: no atomics, bit-reproducible).
#include "common.cuh"
#include "kernels.h"

namespace ga3c {

constexpr int HD_THREADS = 256, HD_CHUNK = 8;

template <int A>
__global__ void __launch_bounds__(HD_THREADS) heads_kernel(HeadsArgs p) {
  constexpr int A1 = A + 1;
  __shared__ __align__(16) float wt[A1][FC];           // wt[k][j]: k < A -> Wp[j][k]; k == A -> Wv[j]
  __shared__ float dzs[HD_CHUNK][A1];                  // (dz_0..dz_{A-1}, dv) per sample of the chunk
  __shared__ __align__(16) uint16_t dd1s[HD_CHUNK][FC];
  __shared__ __align__(16) float d1s[HD_CHUNK][FC];    // dense1 output of the chunk (post bias + ReLU)
  __shared__ __align__(16) float b1s[FC];
  __shared__ float bias_s[A1];
  __shared__ float loss_s[HD_THREADS / 32][3];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // The head weights are written by the optimizer launch of the previous step.  A prologue may run while a kernel two or more
  // launches back is still executing (only code after the wait is ordered transitively, common.cuh), so reading them before
  // the dependency wait is safe only when something in between cannot be resident next to that optimizer launch: p.preload
  // (batch >= num_sms: the conv forward in front of this kernel fills every SM, see rmsprop_reduce_kernel).  Otherwise
  // they are read after the wait.
  auto load_weights = [&]() {
    for (int i = tid; i < A1 * FC; i += HD_THREADS) {
      const int k = i / FC, jx = i - k * FC;
      wt[k][jx] = (k < A) ? p.wp[jx * A + k] : p.wv[jx];
    }
    if (tid < A1) bias_s[tid] = (tid < A) ? p.bp[tid] : p.bv[0];
    b1s[tid] = p.b1[tid];
  };
  if (p.preload) load_weights();
  griddep_launch();
  griddep_wait(K_HEADS);               // d1_part comes from the dense1 GEMM that precedes this kernel
  if (!p.preload) load_weights();
  __syncthreads();

  float acc[A1] = {};      // thread j: dWp[j][0..A-1], dWv[j]
  float acc_b1 = 0.f;      // thread j: db1[j]
  float acc_bh = 0.f;      // thread k <= A: dbp[k] / dbv
  float l1 = 0.f, l2 = 0.f, lv = 0.f;   // lane 0 of each warp
  const float inv_mix = 1.f / (1.f + p.min_policy * (float)A);

  const int n_chunks = (p.batch + HD_CHUNK - 1) / HD_CHUNK;
  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    // ---------------- phase 1: warp per sample ----------------
#pragma unroll
    for (int i = 0; i < HD_CHUNK / 8; ++i) {
      const int sl = warp * (HD_CHUNK / 8) + i;
      const int b = c * HD_CHUNK + sl;
      if (b < p.batch) {
        // dense1 tail: sum the split-K partials in split order, + bias, ReLU
        float4 fa = *reinterpret_cast<const float4*>(&b1s[4 * lane]);
        float4 fb = *reinterpret_cast<const float4*>(&b1s[128 + 4 * lane]);
        {
          float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sb = sa;
#pragma unroll 4
          for (int sp = 0; sp < p.n_split; ++sp) {
            const float* src = p.d1_part + ((size_t)sp * p.batch + b) * FC;
            const float4 qa = *reinterpret_cast<const float4*>(src + 4 * lane);
            const float4 qb = *reinterpret_cast<const float4*>(src + 128 + 4 * lane);
            sa.x += qa.x; sa.y += qa.y; sa.z += qa.z; sa.w += qa.w;
            sb.x += qb.x; sb.y += qb.y; sb.z += qb.z; sb.w += qb.w;
          }
          fa.x = fmaxf(fa.x + sa.x, 0.f); fa.y = fmaxf(fa.y + sa.y, 0.f); fa.z = fmaxf(fa.z + sa.z, 0.f); fa.w = fmaxf(fa.w + sa.w, 0.f);
          fb.x = fmaxf(fb.x + sb.x, 0.f); fb.y = fmaxf(fb.y + sb.y, 0.f); fb.z = fmaxf(fb.z + sb.z, 0.f); fb.w = fmaxf(fb.w + sb.w, 0.f);
        }
        *reinterpret_cast<float4*>(p.d1 + (size_t)b * FC + 4 * lane) = fa;
        *reinterpret_cast<float4*>(p.d1 + (size_t)b * FC + 128 + 4 * lane) = fb;
        if (p.train) {
          *reinterpret_cast<float4*>(&d1s[sl][4 * lane]) = fa;
          *reinterpret_cast<float4*>(&d1s[sl][128 + 4 * lane]) = fb;
        }
        float z[A1];
#pragma unroll
        for (int k = 0; k < A1; ++k) {
          const float4 wa = *reinterpret_cast<const float4*>(&wt[k][4 * lane]);
          const float4 wb = *reinterpret_cast<const float4*>(&wt[k][128 + 4 * lane]);
          float s = fa.x * wa.x;
          s = fmaf(fa.y, wa.y, s); s = fmaf(fa.z, wa.z, s); s = fmaf(fa.w, wa.w, s);
          s = fmaf(fb.x, wb.x, s); s = fmaf(fb.y, wb.y, s); s = fmaf(fb.z, wb.z, s); s = fmaf(fb.w, wb.w, s);
          z[k] = warp_sum(s) + bias_s[k];
        }
        const float v = z[A];
        float mx = z[0];
#pragma unroll
        for (int k = 1; k < A; ++k) mx = fmaxf(mx, z[k]);
        float sm[A], den = 0.f;
#pragma unroll
        for (int k = 0; k < A; ++k) { sm[k] = expf(z[k] - mx); den += sm[k]; }
        const float inv_den = 1.f / den;
        float pr[A];
#pragma unroll
        for (int k = 0; k < A; ++k) { sm[k] *= inv_den; pr[k] = p.log_softmax ? sm[k] : (sm[k] + p.min_policy) * inv_mix; }
        if (p.p_out != nullptr) {
#pragma unroll
          for (int k = 0; k < A; ++k) if (lane == k) p.p_out[(size_t)b * A + k] = pr[k];
          if (lane == 0) p.v_out[b] = v;
        }
        if (p.train) {
          const float yr = p.yr[b];
          float av[A], sel = 0.f;
#pragma unroll
          for (int k = 0; k < A; ++k) { av[k] = p.a[(size_t)b * A + k]; sel = fmaf(pr[k], av[k], sel); }
          const float adv = yr - v, dv = v - yr;
          float dz[A], ent = 0.f, c1;
          if (p.log_softmax) {
            // Config.USE_LOG_SOFTMAX (NetworkVP_discrate.py:64-71): lsm = z - max - log(den); cost_p_1 = sum(lsm a) adv,
            // cost_p_2 = -beta sum(lsm s);  dz_k = -adv (a_k - s_k sum(a)) + beta s_k (lsm_k - sum(lsm s))
            const float lden = logf(den);
            float lsm[A], sa = 0.f, sla = 0.f;
#pragma unroll
            for (int k = 0; k < A; ++k) {
              lsm[k] = (z[k] - mx) - lden;
              ent = fmaf(lsm[k], sm[k], ent);
              sa += av[k];
              sla = fmaf(lsm[k], av[k], sla);
            }
#pragma unroll
            for (int k = 0; k < A; ++k) dz[k] = -adv * (av[k] - sm[k] * sa) + p.beta * sm[k] * (lsm[k] - ent);
            c1 = sla * adv;
          } else {
            const float coef = (sel >= p.log_eps) ? adv / sel : 0.f;
            float h[A], sh = 0.f;
#pragma unroll
            for (int k = 0; k < A; ++k) {
              const float lg = logf(fmaxf(pr[k], p.log_eps));
              ent = fmaf(lg, pr[k], ent);
              const float gk = -av[k] * coef + p.beta * (lg + (pr[k] >= p.log_eps ? 1.f : 0.f));
              h[k] = gk * inv_mix;
              sh = fmaf(sm[k], h[k], sh);
            }
#pragma unroll
            for (int k = 0; k < A; ++k) dz[k] = sm[k] * (h[k] - sh);
            c1 = logf(fmaxf(sel, p.log_eps)) * adv;
          }
          // Config.DUAL_RMSPROP: part 1 = gradient of cost_p alone, part 2 = of cost_v alone (0: cost_all)
          const float dvv = p.part == 1 ? 0.f : dv;
          if (p.part == 2) {
#pragma unroll
            for (int k = 0; k < A; ++k) dz[k] = 0.f;
          }
          if (lane == 0) {
            l1 += c1;
            l2 += -p.beta * ent;
            lv += 0.5f * (yr - v) * (yr - v);
#pragma unroll
            for (int k = 0; k < A; ++k) dzs[sl][k] = dz[k];
            dzs[sl][A] = dvv;
          }
          // dd1[j] = relu'(d1[j]) * (sum_k dz_k Wp[j][k] + dv Wv[j]) for this lane's 8 features
          float da[4], db[4];
          {
            const float4 wa = *reinterpret_cast<const float4*>(&wt[A][4 * lane]);
            const float4 wb = *reinterpret_cast<const float4*>(&wt[A][128 + 4 * lane]);
            da[0] = dvv * wa.x; da[1] = dvv * wa.y; da[2] = dvv * wa.z; da[3] = dvv * wa.w;
            db[0] = dvv * wb.x; db[1] = dvv * wb.y; db[2] = dvv * wb.z; db[3] = dvv * wb.w;
          }
#pragma unroll
          for (int k = 0; k < A; ++k) {
            const float4 wa = *reinterpret_cast<const float4*>(&wt[k][4 * lane]);
            const float4 wb = *reinterpret_cast<const float4*>(&wt[k][128 + 4 * lane]);
            da[0] = fmaf(dz[k], wa.x, da[0]); da[1] = fmaf(dz[k], wa.y, da[1]);
            da[2] = fmaf(dz[k], wa.z, da[2]); da[3] = fmaf(dz[k], wa.w, da[3]);
            db[0] = fmaf(dz[k], wb.x, db[0]); db[1] = fmaf(dz[k], wb.y, db[1]);
            db[2] = fmaf(dz[k], wb.z, db[2]); db[3] = fmaf(dz[k], wb.w, db[3]);
          }
          const uint2 qa = make_uint2(pack_bf16(fa.x > 0.f ? da[0] : 0.f, fa.y > 0.f ? da[1] : 0.f),
                                      pack_bf16(fa.z > 0.f ? da[2] : 0.f, fa.w > 0.f ? da[3] : 0.f));
          const uint2 qb = make_uint2(pack_bf16(fb.x > 0.f ? db[0] : 0.f, fb.y > 0.f ? db[1] : 0.f),
                                      pack_bf16(fb.z > 0.f ? db[2] : 0.f, fb.w > 0.f ? db[3] : 0.f));
          *reinterpret_cast<uint2*>(&dd1s[sl][4 * lane]) = qa;
          *reinterpret_cast<uint2*>(&dd1s[sl][128 + 4 * lane]) = qb;
          *reinterpret_cast<uint2*>(p.dd1 + (size_t)b * FC + 4 * lane) = qa;
          *reinterpret_cast<uint2*>(p.dd1 + (size_t)b * FC + 128 + 4 * lane) = qb;
        }
      } else if (p.train) {
        if (lane < A1) dzs[sl][lane] = 0.f;
        *reinterpret_cast<uint2*>(&dd1s[sl][4 * lane]) = make_uint2(0, 0);
        *reinterpret_cast<uint2*>(&dd1s[sl][128 + 4 * lane]) = make_uint2(0, 0);
      }
    }
    if (!p.train) continue;
    __syncthreads();
    // ---------------- phase 2: thread per feature ----------------
    {
      const int jx = tid;
#pragma unroll 4
      for (int s = 0; s < HD_CHUNK; ++s) {
        const int b = c * HD_CHUNK + s;
        const float dval = (b < p.batch) ? d1s[s][jx] : 0.f;
#pragma unroll
        for (int k = 0; k < A1; ++k) acc[k] = fmaf(dval, dzs[s][k], acc[k]);
        acc_b1 += __uint_as_float((uint32_t)dd1s[s][jx] << 16);
      }
      if (tid < A1) {
#pragma unroll 4
        for (int s = 0; s < HD_CHUNK; ++s) acc_bh += dzs[s][tid];
      }
    }
    __syncthreads();
  }

  if (!p.train) { trace_mark(K_HEADS, 2); return; }
  const int64_t slab = (int64_t)blockIdx.x * p.gp_stride;
  {
    const int jx = tid;
#pragma unroll
    for (int k = 0; k < A; ++k) p.g_wp[slab + jx * A + k] = acc[k];
    p.g_wv[slab + jx] = acc[A];
    p.g_b1[slab + jx] = acc_b1;
    if (tid < A) p.g_bp[slab + tid] = acc_bh;
    if (tid == A) p.g_bv[slab] = acc_bh;
  }
  if (lane == 0) { loss_s[warp][0] = l1; loss_s[warp][1] = l2; loss_s[warp][2] = lv; }
  __syncthreads();
  if (tid < 4) {
    float s = 0.f;
    if (tid < 3)
#pragma unroll
      for (int w = 0; w < HD_THREADS / 32; ++w) s += loss_s[w][tid];
    p.loss[slab + tid] = s;
  }
  trace_mark(K_HEADS, 2);
}

GA3C_TRACE_ATTACH(trace_attach_heads)

template <int A>
static int launch_heads_t(const HeadsArgs& args, int grid, cudaStream_t stream) {
  return launch_pdl(heads_kernel<A>, dim3(grid), dim3(HD_THREADS), 0, stream, args);
}

int heads_grid(int batch, int num_sms) { return min((batch + HD_CHUNK - 1) / HD_CHUNK, 2 * num_sms); }

int launch_heads(const HeadsArgs& args, int num_sms, cudaStream_t stream) {
  const int grid = heads_grid(args.batch, num_sms);
  switch (args.num_actions) {
#define GA3C_CASE(N) case N: return launch_heads_t<N>(args, grid, stream);
    GA3C_CASE(1) GA3C_CASE(2) GA3C_CASE(3) GA3C_CASE(4) GA3C_CASE(5) GA3C_CASE(6) GA3C_CASE(7) GA3C_CASE(8)
    GA3C_CASE(9) GA3C_CASE(10) GA3C_CASE(11) GA3C_CASE(12) GA3C_CASE(13) GA3C_CASE(14) GA3C_CASE(15)
    GA3C_CASE(16) GA3C_CASE(17) GA3C_CASE(18)
#undef GA3C_CASE
    default: return (int)cudaErrorInvalidValue;
  }
}

}  // namespace ga3c
