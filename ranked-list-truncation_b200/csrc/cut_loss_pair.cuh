// K3, packed: the cut losses of heads.cu (`cut_loss_kernel`) for logits in, even L -- the hot configuration of every
// criterion -- rewritten to be HBM-bound instead of issue-bound (round 1: 1068 warp instructions per 300-position list
// for DivLoss JS, 0.48 of the copy peak).  Same mathematics (utils/losses.py:84-101, 216-233; Metric_for_Loss,
// utils/metrics.py:85-101), different instruction stream:
//   * a lane owns PAIRS of neighbouring positions (2*lane, 2*lane+1) + 64*i: 64-bit loads / stores and the packed
//     fp32x2 pipe of sm_100 (FADD2 / FMUL2 / FFMA2) for every elementwise step;
//   * everything in the log2 domain: t = z*log2(e) - m*log2(e) is ONE fused multiply-add feeding ex2, and
//     log2 p = t - log2 s is reused; the factors ln 2 and 1/2 of the JS terms are applied once per list;
//   * JS: loss = (ln2/2) [ sum q (lq - lm) + sum p (lp - lm) ] with lm = log2(p+q) - 1: one lg2 per position, no third
//     logarithm, and the second sum IS the <p, dL/dp> of the softmax backward;
//   * F1 reward: 2c/(k+N) = c * rcp((k+N)/2); the prefix count c from ballots (float labels) or straight from the
//     bit masks of rlt_pack_labels (kBits: 4 L/32 bytes of labels per list instead of 4 L);
//   * positions >= L hold z = r = -1e30 (finite): their p and q are exactly 0 and 0 * finite = 0, so no per-element
//     validity predicates; logits are clamped to >= -1e30 for the same reason (a -inf logit has p = 0 either way).
#pragma once
#include <stdint.h>

#include "f32x2.cuh"

namespace rlt {

// float32(1) / float32(log2(j + 2)): this translation unit's copy of heads.cu's table (uploaded by rlt_set_dcg_tables)
__device__ __align__(16) float g_pair_rcoef32[1024];

__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_fast(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_fast(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// NP = ceil(L / 64) pairs per lane.  loss_kind 0 ChoopyLoss | 1 AttnCutLoss (RAML) | 2 DivLoss kl | 3 DivLoss js.
// kBits: `labels` is the uint32 bit-mask array of rlt_pack_labels ([B, ceil(L/32)] words).
template <int NP, int loss_kind, int metric_dcg, bool kBits>
__global__ void __launch_bounds__(128) cut_loss_pair_kernel(const float* __restrict__ in, const void* __restrict__ labels,
                                                            float* __restrict__ probs_out, float* __restrict__ grad,
                                                            float* __restrict__ loss_per_list, int B, int L, float tau,
                                                            float gscale) {
  constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f, kNeg = -1e30f;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const float2* zin = reinterpret_cast<const float2*>(in + size_t(b) * L);
  const int npair = L >> 1;                           // pairs in the list; pair index of (lane, i) is lane + 32 i
  float2 z[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const int pj = lane + 32 * i;
    z[i] = pj < npair ? zin[pj] : f2_dup(kNeg);
  }
  // ---- labels: per pair-slot the masks of the even / odd positions that are relevant (bit = lane)
  uint32_t me[NP], mo[NP];
  float2 y[NP];                                       // only read by the DCG reward
  int n_rel_i = 0;
  if (kBits) {
    const int words = (L + 31) >> 5;
    const uint32_t* wr = static_cast<const uint32_t*>(labels) + size_t(b) * words;
#pragma unroll
    for (int i = 0; i < NP; ++i) {                    // broadcast loads: one 128-byte line per list
      me[i] = 2 * i < words ? __ldg(wr + 2 * i) : 0u;           // here: the raw words of positions 64 i .. 64 i + 63
      mo[i] = 2 * i + 1 < words ? __ldg(wr + 2 * i + 1) : 0u;
      n_rel_i += __popc(me[i]) + __popc(mo[i]);
    }
  } else {
    const float2* yin = reinterpret_cast<const float2*>(static_cast<const float*>(labels) + size_t(b) * L);
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const int pj = lane + 32 * i;
      y[i] = pj < npair ? yin[pj] : f2_dup(0.f);
      me[i] = __ballot_sync(0xffffffffu, y[i].x == 1.f);
      mo[i] = __ballot_sync(0xffffffffu, y[i].y == 1.f);
      n_rel_i += __popc(me[i]) + __popc(mo[i]);
    }
  }
  // ---- softmax over the positions, log2 domain
  float m = kNeg;
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    z[i].x = fmaxf(z[i].x, kNeg);
    z[i].y = fmaxf(z[i].y, kNeg);
    m = fmaxf(m, fmaxf(z[i].x, z[i].y));
  }
  m = warp_max(m);
  const float ml = m * kLog2e;
  float2 p[NP], s2 = f2_dup(0.f);
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const float2 t = f2_fma(z[i], f2_dup(kLog2e), f2_dup(-ml));
    p[i].x = ex2_fast(t.x);
    p[i].y = ex2_fast(t.y);
    s2 = f2_add(s2, p[i]);
  }
  const float s = warp_sum(s2.x + s2.y);
  const float inv = 1.f / s, l2s = lg2_fast(s);
#pragma unroll
  for (int i = 0; i < NP; ++i) p[i] = f2_mul(p[i], f2_dup(inv));
  if (probs_out != nullptr) {
    float2* po = reinterpret_cast<float2*>(probs_out + size_t(b) * L);
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const int pj = lane + 32 * i;
      if (pj < npair) po[pj] = p[i];
    }
  }
  // ---- rewards r_j for cutting after position j (k = j + 1)
  float2 r[NP];
  if (metric_dcg) {
    // float32 prefix sums of +-1/log2(j+2): a lane adds its pair, the pair sums are scanned over the lanes
    const float2* rc = reinterpret_cast<const float2*>(g_pair_rcoef32);
    float carry = 0.f;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const int pj = lane + 32 * i;
      float2 t = pj < npair ? rc[pj] : f2_dup(0.f);
      if (kBits) {
        const uint32_t w = lane < 16 ? me[i] : mo[i];
        const uint32_t two = (w >> ((2 * lane) & 31)) & 3u;
        t.x = (two & 1u) ? t.x : -t.x;
        t.y = (two & 2u) ? t.y : -t.y;
      } else {
        t = f2_mul(t, f2_fma(y[i], f2_dup(2.f), f2_dup(-1.f)));
      }
      const float inc = warp_incl_scan(t.x + t.y, lane) + carry;
      carry = __shfl_sync(0xffffffffu, inc, 31);
      r[i].y = inc;
      r[i].x = inc - t.y;
    }
  } else {
    // F1: 2 c / (k + N) = c * rcp((k + N) / 2); c = 0 gives 0 without a select (utils/metrics.py:85-91)
    const float n_rel = float(n_rel_i);
    float2 hk = make_float2(0.5f * (float(2 * lane + 1) + n_rel), 0.5f * (float(2 * lane + 2) + n_rel));
    int base = 0;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      int ce, co;
      if (kBits) {
        // raw words: positions 64 i + [0, 32) in me, + [32, 64) in mo; this lane's pair sits at bit 2 lane of the 64
        const uint32_t lo_e = lane < 16 ? (0xffffffffu >> (31 - 2 * lane)) : 0xffffffffu;        // bits <= 2 lane (low word)
        const uint32_t hi_e = lane < 16 ? 0u : (0xffffffffu >> (63 - 2 * lane));                 // bits <= 2 lane (high word)
        const uint32_t own = lane < 16 ? me[i] : mo[i];
        ce = base + __popc(me[i] & lo_e) + __popc(mo[i] & hi_e);
        co = ce + int((own >> ((2 * lane + 1) & 31)) & 1u);
        base += __popc(me[i]) + __popc(mo[i]);
      } else {
        const uint32_t le = 0xffffffffu >> (31 - lane), lt = le >> 1;
        const int ae = __popc(me[i] & le);
        ce = base + ae + __popc(mo[i] & lt);
        co = base + ae + __popc(mo[i] & le);
        base += __popc(me[i]) + __popc(mo[i]);
      }
      // int -> float without the conversion pipe: c < 2^23
      const float2 cf = f2_add(make_float2(__int_as_float(ce | 0x4B000000), __int_as_float(co | 0x4B000000)), f2_dup(-8388608.f));
      r[i] = f2_mul(cf, make_float2(rcp_fast(hk.x), rcp_fast(hk.y)));
      hk = f2_add(hk, f2_dup(32.f));
    }
  }
  // ---- loss and gradient with respect to the logits
  float loss;
  if (loss_kind == 0) {
    // ChoopyLoss: L = -sum p r; dL/dz_j = p_j (-r_j - L) = -p_j (r_j + L)
    float2 acc = f2_dup(0.f);
#pragma unroll
    for (int i = 0; i < NP; ++i) acc = f2_fma(p[i], r[i], acc);
    loss = -warp_sum(acc.x + acc.y);
    if (grad != nullptr) {
      float2* go = reinterpret_cast<float2*>(grad + size_t(b) * L);
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const int pj = lane + 32 * i;
        if (pj < npair) go[pj] = f2_mul(f2_mul(p[i], f2_dup(-gscale)), f2_add(r[i], f2_dup(loss)));
      }
    }
  } else {
    // q = softmax(r / tau) over the L positions (losses.py:90-92, 226-228)
    // positions >= L must not take part in q: L may be far below 64 NP (every L in (64, 320] runs NP = 5), so any slot
    // from npair / 32 on can hold them -- a warp-uniform test per slot
#pragma unroll
    for (int i = 0; i < NP; ++i)
      if (32 * (i + 1) > npair && lane + 32 * i >= npair) r[i] = f2_dup(kNeg);
    float rm = kNeg;
#pragma unroll
    for (int i = 0; i < NP; ++i) rm = fmaxf(rm, fmaxf(r[i].x, r[i].y));
    rm = warp_max(rm);
    const float c1 = kLog2e / tau;
    float2 q[NP], qs2 = f2_dup(0.f);
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const float2 u = f2_fma(r[i], f2_dup(c1), f2_dup(-rm * c1));
      q[i].x = ex2_fast(u.x);
      q[i].y = ex2_fast(u.y);
      qs2 = f2_add(qs2, q[i]);
    }
    const float qs = warp_sum(qs2.x + qs2.y);
    const float qinv = 1.f / qs, l2qs = lg2_fast(qs);
#pragma unroll
    for (int i = 0; i < NP; ++i) q[i] = f2_mul(q[i], f2_dup(qinv));
    if (loss_kind == 1 || loss_kind == 2) {
      // RAML: -sum q log p.   KL(q || p): sum q (log q - log p).   Both: dL/dz = p - q
      float2 acc = f2_dup(0.f);
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const float2 l2p = f2_fma(z[i], f2_dup(kLog2e), f2_dup(-(ml + l2s)));
        if (loss_kind == 1) {
          acc = f2_fma(q[i], l2p, acc);
        } else {
          const float2 l2q = f2_fma(r[i], f2_dup(c1), f2_dup(-(rm * c1 + l2qs)));
          acc = f2_fma(q[i], f2_fma(l2p, f2_dup(-1.f), l2q), acc);
        }
      }
      loss = warp_sum(acc.x + acc.y) * (loss_kind == 1 ? -kLn2 : kLn2);
      if (grad != nullptr) {
        float2* go = reinterpret_cast<float2*>(grad + size_t(b) * L);
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          const int pj = lane + 32 * i;
          if (pj < npair) go[pj] = f2_mul(f2_fma(q[i], f2_dup(-1.f), p[i]), f2_dup(gscale));
        }
      }
    } else {
      // JS: 1/2 [ sum q (log q - log m) + sum p (log p - log m) ], m = (p + q) / 2; dL/dp_j = (log p_j - log m_j) / 2
      float2 accq = f2_dup(0.f), accp = f2_dup(0.f), dp[NP];
      const float2 kq = f2_dup(-(rm * c1 + l2qs - 1.f)), kp = f2_dup(-(ml + l2s - 1.f));
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const float2 sm = f2_add(p[i], q[i]);
        const float2 nlg = make_float2(-lg2_fast(fmaxf(sm.x, 1.1754944e-38f)), -lg2_fast(fmaxf(sm.y, 1.1754944e-38f)));
        const float2 dq = f2_add(f2_fma(r[i], f2_dup(c1), kq), nlg);      // log2 q - log2 m
        dp[i] = f2_add(f2_fma(z[i], f2_dup(kLog2e), kp), nlg);            // log2 p - log2 m
        accq = f2_fma(q[i], dq, accq);
        accp = f2_fma(p[i], dp[i], accp);
      }
      const float dsum = warp_sum(accp.x + accp.y);                       // = <p, dL/dp> in units of ln2 / 2
      loss = 0.5f * kLn2 * (warp_sum(accq.x + accq.y) + dsum);
      if (grad != nullptr) {
        float2* go = reinterpret_cast<float2*>(grad + size_t(b) * L);
        const float cg = 0.5f * kLn2 * gscale;
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          const int pj = lane + 32 * i;
          if (pj < npair) go[pj] = f2_mul(p[i], f2_fma(dp[i], f2_dup(cg), f2_dup(-dsum * cg)));
        }
      }
    }
  }
  if (lane == 0 && loss_per_list != nullptr) loss_per_list[b] = loss;
}

}  // namespace rlt
