// Per-list kernels: cut head + reward losses (K3), argmax-cut + F1/DCG evaluation (K4), auxiliary
// heads (classification BCE + batch-global rerank hinge), BiCut head loss, and the d->1 head dots.
//
// Layout: one WARP per ranked list; element j of the list lives in lane (j % 32), slot (j / 32), so
// global loads/stores are fully coalesced 128-byte rows and prefix sums over positions are
// ceil(L/32) warp scans with a running carry.  HBM-bound: 12*L bytes per list (K3 with gradient),
// 8*L + 20 bytes per list (K4).
#include <math.h>

#include <type_traits>

#include "common.h"
#include "warp_utils.cuh"

namespace rlt {

// float32-rounded math.log(j+2, 2) (Metric_for_Loss.dcg builds torch.tensor(DCG_coef_300[:k]), reference
// utils/metrics.py:7,97) — uploaded once by the host from the Python-side table so that it is the SAME
// table, not a device log2.
__device__ float g_dcg_coef32[1024];
// float32(1) / coef32 (correctly rounded on the host): (+-1) / coef of the reference without a device division
__device__ __align__(16) float g_dcg_rcoef32[1024];
// float64 1/math.log(j+2, 2) for Metric.dcg (utils/metrics.py:26-38)
__device__ double g_dcg_term64[1024];

// ------------------------------------------------------------------------------------------
// K3: cut loss.  input_kind 0: `in` holds logits z (softmax over positions fused here, grad = dL/dz)
//                input_kind 1: `in` holds probabilities p (the reference API boundary, grad = dL/dp)
// loss_kind 0 ChoopyLoss | 1 AttnCutLoss (RAML) | 2 DivLoss kl | 3 DivLoss js
// ------------------------------------------------------------------------------------------
// The three configuration words are template parameters: every criterion of the reference gets its own small kernel
// (the run-time-switched kernel spent 15 % of its samples on instruction fetch, profiles/r01_ncu_k3.txt).
template <int NI, int input_kind, int loss_kind, int metric_dcg>
__global__ void __launch_bounds__(128) cut_loss_kernel(const float* __restrict__ in, const float* __restrict__ labels,
                                                       float* __restrict__ probs_out, float* __restrict__ grad,
                                                       float* __restrict__ loss_per_list,
                                                       float* __restrict__ rewards_out, int B, int L, float tau,
                                                       float gscale) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const float* zin = in + size_t(b) * L;
  const float* yin = labels + size_t(b) * L;
  float z[NI], y[NI];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int j = lane + 32 * i;
    z[i] = (j < L && in != nullptr) ? zin[j] : (input_kind == 0 ? -INFINITY : 0.f);
    y[i] = j < L ? yin[j] : 0.f;
  }
  // ---- probabilities and log-probabilities
  float p[NI], logp[NI];
  if (input_kind == 0) {
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < NI; ++i) m = fmaxf(m, z[i]);
    m = warp_max(m);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NI; ++i) { p[i] = __expf(z[i] - m); s += p[i]; }
    s = warp_sum(s);
    const float inv = 1.f / s, ls = __logf(s);
#pragma unroll
    for (int i = 0; i < NI; ++i) { p[i] *= inv; logp[i] = z[i] - m - ls; }
  } else {
#pragma unroll
    for (int i = 0; i < NI; ++i) { p[i] = z[i]; logp[i] = __logf(z[i]); }
  }
  if (probs_out != nullptr) {
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int j = lane + 32 * i;
      if (j < L) probs_out[size_t(b) * L + j] = p[i];
    }
  }
  // ---- rewards r_j for cutting after position j (k = j+1), SURVEY.md A.3
  float r[NI];
  float n_rel = 0.f;
#pragma unroll
  for (int i = 0; i < NI; ++i) n_rel += y[i];
  n_rel = warp_sum(n_rel);
  if (metric_dcg) {
    float carry = 0.f;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int j = lane + 32 * i;
      const float term = (j < L) ? (y[i] == 1.f ? g_dcg_rcoef32[j] : -g_dcg_rcoef32[j]) : 0.f;
      const float inc = warp_incl_scan(term, lane) + carry;
      carry = __shfl_sync(0xffffffffu, inc, 31);
      r[i] = inc;
    }
  } else {
    // labels are 0/1: the prefix count c_k comes from one ballot + popc per 32 positions instead of a shuffle scan.
    // Metric_for_Loss.f1 (utils/metrics.py:85-91): p = c/k, r = c/N (0 if N == 0), 2pr/(p+r) (0 if p+r == 0), which is
    // 2c / (k + N) for c > 0 and 0 otherwise (max abs difference 3e-8): one fast division.
    const uint32_t le = 0xffffffffu >> (31 - lane);
    int carry = 0;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const uint32_t m = __ballot_sync(0xffffffffu, y[i] == 1.f);
      const int c = carry + __popc(m & le);
      carry += __popc(m);
      if (rewards_out != nullptr) {
        // the reward-matrix API keeps the reference's operation order: bit-exact against its goldens
        const float inc = float(c);
        const float prec = __fdiv_rn(inc, float(lane + 32 * i + 1));
        const float rec = n_rel != 0.f ? __fdiv_rn(inc, n_rel) : 0.f;
        const float den = prec + rec;
        r[i] = den != 0.f ? __fdiv_rn(prec * rec * 2.f, den) : 0.f;
      } else {
        r[i] = c > 0 ? __fdividef(2.f * float(c), float(lane + 32 * i + 1) + n_rel) : 0.f;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int j = lane + 32 * i;
    if (j >= L) r[i] = 0.f;
    else if (rewards_out != nullptr) rewards_out[size_t(b) * L + j] = r[i];
  }
  if (in == nullptr) return;  // reward-matrix-only call (warp-uniform)
  // ---- loss and gradient w.r.t. p
  float loss = 0.f;
  float g[NI];  // dL_b/dp_j (unscaled)
  if (loss_kind == 0) {
#pragma unroll
    for (int i = 0; i < NI; ++i) { loss -= p[i] * r[i]; g[i] = -r[i]; }
  } else {
    // q = softmax(r / tau) over the L positions (losses.py:90-92, 226-228)
    const float itau = 1.f / tau;
    float rm = -INFINITY;
#pragma unroll
    for (int i = 0; i < NI; ++i) if (lane + 32 * i < L) rm = fmaxf(rm, r[i]);
    rm = warp_max(rm);
    float q[NI], qs = 0.f;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      q[i] = (lane + 32 * i < L) ? __expf((r[i] - rm) * itau) : 0.f;
      qs += q[i];
    }
    qs = warp_sum(qs);
    const float qinv = 1.f / qs, lqs = __logf(qs);
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const bool ok = lane + 32 * i < L;
      const float logq = (r[i] - rm) * itau - lqs;
      q[i] *= qinv;
      if (loss_kind == 1) {         // RAML: -sum q log p
        if (ok) loss -= q[i] * logp[i];
        g[i] = ok ? -__fdividef(q[i], p[i]) : 0.f;
      } else if (loss_kind == 2) {  // KL(q || p) = sum q (log q - log p)
        if (ok && q[i] > 0.f) loss += q[i] * (logq - logp[i]);
        g[i] = ok ? -__fdividef(q[i], p[i]) : 0.f;
      } else {                      // JS: 1/2 [ sum q (log q - log m) + sum p (log p - log m) ],  m = (p+q)/2
        const float mm = 0.5f * (p[i] + q[i]);
        const float logm = __logf(mm);
        if (ok && q[i] > 0.f) loss += 0.5f * q[i] * (logq - logm);
        if (ok && p[i] > 0.f) loss += 0.5f * p[i] * (logp[i] - logm);
        g[i] = (ok && p[i] > 0.f) ? 0.5f * (logp[i] - logm) : 0.f;
      }
      if (input_kind == 0 && (loss_kind == 1 || loss_kind == 2)) g[i] = ok ? (p[i] - q[i]) : 0.f;  // already dL/dz
    }
  }
  loss = warp_sum(loss);
  if (lane == 0 && loss_per_list != nullptr) loss_per_list[b] = loss;
  if (grad != nullptr) {
    if (input_kind == 0 && !(loss_kind == 1 || loss_kind == 2)) {
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < NI; ++i) dot += p[i] * g[i];
      dot = warp_sum(dot);
#pragma unroll
      for (int i = 0; i < NI; ++i) g[i] = p[i] * (g[i] - dot);
    }
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int j = lane + 32 * i;
      if (j < L) grad[size_t(b) * L + j] = g[i] * gscale;
    }
  }
}

// deterministic sum of n values by ONE CTA: out = (accumulate ? out : 0) + scale * sum
__global__ void __launch_bounds__(256) reduce_scale_kernel(const float* __restrict__ v, int n, float scale,
                                                           float* __restrict__ out, int accumulate) {
  __shared__ float red[256];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) acc += v[i];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = (accumulate ? *out : 0.f) + scale * red[0];
}

// ------------------------------------------------------------------------------------------
// softmax over the L positions of each list, and its backward (model output boundary)
// ------------------------------------------------------------------------------------------
template <int NI>
__global__ void __launch_bounds__(128) softmax_lists_kernel(const float* __restrict__ z, float* __restrict__ p, int B,
                                                            int L) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  float v[NI];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int j = lane + 32 * i;
    v[i] = j < L ? z[size_t(b) * L + j] : -INFINITY;
    m = fmaxf(m, v[i]);
  }
  m = warp_max(m);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NI; ++i) { v[i] = __expf(v[i] - m); s += v[i]; }
  s = warp_sum(s);
  const float inv = 1.f / s;
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int j = lane + 32 * i;
    if (j < L) p[size_t(b) * L + j] = v[i] * inv;
  }
}

// dz = p * (dp - <p, dp>)  (in place allowed: dz may alias dp)
template <int NI>
__global__ void __launch_bounds__(128) softmax_lists_bwd_kernel(const float* __restrict__ p, const float* __restrict__ dp,
                                                                float* __restrict__ dz, int B, int L) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  float pv[NI], gv[NI];
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int j = lane + 32 * i;
    pv[i] = j < L ? p[size_t(b) * L + j] : 0.f;
    gv[i] = j < L ? dp[size_t(b) * L + j] : 0.f;
    dot += pv[i] * gv[i];
  }
  dot = warp_sum(dot);
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int j = lane + 32 * i;
    if (j < L) dz[size_t(b) * L + j] = pv[i] * (gv[i] - dot);
  }
}

// ------------------------------------------------------------------------------------------
// K4: argmax cut + per-list F1 / DCG with the reference's numpy arithmetic:
//   k      = first argmax of the list + 1                                   (run.py:140-142)
//   F1     = ((2*p)*r)/(p+r), p = double(count)/double(k), r = float(count)/float(N_D) (float32!)
//   DCG    = numpy pairwise float64 sum of +-1/log(j+2, 2) over j < k       (utils/metrics.py:26-38)
// mode 1 (BiCut, run.py:132-136): `probs` is [B, L, 2]; predicted class = argmax over the 2 classes
//   (tie -> 0); k = L with Python-int semantics (float32 precision) if no position predicts class 0,
//   else first such position + 1.
// A warp processes 32 lists: phase 1 cooperatively (coalesced rows -> argmax, label bit masks in
// shared memory), phase 2 one list per lane (the order-exact float64 summation).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double np_pairwise_block(const uint32_t* bits, int a0, int n) {
  // numpy pairwise_sum for n <= 128 contiguous terms starting at index a0 (n >= 8), or the plain loop (n < 8)
  auto term = [&](int j) -> double {
    const double t = g_dcg_term64[j];
    return ((bits[j >> 5] >> (j & 31)) & 1u) ? t : -t;
  };
  if (n < 8) {
    double s = term(a0);
    for (int i = 1; i < n; ++i) s = __dadd_rn(s, term(a0 + i));
    return s;
  }
  double r0 = term(a0), r1 = term(a0 + 1), r2 = term(a0 + 2), r3 = term(a0 + 3), r4 = term(a0 + 4),
         r5 = term(a0 + 5), r6 = term(a0 + 6), r7 = term(a0 + 7);
  int i = 8;
  for (; i < n - (n % 8); i += 8) {
    r0 = __dadd_rn(r0, term(a0 + i)); r1 = __dadd_rn(r1, term(a0 + i + 1));
    r2 = __dadd_rn(r2, term(a0 + i + 2)); r3 = __dadd_rn(r3, term(a0 + i + 3));
    r4 = __dadd_rn(r4, term(a0 + i + 4)); r5 = __dadd_rn(r5, term(a0 + i + 5));
    r6 = __dadd_rn(r6, term(a0 + i + 6)); r7 = __dadd_rn(r7, term(a0 + i + 7));
  }
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r0, r1), __dadd_rn(r2, r3)), __dadd_rn(__dadd_rn(r4, r5), __dadd_rn(r6, r7)));
  for (; i < n; ++i) res = __dadd_rn(res, term(a0 + i));
  return res;
}

// recursion of numpy's pairwise sum (n > 128: split at n/2 rounded down to a multiple of 8), done with an
// explicit stack; depth <= 4 for n <= 1024.
__device__ double np_pairwise(const uint32_t* bits, int n) {
  if (n <= 128) return np_pairwise_block(bits, 0, n);
  int st_a[8], st_n[8], st_state[8];
  double st_left[8];
  int sp = 0;
  st_a[0] = 0; st_n[0] = n; st_state[0] = 0;
  double ret = 0.0;
  while (sp >= 0) {
    const int a = st_a[sp], m = st_n[sp];
    if (m <= 128) {
      ret = np_pairwise_block(bits, a, m);
      --sp;
      continue;
    }
    int h = m / 2;
    h -= h % 8;
    if (st_state[sp] == 0) {         // descend left
      st_state[sp] = 1;
      ++sp; st_a[sp] = a; st_n[sp] = h; st_state[sp] = 0;
    } else if (st_state[sp] == 1) {  // left done -> descend right
      st_left[sp] = ret;
      st_state[sp] = 2;
      ++sp; st_a[sp] = a + h; st_n[sp] = m - h; st_state[sp] = 0;
    } else {                         // both done
      ret = __dadd_rn(st_left[sp], ret);
      --sp;
    }
  }
  return ret;
}

// NI = ceil(L / 32) elements per lane.  Phase 1 keeps the NEXT list's probabilities and labels in flight (2 x NI
// coalesced loads per lane) while the current one is reduced; the arg-max uses two redux.sync operations on an
// order-preserving integer key instead of a 10-shuffle butterfly (the first version of this kernel ran one dependent
// load-compare loop per list and reached 0.27 of the HBM peak).
template <int NI>
__global__ void __launch_bounds__(128) eval_cut_kernel(const float* __restrict__ probs, const float* __restrict__ labels,
                                                       const int32_t* __restrict__ k_in,
                                                       const int32_t* __restrict__ pyint_in, int B, int L, int mode, int32_t* __restrict__ k_out,
                                                       int32_t* __restrict__ count_out, int32_t* __restrict__ nrel_out,
                                                       double* __restrict__ f1_out, double* __restrict__ dcg_out) {
  __shared__ uint32_t s_bits[4][32][33];  // [warp][list in warp][word]; 33: phase 2 reads one ROW per lane, conflict-free
  __shared__ int s_k[4][32];
  __shared__ int s_pyint[4][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwords = (L + 31) / 32;
  const long long base = (long long)(blockIdx.x * 4 + warp) * 32;
  const bool argmax_path = k_in == nullptr && mode == 0;
  // ---- phase 1: cooperative, one list at a time, the next list's rows already in flight
  float pv[NI], yv[NI];
  auto fetch = [&](long long b, float (&pd)[NI], float (&yd)[NI]) {
    const float* yr = labels + size_t(b) * L;
    const float* pr = probs + size_t(b) * L;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int j = lane + 32 * i;
      yd[i] = j < L ? __ldg(yr + j) : 0.f;
      pd[i] = (argmax_path && j < L) ? __ldg(pr + j) : -INFINITY;
    }
  };
  if (base < B) fetch(base, pv, yv);
  for (int t = 0; t < 32; ++t) {
    const long long b = base + t;
    if (b >= B) break;  // warp-uniform
    float pn[NI], yn[NI];
    if (t + 1 < 32 && b + 1 < B) fetch(b + 1, pn, yn);
    int best_j = 0x7fffffff;
    if (k_in != nullptr) {  // cut positions supplied by the caller (Metric.f1 / Metric.dcg API)
      if (lane == 0) { s_k[warp][t] = k_in[b]; s_pyint[warp][t] = pyint_in ? pyint_in[b] : 0; }
    } else if (mode == 0) {
      float best = -INFINITY;
#pragma unroll
      for (int i = 0; i < NI; ++i)
        if (pv[i] > best) { best = pv[i]; best_j = lane + 32 * i; }  // strict >: the first maximum of this lane wins
      // order-preserving key (negative floats: all bits flipped; others: sign bit set); 0 = "nothing above -inf / NaN"
      const uint32_t u = __float_as_uint(best);
      const uint32_t key = best_j == 0x7fffffff ? 0u : ((u & 0x80000000u) ? ~u : (u | 0x80000000u));
      const uint32_t kmax = __reduce_max_sync(0xffffffffu, key);
      const uint32_t cand = (key == kmax && best_j != 0x7fffffff) ? uint32_t(best_j) : 0x7fffffffu;
      best_j = int(__reduce_min_sync(0xffffffffu, cand));          // the first position among the lanes that hold the maximum
      if (best_j == 0x7fffffff) best_j = 0;  // all -inf / NaN row
      if (lane == 0) { s_k[warp][t] = best_j + 1; s_pyint[warp][t] = 0; }
    } else {
      const float2* pr = reinterpret_cast<const float2*>(probs) + size_t(b) * L;
      for (int j = lane; j < L; j += 32) {
        const float2 v = pr[j];
        if (!(v.y > v.x)) { best_j = j; break; }  // class 0 ("truncate") wins ties; first such position
      }
      best_j = int(__reduce_min_sync(0xffffffffu, uint32_t(best_j)));
      if (lane == 0) {
        const bool none = best_j == 0x7fffffff;
        s_k[warp][t] = none ? L : best_j + 1;
        s_pyint[warp][t] = none ? 1 : 0;
      }
    }
#pragma unroll
    for (int w = 0; w < NI; ++w) {
      const uint32_t m = __ballot_sync(0xffffffffu, yv[w] == 1.f);   // positions >= L hold 0
      if (lane == 0 && w < nwords) s_bits[warp][t][w] = m;
    }
#pragma unroll
    for (int i = 0; i < NI; ++i) { pv[i] = pn[i]; yv[i] = yn[i]; }
  }
  __syncwarp();
  // ---- phase 2: one list per lane
  const long long b = base + lane;
  if (b >= B) return;
  const uint32_t* bits = s_bits[warp][lane];
  const int k = s_k[warp][lane];
  int count = 0, nrel = 0;
  for (int w = 0; w < nwords; ++w) {
    const uint32_t m = bits[w];
    nrel += __popc(m);
    const int lo = w * 32;
    if (k >= lo + 32) count += __popc(m);
    else if (k > lo) count += __popc(m & ((1u << (k - lo)) - 1u));
  }
  double f1;
  const float r32 = nrel != 0 ? __fdiv_rn(float(count), float(nrel)) : 0.f;
  if (s_pyint[warp][lane]) {
    // k is a Python int: numpy keeps float32 (count is np.float32): p, 2*p*r and p+r all in float32
    const float p32 = __fdiv_rn(float(count), float(k));
    const float den = __fadd_rn(p32, r32);
    f1 = den != 0.f ? double(__fdiv_rn(__fmul_rn(__fmul_rn(2.f, p32), r32), den)) : 0.0;
  } else {
    const double p64 = __ddiv_rn(double(count), double(k));
    const double den = __dadd_rn(p64, double(r32));
    f1 = den != 0.0 ? __ddiv_rn(__dmul_rn(__dmul_rn(2.0, p64), double(r32)), den) : 0.0;
  }
  if (k_out) k_out[b] = k;
  if (count_out) count_out[b] = count;
  if (nrel_out) nrel_out[b] = nrel;
  if (f1_out) f1_out[b] = f1;
  if (dcg_out) dcg_out[b] = np_pairwise(bits, k);
}

// ------------------------------------------------------------------------------------------
// K4, the arg-max path at even L (the inference sweep of BASELINE config 5), second version.  Same results bit for bit;
// what changed against eval_cut_kernel above (576 warp instructions per 300-position list, 0.56 of the copy peak):
//   * phase 1 reads the probabilities as 64-bit pairs (position 2 lane + 64 i), labels stay one position per lane so
//     that a ballot IS the bit word of 32 consecutive positions; with kBits the words come from rlt_pack_labels
//     (4 L + L/8 bytes per list instead of 8 L) and phase 1 is the arg-max alone;
//   * phase 2 (numpy's pairwise float64 order, one list per lane) took 12.7 instructions per term: the 1/log2(j+2)
//     table now sits in shared memory (one broadcast LDS.64 per term), every leaf of the pairwise tree starts at a
//     multiple of 8, so the 8 label bits of an unrolled step are ONE byte of the mask and the sign is one shift + one
//     LOP3 on the high word: 4 instructions per term.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double signed_term(double t, uint32_t nbits, int q) {
  // nbits = ~label bits; bit q set -> the label is 0 -> the term is negative
  const int hi = __double2hiint(t) ^ int((nbits << (31 - q)) & 0x80000000u);
  return __hiloint2double(hi, __double2loint(t));
}
__device__ __noinline__ double np_leaf(const uint32_t* bits, const double* term, int a0, int n) {
  auto one = [&](int j) -> double {
    const double t = term[j];
    return ((bits[j >> 5] >> (j & 31)) & 1u) ? t : -t;
  };
  if (n < 8) {
    double s = one(a0);
    for (int i = 1; i < n; ++i) s = __dadd_rn(s, one(a0 + i));
    return s;
  }
  // a0 is a multiple of 8 (leaves start at 0 or at a split point rounded to 8): 8 consecutive labels = one byte
  uint32_t nb = ~(bits[a0 >> 5] >> (a0 & 31));
  double r0 = signed_term(term[a0], nb, 0), r1 = signed_term(term[a0 + 1], nb, 1), r2 = signed_term(term[a0 + 2], nb, 2),
         r3 = signed_term(term[a0 + 3], nb, 3), r4 = signed_term(term[a0 + 4], nb, 4), r5 = signed_term(term[a0 + 5], nb, 5),
         r6 = signed_term(term[a0 + 6], nb, 6), r7 = signed_term(term[a0 + 7], nb, 7);
  int i = 8;
  const int n8 = n - (n % 8);
  for (; i < n8; i += 8) {
    const int j = a0 + i;
    nb = ~(bits[j >> 5] >> (j & 31));
    r0 = __dadd_rn(r0, signed_term(term[j], nb, 0)); r1 = __dadd_rn(r1, signed_term(term[j + 1], nb, 1));
    r2 = __dadd_rn(r2, signed_term(term[j + 2], nb, 2)); r3 = __dadd_rn(r3, signed_term(term[j + 3], nb, 3));
    r4 = __dadd_rn(r4, signed_term(term[j + 4], nb, 4)); r5 = __dadd_rn(r5, signed_term(term[j + 5], nb, 5));
    r6 = __dadd_rn(r6, signed_term(term[j + 6], nb, 6)); r7 = __dadd_rn(r7, signed_term(term[j + 7], nb, 7));
  }
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r0, r1), __dadd_rn(r2, r3)), __dadd_rn(__dadd_rn(r4, r5), __dadd_rn(r6, r7)));
  for (; i < n; ++i) res = __dadd_rn(res, one(a0 + i));
  return res;
}
// numpy's pairwise sum (pairwise_sum in numpy/core/src/umath/loops_utils.h.src): n <= 128 one leaf; else split at n/2
// rounded down to a multiple of 8 and recurse.  The larger part has at most n/2 + 7.5 terms, so n <= 1024 needs at most
// four splits (1023 -> 519 -> 263 -> 135 -> 71); the recursion is a template over the remaining depth, no local-memory
// stack.
template <int D>
__device__ __forceinline__ double np_pairwise_fast(const uint32_t* bits, const double* term, int a, int m) {
  if constexpr (D == 0) {
    return np_leaf(bits, term, a, m);
  } else {
    if (m <= 128) return np_leaf(bits, term, a, m);
    int h = m / 2;
    h -= h % 8;
    const double l = np_pairwise_fast<D - 1>(bits, term, a, h);
    return __dadd_rn(l, np_pairwise_fast<D - 1>(bits, term, a + h, m - h));
  }
}

template <int NP, bool kBits, bool kPrefetch>
__global__ void __launch_bounds__(128) eval_cut_fast_kernel(const float* __restrict__ probs, const void* __restrict__ labels,
                                                            int B, int L, int32_t* __restrict__ k_out,
                                                            int32_t* __restrict__ count_out, int32_t* __restrict__ nrel_out,
                                                            double* __restrict__ f1_out, double* __restrict__ dcg_out) {
  constexpr int NI = 2 * NP;
  __shared__ uint32_t s_bits[4][32][33];  // [warp][list in warp][word]; 33: phase 2 reads one ROW per lane, conflict-free
  __shared__ int s_k[4][32];
  __shared__ double s_term[64 * NP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwords = (L + 31) / 32, npair = L >> 1;
  const long long base = (long long)(blockIdx.x * 4 + warp) * 32;
  if (dcg_out != nullptr)
    for (int j = threadIdx.x; j < L; j += 128) s_term[j] = g_dcg_term64[j];
  // ---- phase 1: cooperative, one list at a time
  float2 pv[NP];
  float yv[kBits ? 1 : NI];
  auto fetch = [&](long long b, float2 (&pd)[NP], float (&yd)[kBits ? 1 : NI]) {
    const float2* pr = reinterpret_cast<const float2*>(probs + size_t(b) * L);
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const int pj = lane + 32 * i;
      pd[i] = pj < npair ? __ldg(pr + pj) : make_float2(-INFINITY, -INFINITY);
    }
    if (!kBits) {
      const float* yr = static_cast<const float*>(labels) + size_t(b) * L;
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int j = lane + 32 * i;
        yd[i] = j < L ? __ldg(yr + j) : 0.f;
      }
    }
  };
  if (kBits) {   // the warp's 32 x nwords label words: one coalesced copy into the rows phase 2 reads
    const uint32_t* src = static_cast<const uint32_t*>(labels) + size_t(base) * nwords;
    const long long left = (long long)(B - base) * nwords;
    for (int idx = lane; idx < 32 * nwords && idx < left; idx += 32) s_bits[warp][idx / nwords][idx % nwords] = __ldg(src + idx);
  }
  if (base < B) fetch(base, pv, yv);
  for (int t = 0; t < 32; ++t) {
    const long long b = base + t;
    if (b >= B) break;  // warp-uniform
    float2 pn[NP];
    float yn[kBits ? 1 : NI];
    if (kPrefetch && t + 1 < 32 && b + 1 < B) fetch(b + 1, pn, yn);
    float best = -INFINITY;
    int best_j = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < NP; ++i) {          // strict >: the first maximum of this lane wins (positions ascend with i)
      if (pv[i].x > best) { best = pv[i].x; best_j = 2 * lane + 64 * i; }
      if (pv[i].y > best) { best = pv[i].y; best_j = 2 * lane + 64 * i + 1; }
    }
    // order-preserving key (negative floats: all bits flipped; others: sign bit set); 0 = "nothing above -inf / NaN"
    const uint32_t u = __float_as_uint(best);
    const uint32_t key = best_j == 0x7fffffff ? 0u : ((u & 0x80000000u) ? ~u : (u | 0x80000000u));
    const uint32_t kmax = __reduce_max_sync(0xffffffffu, key);
    const uint32_t cand = (key == kmax && best_j != 0x7fffffff) ? uint32_t(best_j) : 0x7fffffffu;
    best_j = int(__reduce_min_sync(0xffffffffu, cand));          // the first position among the lanes that hold the maximum
    if (best_j == 0x7fffffff) best_j = 0;  // all -inf / NaN row
    if (lane == 0) s_k[warp][t] = best_j + 1;
    if (!kBits) {
      uint32_t mine = 0u;
#pragma unroll
      for (int w = 0; w < NI; ++w) {
        const uint32_t m = __ballot_sync(0xffffffffu, yv[w] == 1.f);   // positions >= L hold 0
        if (lane == w) mine = m;
      }
      if (lane < nwords) s_bits[warp][t][lane] = mine;
    }
    if (kPrefetch) {
#pragma unroll
      for (int i = 0; i < NP; ++i) pv[i] = pn[i];
      if (!kBits) {
#pragma unroll
        for (int i = 0; i < NI; ++i) yv[i] = yn[i];
      }
    } else if (t + 1 < 32 && b + 1 < B) {
      fetch(b + 1, pv, yv);
    }
  }
  __syncthreads();      // s_term (all warps) and this warp's s_bits / s_k rows
  // ---- phase 2: one list per lane
  const long long b = base + lane;
  if (b >= B) return;
  const uint32_t* bits = s_bits[warp][lane];
  const int k = s_k[warp][lane];
  int count = 0, nrel = 0;
  for (int w = 0; w < nwords; ++w) {
    const uint32_t m = bits[w];
    nrel += __popc(m);
    const int lo = w * 32;
    if (k >= lo + 32) count += __popc(m);
    else if (k > lo) count += __popc(m & ((1u << (k - lo)) - 1u));
  }
  // F1 with numpy's promotion (k is an np.int64 here): p in float64, r in float32 (utils/metrics.py:15-24)
  const float r32 = nrel != 0 ? __fdiv_rn(float(count), float(nrel)) : 0.f;
  const double p64 = __ddiv_rn(double(count), double(k));
  const double den = __dadd_rn(p64, double(r32));
  const double f1 = den != 0.0 ? __ddiv_rn(__dmul_rn(__dmul_rn(2.0, p64), double(r32)), den) : 0.0;
  if (k_out) k_out[b] = k;
  if (count_out) count_out[b] = count;
  if (nrel_out) nrel_out[b] = nrel;
  if (f1_out) f1_out[b] = f1;
  if (dcg_out) dcg_out[b] = np_pairwise_fast<4>(bits, s_term, 0, k);
}

// ------------------------------------------------------------------------------------------
// d -> n_heads dot products per token (Linear(d, 1) heads): z[h, t] = x[t,:] . w[h,:] + b[h]
// ------------------------------------------------------------------------------------------
template <int D, int NH>
__global__ void __launch_bounds__(256) head_dots_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias, float* __restrict__ z,
                                                            int T) {
  constexpr int V4 = D / 128;
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= T) return;
  float acc[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) acc[h] = 0.f;
#pragma unroll
  for (int i = 0; i < V4; ++i) {
    const float4 a = reinterpret_cast<const float4*>(x + size_t(t) * D)[lane + 32 * i];
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      const float4 ww = __ldg(reinterpret_cast<const float4*>(w + h * D) + lane + 32 * i);
      acc[h] += (a.x * ww.x + a.y * ww.y) + (a.z * ww.z + a.w * ww.w);
    }
  }
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    const float s = warp_sum(acc[h]);
    if (lane == 0) z[size_t(h) * T + t] = s + bias[h];
  }
}

// dx[t,:] (+)= sum_h dz[h,t] w[h,:] ; dw[h,:] += sum_t dz[h,t] x[t,:] ; db[h] += sum_t dz[h,t]
template <int D, int NH>
__global__ void __launch_bounds__(256) head_dots_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ dz, float* __restrict__ dx,
                                                            float* __restrict__ dw, float* __restrict__ db, int T,
                                                            int accumulate_dx, int relu_gate) {
  constexpr int V4 = D / 128;
  __shared__ float red[NH][D];
  __shared__ float redb[NH];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < NH * D; i += blockDim.x) (&red[0][0])[i] = 0.f;
  if (threadIdx.x < NH) redb[threadIdx.x] = 0.f;
  __syncthreads();
  float4 ww[NH][V4], aw[NH][V4];
  float ab[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    ab[h] = 0.f;
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      ww[h][i] = __ldg(reinterpret_cast<const float4*>(w + h * D) + lane + 32 * i);
      aw[h][i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  for (int t = blockIdx.x * nwarps + warp; t < T; t += gridDim.x * nwarps) {
    float g[NH];
#pragma unroll
    for (int h = 0; h < NH; ++h) { g[h] = dz[size_t(h) * T + t]; ab[h] += g[h]; }
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      const float4 a = reinterpret_cast<const float4*>(x + size_t(t) * D)[lane + 32 * i];
      float4 o = accumulate_dx ? reinterpret_cast<float4*>(dx + size_t(t) * D)[lane + 32 * i]
                               : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        o.x = fmaf(g[h], ww[h][i].x, o.x); o.y = fmaf(g[h], ww[h][i].y, o.y);
        o.z = fmaf(g[h], ww[h][i].z, o.z); o.w = fmaf(g[h], ww[h][i].w, o.w);
        aw[h][i].x = fmaf(g[h], a.x, aw[h][i].x); aw[h][i].y = fmaf(g[h], a.y, aw[h][i].y);
        aw[h][i].z = fmaf(g[h], a.z, aw[h][i].z); aw[h][i].w = fmaf(g[h], a.w, aw[h][i].w);
      }
      if (relu_gate) {  // x is a ReLU output: pass the gradient only where it was active
        o.x = a.x > 0.f ? o.x : 0.f; o.y = a.y > 0.f ? o.y : 0.f; o.z = a.z > 0.f ? o.z : 0.f; o.w = a.w > 0.f ? o.w : 0.f;
      }
      reinterpret_cast<float4*>(dx + size_t(t) * D)[lane + 32 * i] = o;
    }
  }
#pragma unroll
  for (int h = 0; h < NH; ++h) {
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      const int c = (lane + 32 * i) * 4;
      atomicAdd(&red[h][c], aw[h][i].x); atomicAdd(&red[h][c + 1], aw[h][i].y);
      atomicAdd(&red[h][c + 2], aw[h][i].z); atomicAdd(&red[h][c + 3], aw[h][i].w);
    }
    if (lane == 0) atomicAdd(&redb[h], ab[h]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NH * D; i += blockDim.x) atomicAdd(dw + i, (&red[0][0])[i]);
  if (threadIdx.x < NH) atomicAdd(db + threadIdx.x, redb[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------
// Auxiliary heads of MtCutLoss (utils/losses.py:180-191), one CTA per attention group (= reference batch):
//   classification: BCE(sigmoid(zc), y) mean over the S*L cells of the group, log clamped at -100
//   rerank: e = mean_{y==0} u - mean_{y==1} u + margin over the group; loss = max(e, 0)   (losses.py:127-141)
//           u = zr (MtChoopy/MtAttnCut) or softmax_L(zr) (MMOECut towers; rerank_softmax = 1)
// Writes per-group losses and the gradients dzc, dzr (scaled by the task weights and gscale).
// status[g]: bit 0 = the group has no relevant or no irrelevant document (the reference raises); bit 1 = the rerank
// hinge is active (e > 0).  While it is inactive the reference's criterion returns a constant (losses.py:141), the
// rerank head receives NO gradient and torch's Adam skips those parameters -- FusedAdam reads this bit to do the same.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) aux_heads_kernel(const float* __restrict__ zc, const float* __restrict__ zr,
                                                        const float* __restrict__ labels, int S, int L,
                                                        int rerank_softmax, int class_probs, float margin, float class_w, float rerank_w,
                                                        float gscale, float* __restrict__ probs_c,
                                                        float* __restrict__ out_r, float* __restrict__ dzc,
                                                        float* __restrict__ dzr, float* __restrict__ loss_group,
                                                        int32_t* __restrict__ status) {
  __shared__ float bc[4];
  __shared__ float part[8][5];
  const int g = blockIdx.x;
  const size_t off = size_t(g) * S * L;
  const int n = S * L;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // pass 0 (MMOECut): u = softmax over L per list, written to out_r (one warp per list)
  if (zr != nullptr && rerank_softmax) {
    for (int s = warp; s < S; s += 8) {
      const float* row = zr + off + size_t(s) * L;
      float m = -INFINITY;
      for (int j = lane; j < L; j += 32) m = fmaxf(m, row[j]);
      m = warp_max(m);
      float sum = 0.f;
      for (int j = lane; j < L; j += 32) sum += __expf(row[j] - m);
      sum = warp_sum(sum);
      const float inv = 1.f / sum;
      for (int j = lane; j < L; j += 32) out_r[off + size_t(s) * L + j] = __expf(row[j] - m) * inv;
    }
    __syncthreads();
  }
  const float* u = (zr == nullptr) ? nullptr : (rerank_softmax ? out_r + off : zr + off);
  // pass 1: group reductions
  float bce = 0.f, spos = 0.f, sneg = 0.f, npos = 0.f, nneg = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) {
    const float y = labels[off + i];
    if (zc != nullptr) {
      const float z = zc[off + i];
      const float pr = class_probs ? z : 1.f / (1.f + __expf(-z));
      if (probs_c) probs_c[off + i] = pr;
      // torch BCELoss: -(y*max(log p, -100) + (1-y)*max(log(1-p), -100))
      const float lp = fmaxf(__logf(pr), -100.f), lq = fmaxf(__logf(1.f - pr), -100.f);
      bce -= y * lp + (1.f - y) * lq;
    }
    if (u != nullptr) {
      const float v = u[i];
      if (y == 1.f) { spos += v; npos += 1.f; }
      if (y == 0.f) { sneg += v; nneg += 1.f; }
    }
  }
  float vals[5] = {bce, spos, sneg, npos, nneg};
#pragma unroll
  for (int k = 0; k < 5; ++k) vals[k] = warp_sum(vals[k]);
  if (lane == 0)
    for (int k = 0; k < 5; ++k) part[warp][k] = vals[k];
  __syncthreads();
  if (threadIdx.x == 0) {
    float t[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int w = 0; w < 8; ++w)
      for (int k = 0; k < 5; ++k) t[k] += part[w][k];
    float loss = 0.f;
    float active = 0.f, inv_pos = 0.f, inv_neg = 0.f;
    int bad = 0;
    if (zc != nullptr) loss += class_w * t[0] / float(n);
    if (u != nullptr) {
      if (t[3] == 0.f || t[4] == 0.f) {
        bad = 1;
      } else {
        const float e = t[2] / t[4] - t[1] / t[3] + margin;
        if (e > 0.f) { loss += rerank_w * e; active = 1.f; }
        inv_pos = 1.f / t[3];
        inv_neg = 1.f / t[4];
      }
    }
    loss_group[g] = loss;
    if (status) status[g] = bad | (active != 0.f ? 2 : 0);   // bit 0: degenerate group, bit 1: the rerank hinge is active
    bc[0] = active; bc[1] = inv_pos; bc[2] = inv_neg;
  }
  __syncthreads();
  const float active = bc[0], inv_pos = bc[1], inv_neg = bc[2];
  // pass 2: gradients
  if (dzc != nullptr && zc != nullptr) {
    const float sc = class_w * gscale / float(n);
    for (int i = threadIdx.x; i < n; i += 256) {
      const float z = zc[off + i];
      const float y = labels[off + i];
      if (class_probs) {  // gradient w.r.t. the probability itself: (p - y) / (p (1 - p)), the clamped-log form
        const float gp = (y == 1.f) ? (z > 3.7200759e-44f ? -1.f / z : 0.f)          // log p clamped at -100
                                    : ((1.f - z) > 3.7200759e-44f ? 1.f / (1.f - z) : 0.f);
        dzc[off + i] = sc * (y == 1.f || y == 0.f ? gp : (z - y) / fmaxf(z * (1.f - z), 1e-30f));
      } else {
        const float pr = 1.f / (1.f + __expf(-z));
        dzc[off + i] = sc * (pr - y);
      }
    }
  }
  if (dzr != nullptr && u != nullptr) {
    const float sc = rerank_w * gscale * active;
    if (!rerank_softmax) {
      for (int i = threadIdx.x; i < n; i += 256) {
        const float y = labels[off + i];
        dzr[off + i] = sc * ((y == 0.f ? inv_neg : 0.f) - (y == 1.f ? inv_pos : 0.f));
      }
    } else {
      // du -> dz through the per-list softmax: dz = u * (du - <u, du>)
      for (int s = warp; s < S; s += 8) {
        const size_t ro = off + size_t(s) * L;
        float dot = 0.f;
        for (int j = lane; j < L; j += 32) {
          const float y = labels[ro + j];
          const float du = (y == 0.f ? inv_neg : 0.f) - (y == 1.f ? inv_pos : 0.f);
          dot += out_r[ro + j] * du;
        }
        dot = warp_sum(dot);
        for (int j = lane; j < L; j += 32) {
          const float y = labels[ro + j];
          const float du = (y == 0.f ? inv_neg : 0.f) - (y == 1.f ? inv_pos : 0.f);
          dzr[ro + j] = sc * out_r[ro + j] * (du - dot);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// BiCut head loss (utils/losses.py:11-45).  u: [B, L, 2] logits (after the optional
// dropout), out = softmax over the 2 classes; idx = LAST position whose argmax is class 0 (ties -> 0),
// L if none; mask keeps positions <= idx; weights ((1-a)/r, 0) for relevant, (0, a/(1-r)) otherwise
// ('nci' metric: (0, -1/log2(j+2)) / (0, (j+1)/a)).  input_kind 1: `u` already holds probabilities and
// the gradient is w.r.t. them (API boundary).  One warp per list.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) bicut_loss_kernel(const float* __restrict__ u, const float* __restrict__ labels,
                                                         int B, int L, int input_kind, int metric_nci, float alpha,
                                                         float rr, float gscale, float* __restrict__ probs_out,
                                                         float* __restrict__ grad, float* __restrict__ loss_per_list) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const float2* ur = reinterpret_cast<const float2*>(u) + size_t(b) * L;
  const float* yr = labels + size_t(b) * L;
  // pass 1: last position predicting class 0
  int last = -1;
  for (int j = lane; j < L; j += 32) {
    const float2 v = ur[j];
    if (!(v.y > v.x)) last = j;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
  const int idx = last < 0 ? L : last;
  float loss = 0.f;
  for (int j = lane; j < L; j += 32) {
    const float2 v = ur[j];
    float o0, o1;
    if (input_kind == 0) {
      const float m = fmaxf(v.x, v.y);
      const float e0 = __expf(v.x - m), e1 = __expf(v.y - m);
      const float inv = 1.f / (e0 + e1);
      o0 = e0 * inv; o1 = e1 * inv;
    } else {
      o0 = v.x; o1 = v.y;
    }
    if (probs_out) reinterpret_cast<float2*>(probs_out)[size_t(b) * L + j] = make_float2(o0, o1);
    const float y = yr[j];
    float w0, w1;
    if (metric_nci) {
      w0 = 0.f;
      w1 = (y == 1.f) ? -1.f / log2f(float(j + 2)) : float(j + 1) / alpha;
    } else {
      w0 = (y == 1.f) ? (1.f - alpha) / rr : 0.f;
      w1 = (y == 1.f) ? 0.f : alpha / (1.f - rr);
    }
    const float mk = j <= idx ? 1.f : 0.f;
    loss += mk * (o0 * w0 + o1 * w1);
    if (grad) {
      float g0 = mk * w0, g1 = mk * w1;  // dL_b/do
      if (input_kind == 0) {             // through the 2-way softmax
        const float dot = o0 * g0 + o1 * g1;
        g0 = o0 * (g0 - dot);
        g1 = o1 * (g1 - dot);
      }
      reinterpret_cast<float2*>(grad)[size_t(b) * L + j] = make_float2(g0 * gscale, g1 * gscale);
    }
  }
  loss = warp_sum(loss);
  if (lane == 0 && loss_per_list) loss_per_list[b] = loss;
}

// ------------------------------------------------------------------------------------------
// Choopy input (models/Choopy.py:18-20, MtChoopy.py:24-25): X[b, l, :] = [score[b,l] | PE[l, 0:127]]
// and the gradient of the learned table: dPE[l, c] += sum_b dX[b, l, 1 + c].
// ------------------------------------------------------------------------------------------
// x[t, 0] = score[t], x[t, 1:128] = PE[l, :]  (models/Choopy.py:19-20).  One thread per 16-byte unit of x, grid-stride:
// the first version launched one 128-thread CTA per token (1.2 M CTAs, 1 TB/s).
__global__ void __launch_bounds__(256) choopy_embed_kernel(const float* __restrict__ score, const float* __restrict__ pe,
                                                           float* __restrict__ x, size_t T, int L) {
  const size_t n4 = T * 32, stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const size_t t = i >> 5;
    const int c = int(i & 31) * 4;
    const float* pr = pe + size_t(t % L) * 127 + c - 1;       // PE column c-1 .. c+2 (rows of 127 floats: unaligned)
    float4 v;
    v.x = c == 0 ? __ldg(score + t) : __ldg(pr);
    v.y = __ldg(pr + 1); v.z = __ldg(pr + 2); v.w = __ldg(pr + 3);
    reinterpret_cast<float4*>(x)[i] = v;
  }
}
// dPE[l, c-1] += sum over lists of dx[b, l, c]: CTA (l, slice of the lists), 4 independent loads in flight per thread
__global__ void __launch_bounds__(128) choopy_embed_bwd_kernel(const float* __restrict__ dx, float* __restrict__ dpe,
                                                               int B, int L) {
  const int l = blockIdx.x, c = threadIdx.x;
  const size_t pitch = size_t(L) * 128;
  const float* base = dx + size_t(l) * 128 + c;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int b = blockIdx.y;
  const int g = gridDim.y;
  for (; b + 3 * g < B; b += 4 * g) {
    a0 += __ldg(base + size_t(b) * pitch); a1 += __ldg(base + size_t(b + g) * pitch);
    a2 += __ldg(base + size_t(b + 2 * g) * pitch); a3 += __ldg(base + size_t(b + 3 * g) * pitch);
  }
  for (; b < B; b += g) a0 += __ldg(base + size_t(b) * pitch);
  if (c > 0) atomicAdd(dpe + size_t(l) * 127 + c - 1, (a0 + a1) + (a2 + a3));
}

// BiCut output head (models/Bicut.py:11-16): 2-class softmax of the logit planes z[0,:], z[1,:] -> o[t, 0:2]
// Train mode: nn.Dropout(p) acts on the two logits BEFORE the softmax (models/Bicut.py:14); element index = c * T + t.
__global__ void pair_softmax_fwd_kernel(const float* __restrict__ z, float* __restrict__ o, size_t T, DropCfg drop) {
  const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= T) return;
  float a = z[t], b = z[T + t];
  if (drop.thr) {
    a *= drop_factor(drop_bits(drop.seed, DROP_LOGITS, t >> 2), int(t & 3), drop.thr, drop.scale);
    b *= drop_factor(drop_bits(drop.seed, DROP_LOGITS, (T + t) >> 2), int((T + t) & 3), drop.thr, drop.scale);
  }
  const float m = fmaxf(a, b);
  const float ea = __expf(a - m), eb = __expf(b - m);
  const float inv = 1.f / (ea + eb);
  reinterpret_cast<float2*>(o)[t] = make_float2(ea * inv, eb * inv);
}
// dz[c, t] = o_c (do_c - <o, do>)
__global__ void pair_softmax_bwd_kernel(const float* __restrict__ o, const float* __restrict__ d_o, float* __restrict__ dz,
                                        size_t T, DropCfg drop) {
  const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const float2 p = reinterpret_cast<const float2*>(o)[t], g = reinterpret_cast<const float2*>(d_o)[t];
  const float dot = p.x * g.x + p.y * g.y;
  float da = p.x * (g.x - dot), db = p.y * (g.y - dot);
  if (drop.thr) {
    da *= drop_factor(drop_bits(drop.seed, DROP_LOGITS, t >> 2), int(t & 3), drop.thr, drop.scale);
    db *= drop_factor(drop_bits(drop.seed, DROP_LOGITS, (T + t) >> 2), int((T + t) & 3), drop.thr, drop.scale);
  }
  dz[t] = da;
  dz[T + t] = db;
}

// one cut_loss_kernel instantiation per (input kind, loss kind, metric)
#define RLT_K3(IK, LK, MD)                                                                                             \
  case (IK) * 8 + (LK) * 2 + (MD):                                                                                     \
    cut_loss_kernel<NI, IK, LK, MD><<<grid, 128, 0, stream>>>(in, labels, probs_out, grad, loss_per_list, nullptr, B, L, \
                                                              tau, gscale);                                            \
    break;
#define RLT_K3_CASES                                                                                                  \
  RLT_K3(0, 0, 0) RLT_K3(0, 0, 1) RLT_K3(0, 1, 0) RLT_K3(0, 1, 1) RLT_K3(0, 2, 0) RLT_K3(0, 2, 1) RLT_K3(0, 3, 0) RLT_K3(0, 3, 1) \
  RLT_K3(1, 0, 0) RLT_K3(1, 0, 1) RLT_K3(1, 1, 0) RLT_K3(1, 1, 1) RLT_K3(1, 2, 0) RLT_K3(1, 2, 1) RLT_K3(1, 3, 0) RLT_K3(1, 3, 1)

template <bool kBits>
static void eval_cut_fast_launch(const float* probs, const void* labels, int n_lists, int L, int32_t* k_out, int32_t* count_out,
                                 int32_t* nrel_out, double* f1_out, double* dcg_out, cudaStream_t stream) {
  const int grid = (n_lists + 127) / 128;
#define RLT_K4F(NP_, PF_) eval_cut_fast_kernel<NP_, kBits, PF_><<<grid, 128, 0, stream>>>(probs, labels, n_lists, L, k_out, count_out, nrel_out, f1_out, dcg_out)
  if (L <= 64) RLT_K4F(1, true);
  else if (L <= 320) RLT_K4F(5, true);
  else if (L <= 512) RLT_K4F(8, true);
  else RLT_K4F(16, false);       // 1024 positions: no second register set for the next list, occupancy hides the latency
#undef RLT_K4F
}

template <typename F>
static int dispatch_ni(int L, F&& f) {
  if (L <= 64) return f(std::integral_constant<int, 2>{});
  if (L <= 320) return f(std::integral_constant<int, 10>{});
  if (L <= 512) return f(std::integral_constant<int, 16>{});
  if (L <= 1024) return f(std::integral_constant<int, 32>{});
  return set_error(RLT_UNSUPPORTED_SHAPE, "list length %d exceeds 1024", L);
}

}  // namespace rlt

using namespace rlt;

extern "C" {

int rlt_set_dcg_tables(const float* coef32_host, const double* term64_host, int n) {
  RLT_REQUIRE(coef32_host && term64_host && n > 0 && n <= 1024, RLT_INVALID_ARG, "rlt_set_dcg_tables: n must be in [1,1024]");
  RLT_CHECK_CUDA(cudaMemcpyToSymbol(g_dcg_coef32, coef32_host, sizeof(float) * n));
  {
    float rc[1024];
    for (int i = 0; i < n; ++i) rc[i] = 1.0f / coef32_host[i];   // IEEE float32 division on the host
    RLT_CHECK_CUDA(cudaMemcpyToSymbol(g_dcg_rcoef32, rc, sizeof(float) * n));
    RLT_TRY(cut_loss_pair_set_rcoef(rc, n));
  }
  RLT_CHECK_CUDA(cudaMemcpyToSymbol(g_dcg_term64, term64_host, sizeof(double) * n));
  return RLT_OK;
}

int rlt_choopy_embed_fwd(const float* score, const float* pe, float* x, int n_lists, int seq_len, rlt_stream_t stream_) {
  RLT_REQUIRE(score && pe && x && n_lists > 0 && seq_len > 0, RLT_INVALID_ARG, "rlt_choopy_embed_fwd: bad arguments");
  const size_t T = size_t(n_lists) * seq_len;
  size_t blocks = (T * 32 + 255) / 256;
  if (blocks > size_t(num_sms()) * 16) blocks = size_t(num_sms()) * 16;
  choopy_embed_kernel<<<unsigned(blocks), 256, 0, static_cast<cudaStream_t>(stream_)>>>(score, pe, x, T, seq_len);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

int rlt_choopy_embed_bwd(const float* dx, float* dpe, int n_lists, int seq_len, rlt_stream_t stream_) {
  RLT_REQUIRE(dx && dpe && n_lists > 0 && seq_len > 0, RLT_INVALID_ARG, "rlt_choopy_embed_bwd: bad arguments");
  int gy = (n_lists + 31) / 32;
  if (gy > 16) gy = 16;     // seq_len x gy CTAs: 4800 at L = 300, each with 4 x 512 B loads in flight per warp group
  choopy_embed_bwd_kernel<<<dim3(seq_len, gy), 128, 0, static_cast<cudaStream_t>(stream_)>>>(dx, dpe, n_lists, seq_len);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

int rlt_cut_loss(const rlt_cut_loss_desc* c, const float* in, const float* labels, float* probs_out, float* grad,
                 float* loss_per_list, float* loss_out, rlt_stream_t stream_) {
  RLT_REQUIRE(c && in && labels, RLT_INVALID_ARG, "rlt_cut_loss: null pointer");
  RLT_REQUIRE(c->n_lists > 0 && c->seq_len > 0, RLT_INVALID_ARG, "rlt_cut_loss: n_lists=%d seq_len=%d", c->n_lists, c->seq_len);
  RLT_REQUIRE(c->loss_kind >= 0 && c->loss_kind <= 3 && (c->input_kind == 0 || c->input_kind == 1), RLT_INVALID_ARG,
              "rlt_cut_loss: loss_kind=%d input_kind=%d", c->loss_kind, c->input_kind);
  RLT_REQUIRE(c->loss_kind == 0 || c->tau > 0.f, RLT_INVALID_ARG, "rlt_cut_loss: tau must be positive");
  RLT_REQUIRE(loss_out == nullptr || loss_per_list != nullptr, RLT_INVALID_ARG,
              "rlt_cut_loss: loss_out needs the loss_per_list scratch");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int B = c->n_lists, L = c->seq_len;
  const int grid = (B + 3) / 4;
  RLT_REQUIRE(c->input_kind >= 0 && c->input_kind <= 1 && c->loss_kind >= 0 && c->loss_kind <= 3, RLT_INVALID_ARG,
              "rlt_cut_loss: input_kind %d / loss_kind %d out of range", c->input_kind, c->loss_kind);
  const int cfg = c->input_kind * 8 + c->loss_kind * 2 + (c->metric_dcg ? 1 : 0);
  const float tau = c->tau, gscale = c->grad_scale;
  if (c->input_kind == 0 && cut_loss_pair_ok(L, in, labels, probs_out, grad)) {      // the packed kernel (cut_loss_pair.cu)
    RLT_TRY(cut_loss_pair_launch(c, in, labels, false, probs_out, grad, loss_per_list, stream));
    RLT_CHECK_LAUNCH();
    if (loss_out != nullptr) {
      reduce_scale_kernel<<<1, 256, 0, stream>>>(loss_per_list, B, c->loss_scale, loss_out, c->accumulate_loss);
      RLT_CHECK_LAUNCH();
    }
    return RLT_OK;
  }
  auto launch = [&](auto ni) {
    constexpr int NI = decltype(ni)::value;
    switch (cfg) {
      RLT_K3_CASES
      default: break;
    }
    return RLT_OK;
  };
  RLT_TRY(dispatch_ni(L, launch));
  RLT_CHECK_LAUNCH();
  if (loss_out != nullptr) {
    reduce_scale_kernel<<<1, 256, 0, stream>>>(loss_per_list, B, c->loss_scale, loss_out, c->accumulate_loss);
    RLT_CHECK_LAUNCH();
  }
  return RLT_OK;
}

int rlt_reward_matrix(const float* labels, float* rewards, int n_lists, int seq_len, int metric_dcg, rlt_stream_t stream_) {
  RLT_REQUIRE(labels && rewards && n_lists > 0 && seq_len > 0, RLT_INVALID_ARG, "rlt_reward_matrix: bad arguments");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RLT_TRY(dispatch_ni(seq_len, [&](auto ni) {
    if (metric_dcg)
      cut_loss_kernel<decltype(ni)::value, 1, 0, 1><<<(n_lists + 3) / 4, 128, 0, stream>>>(nullptr, labels, nullptr, nullptr, nullptr,
                                                                                         rewards, n_lists, seq_len, 1.f, 1.f);
    else
      cut_loss_kernel<decltype(ni)::value, 1, 0, 0><<<(n_lists + 3) / 4, 128, 0, stream>>>(nullptr, labels, nullptr, nullptr, nullptr,
                                                                                         rewards, n_lists, seq_len, 1.f, 1.f);
    return RLT_OK;
  }));
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

int rlt_softmax_lists(const float* z, float* p, int n_lists, int seq_len, rlt_stream_t stream_) {
  RLT_REQUIRE(z && p && n_lists > 0 && seq_len > 0, RLT_INVALID_ARG, "rlt_softmax_lists: bad arguments");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RLT_TRY(dispatch_ni(seq_len, [&](auto ni) {
    softmax_lists_kernel<decltype(ni)::value><<<(n_lists + 3) / 4, 128, 0, stream>>>(z, p, n_lists, seq_len);
    return RLT_OK;
  }));
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

int rlt_softmax_lists_bwd(const float* p, const float* dp, float* dz, int n_lists, int seq_len, rlt_stream_t stream_) {
  RLT_REQUIRE(p && dp && dz && n_lists > 0 && seq_len > 0, RLT_INVALID_ARG, "rlt_softmax_lists_bwd: bad arguments");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RLT_TRY(dispatch_ni(seq_len, [&](auto ni) {
    softmax_lists_bwd_kernel<decltype(ni)::value><<<(n_lists + 3) / 4, 128, 0, stream>>>(p, dp, dz, n_lists, seq_len);
    return RLT_OK;
  }));
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

int rlt_eval_cut(const float* probs, const float* labels, int n_lists, int seq_len, int mode, int32_t* k_out,
                 int32_t* count_out, int32_t* nrel_out, double* f1_out, double* dcg_out, rlt_stream_t stream_) {
  RLT_REQUIRE(probs && labels && n_lists > 0 && seq_len > 0, RLT_INVALID_ARG, "rlt_eval_cut: bad arguments");
  RLT_REQUIRE(seq_len <= 1024, RLT_UNSUPPORTED_SHAPE, "rlt_eval_cut: seq_len %d exceeds 1024", seq_len);
  RLT_REQUIRE(mode == 0 || mode == 1, RLT_INVALID_ARG, "rlt_eval_cut: mode must be 0 (argmax cut) or 1 (BiCut rule)");
  const int grid = (n_lists + 127) / 128;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (mode == 0 && seq_len % 2 == 0 && ((reinterpret_cast<uintptr_t>(probs) | reinterpret_cast<uintptr_t>(labels)) & 7u) == 0) {
    eval_cut_fast_launch<false>(probs, labels, n_lists, seq_len, k_out, count_out, nrel_out, f1_out, dcg_out, stream);
    RLT_CHECK_LAUNCH();
    return RLT_OK;
  }
  RLT_TRY(dispatch_ni(seq_len, [&](auto ni) {
    eval_cut_kernel<decltype(ni)::value><<<grid, 128, 0, stream>>>(probs, labels, nullptr, nullptr, n_lists, seq_len, mode, k_out,
                                                                   count_out, nrel_out, f1_out, dcg_out);
    return RLT_OK;
  }));
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

int rlt_eval_cut_bits(const float* probs, const uint32_t* label_bits, int n_lists, int seq_len, int32_t* k_out,
                      int32_t* count_out, int32_t* nrel_out, double* f1_out, double* dcg_out, rlt_stream_t stream_) {
  RLT_REQUIRE(probs && label_bits && n_lists > 0 && seq_len > 0, RLT_INVALID_ARG, "rlt_eval_cut_bits: bad arguments");
  RLT_REQUIRE(seq_len <= 1024, RLT_UNSUPPORTED_SHAPE, "rlt_eval_cut_bits: seq_len %d exceeds 1024", seq_len);
  RLT_REQUIRE(seq_len % 2 == 0 && (reinterpret_cast<uintptr_t>(probs) & 7u) == 0, RLT_UNSUPPORTED_SHAPE,
              "rlt_eval_cut_bits: seq_len %d must be even and probs 8-byte aligned", seq_len);
  eval_cut_fast_launch<true>(probs, label_bits, n_lists, seq_len, k_out, count_out, nrel_out, f1_out, dcg_out,
                             static_cast<cudaStream_t>(stream_));
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

int rlt_eval_given_k(const float* labels, const int32_t* k_in, const int32_t* pyint_in, int n_lists, int seq_len,
                     int32_t* count_out, int32_t* nrel_out, double* f1_out, double* dcg_out, rlt_stream_t stream_) {
  RLT_REQUIRE(labels && k_in && n_lists > 0 && seq_len > 0, RLT_INVALID_ARG, "rlt_eval_given_k: bad arguments");
  RLT_REQUIRE(seq_len <= 1024, RLT_UNSUPPORTED_SHAPE, "rlt_eval_given_k: seq_len %d exceeds 1024", seq_len);
  const int grid = (n_lists + 127) / 128;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RLT_TRY(dispatch_ni(seq_len, [&](auto ni) {
    eval_cut_kernel<decltype(ni)::value><<<grid, 128, 0, stream>>>(nullptr, labels, k_in, pyint_in, n_lists, seq_len, 0, nullptr,
                                                                   count_out, nrel_out, f1_out, dcg_out);
    return RLT_OK;
  }));
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

int rlt_head_dots_fwd(const float* x, const float* w, const float* bias, float* z, int n_tokens, int d, int n_heads,
                      rlt_stream_t stream_) {
  RLT_REQUIRE(x && w && bias && z && n_tokens > 0, RLT_INVALID_ARG, "rlt_head_dots_fwd: bad arguments");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int grid = (n_tokens + 7) / 8;
#define RLT_HD(D, NH) head_dots_fwd_kernel<D, NH><<<grid, 256, 0, stream>>>(x, w, bias, z, n_tokens)
  if (d == 128 && n_heads == 1) RLT_HD(128, 1);
  else if (d == 128 && n_heads == 2) RLT_HD(128, 2);
  else if (d == 128 && n_heads == 3) RLT_HD(128, 3);
  else if (d == 256 && n_heads == 1) RLT_HD(256, 1);
  else if (d == 256 && n_heads == 2) RLT_HD(256, 2);
  else if (d == 256 && n_heads == 3) RLT_HD(256, 3);
  else return set_error(RLT_UNSUPPORTED_SHAPE, "rlt_head_dots: d=%d n_heads=%d unsupported", d, n_heads);
#undef RLT_HD
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

int rlt_head_dots_bwd(const float* x, const float* w, const float* dz, float* dx, float* dw, float* db, int n_tokens,
                      int d, int n_heads, int accumulate_dx, int relu_gate, rlt_stream_t stream_) {
  RLT_REQUIRE(x && w && dz && dx && dw && db && n_tokens > 0, RLT_INVALID_ARG, "rlt_head_dots_bwd: bad arguments");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int grid = (n_tokens + 63) / 64;
  const int cap = num_sms() * 4;
  if (grid > cap) grid = cap;
#define RLT_HB(D, NH) \
  head_dots_bwd_kernel<D, NH><<<grid, 256, 0, stream>>>(x, w, dz, dx, dw, db, n_tokens, accumulate_dx, relu_gate)
  if (d == 128 && n_heads == 1) RLT_HB(128, 1);
  else if (d == 128 && n_heads == 2) RLT_HB(128, 2);
  else if (d == 128 && n_heads == 3) RLT_HB(128, 3);
  else if (d == 256 && n_heads == 1) RLT_HB(256, 1);
  else if (d == 256 && n_heads == 2) RLT_HB(256, 2);
  else if (d == 256 && n_heads == 3) RLT_HB(256, 3);
  else return set_error(RLT_UNSUPPORTED_SHAPE, "rlt_head_dots: d=%d n_heads=%d unsupported", d, n_heads);
#undef RLT_HB
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

int rlt_pair_softmax_fwd(const float* z, float* o, size_t n_tokens, float dropout_p, uint64_t dropout_seed,
                         rlt_stream_t stream_) {
  RLT_REQUIRE(z && o && n_tokens > 0 && dropout_p >= 0.f && dropout_p < 1.f, RLT_INVALID_ARG, "rlt_pair_softmax_fwd: bad arguments");
  const DropCfg drop = dropout_p > 0.f ? make_drop(dropout_p, dropout_seed) : DropCfg{0, 0, 1.f};
  pair_softmax_fwd_kernel<<<unsigned((n_tokens + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(z, o, n_tokens, drop);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}
int rlt_pair_softmax_bwd(const float* o, const float* d_o, float* dz, size_t n_tokens, float dropout_p,
                         uint64_t dropout_seed, rlt_stream_t stream_) {
  RLT_REQUIRE(o && d_o && dz && n_tokens > 0 && dropout_p >= 0.f && dropout_p < 1.f, RLT_INVALID_ARG, "rlt_pair_softmax_bwd: bad arguments");
  const DropCfg drop = dropout_p > 0.f ? make_drop(dropout_p, dropout_seed) : DropCfg{0, 0, 1.f};
  pair_softmax_bwd_kernel<<<unsigned((n_tokens + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(o, d_o, dz, n_tokens, drop);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

int rlt_aux_heads_loss(const rlt_aux_loss_desc* a, const float* zc, const float* zr, const float* labels,
                       float* probs_c, float* out_r, float* dzc, float* dzr, float* loss_group, int32_t* status,
                       float* loss_out, rlt_stream_t stream_) {
  RLT_REQUIRE(a && labels && loss_group, RLT_INVALID_ARG, "rlt_aux_heads_loss: null pointer");
  RLT_REQUIRE(a->n_groups > 0 && a->group_size > 0 && a->seq_len > 0, RLT_INVALID_ARG, "rlt_aux_heads_loss: bad sizes");
  RLT_REQUIRE(!(zr && a->rerank_softmax && !out_r), RLT_INVALID_ARG, "rlt_aux_heads_loss: softmax rerank head needs out_r");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  aux_heads_kernel<<<a->n_groups, 256, 0, stream>>>(zc, zr, labels, a->group_size, a->seq_len, a->rerank_softmax,
                                                    a->class_probs, a->margin, a->class_weight, a->rerank_weight, a->grad_scale, probs_c,
                                                    out_r, dzc, dzr, loss_group, status);
  RLT_CHECK_LAUNCH();
  if (loss_out != nullptr) {
    reduce_scale_kernel<<<1, 256, 0, stream>>>(loss_group, a->n_groups, a->loss_scale, loss_out, a->accumulate_loss);
    RLT_CHECK_LAUNCH();
  }
  return RLT_OK;
}

int rlt_bicut_loss(const rlt_bicut_loss_desc* c, const float* u, const float* labels, float* probs_out, float* grad,
                   float* loss_per_list, float* loss_out, rlt_stream_t stream_) {
  RLT_REQUIRE(c && u && labels, RLT_INVALID_ARG, "rlt_bicut_loss: null pointer");
  RLT_REQUIRE(c->n_lists > 0 && c->seq_len > 0, RLT_INVALID_ARG, "rlt_bicut_loss: bad sizes");
  RLT_REQUIRE(loss_out == nullptr || loss_per_list != nullptr, RLT_INVALID_ARG, "rlt_bicut_loss: loss_out needs loss_per_list");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  bicut_loss_kernel<<<(c->n_lists + 3) / 4, 128, 0, stream>>>(u, labels, c->n_lists, c->seq_len, c->input_kind,
                                                              c->metric_nci, c->alpha, c->r, c->grad_scale, probs_out,
                                                              grad, loss_per_list);
  RLT_CHECK_LAUNCH();
  if (loss_out != nullptr) {
    reduce_scale_kernel<<<1, 256, 0, stream>>>(loss_per_list, c->n_lists, c->loss_scale, loss_out, c->accumulate_loss);
    RLT_CHECK_LAUNCH();
  }
  return RLT_OK;
}

}  // extern "C"
