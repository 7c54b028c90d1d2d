// Transformer encoder layer (post-norm, ReLU) for the truncation models: forward and backward.
//
// Reference semantics (torch.nn.TransformerEncoderLayer as built at models/Choopy.py:11,
// AttnCut.py:9, MtChoopy.py:11, MtAttnCut.py:9, MMOECut.py:9 — WITHOUT batch_first): the input is
// [B, L, d]; attention runs over dim 0, i.e. across the S = B lists of one forward call ("group"),
// independently for every position l and head h.  This file keeps that behaviour (attend_axis = 0) and
// processes G independent groups per call.
//
// Kernels here: row LayerNorm fwd / bwd (one warp per token), cross-list attention fwd / bwd (one CTA
// per (group, position, head)), column sums.  The dense contractions go through gemm_tn / gemm_nn /
// gemm_dw (tcgen05).  Layer sequencing:
//   fwd : qkv = x Win^T + b        -> attn -> o
//         u1  = x + o Wo^T + bo    -> y = LN1(u1)
//         h   = relu(y W1^T + b1)
//         u2  = y + h W2^T + b2    -> out = LN2(u2)
//   bwd : the chain of SURVEY.md A.2 in reverse; weight grads are token contractions (gemm_dw).
#include <math.h>

#include "attention_mma.cuh"
#include "common.h"
#include "gemm_tc.cuh"

namespace rlt {

// ------------------------------------------------------------------------------------------
// LayerNorm over the last dim (d = 32 * VPL * 4... i.e. d in {128, 256}), one warp per row.
// ------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ u, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, float* __restrict__ y,
                                                     float* __restrict__ stats, int T, float eps,
                                                     __half* __restrict__ y16 /* optional fp16 copy of y */) {
  constexpr int V4 = D / 128;  // float4 per lane
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= T) return;
  const float4* src = reinterpret_cast<const float4*>(u + size_t(row) * D);
  float4 v[V4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V4; ++i) {
    v[i] = src[lane + 32 * i];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mu = s * (1.f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V4; ++i) {
    v[i].x -= mu; v[i].y -= mu; v[i].z -= mu; v[i].w -= mu;
    q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q * (1.f / D) + eps);
  float4* dst = reinterpret_cast<float4*>(y + size_t(row) * D);
#pragma unroll
  for (int i = 0; i < V4; ++i) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * i);
    const float4 r = make_float4(v[i].x * rstd * g.x + b.x, v[i].y * rstd * g.y + b.y,
                                 v[i].z * rstd * g.z + b.z, v[i].w * rstd * g.w + b.w);
    dst[lane + 32 * i] = r;
    if (y16 != nullptr) {
      const __half2 lo = __floats2half2_rn(r.x, r.y), hi = __floats2half2_rn(r.z, r.w);
      reinterpret_cast<uint2*>(y16 + size_t(row) * D)[lane + 32 * i] =
          make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    }
  }
  if (stats != nullptr && lane == 0) {
    stats[2 * size_t(row)] = mu;
    stats[2 * size_t(row) + 1] = rstd;
  }
}

// du = LN backward of dy; dgamma += sum dy*xhat; dbeta += sum dy; dbias_prev += sum du (bias of the
// Linear whose output fed the residual sum).  Persistent CTAs, per-lane column accumulators.
template <int D>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ u,
                                                     const float* __restrict__ stats,
                                                     const float* __restrict__ gamma, float* __restrict__ du,
                                                     float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                     float* __restrict__ dbias_prev, int T,
                                                     unsigned int* __restrict__ amax_out /* optional: max |du| (float bits) */) {
  constexpr int V4 = D / 128;
  __shared__ float red[3][D];
  float amax = 0.f;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) (&red[0][0])[i] = 0.f;
  __syncthreads();
  float4 ag[V4], ab[V4], ad[V4];
  float4 g[V4];
#pragma unroll
  for (int i = 0; i < V4; ++i) {
    ag[i] = ab[i] = ad[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    g[i] = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
  }
  for (int row = blockIdx.x * nwarps + warp; row < T; row += gridDim.x * nwarps) {
    const float mu = stats[2 * size_t(row)], rstd = stats[2 * size_t(row) + 1];
    const float4* pu = reinterpret_cast<const float4*>(u + size_t(row) * D);
    const float4* pd = reinterpret_cast<const float4*>(dy + size_t(row) * D);
    float4 xh[V4], dg[V4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      const float4 a = pu[lane + 32 * i], d = pd[lane + 32 * i];
      xh[i] = make_float4((a.x - mu) * rstd, (a.y - mu) * rstd, (a.z - mu) * rstd, (a.w - mu) * rstd);
      dg[i] = make_float4(d.x * g[i].x, d.y * g[i].y, d.z * g[i].z, d.w * g[i].w);
      s1 += (dg[i].x + dg[i].y) + (dg[i].z + dg[i].w);
      s2 += (dg[i].x * xh[i].x + dg[i].y * xh[i].y) + (dg[i].z * xh[i].z + dg[i].w * xh[i].w);
      ag[i].x += d.x * xh[i].x; ag[i].y += d.y * xh[i].y; ag[i].z += d.z * xh[i].z; ag[i].w += d.w * xh[i].w;
      ab[i].x += d.x; ab[i].y += d.y; ab[i].z += d.z; ab[i].w += d.w;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float m1 = s1 * (1.f / D), m2 = s2 * (1.f / D);
    float4* po = reinterpret_cast<float4*>(du + size_t(row) * D);
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      const float4 r = make_float4(rstd * (dg[i].x - m1 - xh[i].x * m2), rstd * (dg[i].y - m1 - xh[i].y * m2),
                                   rstd * (dg[i].z - m1 - xh[i].z * m2), rstd * (dg[i].w - m1 - xh[i].w * m2));
      po[lane + 32 * i] = r;
      ad[i].x += r.x; ad[i].y += r.y; ad[i].z += r.z; ad[i].w += r.w;
      amax = fmaxf(amax, fmaxf(fmaxf(fabsf(r.x), fabsf(r.y)), fmaxf(fabsf(r.z), fabsf(r.w))));
    }
  }
  if (amax_out != nullptr) {   // the fp16 scale of the FFN gradient branch comes from max |dU2|: no separate pass over du
#pragma unroll
    for (int o = 16; o; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if (lane == 0 && amax > 0.f) atomicMax(amax_out, __float_as_uint(amax));
  }
#pragma unroll
  for (int i = 0; i < V4; ++i) {
    const int c = (lane + 32 * i) * 4;
    atomicAdd(&red[0][c], ag[i].x); atomicAdd(&red[0][c + 1], ag[i].y);
    atomicAdd(&red[0][c + 2], ag[i].z); atomicAdd(&red[0][c + 3], ag[i].w);
    atomicAdd(&red[1][c], ab[i].x); atomicAdd(&red[1][c + 1], ab[i].y);
    atomicAdd(&red[1][c + 2], ab[i].z); atomicAdd(&red[1][c + 3], ab[i].w);
    atomicAdd(&red[2][c], ad[i].x); atomicAdd(&red[2][c + 1], ad[i].y);
    atomicAdd(&red[2][c + 2], ad[i].z); atomicAdd(&red[2][c + 3], ad[i].w);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    if (dgamma) atomicAdd(dgamma + c, red[0][c]);
    if (dbeta) atomicAdd(dbeta + c, red[1][c]);
    if (dbias_prev) atomicAdd(dbias_prev + c, red[2][c]);
  }
}

// out[c] += sum_t src[t, c]   (C a multiple of 4)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ src, float* __restrict__ out, int T,
                                                     int C, DropCfg drop, uint32_t site) {
  // thread -> one float4 column group; rows strided over (blockIdx.y, threadIdx.y)
  const int c4 = blockIdx.x * 32 + threadIdx.x;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c4 * 4 < C) {
    for (int t = blockIdx.y * blockDim.y + threadIdx.y; t < T; t += gridDim.y * blockDim.y) {
      float4 v = *reinterpret_cast<const float4*>(src + size_t(t) * C + c4 * 4);
      if (drop.thr) {   // column sums of the dropout-masked tensor (bias gradient of a Linear that sits inside a dropout)
        const uint64_t bits = drop_bits(drop.seed, site, (size_t(t) * C + c4 * 4) >> 2);
        v.x *= drop_factor(bits, 0, drop.thr, drop.scale); v.y *= drop_factor(bits, 1, drop.thr, drop.scale);
        v.z *= drop_factor(bits, 2, drop.thr, drop.scale); v.w *= drop_factor(bits, 3, drop.thr, drop.scale);
      }
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  __shared__ float4 red[8][32];
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c4 * 4 < C) {
    for (int i = 1; i < 8; ++i) {
      const float4 v = red[i][threadIdx.x];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    atomicAdd(out + c4 * 4, acc.x); atomicAdd(out + c4 * 4 + 1, acc.y);
    atomicAdd(out + c4 * 4 + 2, acc.z); atomicAdd(out + c4 * 4 + 3, acc.w);
  }
}

int colsum_masked(const float* src, float* out, int T, int C, DropCfg drop, uint32_t site, cudaStream_t stream) {
  RLT_REQUIRE(C % 4 == 0, RLT_UNSUPPORTED_SHAPE, "colsum: C=%d must be a multiple of 4", C);
  int gy = (T + 8 * 64 - 1) / (8 * 64);
  if (gy < 1) gy = 1;
  const int maxy = num_sms() * 4;
  if (gy > maxy) gy = maxy;
  colsum_kernel<<<dim3((C / 4 + 31) / 32, gy), dim3(32, 8), 0, stream>>>(src, out, T, C, drop, site);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}
int colsum(const float* src, float* out, int T, int C, cudaStream_t stream) {
  RLT_REQUIRE(C % 4 == 0, RLT_UNSUPPORTED_SHAPE, "colsum: C=%d must be a multiple of 4", C);
  int gy = (T + 8 * 64 - 1) / (8 * 64);
  if (gy < 1) gy = 1;
  const int maxy = num_sms() * 4;
  if (gy > maxy) gy = maxy;
  colsum_kernel<<<dim3((C / 4 + 31) / 32, gy), dim3(32, 8), 0, stream>>>(src, out, T, C, DropCfg{0, 0, 1.f}, 0);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

// ------------------------------------------------------------------------------------------
// Cross-list attention.  One CTA per (position l, group g, head h): the S lists of the group attend to
// each other at that position.  qkv is [T, 3d] (q | k | v), token t = (g*S + s)*L + l.
// Two threads per query row split the keys; online softmax in registers; lse saved for backward.
// ------------------------------------------------------------------------------------------
// attn_mask(): keep-and-scale factor of attention probability (query i, key j) of a work item -- the same element
// numbering as the tensor-core kernels (attention_mma.cuh) and the rlt_dropout_mask test hook.
__device__ __forceinline__ float attn_mask(const DropCfg& drop, uint64_t item, int S, int i, int j) {
  const uint64_t e = (item * uint64_t(S) + uint64_t(i)) * uint64_t(S) + uint64_t(j & ~1);
  return drop_factor(drop_bits(drop.seed, DROP_ATTN, e), j & 1, drop.thr, drop.scale);
}

template <int DH, bool kDrop>
__global__ void __launch_bounds__(128) attn_lists_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ o,
                                                             float* __restrict__ lse, int S, int L, int d,
                                                             int n_head, float scale, DropCfg drop, int group_tokens) {
  // Token of attended element s of work item (l, g): g * group_tokens + l + s * L, where the caller passes
  //   attend across lists (reference):  S = lists per group, L = seq_len (stride between lists), group_tokens = S * seq_len,
  //                                     gridDim.x = seq_len positions;
  //   attend within a list:             S = seq_len, L = 1 (consecutive tokens), group_tokens = seq_len, gridDim.x = 1.
  extern __shared__ float sm[];
  float* sK = sm;                 // [S][DH+1]
  float* sV = sm + S * (DH + 1);  // [S][DH+1]
  const int l = blockIdx.x, g = blockIdx.y, h = blockIdx.z;
  const uint64_t item = (uint64_t(g) * gridDim.x + l) * n_head + h;   // work-item number of the dropout hash (head fastest)
  const size_t tok0 = size_t(g) * group_tokens + l;  // token of element s: tok0 + s*L
  const int ld = 3 * d;
  for (int i = threadIdx.x; i < S * (DH / 4); i += blockDim.x) {
    const int s = i / (DH / 4), c = (i % (DH / 4)) * 4;
    const float* base = qkv + (tok0 + size_t(s) * L) * ld + h * DH + c;
    const float4 k4 = *reinterpret_cast<const float4*>(base + d);
    const float4 v4 = *reinterpret_cast<const float4*>(base + 2 * d);
    float* pk = sK + s * (DH + 1) + c;
    float* pv = sV + s * (DH + 1) + c;
    pk[0] = k4.x; pk[1] = k4.y; pk[2] = k4.z; pk[3] = k4.w;
    pv[0] = v4.x; pv[1] = v4.y; pv[2] = v4.z; pv[3] = v4.w;
  }
  __syncthreads();
  const int half = threadIdx.x & 1;
  const int jmid = (S + 1) / 2;
  const int j0 = half ? jmid : 0, j1 = half ? S : jmid;
  for (int i0 = 0; i0 < S; i0 += 64) {
    const int i = i0 + (threadIdx.x >> 1);
    const bool active = i < S;
    const int ii = active ? i : 0;
    float q[DH], acc[DH];
    const float* qp = qkv + (tok0 + size_t(ii) * L) * ld + h * DH;
#pragma unroll
    for (int c = 0; c < DH; c += 4) {
      const float4 t4 = *reinterpret_cast<const float4*>(qp + c);
      q[c] = t4.x * scale; q[c + 1] = t4.y * scale; q[c + 2] = t4.z * scale; q[c + 3] = t4.w * scale;
    }
#pragma unroll
    for (int c = 0; c < DH; ++c) acc[c] = 0.f;
    float m = -INFINITY, den = 0.f;
    for (int j = j0; j < j1; ++j) {
      const float* kr = sK + j * (DH + 1);
      float sc = 0.f;
#pragma unroll
      for (int c = 0; c < DH; ++c) sc = fmaf(q[c], kr[c], sc);
      const float mn = fmaxf(m, sc);
      const float corr = __expf(m - mn);  // exp(-inf) = 0 on the first key
      const float p = __expf(sc - mn);
      den = den * corr + p;                                   // the row sum stays undropped
      const float pd = kDrop ? p * attn_mask(drop, item, S, ii, j) : p;
      const float* vr = sV + j * (DH + 1);
#pragma unroll
      for (int c = 0; c < DH; ++c) acc[c] = fmaf(acc[c], corr, pd * vr[c]);
      m = mn;
    }
    // merge the two halves of the row (partner = lane ^ 1)
    const float m2 = __shfl_xor_sync(0xffffffffu, m, 1);
    const float d2 = __shfl_xor_sync(0xffffffffu, den, 1);
    const float mm = fmaxf(m, m2);
    const float c1 = (m == -INFINITY) ? 0.f : __expf(m - mm);
    const float c2 = (m2 == -INFINITY) ? 0.f : __expf(m2 - mm);
    const float dtot = den * c1 + d2 * c2;
    const float inv = 1.f / dtot;
    float* op = o + (tok0 + size_t(ii) * L) * d + h * DH;
#pragma unroll
    for (int c = 0; c < DH; ++c) {
      const float other = __shfl_xor_sync(0xffffffffu, acc[c], 1);
      acc[c] = (acc[c] * c1 + other * c2) * inv;
    }
    if (active) {
      // each of the two threads stores half of the head's columns
#pragma unroll
      for (int c = 0; c < DH / 2; c += 4) {
        const int cc = half * (DH / 2) + c;
        *reinterpret_cast<float4*>(op + cc) = make_float4(acc[cc], acc[cc + 1], acc[cc + 2], acc[cc + 3]);
      }
      if (half == 0 && lse != nullptr) lse[(tok0 + size_t(ii) * L) * n_head + h] = mm + __logf(dtot);
    }
  }
}

// Backward: dqkv from (qkv, o, lse, do).  Phase A: thread pair per query row -> dQ.  Phase B: thread pair
// per key row -> dK, dV (scores recomputed; no atomics).
template <int DH, bool kDrop>
__global__ void __launch_bounds__(128) attn_lists_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ o,
                                                             const float* __restrict__ lse, const float* __restrict__ d_o,
                                                             float* __restrict__ dqkv, int S, int L, int d, int n_head,
                                                             float scale, DropCfg drop, int group_tokens) {
  extern __shared__ float sm[];
  constexpr int P = DH + 1;
  float* sQ = sm;            // [S][P]  (q * scale)
  float* sK = sQ + S * P;
  float* sV = sK + S * P;
  float* sG = sV + S * P;    // dO
  float* sL = sG + S * P;    // lse  [S]
  float* sD = sL + S;        // D_i = dO_i . O_i  [S]
  const int l = blockIdx.x, g = blockIdx.y, h = blockIdx.z;
  const uint64_t item = (uint64_t(g) * gridDim.x + l) * n_head + h;
  const size_t tok0 = size_t(g) * group_tokens + l;       // see attn_lists_fwd_kernel
  const int ld = 3 * d;
  for (int i = threadIdx.x; i < S * (DH / 4); i += blockDim.x) {
    const int s = i / (DH / 4), c = (i % (DH / 4)) * 4;
    const size_t tok = tok0 + size_t(s) * L;
    const float* base = qkv + tok * ld + h * DH + c;
    const float4 q4 = *reinterpret_cast<const float4*>(base);
    const float4 k4 = *reinterpret_cast<const float4*>(base + d);
    const float4 v4 = *reinterpret_cast<const float4*>(base + 2 * d);
    const float4 g4 = *reinterpret_cast<const float4*>(d_o + tok * d + h * DH + c);
    float* p;
    p = sQ + s * P + c; p[0] = q4.x * scale; p[1] = q4.y * scale; p[2] = q4.z * scale; p[3] = q4.w * scale;
    p = sK + s * P + c; p[0] = k4.x; p[1] = k4.y; p[2] = k4.z; p[3] = k4.w;
    p = sV + s * P + c; p[0] = v4.x; p[1] = v4.y; p[2] = v4.z; p[3] = v4.w;
    p = sG + s * P + c; p[0] = g4.x; p[1] = g4.y; p[2] = g4.z; p[3] = g4.w;
  }
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const size_t tok = tok0 + size_t(s) * L;
    sL[s] = lse[tok * n_head + h];
    const float* po = o + tok * d + h * DH;
    const float* pg = d_o + tok * d + h * DH;
    float acc = 0.f;
    for (int c = 0; c < DH; ++c) acc = fmaf(po[c], pg[c], acc);
    sD[s] = acc;
  }
  __syncthreads();
  const int half = threadIdx.x & 1;
  const int mid = (S + 1) / 2;
  const int b0 = half ? mid : 0, b1 = half ? S : mid;
  // ---- phase A: dQ_i = scale * sum_j dS_ij K_j
  for (int i0 = 0; i0 < S; i0 += 64) {
    const int i = i0 + (threadIdx.x >> 1);
    const bool active = i < S;
    const int ii = active ? i : 0;
    float acc[DH];
#pragma unroll
    for (int c = 0; c < DH; ++c) acc[c] = 0.f;
    const float* qr = sQ + ii * P;
    const float* gr = sG + ii * P;
    const float li = sL[ii], Di = sD[ii];
    for (int j = b0; j < b1; ++j) {
      const float* kr = sK + j * P;
      const float* vr = sV + j * P;
      float sc = 0.f, dp = 0.f;
#pragma unroll
      for (int c = 0; c < DH; ++c) { sc = fmaf(qr[c], kr[c], sc); dp = fmaf(gr[c], vr[c], dp); }
      if (kDrop) dp *= attn_mask(drop, item, S, ii, j);      // d(attn) = mask/(1-p) * d(dropped attn)
      const float ds = __expf(sc - li) * (dp - Di);
#pragma unroll
      for (int c = 0; c < DH; ++c) acc[c] = fmaf(ds, kr[c], acc[c]);
    }
    float* out = dqkv + (tok0 + size_t(ii) * L) * ld + h * DH;
#pragma unroll
    for (int c = 0; c < DH; ++c) acc[c] = (acc[c] + __shfl_xor_sync(0xffffffffu, acc[c], 1)) * scale;
    if (active) {
#pragma unroll
      for (int c = 0; c < DH / 2; c += 4) {
        const int cc = half * (DH / 2) + c;
        *reinterpret_cast<float4*>(out + cc) = make_float4(acc[cc], acc[cc + 1], acc[cc + 2], acc[cc + 3]);
      }
    }
  }
  // ---- phase B: dK_j = sum_i dS_ij (scale Q_i) ; dV_j = sum_i P_ij dO_i
  for (int j0 = 0; j0 < S; j0 += 64) {
    const int j = j0 + (threadIdx.x >> 1);
    const bool active = j < S;
    const int jj = active ? j : 0;
    float ak[DH], av[DH];
#pragma unroll
    for (int c = 0; c < DH; ++c) { ak[c] = 0.f; av[c] = 0.f; }
    const float* kr = sK + jj * P;
    const float* vr = sV + jj * P;
    for (int i = b0; i < b1; ++i) {
      const float* qr = sQ + i * P;
      const float* gr = sG + i * P;
      float sc = 0.f, dp = 0.f;
#pragma unroll
      for (int c = 0; c < DH; ++c) { sc = fmaf(qr[c], kr[c], sc); dp = fmaf(gr[c], vr[c], dp); }
      const float p = __expf(sc - sL[i]);
      const float mk = kDrop ? attn_mask(drop, item, S, i, jj) : 1.f;
      const float ds = p * (dp * mk - sD[i]);
      const float pd = p * mk;                                // dV sees the dropped probabilities
#pragma unroll
      for (int c = 0; c < DH; ++c) { ak[c] = fmaf(ds, qr[c], ak[c]); av[c] = fmaf(pd, gr[c], av[c]); }
    }
    float* outk = dqkv + (tok0 + size_t(jj) * L) * ld + d + h * DH;
    float* outv = outk + d;
#pragma unroll
    for (int c = 0; c < DH; ++c) {
      ak[c] += __shfl_xor_sync(0xffffffffu, ak[c], 1);
      av[c] += __shfl_xor_sync(0xffffffffu, av[c], 1);
    }
    if (active) {
#pragma unroll
      for (int c = 0; c < DH / 2; c += 4) {
        const int cc = half * (DH / 2) + c;
        *reinterpret_cast<float4*>(outk + cc) = make_float4(ak[cc], ak[cc + 1], ak[cc + 2], ak[cc + 3]);
        *reinterpret_cast<float4*>(outv + cc) = make_float4(av[cc], av[cc + 1], av[cc + 2], av[cc + 3]);
      }
    }
  }
}

// persistent pipelined kernels: as many CTAs as are co-resident (occupancy API), each walking items with stride gridDim.x
template <int DH, int NT, bool kDrop, bool kFull>
static int attention_fwd_launch(const float* qkv, float* o, float* lse, int G, int S, int L, int d, int n_head, float scale,
                                cudaStream_t stream, DropCfg drop) {
  const size_t smem = size_t(2) * 3 * NT * 8 * (DH + 4) * sizeof(float);
  static DeviceOnce attr_set;
  if (attr_set.first()) {
    RLT_CHECK_CUDA(cudaFuncSetAttribute(attn_lists_fwd_pipe_kernel<DH, NT, kDrop, kFull>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int per_sm_q = 0;
    RLT_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_q, attn_lists_fwd_pipe_kernel<DH, NT, kDrop, kFull>, 128, smem));
    attr_set.value[attr_set.dev()] = per_sm_q < 1 ? 1 : per_sm_q;
  }
  const long long n_items = (long long)G * L * n_head;
  RLT_REQUIRE(n_items < (1ll << 31) && size_t(S) * L * 3 * d < (size_t(1) << 32), RLT_UNSUPPORTED_SHAPE,
              "attention: %lld work items / group span exceed the 32-bit index range", n_items);
  long long grid = (long long)num_sms() * attr_set.value[attr_set.dev()];
  if (grid > n_items) grid = n_items;
  time_begin(TAG_ATTN_FWD, stream);
  attn_lists_fwd_pipe_kernel<DH, NT, kDrop, kFull><<<int(grid), 128, smem, stream>>>(qkv, o, lse, S, L, d, n_head, scale, int(n_items), drop);
  time_end(TAG_ATTN_FWD, stream);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}
template <int DH, int NT>
static int attention_fwd_mma(const float* qkv, float* o, float* lse, int G, int S, int L, int d, int n_head, float scale,
                             cudaStream_t stream, DropCfg drop) {
  if (S == NT * 8)
    return drop.thr ? attention_fwd_launch<DH, NT, true, true>(qkv, o, lse, G, S, L, d, n_head, scale, stream, drop)
                    : attention_fwd_launch<DH, NT, false, true>(qkv, o, lse, G, S, L, d, n_head, scale, stream, drop);
  return drop.thr ? attention_fwd_launch<DH, NT, true, false>(qkv, o, lse, G, S, L, d, n_head, scale, stream, drop)
                  : attention_fwd_launch<DH, NT, false, false>(qkv, o, lse, G, S, L, d, n_head, scale, stream, drop);
}
template <int DH, int NT, bool kDrop, bool kFull>
static int attention_bwd_launch(const float* qkv, const float* lse, const float* d_o, float* dqkv, int G, int S, int L, int d,
                                int n_head, float scale, cudaStream_t stream, DropCfg drop) {
  // double-buffered staging when two buffers still leave room for two CTAs per SM, else one buffer
  constexpr size_t kBuf = (4 * NT * 8 * (DH + 4) + NT * 8) * sizeof(float);
  constexpr size_t kPS = size_t(2) * NT * 8 * (NT * 8) * sizeof(float);
  constexpr int kBufs = (2 * kBuf + kPS <= 110 * 1024) ? 2 : 1;
  constexpr size_t smem = kBufs * kBuf + kPS;
  static DeviceOnce attr_set;
  if (attr_set.first()) {
    RLT_CHECK_CUDA(cudaFuncSetAttribute(attn_lists_bwd_pipe_kernel<DH, NT, kDrop, kBufs, kFull>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int per_sm_q = 0;
    RLT_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_q, attn_lists_bwd_pipe_kernel<DH, NT, kDrop, kBufs, kFull>, 128, smem));
    attr_set.value[attr_set.dev()] = per_sm_q < 1 ? 1 : per_sm_q;
  }
  const long long n_items = (long long)G * L * n_head;
  RLT_REQUIRE(n_items < (1ll << 31) && size_t(S) * L * 3 * d < (size_t(1) << 32), RLT_UNSUPPORTED_SHAPE,
              "attention: %lld work items / group span exceed the 32-bit index range", n_items);
  long long grid = (long long)num_sms() * attr_set.value[attr_set.dev()];
  if (grid > n_items) grid = n_items;
  time_begin(TAG_ATTN_BWD, stream);
  attn_lists_bwd_pipe_kernel<DH, NT, kDrop, kBufs, kFull><<<int(grid), 128, smem, stream>>>(qkv, lse, d_o, dqkv, S, L, d, n_head, scale, int(n_items), drop);
  time_end(TAG_ATTN_BWD, stream);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}
template <int DH, int NT>
static int attention_bwd_mma(const float* qkv, const float* o, const float* lse, const float* d_o, float* dqkv, int G,
                             int S, int L, int d, int n_head, float scale, cudaStream_t stream, DropCfg drop) {
  (void)o;   // D = rowsum(P * dP) is recomputed from the fragments; the attention output is not needed
  if (S == NT * 8)
    return drop.thr ? attention_bwd_launch<DH, NT, true, true>(qkv, lse, d_o, dqkv, G, S, L, d, n_head, scale, stream, drop)
                    : attention_bwd_launch<DH, NT, false, true>(qkv, lse, d_o, dqkv, G, S, L, d, n_head, scale, stream, drop);
  return drop.thr ? attention_bwd_launch<DH, NT, true, false>(qkv, lse, d_o, dqkv, G, S, L, d, n_head, scale, stream, drop)
                  : attention_bwd_launch<DH, NT, false, false>(qkv, lse, d_o, dqkv, G, S, L, d, n_head, scale, stream, drop);
}
// tensor-core path available for S <= 128 and dh in {16, 32, 64}
static bool attention_mma_ok(int S, int dh) { return gemm_backend() == 0 && S <= 128 && (dh == 16 || dh == 32 || dh == 64); }

static int attention_fwd(const float* qkv, float* o, float* lse, int G, int S, int L, int d, int n_head,
                         cudaStream_t stream, DropCfg drop, int attend_axis = 0) {
  const int dh = d / n_head;
  const float scale = 1.0f / sqrtf(float(dh));
  // attend_axis 1 (attention WITHIN each list, the batch_first behaviour the reference does not have): every list is its
  // own group of seq_len consecutive tokens; the generic kernels below take it as (S = seq_len, stride 1)
  const int n_lists = G * S, seq_len = L;
  int group_tokens = S * L, grid_x = L;
  if (attend_axis == 1) { G = n_lists; S = seq_len; L = 1; group_tokens = seq_len; grid_x = 1; }
  if (attend_axis == 0 && drop.thr == 0 && attention_fwd_tc_ok(S, dh))       // tcgen05 path (attention_tc.cuh): head dim 16, groups of <= 64 lists
    return attention_fwd_tc(qkv, o, lse, G, S, L, d, n_head, scale, stream, TAG_ATTN_FWD);
  if (attend_axis == 0 && attention_mma_ok(S, dh)) {
#define RLT_AF(DH_)                                                                                            \
  return S <= 64 ? attention_fwd_mma<DH_, 8>(qkv, o, lse, G, S, L, d, n_head, scale, stream, drop)            \
                 : attention_fwd_mma<DH_, 16>(qkv, o, lse, G, S, L, d, n_head, scale, stream, drop)
    if (dh == 16) RLT_AF(16);
    if (dh == 32) RLT_AF(32);
    RLT_AF(64);
#undef RLT_AF
  }
  RLT_REQUIRE(G <= 65535, RLT_UNSUPPORTED_SHAPE, "attention: %d groups exceed gridDim.y", G);
  const dim3 grid(grid_x, G, n_head);
  const size_t smem = size_t(2) * S * (dh + 1) * sizeof(float);
  RLT_REQUIRE(smem <= 200 * 1024, RLT_UNSUPPORTED_SHAPE,
              "attention: %d attended elements of head dim %d do not fit in shared memory", S, dh);
#define RLT_ATTN_FWD(DH)                                                                                        \
  do {                                                                                                          \
    RLT_CHECK_CUDA(cudaFuncSetAttribute(attn_lists_fwd_kernel<DH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                        int(smem)));                                                            \
    RLT_CHECK_CUDA(cudaFuncSetAttribute(attn_lists_fwd_kernel<DH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                        int(smem)));                                                            \
    time_begin(TAG_ATTN_FWD, stream);                                                                           \
    if (drop.thr) attn_lists_fwd_kernel<DH, true><<<grid, 128, smem, stream>>>(qkv, o, lse, S, L, d, n_head, scale, drop, group_tokens); \
    else attn_lists_fwd_kernel<DH, false><<<grid, 128, smem, stream>>>(qkv, o, lse, S, L, d, n_head, scale, drop, group_tokens);         \
    time_end(TAG_ATTN_FWD, stream);                                                                             \
  } while (0)
  if (dh == 16) RLT_ATTN_FWD(16);
  else if (dh == 32) RLT_ATTN_FWD(32);
  else if (dh == 64) RLT_ATTN_FWD(64);
  else if (dh == 128) RLT_ATTN_FWD(128);
  else return set_error(RLT_UNSUPPORTED_SHAPE, "attention: head dim %d not in {16,32,64,128}", dh);
#undef RLT_ATTN_FWD
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

static int attention_bwd(const float* qkv, const float* o, const float* lse, const float* d_o, float* dqkv, int G,
                         int S, int L, int d, int n_head, cudaStream_t stream, DropCfg drop, int attend_axis = 0) {
  const int dh = d / n_head;
  const float scale = 1.0f / sqrtf(float(dh));
  const int n_lists = G * S, seq_len = L;
  int group_tokens = S * L, grid_x = L;
  if (attend_axis == 1) { G = n_lists; S = seq_len; L = 1; group_tokens = seq_len; grid_x = 1; }
  if (attend_axis == 0 && attention_mma_ok(S, dh)) {
#define RLT_AB(DH_)                                                                                                  \
  return S <= 64 ? attention_bwd_mma<DH_, 8>(qkv, o, lse, d_o, dqkv, G, S, L, d, n_head, scale, stream, drop)       \
                 : attention_bwd_mma<DH_, 16>(qkv, o, lse, d_o, dqkv, G, S, L, d, n_head, scale, stream, drop)
    if (dh == 16) RLT_AB(16);
    if (dh == 32) RLT_AB(32);
    RLT_AB(64);
#undef RLT_AB
  }
  RLT_REQUIRE(G <= 65535, RLT_UNSUPPORTED_SHAPE, "attention bwd: %d groups exceed gridDim.y", G);
  const dim3 grid(grid_x, G, n_head);
  const size_t smem = (size_t(4) * S * (dh + 1) + 2 * S) * sizeof(float);
  RLT_REQUIRE(smem <= 200 * 1024, RLT_UNSUPPORTED_SHAPE,
              "attention bwd: %d attended elements of head dim %d do not fit in shared memory (attend_axis = 1 is built for "
              "head dims up to 32 at seq_len 300)", S, dh);
#define RLT_ATTN_BWD(DH)                                                                                        \
  do {                                                                                                          \
    RLT_CHECK_CUDA(cudaFuncSetAttribute(attn_lists_bwd_kernel<DH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                        int(smem)));                                                            \
    RLT_CHECK_CUDA(cudaFuncSetAttribute(attn_lists_bwd_kernel<DH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                        int(smem)));                                                            \
    time_begin(TAG_ATTN_BWD, stream);                                                                           \
    if (drop.thr) attn_lists_bwd_kernel<DH, true><<<grid, 128, smem, stream>>>(qkv, o, lse, d_o, dqkv, S, L, d, n_head, scale, drop, group_tokens); \
    else attn_lists_bwd_kernel<DH, false><<<grid, 128, smem, stream>>>(qkv, o, lse, d_o, dqkv, S, L, d, n_head, scale, drop, group_tokens);         \
    time_end(TAG_ATTN_BWD, stream);                                                                             \
  } while (0)
  if (dh == 16) RLT_ATTN_BWD(16);
  else if (dh == 32) RLT_ATTN_BWD(32);
  else if (dh == 64) RLT_ATTN_BWD(64);
  else if (dh == 128) RLT_ATTN_BWD(128);
  else return set_error(RLT_UNSUPPORTED_SHAPE, "attention: head dim %d not in {16,32,64,128}", dh);
#undef RLT_ATTN_BWD
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

static int layer_norm_fwd(const float* u, const float* gamma, const float* beta, float* y, float* stats, int T, int d,
                          float eps, cudaStream_t stream, __half* y16 = nullptr) {
  const int rows_per_cta = 8;
  const int grid = (T + rows_per_cta - 1) / rows_per_cta;
  if (d == 128) ln_fwd_kernel<128><<<grid, 256, 0, stream>>>(u, gamma, beta, y, stats, T, eps, y16);
  else if (d == 256) ln_fwd_kernel<256><<<grid, 256, 0, stream>>>(u, gamma, beta, y, stats, T, eps, y16);
  else return set_error(RLT_UNSUPPORTED_SHAPE, "layer_norm: d_model %d not in {128, 256}", d);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

static int layer_norm_bwd(const float* dy, const float* u, const float* stats, const float* gamma, float* du,
                          float* dgamma, float* dbeta, float* dbias_prev, int T, int d, cudaStream_t stream,
                          unsigned int* amax_out = nullptr) {
  int grid = (T + 63) / 64;
  const int cap = num_sms() * 4;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  if (d == 128) ln_bwd_kernel<128><<<grid, 256, 0, stream>>>(dy, u, stats, gamma, du, dgamma, dbeta, dbias_prev, T, amax_out);
  else if (d == 256) ln_bwd_kernel<256><<<grid, 256, 0, stream>>>(dy, u, stats, gamma, du, dgamma, dbeta, dbias_prev, T, amax_out);
  else return set_error(RLT_UNSUPPORTED_SHAPE, "layer_norm: d_model %d not in {128, 256}", d);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

// ------------------------------------------------------------------------------------------
// saved-for-backward layout of one layer (floats per token): qkv 3d | o d | u1 d | y d | h dff | u2 d |
// stats1 2 | stats2 2 | lse n_head | y16 d/2, plus fp16 copies of the FFN weights.
// On the tensor-core backend the FFN hidden h (and dH in the backward workspace) are stored in fp16 — the 11
// significant bits the TF32 contraction would keep anyway — inside the fp32-sized regions (first half used):
// this halves the HBM traffic of the six hidden-sized GEMMs, which bound the layer.
// ------------------------------------------------------------------------------------------
struct SavedLayout {
  size_t qkv, o, u1, y, h, u2, st1, st2, lse, y16, wh, total;
};
static SavedLayout saved_layout(const rlt_encoder_desc& e) {
  const size_t T = size_t(e.n_groups) * e.group_size * e.seq_len;
  const size_t d = e.d_model, f = e.d_ff;
  SavedLayout s;
  size_t off = 0;
  auto take = [&](size_t n) { size_t r = off; off += (n + 63) / 64 * 64; return r; };  // 256 B aligned
  s.qkv = take(T * 3 * d);
  s.o = take(T * d);
  s.u1 = take(T * d);
  s.y = take(T * d);
  s.h = take(T * f);
  s.u2 = take(T * d);
  s.st1 = take(T * 2);
  s.st2 = take(T * 2);
  s.lse = take(T * e.n_head);
  s.y16 = take(T * d / 2);      // fp16 copy of y (operand of the dW1 contraction)
  s.wh = take(2 * d * f);       // fp16 copies of the FFN weights: W2 [d, f], W1^T [d, f], W1 [f, d], W2^T [f, d]
  s.total = off;
  return s;
}

// fp16 hidden activations: tensor-core backend, shapes the fp16 GEMMs take (d_ff multiple of 128)
static bool hidden_f16(const rlt_encoder_desc& e) { return gemm_backend() == 0 && e.d_ff % 128 == 0 && e.d_model % 128 == 0; }

static int check_desc(const rlt_encoder_desc* e) {
  RLT_REQUIRE(e != nullptr, RLT_INVALID_ARG, "encoder: null descriptor");
  RLT_REQUIRE(e->n_groups > 0 && e->group_size > 0 && e->seq_len > 0, RLT_INVALID_ARG,
              "encoder: n_groups=%d group_size=%d seq_len=%d must be positive", e->n_groups, e->group_size, e->seq_len);
  RLT_REQUIRE(e->d_model == 128 || e->d_model == 256, RLT_UNSUPPORTED_SHAPE, "encoder: d_model %d not in {128,256}",
              e->d_model);
  RLT_REQUIRE(e->n_head > 0 && e->d_model % e->n_head == 0, RLT_INVALID_ARG, "encoder: n_head %d does not divide d_model",
              e->n_head);
  RLT_REQUIRE(e->d_ff > 0 && e->d_ff % 32 == 0, RLT_UNSUPPORTED_SHAPE, "encoder: d_ff %d must be a multiple of 32", e->d_ff);
  RLT_REQUIRE(e->attend_axis == 0 || e->attend_axis == 1, RLT_INVALID_ARG,
              "encoder: attend_axis %d (0 = across the lists of a group, as the reference; 1 = within each list)", e->attend_axis);
  RLT_REQUIRE(e->dropout_p >= 0.f && e->dropout_p < 1.f, RLT_INVALID_ARG, "encoder: dropout_p %f outside [0, 1)", e->dropout_p);
  RLT_REQUIRE(e->dropout_p == 0.f || (gemm_backend() == 0 && e->d_ff % 128 == 0 && e->d_model % 128 == 0), RLT_UNSUPPORTED_SHAPE,
              "encoder: train-mode dropout runs on the tensor-core backend only (the validation kernels have no dropout)");
  RLT_REQUIRE(size_t(e->n_groups) * e->group_size * e->seq_len < (size_t(1) << 31), RLT_UNSUPPORTED_SHAPE,
              "encoder: token count overflows int32");
  return RLT_OK;
}

}  // namespace rlt

using namespace rlt;

extern "C" {

int rlt_colsum(const float* src, float* out, int n_rows, int n_cols, rlt_stream_t stream) {
  RLT_REQUIRE(src && out && n_rows > 0 && n_cols > 0, RLT_INVALID_ARG, "rlt_colsum: bad arguments");
  return colsum(src, out, n_rows, n_cols, static_cast<cudaStream_t>(stream));
}

int rlt_ffn_fused_fwd(const void* y16, const float* y, const void* w1_h, const float* b1, const void* w2_h, const float* b2,
                      const float* gamma, const float* beta, float* out, float* u2, float* stats, void* h_out, int n_tokens,
                      int d_model, int d_ff, float ln_eps, rlt_stream_t stream) {
  RLT_REQUIRE(y16 && y && w1_h && b1 && w2_h && b2 && gamma && beta && out && n_tokens > 0, RLT_INVALID_ARG,
              "rlt_ffn_fused_fwd: null pointer or empty problem");
  RLT_REQUIRE(ffn_fwd_fused_ok(d_model, d_ff), RLT_UNSUPPORTED_SHAPE,
              "rlt_ffn_fused_fwd: d_model=%d d_ff=%d (needs d_model 128 or 256, d_ff a multiple of 128 up to 2048, tensor-core backend)",
              d_model, d_ff);
  return ffn_fwd_fused(static_cast<const __half*>(y16), y, static_cast<const __half*>(w1_h), b1,
                       static_cast<const __half*>(w2_h), b2, gamma, beta, out, u2, stats, static_cast<__half*>(h_out),
                       n_tokens, d_model, d_ff, ln_eps, static_cast<cudaStream_t>(stream), TAG_FFN_FUSED);
}

int rlt_attention_lists_fwd(const float* qkv, float* o, float* lse, int n_groups, int group_size, int seq_len, int d_model,
                            int n_head, rlt_stream_t stream) {
  RLT_REQUIRE(qkv && o && n_groups > 0 && group_size > 0 && seq_len > 0 && n_head > 0 && d_model % n_head == 0, RLT_INVALID_ARG,
              "rlt_attention_lists_fwd: bad arguments");
  return attention_fwd(qkv, o, lse, n_groups, group_size, seq_len, d_model, n_head, static_cast<cudaStream_t>(stream),
                       DropCfg{0, 0, 1.f});
}

int rlt_ffn_fused_set_timeline(long long* device_buffer) {
  ffn_fwd_set_timeline(device_buffer);
  return RLT_OK;
}

size_t rlt_encoder_layer_saved_bytes(const rlt_encoder_desc* e) {
  if (check_desc(e) != RLT_OK) return 0;
  return saved_layout(*e).total * sizeof(float);
}

size_t rlt_encoder_layer_workspace_bytes(const rlt_encoder_desc* e) {
  if (check_desc(e) != RLT_OK) return 0;
  // backward scratch: d_u (T*d) + d_h / d_qkv (T*max(dff,3d)) + d_y (T*d)
  const size_t T = size_t(e->n_groups) * e->group_size * e->seq_len;
  const size_t wide = e->d_ff > 3 * e->d_model ? e->d_ff : 3 * e->d_model;
  return (T * (2 * size_t(e->d_model) + wide) + 256) * sizeof(float);
}

int rlt_encoder_layer_fwd(const rlt_encoder_desc* e, const rlt_encoder_weights* w, const float* x, float* out,
                          void* saved, size_t saved_bytes, rlt_stream_t stream_) {
  RLT_TRY(check_desc(e));
  RLT_REQUIRE(w && x && out && saved, RLT_INVALID_ARG, "encoder fwd: null pointer");
  const SavedLayout sl = saved_layout(*e);
  RLT_REQUIRE(saved_bytes >= sl.total * sizeof(float), RLT_WORKSPACE_TOO_SMALL,
              "encoder fwd: saved buffer has %zu bytes, needs %zu", saved_bytes, sl.total * sizeof(float));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  float* sv = static_cast<float*>(saved);
  const int T = e->n_groups * e->group_size * e->seq_len, d = e->d_model, f = e->d_ff;

  // train-mode dropout (torch: attention probabilities, dropout1 after the attention block, dropout inside the FFN,
  // dropout2 after it); the masks are functions of (seed, site, element) and are regenerated in the backward
  const DropCfg drop = e->dropout_p > 0.f ? make_drop(e->dropout_p, e->dropout_seed) : DropCfg{0, 0, 1.f};
  EpiParams ep{};
  ep.alpha = 1.f;
  // qkv = x Win^T + b_in
  ep.out = sv + sl.qkv; ep.ldo = 3 * d; ep.bias = w->in_proj_b; ep.tag = TAG_QKV;
  RLT_TRY(gemm_tn(x, d, w->in_proj_w, d, T, 3 * d, d, ep, stream));
  RLT_TRY(attention_fwd(sv + sl.qkv, sv + sl.o, sv + sl.lse, e->n_groups, e->group_size, e->seq_len, d, e->n_head,
                        stream, drop, e->attend_axis));
  // u1 = x + drop(o Wo^T + b_o) ; y = LN1(u1)
  ep = EpiParams{}; ep.alpha = 1.f; ep.out = sv + sl.u1; ep.ldo = d; ep.bias = w->out_proj_b; ep.residual = x; ep.tag = TAG_OUT_PROJ;
  ep.drop = drop; ep.drop_site = DROP_AFTER_ATTN;
  RLT_TRY(gemm_tn(sv + sl.o, d, w->out_proj_w, d, T, d, d, ep, stream));
  const bool h16 = hidden_f16(*e);
  __half* y16 = reinterpret_cast<__half*>(sv + sl.y16);
  __half* hh = reinterpret_cast<__half*>(sv + sl.h);
  __half* w2h = reinterpret_cast<__half*>(sv + sl.wh);
  __half* w1th = w2h + size_t(d) * f;
  RLT_TRY(layer_norm_fwd(sv + sl.u1, w->norm1_w, w->norm1_b, sv + sl.y, sv + sl.st1, T, d, e->ln_eps, stream,
                         h16 ? y16 : nullptr));
  __half* w1h = w1th + size_t(d) * f;
  __half* w2th = w1h + size_t(d) * f;
  if (h16 && drop.thr == 0 && ffn_fwd_fused_ok(d, f)) {
    // the whole feed-forward block + LayerNorm2 in one kernel (ffn_fwd_fused.cuh): the hidden stays on chip; in
    // training it is written once for the backward, which also needs the transposed fp16 weight copies
    const bool keep = e->inference == 0;
    RLT_TRY(convert_f16(w->lin1_w, w1h, size_t(d) * f, nullptr, stream));
    RLT_TRY(convert_f16(w->lin2_w, w2h, size_t(d) * f, nullptr, stream));
    if (keep) {
      RLT_TRY(transpose_f16(w->lin2_w, w2th, d, f, stream));                   // [d, f] -> [f, d] (backward dH)
      RLT_TRY(transpose_f16(w->lin1_w, w1th, f, d, stream));                   // [f, d] -> [d, f] (backward dY)
    }
    return ffn_fwd_fused(y16, sv + sl.y, w1h, w->lin1_b, w2h, w->lin2_b, w->norm2_w, w->norm2_b, out, keep ? sv + sl.u2 : nullptr,
                         keep ? sv + sl.st2 : nullptr, keep ? hh : nullptr, T, d, f, e->ln_eps, stream, TAG_FFN_FUSED);
  }
  // h = relu(y W1^T + b1)
  ep = EpiParams{}; ep.alpha = 1.f; ep.ldo = f; ep.bias = w->lin1_b; ep.relu = 1; ep.tag = TAG_FFN1;
  ep.drop = drop; ep.drop_site = DROP_FFN;       // the stored hidden is the DROPPED one (its sign pattern gates the backward)
  if (h16) {
    // fp16 operands (y16 is written by LN1, W1 is copied once per call): half the shared-memory operand traffic of
    // the TF32 form - the store-bound K = 128 GEMM is limited by the SM's shared-memory pipe, not by the tensor core
    ep.out_h = hh;
    RLT_TRY(convert_f16(w->lin1_w, w1h, size_t(d) * f, nullptr, stream));
    RLT_TRY(transpose_f16(w->lin2_w, w2th, d, f, stream));                     // [d, f] -> [f, d] (backward dH)
    RLT_TRY(gemm_tn_h(y16, d, w1h, d, T, f, d, ep, stream));
  } else {
    ep.out = sv + sl.h;
    RLT_TRY(gemm_tn(sv + sl.y, d, w->lin1_w, d, T, f, d, ep, stream));
  }
  // u2 = y + h W2^T + b2 ; out = LN2(u2)
  ep = EpiParams{}; ep.alpha = 1.f; ep.out = sv + sl.u2; ep.ldo = d; ep.bias = w->lin2_b; ep.residual = sv + sl.y; ep.tag = TAG_FFN2;
  ep.drop = drop; ep.drop_site = DROP_AFTER_FFN;
  if (h16) {
    RLT_TRY(convert_f16(w->lin2_w, w2h, size_t(d) * f, nullptr, stream));      // [d, f] as stored
    RLT_TRY(transpose_f16(w->lin1_w, w1th, f, d, stream));                     // [f, d] -> [d, f] (backward dY)
    RLT_TRY(gemm_tn_h(hh, f, w2h, f, T, d, f, ep, stream));
  } else {
    RLT_TRY(gemm_tn(sv + sl.h, f, w->lin2_w, f, T, d, f, ep, stream));
  }
  RLT_TRY(layer_norm_fwd(sv + sl.u2, w->norm2_w, w->norm2_b, out, sv + sl.st2, T, d, e->ln_eps, stream));
  return RLT_OK;
}

int rlt_encoder_layer_bwd(const rlt_encoder_desc* e, const rlt_encoder_weights* w, const rlt_encoder_grads* gw,
                          const float* x, const void* saved, const float* d_out, float* d_x, void* workspace,
                          size_t workspace_bytes, rlt_stream_t stream_) {
  RLT_TRY(check_desc(e));
  RLT_REQUIRE(w && gw && x && saved && d_out && d_x && workspace, RLT_INVALID_ARG, "encoder bwd: null pointer");
  RLT_REQUIRE(workspace_bytes >= rlt_encoder_layer_workspace_bytes(e), RLT_WORKSPACE_TOO_SMALL,
              "encoder bwd: workspace has %zu bytes, needs %zu", workspace_bytes, rlt_encoder_layer_workspace_bytes(e));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const SavedLayout sl = saved_layout(*e);
  const float* sv = static_cast<const float*>(saved);
  const int T = e->n_groups * e->group_size * e->seq_len, d = e->d_model, f = e->d_ff;
  const DropCfg drop = e->dropout_p > 0.f ? make_drop(e->dropout_p, e->dropout_seed) : DropCfg{0, 0, 1.f};
  float* ws = static_cast<float*>(workspace);
  float* d_u = ws;                          // [T, d]   dU2, later dU1
  float* d_y = ws + size_t(T) * d;          // [T, d]
  float* wide = ws + size_t(2) * T * d;     // [T, max(dff, 3d)]  dHpre, later dO (first T*d) / dQKV

  // LN2 backward: dU2, dgamma2, dbeta2, db2 (= column sums of dU2)
  // fp16 path: max |dU2| (for the power-of-two gradient scale) is reduced inside the LayerNorm backward
  float* scale = ws + size_t(T) * (2 * size_t(d) + (size_t(f) > 3 * size_t(d) ? f : 3 * d));   // {s, 1/s}, then the amax word
  unsigned int* amax = reinterpret_cast<unsigned int*>(scale + 2);
  if (hidden_f16(*e)) RLT_CHECK_CUDA(cudaMemsetAsync(amax, 0, sizeof(unsigned int), stream));
  RLT_TRY(layer_norm_bwd(d_out, sv + sl.u2, sv + sl.st2, w->norm2_w, d_u, gw->norm2_w, gw->norm2_b,
                         drop.thr ? nullptr : gw->lin2_b, T, d, stream, hidden_f16(*e) ? amax : nullptr));
  if (drop.thr) RLT_TRY(colsum_masked(d_u, gw->lin2_b, T, d, drop, DROP_AFTER_FFN, stream));   // b2 sits inside dropout2
  if (hidden_f16(*e)) {
    const __half* hh = reinterpret_cast<const __half*>(sv + sl.h);
    const __half* y16 = reinterpret_cast<const __half*>(sv + sl.y16);
    const __half* w1th = reinterpret_cast<const __half*>(sv + sl.wh) + size_t(d) * f;
    __half* dh16 = reinterpret_cast<__half*>(wide);                           // [T, f] fp16, scaled by s
    __half* du16 = reinterpret_cast<__half*>(wide + size_t(T) * f / 2);       // [T, d] fp16, scaled by s
    // power-of-two scale s: max|dU2| * s in [32, 64) -> dU2, dH = dU2 W2 stay far from fp16's limits on both sides
    RLT_TRY(pow2_scale(amax, scale, 6, stream));
    // the FFN branch sees dropout2(dU2); the residual branch below keeps the undropped d_u
    RLT_TRY(convert_f16(d_u, du16, size_t(T) * d, scale, stream, drop, DROP_AFTER_FFN));
    const __half* w2th = w1th + 2 * size_t(d) * f;
    EpiParams ep{};
    if (ffn_bwd_fused_ok(d, f)) {
      // one pass over h: dHpre (fp16), db1 += colsum(dHpre) / s, dW2 += (s dU2)^T h / s
      RLT_TRY(ffn_bwd_fused(du16, w2th, hh, dh16, T, d, f, drop.scale, scale, gw->lin1_b, gw->lin2_w, stream, TAG_D_FFN2));
    } else {
      // dW2 += dU2^T h
      RLT_TRY(gemm_dw_h(du16, d, hh, f, T, d, f, gw->lin2_w, f, 1.f, scale + 1, stream, TAG_DW_FFN2));
      // dHpre = ((s dU2) W2) * (h > 0) -> fp16 (the scale rides on the fp16 operand) ; db1 += colsum(dHpre) / s
      ep.alpha = drop.scale; ep.out_h = dh16; ep.ldo = f; ep.gate_h = hh; ep.colsum = gw->lin1_b; ep.scale_ptr = scale;
      ep.scale_mode = 3; ep.tag = TAG_D_FFN2;     // h is the dropped hidden: [h > 0] is relu-mask AND keep-mask; 1/(1-p) via alpha
      RLT_TRY(gemm_tn_h(du16, d, w2th, d, T, f, d, ep, stream));
    }
    // dW1 += dHpre^T y
    RLT_TRY(gemm_dw_h(dh16, f, y16, d, T, f, d, gw->lin1_w, d, 1.f, scale + 1, stream, TAG_DW_FFN1));
    // dY = dU2 + (dHpre W1) / s
    ep = EpiParams{}; ep.alpha = 1.f; ep.out = d_y; ep.ldo = d; ep.residual = d_u; ep.scale_ptr = scale; ep.scale_mode = 2;
    ep.tag = TAG_D_FFN1;
    RLT_TRY(gemm_tn_h(dh16, f, w1th, f, T, d, f, ep, stream));
  } else {
    // dW2 += dU2^T h
    RLT_TRY(gemm_dw(d_u, d, sv + sl.h, f, T, d, f, gw->lin2_w, f, 1.f, stream, TAG_DW_FFN2));
    // dHpre = (dU2 W2) * (h > 0) ; db1 += colsum(dHpre)
    EpiParams ep{};
    ep.alpha = 1.f; ep.out = wide; ep.ldo = f; ep.gate_src = sv + sl.h; ep.colsum = gw->lin1_b; ep.tag = TAG_D_FFN2;
    RLT_TRY(gemm_nn(d_u, d, w->lin2_w, f, T, f, d, ep, stream));
    // dW1 += dHpre^T y
    RLT_TRY(gemm_dw(wide, f, sv + sl.y, d, T, f, d, gw->lin1_w, d, 1.f, stream, TAG_DW_FFN1));
    // dY = dU2 + dHpre W1
    ep = EpiParams{}; ep.alpha = 1.f; ep.out = d_y; ep.ldo = d; ep.residual = d_u; ep.tag = TAG_D_FFN1;
    RLT_TRY(gemm_nn(wide, f, w->lin1_w, d, T, d, f, ep, stream));
  }
  EpiParams ep{};
  // LN1 backward: dU1 (into d_u), dgamma1, dbeta1, db_o
  RLT_TRY(layer_norm_bwd(d_y, sv + sl.u1, sv + sl.st1, w->norm1_w, d_u, gw->norm1_w, gw->norm1_b,
                         drop.thr ? nullptr : gw->out_proj_b, T, d, stream));
  if (drop.thr) RLT_TRY(colsum_masked(d_u, gw->out_proj_b, T, d, drop, DROP_AFTER_ATTN, stream));   // b_o sits inside dropout1
  // dWo += dU1m^T o ; dO = dU1m Wo with dU1m = dropout1 mask applied to dU1 (the residual branch keeps d_u itself);
  // `wide` is free here (dH / dU16 are dead, dQKV is written after the last use of the masked copy)
  const float* d_um = d_u;
  if (drop.thr) {
    RLT_TRY(dropout_apply(d_u, wide, size_t(T) * d, drop, DROP_AFTER_ATTN, stream));
    d_um = wide;
  }
  RLT_TRY(gemm_dw(d_um, d, sv + sl.o, d, T, d, d, gw->out_proj_w, d, 1.f, stream));
  float* d_o = d_y;  // dY is dead
  ep = EpiParams{}; ep.alpha = 1.f; ep.out = d_o; ep.ldo = d;
  RLT_TRY(gemm_nn(d_um, d, w->out_proj_w, d, T, d, d, ep, stream));
  // attention backward -> dQKV
  float* d_qkv = wide;
  RLT_TRY(attention_bwd(sv + sl.qkv, sv + sl.o, sv + sl.lse, d_o, d_qkv, e->n_groups, e->group_size, e->seq_len, d,
                        e->n_head, stream, drop, e->attend_axis));
  // db_in += colsum(dQKV) ; dWin += dQKV^T x ; dX = dU1 + dQKV Win
  RLT_TRY(gemm_dw(d_qkv, 3 * d, x, d, T, 3 * d, d, gw->in_proj_w, d, 1.f, stream, 0, gw->in_proj_b));
  ep = EpiParams{}; ep.alpha = 1.f; ep.out = d_x; ep.ldo = d; ep.residual = d_u; ep.accumulate = e->accumulate_dx ? 1 : 0;
  RLT_TRY(gemm_nn(d_qkv, 3 * d, w->in_proj_w, d, T, d, 3 * d, ep, stream));
  return RLT_OK;
}

}  // extern "C"
