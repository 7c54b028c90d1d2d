// K5 — multi-gate mixture-of-experts heads of MMOECut (reference models/MMOECut.py:86-110).
//
//   gate_t[b, :]  = softmax_E( flat(H_lstm[b]) . W_t )            flat: L*256 = 76 800 inputs, true fp32 (the
//                                                                  logits reach |22|; TF32 would move the gates)
//   z_t[b, l]     = w_t . ( sum_e gate_t[b,e] X_e[b,l,:] ) + b_t   tower Linear(d,1) on the gate-weighted mixture
//                 = sum_e gate_t[b,e] (w_t . X_e[b,l,:]) + b_t     -> the [B,L,d] mixtures are never materialised
// Backward returns dX_e, dw_t, db_t, dW_t and the gradient that flows into H_lstm through the gates.
#include <math.h>

#include "common.h"

namespace rlt {

constexpr int MAX_E = 4;   // experts
constexpr int MAX_T = 3;   // tasks

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// logits[t, b, e] = sum_j flat[b, j] * W[t][j, e]; one CTA per list, 256 threads stride over j.
constexpr int GATE_LB = 8;   // lists per CTA of the gate kernels: every gate weight fetched from L2 serves 8 lists
// gates[t, b, :] = softmax_e( flat[b, :] . W_t[:, e] )  in true fp32 (K = 76 800).  One CTA per GATE_LB
// lists (the first version used one CTA per list and re-read the 2.8 MB of gate weights from L2 for each of them: 0.83 ms
// for 2048 lists, L2-bandwidth bound).
__global__ void __launch_bounds__(256) moe_gate_logits_kernel(const float* __restrict__ flat, const float* __restrict__ wg,
                                                              float* __restrict__ gates, int B, int J, int E, int Tk) {
  const int b0 = blockIdx.x * GATE_LB;
  float acc[GATE_LB][MAX_T * MAX_E];
#pragma unroll
  for (int l = 0; l < GATE_LB; ++l)
#pragma unroll
    for (int i = 0; i < MAX_T * MAX_E; ++i) acc[l][i] = 0.f;
  for (int j = threadIdx.x; j < J; j += blockDim.x) {
    float w[MAX_T * MAX_E];
#pragma unroll
    for (int t = 0; t < MAX_T; ++t)
#pragma unroll
      for (int e = 0; e < MAX_E; ++e) w[t * MAX_E + e] = (t < Tk && e < E) ? __ldg(wg + (size_t(t) * J + j) * E + e) : 0.f;
#pragma unroll
    for (int l = 0; l < GATE_LB; ++l) {
      const float v = b0 + l < B ? __ldg(flat + size_t(b0 + l) * J + j) : 0.f;
#pragma unroll
      for (int i = 0; i < MAX_T * MAX_E; ++i) acc[l][i] = fmaf(v, w[i], acc[l][i]);
    }
  }
  __shared__ float red[8][GATE_LB][MAX_T * MAX_E];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int l = 0; l < GATE_LB; ++l)
#pragma unroll
    for (int i = 0; i < MAX_T * MAX_E; ++i) {
      const float sacc = wsum(acc[l][i]);
      if (lane == 0) red[warp][l][i] = sacc;
    }
  __syncthreads();
  if (threadIdx.x < GATE_LB * Tk) {
    const int l = threadIdx.x / Tk, t = threadIdx.x % Tk;
    if (b0 + l < B) {
      float lg[MAX_E], m = -INFINITY;
      for (int e = 0; e < E; ++e) {
        float sacc = 0.f;
        for (int wq = 0; wq < 8; ++wq) sacc += red[wq][l][t * MAX_E + e];
        lg[e] = sacc;
        m = fmaxf(m, sacc);
      }
      float den = 0.f;
      for (int e = 0; e < E; ++e) { lg[e] = expf(lg[e] - m); den += lg[e]; }
      for (int e = 0; e < E; ++e) gates[(size_t(t) * B + b0 + l) * E + e] = lg[e] / den;
    }
  }
}

struct ExpertPtrs { const float* x[MAX_E]; };
struct ExpertGradPtrs { float* dx[MAX_E]; };

// z[t, b, l] = sum_e gate[t,b,e] * (w[t] . X_e[b,l,:]) + bias[t]; one warp per token.
template <int D>
__global__ void __launch_bounds__(256) moe_heads_fwd_kernel(ExpertPtrs xs, const float* __restrict__ gates,
                                                            const float* __restrict__ w, const float* __restrict__ bias,
                                                            float* __restrict__ z, int B, int L, int E, int Tk) {
  constexpr int V4 = D / 128;
  const int lane = threadIdx.x & 31;
  const size_t tok = size_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tok >= size_t(B) * L) return;
  const int b = int(tok / L);
  float zt[MAX_T] = {0.f, 0.f, 0.f};
  for (int e = 0; e < E; ++e) {
    float4 xv[V4];
#pragma unroll
    for (int i = 0; i < V4; ++i) xv[i] = reinterpret_cast<const float4*>(xs.x[e] + tok * D)[lane + 32 * i];
#pragma unroll
    for (int t = 0; t < MAX_T; ++t)
      if (t < Tk) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < V4; ++i) {
          const float4 ww = __ldg(reinterpret_cast<const float4*>(w + t * D) + lane + 32 * i);
          s += (xv[i].x * ww.x + xv[i].y * ww.y) + (xv[i].z * ww.z + xv[i].w * ww.w);
        }
        s = wsum(s);
        zt[t] = fmaf(gates[(size_t(t) * B + b) * E + e], s, zt[t]);
      }
  }
  if (lane == 0)
    for (int t = 0; t < Tk; ++t) z[size_t(t) * B * L + tok] = zt[t] + bias[t];
}

// Per token: dX_e = sum_t gate[t,b,e] dz_t w_t ; dw_t += dz_t * mix_t ; db_t += dz_t ; dG[t,b,e] += dz_t * s[t,e]
template <int D>
__global__ void __launch_bounds__(256) moe_heads_bwd_kernel(ExpertPtrs xs, ExpertGradPtrs dxs, const float* __restrict__ gates,
                                                            const float* __restrict__ w, const float* __restrict__ dz,
                                                            float* __restrict__ dw, float* __restrict__ db,
                                                            float* __restrict__ dG, int B, int L, int E, int Tk) {
  constexpr int V4 = D / 128;
  __shared__ float red[MAX_T][D];
  __shared__ float redb[MAX_T];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < MAX_T * D; i += blockDim.x) (&red[0][0])[i] = 0.f;
  if (threadIdx.x < MAX_T) redb[threadIdx.x] = 0.f;
  __syncthreads();
  float4 ww[MAX_T][V4], aw[MAX_T][V4];
  float ab[MAX_T];
#pragma unroll
  for (int t = 0; t < MAX_T; ++t) {
    ab[t] = 0.f;
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      ww[t][i] = t < Tk ? __ldg(reinterpret_cast<const float4*>(w + t * D) + lane + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
      aw[t][i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const size_t ntok = size_t(B) * L;
  for (size_t tok = size_t(blockIdx.x) * nwarps + warp; tok < ntok; tok += size_t(gridDim.x) * nwarps) {
    const int b = int(tok / L);
    float g[MAX_T];
#pragma unroll
    for (int t = 0; t < MAX_T; ++t) { g[t] = t < Tk ? dz[size_t(t) * ntok + tok] : 0.f; ab[t] += g[t]; }
    for (int e = 0; e < E; ++e) {
      float gt[MAX_T];
#pragma unroll
      for (int t = 0; t < MAX_T; ++t) gt[t] = t < Tk ? gates[(size_t(t) * B + b) * E + e] : 0.f;
#pragma unroll
      for (int i = 0; i < V4; ++i) {
        const float4 xv = reinterpret_cast<const float4*>(xs.x[e] + tok * D)[lane + 32 * i];
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int t = 0; t < MAX_T; ++t) {
          const float c = gt[t] * g[t];
          o.x = fmaf(c, ww[t][i].x, o.x); o.y = fmaf(c, ww[t][i].y, o.y);
          o.z = fmaf(c, ww[t][i].z, o.z); o.w = fmaf(c, ww[t][i].w, o.w);
          aw[t][i].x = fmaf(c, xv.x, aw[t][i].x); aw[t][i].y = fmaf(c, xv.y, aw[t][i].y);
          aw[t][i].z = fmaf(c, xv.z, aw[t][i].z); aw[t][i].w = fmaf(c, xv.w, aw[t][i].w);
        }
        reinterpret_cast<float4*>(dxs.dx[e] + tok * D)[lane + 32 * i] = o;
      }
      // dG[t,b,e] += dz_t * (w_t . X_e[tok])   (the row was just read: the second pass hits L1)
#pragma unroll
      for (int t = 0; t < MAX_T; ++t)
        if (t < Tk) {
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < V4; ++i) {
            const float4 xv = reinterpret_cast<const float4*>(xs.x[e] + tok * D)[lane + 32 * i];
            s += (xv.x * ww[t][i].x + xv.y * ww[t][i].y) + (xv.z * ww[t][i].z + xv.w * ww[t][i].w);
          }
          s = wsum(s);
          if (lane == 0) atomicAdd(dG + (size_t(t) * B + b) * E + e, g[t] * s);
        }
    }
  }
#pragma unroll
  for (int t = 0; t < MAX_T; ++t) {
    if (t >= Tk) continue;
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      const int c = (lane + 32 * i) * 4;
      atomicAdd(&red[t][c], aw[t][i].x); atomicAdd(&red[t][c + 1], aw[t][i].y);
      atomicAdd(&red[t][c + 2], aw[t][i].z); atomicAdd(&red[t][c + 3], aw[t][i].w);
    }
    if (lane == 0) atomicAdd(&redb[t], ab[t]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Tk * D; i += blockDim.x) atomicAdd(dw + i, (&red[0][0])[i]);
  if (threadIdx.x < Tk) atomicAdd(db + threadIdx.x, redb[threadIdx.x]);
}

// dlogit[t,b,e] = gate * (dG - sum_e gate dG)   (softmax over experts), in place over dG
__global__ void moe_gate_softmax_bwd_kernel(const float* __restrict__ gates, float* __restrict__ dG, int n, int E) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // (t, b) pair
  if (i >= n) return;
  float dot = 0.f;
  for (int e = 0; e < E; ++e) dot += gates[size_t(i) * E + e] * dG[size_t(i) * E + e];
  for (int e = 0; e < E; ++e) dG[size_t(i) * E + e] = gates[size_t(i) * E + e] * (dG[size_t(i) * E + e] - dot);
}

// dW[t][j, e] += sum_b flat[b, j] * dlogit[t, b, e]: thread = gate input j, blockIdx.y = slice of the lists (the first
// version walked ALL lists in one dependent loop per thread: 1.4 ms for 2048 lists); the dlogit addresses are uniform
// per warp (broadcast), four independent flat loads are in flight, partial sums go out with one red.add per element.
__global__ void __launch_bounds__(256) moe_gate_dw_kernel(const float* __restrict__ flat, const float* __restrict__ dlogit,
                                                          float* __restrict__ dwg, int B, int J, int E, int Tk) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= J) return;
  const int per = (B + gridDim.y - 1) / gridDim.y;
  const int bs = blockIdx.y * per, be = min(B, bs + per);
  float acc[MAX_T * MAX_E];
#pragma unroll
  for (int i = 0; i < MAX_T * MAX_E; ++i) acc[i] = 0.f;
  int b = bs;
  for (; b + 3 < be; b += 4) {
    float v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = __ldg(flat + size_t(b + q) * J + j);
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int t = 0; t < MAX_T; ++t)
        if (t < Tk)
#pragma unroll
          for (int e = 0; e < MAX_E; ++e)
            if (e < E) acc[t * MAX_E + e] = fmaf(v[q], __ldg(dlogit + (size_t(t) * B + b + q) * E + e), acc[t * MAX_E + e]);
  }
  for (; b < be; ++b) {
    const float v = __ldg(flat + size_t(b) * J + j);
#pragma unroll
    for (int t = 0; t < MAX_T; ++t)
      if (t < Tk)
#pragma unroll
        for (int e = 0; e < MAX_E; ++e)
          if (e < E) acc[t * MAX_E + e] = fmaf(v, __ldg(dlogit + (size_t(t) * B + b) * E + e), acc[t * MAX_E + e]);
  }
  for (int t = 0; t < Tk; ++t)
    for (int e = 0; e < E; ++e) atomicAdd(dwg + (size_t(t) * J + j) * E + e, acc[t * MAX_E + e]);
}

// dflat[b, j] (+)= sum_{t,e} dlogit[t,b,e] * W[t][j,e]: thread = gate input j for GATE_LB lists (weights loaded once)
__global__ void __launch_bounds__(256) moe_gate_dflat_kernel(const float* __restrict__ dlogit, const float* __restrict__ wg,
                                                             float* __restrict__ dflat, int B, int J, int E, int Tk,
                                                             int accumulate) {
  const int b0 = blockIdx.y * GATE_LB;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= J) return;
  float w[MAX_T * MAX_E];
#pragma unroll
  for (int t = 0; t < MAX_T; ++t)
#pragma unroll
    for (int e = 0; e < MAX_E; ++e) w[t * MAX_E + e] = (t < Tk && e < E) ? __ldg(wg + (size_t(t) * J + j) * E + e) : 0.f;
#pragma unroll
  for (int l = 0; l < GATE_LB; ++l) {
    const int b = b0 + l;
    if (b >= B) break;
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < MAX_T; ++t)
      if (t < Tk)
#pragma unroll
        for (int e = 0; e < MAX_E; ++e)
          if (e < E) acc = fmaf(__ldg(dlogit + (size_t(t) * B + b) * E + e), w[t * MAX_E + e], acc);
    float* o = dflat + size_t(b) * J + j;
    *o = accumulate ? *o + acc : acc;
  }
}

}  // namespace rlt

using namespace rlt;

extern "C" {

int rlt_moe_heads_fwd(const rlt_moe_desc* m, const float* h_lstm, const float* w_gates, const float* const* experts,
                      const float* tower_w, const float* tower_b, float* gates, float* z, rlt_stream_t stream_) {
  RLT_REQUIRE(m && h_lstm && w_gates && experts && tower_w && tower_b && gates && z, RLT_INVALID_ARG, "moe fwd: null pointer");
  RLT_REQUIRE(m->n_experts >= 1 && m->n_experts <= MAX_E && m->n_tasks >= 1 && m->n_tasks <= MAX_T, RLT_UNSUPPORTED_SHAPE,
              "moe: n_experts=%d n_tasks=%d (max %d / %d)", m->n_experts, m->n_tasks, MAX_E, MAX_T);
  RLT_REQUIRE(m->d_model == 128 || m->d_model == 256, RLT_UNSUPPORTED_SHAPE, "moe: d_model %d", m->d_model);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int B = m->n_lists, L = m->seq_len, J = L * m->d_lstm, E = m->n_experts, Tk = m->n_tasks;
  moe_gate_logits_kernel<<<(B + GATE_LB - 1) / GATE_LB, 256, 0, stream>>>(h_lstm, w_gates, gates, B, J, E, Tk);
  RLT_CHECK_LAUNCH();
  ExpertPtrs xs{};
  for (int e = 0; e < E; ++e) xs.x[e] = experts[e];
  const int grid = int((size_t(B) * L + 7) / 8);
  if (m->d_model == 128) moe_heads_fwd_kernel<128><<<grid, 256, 0, stream>>>(xs, gates, tower_w, tower_b, z, B, L, E, Tk);
  else moe_heads_fwd_kernel<256><<<grid, 256, 0, stream>>>(xs, gates, tower_w, tower_b, z, B, L, E, Tk);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

int rlt_moe_heads_bwd(const rlt_moe_desc* m, const float* h_lstm, const float* w_gates, const float* const* experts,
                      const float* tower_w, const float* gates, const float* dz, float* const* d_experts,
                      float* d_tower_w, float* d_tower_b, float* d_w_gates, float* d_h_lstm, int accumulate_dh,
                      float* dgate_scratch, rlt_stream_t stream_) {
  RLT_REQUIRE(m && h_lstm && w_gates && experts && tower_w && gates && dz && d_experts && d_tower_w && d_tower_b &&
                  d_w_gates && d_h_lstm && dgate_scratch, RLT_INVALID_ARG, "moe bwd: null pointer");
  RLT_REQUIRE(m->n_experts >= 1 && m->n_experts <= MAX_E && m->n_tasks >= 1 && m->n_tasks <= MAX_T, RLT_UNSUPPORTED_SHAPE,
              "moe: n_experts=%d n_tasks=%d", m->n_experts, m->n_tasks);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int B = m->n_lists, L = m->seq_len, J = L * m->d_lstm, E = m->n_experts, Tk = m->n_tasks;
  RLT_CHECK_CUDA(cudaMemsetAsync(dgate_scratch, 0, sizeof(float) * size_t(Tk) * B * E, stream));
  ExpertPtrs xs{};
  ExpertGradPtrs dxs{};
  for (int e = 0; e < E; ++e) { xs.x[e] = experts[e]; dxs.dx[e] = d_experts[e]; }
  size_t ntok = size_t(B) * L;
  int grid = int((ntok + 63) / 64);
  if (grid > num_sms() * 4) grid = num_sms() * 4;
  if (m->d_model == 128)
    moe_heads_bwd_kernel<128><<<grid, 256, 0, stream>>>(xs, dxs, gates, tower_w, dz, d_tower_w, d_tower_b, dgate_scratch, B, L, E, Tk);
  else
    moe_heads_bwd_kernel<256><<<grid, 256, 0, stream>>>(xs, dxs, gates, tower_w, dz, d_tower_w, d_tower_b, dgate_scratch, B, L, E, Tk);
  RLT_CHECK_LAUNCH();
  moe_gate_softmax_bwd_kernel<<<(Tk * B + 127) / 128, 128, 0, stream>>>(gates, dgate_scratch, Tk * B, E);
  RLT_CHECK_LAUNCH();
  {
    int slices = (B + 127) / 128;      // >= 128 lists per slice; (J / 256) x slices CTAs
    if (slices > 16) slices = 16;
    moe_gate_dw_kernel<<<dim3((J + 255) / 256, slices), 256, 0, stream>>>(h_lstm, dgate_scratch, d_w_gates, B, J, E, Tk);
  }
  RLT_CHECK_LAUNCH();
  moe_gate_dflat_kernel<<<dim3((J + 255) / 256, (B + GATE_LB - 1) / GATE_LB), 256, 0, stream>>>(dgate_scratch, w_gates, d_h_lstm, B, J, E,
                                                                                             Tk, accumulate_dh);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

}  // extern "C"
