// 2-layer bidirectional LSTM (H = 128) forward and BPTT backward.
//
// Replaces torch.nn.LSTM(input_size=F, hidden_size=128, num_layers=2, batch_first=True, bidirectional=True)
// as used at reference models/Bicut.py:8-9,19, AttnCut.py:8,17, MtAttnCut.py:8,22, MMOECut.py:63,88.
// Math: SURVEY.md A.1 (gate row order i,f,g,o; h0 = c0 = 0; layer 1 consumes [fwd || bwd] of layer 0).
//
// Structure per layer:
//   P[t, dir*512 + :]  = x_t W_ih^T + b_ih + b_hh        one dense GEMM over all tokens (tcgen05 when K % 4 == 0,
//                                                        a small-K kernel for the 3..47-feature first layer)
//   recurrence         : a_t = P_t + h_{t-1} W_hh^T -> gates -> (c_t, h_t), both directions concurrently
//   backward recurrence: da_t from (dy_t + dh_rec, dc_rec), dh_rec = da_t W_hh
//   deferred GEMMs     : dW_hh += dA^T Hprev, dW_ih += dA^T X, dX = dA W_ih, db = colsum(dA)   (tcgen05)
// Saved per (list, t, dir): i, f, g, o, c_t, h_{t-1}  (6 x 128 fp32).
#include <math.h>
#include <stdlib.h>

#include "common.h"
#include "gemm_tc.cuh"
#include "lstm_um.cuh"

namespace rlt {

int colsum(const float* src, float* out, int T, int C, cudaStream_t stream);  // encoder.cu

constexpr int H = 128;      // hidden size
constexpr int G4 = 4 * H;   // gate rows per direction
constexpr int SAVE = 6;     // saved planes per (list, t, dir)

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.f / (1.f + expf(-x)); }

// P[t, n] = sum_k x[t,k] W[n,k] + b1[n] + b2[n] for small K (first layer, K = 3 / 25 / 47).  n in [0, 512).
__global__ void __launch_bounds__(256) lstm_inproj_small_kernel(const float* __restrict__ x, int K,
                                                                const float* __restrict__ w, const float* __restrict__ b1,
                                                                const float* __restrict__ b2, float* __restrict__ P,
                                                                int ldp, int T) {
  extern __shared__ float sw[];  // [K][512] transposed weights + [512] bias
  float* sb = sw + K * G4;
  for (int i = threadIdx.x; i < K * G4; i += blockDim.x) {
    const int n = i / K, k = i % K;
    sw[k * G4 + n] = w[i];
  }
  for (int n = threadIdx.x; n < G4; n += blockDim.x) sb[n] = b1[n] + b2[n];
  __syncthreads();
  for (int t = blockIdx.x; t < T; t += gridDim.x) {
    const float* xr = x + size_t(t) * K;
    for (int n = threadIdx.x; n < G4; n += blockDim.x) {
      float acc = sb[n];
      for (int k = 0; k < K; ++k) acc = fmaf(xr[k], sw[k * G4 + n], acc);
      P[size_t(t) * ldp + n] = acc;
    }
  }
}

// dW[n, k] += sum_t dA[t, n] x[t, k] for small K.  One CTA per token chunk; thread n keeps K accumulators.
template <int KMAX>
__global__ void __launch_bounds__(512) lstm_dwih_small_kernel(const float* __restrict__ dA, int lda,
                                                              const float* __restrict__ x, int K,
                                                              float* __restrict__ dW, int T, int chunk) {
  const int n = threadIdx.x;
  const int t0 = blockIdx.x * chunk, t1 = min(T, t0 + chunk);
  float acc[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) acc[k] = 0.f;
  for (int t = t0; t < t1; ++t) {
    const float a = dA[size_t(t) * lda + n];
    const float* xr = x + size_t(t) * K;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < K) acc[k] = fmaf(a, __ldg(xr + k), acc[k]);
  }
#pragma unroll
  for (int k = 0; k < KMAX; ++k)
    if (k < K) atomicAdd(dW + size_t(n) * K + k, acc[k]);
}

// dx[t, k] = sum_n dA[t, n] W[n, k] over both directions, small K (only when the caller wants input grads)
__global__ void __launch_bounds__(128) lstm_dx_small_kernel(const float* __restrict__ dA, const float* __restrict__ wf,
                                                            const float* __restrict__ wr, int K, float* __restrict__ dx,
                                                            int T) {
  const int t = blockIdx.x;
  if (t >= T) return;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float acc = 0.f;
    const float* a = dA + size_t(t) * 2 * G4;
    for (int n = 0; n < G4; ++n) acc = fmaf(a[n], wf[size_t(n) * K + k], acc);
    for (int n = 0; n < G4; ++n) acc = fmaf(a[G4 + n], wr[size_t(n) * K + k], acc);
    dx[size_t(t) * K + k] = acc;
  }
}

// WT[k][n] = W[n][k]  ([512,128] -> [128,512], exact copy)
__global__ void transpose_whh_kernel(const float* __restrict__ w, float* __restrict__ wt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < G4 * H) {
    const int n = i / H, k = i % H;
    wt[k * G4 + n] = w[i];
  }
}

// ------------------------------------------------------------------------------------------
// Recurrence, plain kernel: one CTA per (list, direction), thread u owns hidden unit u.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(H) lstm_rec_fwd_kernel(const float* __restrict__ P, const float* __restrict__ wt_f,
                                                         const float* __restrict__ wt_r, float* __restrict__ y,
                                                         float* __restrict__ saved, int L) {
  __shared__ float sh[H];
  const int b = blockIdx.x, dir = blockIdx.y, u = threadIdx.x;
  const float* wt = dir ? wt_r : wt_f;  // [128][512]
  float c = 0.f, hprev = 0.f;
  sh[u] = 0.f;
  __syncthreads();
  for (int step = 0; step < L; ++step) {
    const int t = dir ? (L - 1 - step) : step;
    const size_t tok = size_t(b) * L + t;
    const float* p = P + tok * (2 * G4) + dir * G4;
    float ai = p[u], af = p[H + u], ag = p[2 * H + u], ao = p[3 * H + u];
#pragma unroll 8
    for (int k = 0; k < H; ++k) {
      const float hk = sh[k];
      const float* wr = wt + k * G4;
      ai = fmaf(hk, wr[u], ai);
      af = fmaf(hk, wr[H + u], af);
      ag = fmaf(hk, wr[2 * H + u], ag);
      ao = fmaf(hk, wr[3 * H + u], ao);
    }
    const float gi = sigmoid_acc(ai), gf = sigmoid_acc(af), gg = tanhf(ag), go = sigmoid_acc(ao);
    c = gf * c + gi * gg;
    const float hn = go * tanhf(c);
    if (saved != nullptr) {
      float* s = saved + (tok * 2 + dir) * (SAVE * H);
      s[u] = gi; s[H + u] = gf; s[2 * H + u] = gg; s[3 * H + u] = go; s[4 * H + u] = c; s[5 * H + u] = hprev;
    }
    y[tok * (2 * H) + dir * H + u] = hn;
    hprev = hn;
    __syncthreads();
    sh[u] = hn;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(H) lstm_rec_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ saved,
                                                         const float* __restrict__ w_f, const float* __restrict__ w_r,
                                                         float* __restrict__ dA, int L) {
  __shared__ float sda[G4];
  __shared__ float sdh[H];
  const int b = blockIdx.x, dir = blockIdx.y, u = threadIdx.x;
  const float* w = dir ? w_r : w_f;  // [512][128]
  float dc_rec = 0.f;
  sdh[u] = 0.f;
  __syncthreads();
  for (int step = L - 1; step >= 0; --step) {
    const int t = dir ? (L - 1 - step) : step;
    const size_t tok = size_t(b) * L + t;
    const float* s = saved + (tok * 2 + dir) * (SAVE * H);
    const float gi = s[u], gf = s[H + u], gg = s[2 * H + u], go = s[3 * H + u], c = s[4 * H + u];
    float cprev = 0.f;
    if (step > 0) {
      const int tp = dir ? (t + 1) : (t - 1);
      cprev = saved[((size_t(b) * L + tp) * 2 + dir) * (SAVE * H) + 4 * H + u];
    }
    const float dh = dy[tok * (2 * H) + dir * H + u] + sdh[u];
    const float tc = tanhf(c);
    const float d_o = dh * tc;
    const float dc = dc_rec + dh * go * (1.f - tc * tc);
    const float d_i = dc * gg, d_g = dc * gi, d_f = dc * cprev;
    dc_rec = dc * gf;
    const float dai = d_i * gi * (1.f - gi), daf = d_f * gf * (1.f - gf), dag = d_g * (1.f - gg * gg),
                dao = d_o * go * (1.f - go);
    float* out = dA + tok * (2 * G4) + dir * G4;
    out[u] = dai; out[H + u] = daf; out[2 * H + u] = dag; out[3 * H + u] = dao;
    sda[u] = dai; sda[H + u] = daf; sda[2 * H + u] = dag; sda[3 * H + u] = dao;
    __syncthreads();
    float acc = 0.f;
#pragma unroll 8
    for (int n = 0; n < G4; ++n) acc = fmaf(sda[n], w[size_t(n) * H + u], acc);
    __syncthreads();
    sdh[u] = acc;
    __syncthreads();
  }
}

// out1[c] += colsum, out2[c] += colsum (b_ih and b_hh receive the same gradient)
__global__ void copy_add_kernel(const float* __restrict__ src, float* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}

static bool small_k(int K) { return K % 4 != 0 || K < 32; }

// 0 = persistent tcgen05 recurrence (default), 1 = plain validation kernels
static int g_lstm_backend = -1;
int lstm_backend() {
  if (g_lstm_backend < 0) {
    const char* v = getenv("RLT_LSTM_BACKEND");
    g_lstm_backend = v ? atoi(v) : 0;
  }
  return gemm_backend() == 1 ? 1 : g_lstm_backend;
}
void set_lstm_backend(int v) { g_lstm_backend = v; }

static int check_lstm(const rlt_bilstm_desc* d) {
  RLT_REQUIRE(d != nullptr, RLT_INVALID_ARG, "bilstm: null descriptor");
  RLT_REQUIRE(d->n_lists > 0 && d->seq_len > 0 && d->input_size > 0, RLT_INVALID_ARG, "bilstm: bad sizes");
  RLT_REQUIRE(d->hidden == H, RLT_UNSUPPORTED_SHAPE, "bilstm: hidden size %d is not 128", d->hidden);
  RLT_REQUIRE(d->num_layers == 2, RLT_UNSUPPORTED_SHAPE, "bilstm: num_layers %d is not 2", d->num_layers);
  RLT_REQUIRE(d->input_size <= 64 || d->input_size % 4 == 0, RLT_UNSUPPORTED_SHAPE,
              "bilstm: input_size %d unsupported (<= 64, or a multiple of 4)", d->input_size);
  RLT_REQUIRE(size_t(d->n_lists) * d->seq_len < (size_t(1) << 31) / 1024, RLT_UNSUPPORTED_SHAPE, "bilstm: too many tokens per call");
  return RLT_OK;
}

struct LstmSaved {
  size_t s0, s1, y0, total;  // float offsets
};
static LstmSaved lstm_saved(const rlt_bilstm_desc& d) {
  const size_t T = size_t(d.n_lists) * d.seq_len;
  LstmSaved s;
  s.s0 = 0;
  s.s1 = T * 2 * SAVE * H;
  s.y0 = 2 * s.s1;
  s.total = s.y0 + T * 2 * H;
  return s;
}
// workspace (floats): P / dA [T, 1024] | dY0 [T, 256] | WT [4][128*512] | bias scratch [1024]
// lists per CTA of the tcgen05 recurrence: 32 while 2 * ceil(B / 32) CTAs still fit in one wave, else 64
static int g_lstm_tile = 0;          // 0 = automatic (rlt_set_option("lstm_tile", 32 | 64) pins it: A/B measurements, tests)
void set_lstm_tile(int v) { g_lstm_tile = (v == 32 || v == 64) ? v : 0; }
int get_lstm_tile() { return g_lstm_tile; }
static int lstm_tile(int B) {
  if (g_lstm_tile != 0) return g_lstm_tile;
  return 2 * ((B + 31) / 32) <= num_sms() ? 32 : 64;
}

static size_t lstm_ws_floats(const rlt_bilstm_desc& d) {
  const size_t T = size_t(d.n_lists) * d.seq_len;
  return T * 2 * G4 + T * 2 * H + 4 * size_t(H) * G4 + 2 * G4 + 256;
}

static int input_projection(const float* x, int K, const float* w, const float* b_ih, const float* b_hh, float* P,
                            float* bias_scratch, int T, cudaStream_t stream) {
  if (small_k(K)) {
    int grid = T < num_sms() * 8 ? T : num_sms() * 8;
    const size_t smem = (size_t(K) * G4 + G4) * sizeof(float);
    RLT_CHECK_CUDA(cudaFuncSetAttribute(lstm_inproj_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    lstm_inproj_small_kernel<<<grid, 256, smem, stream>>>(x, K, w, b_ih, b_hh, P, 2 * G4, T);
    RLT_CHECK_LAUNCH();
    return RLT_OK;
  }
  // bias = b_ih + b_hh
  RLT_CHECK_CUDA(cudaMemcpyAsync(bias_scratch, b_ih, G4 * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  copy_add_kernel<<<(G4 + 255) / 256, 256, 0, stream>>>(b_hh, bias_scratch, G4);
  RLT_CHECK_LAUNCH();
  EpiParams ep{};
  ep.alpha = 1.f; ep.out = P; ep.ldo = 2 * G4; ep.bias = bias_scratch;
  return gemm_tn(x, K, w, K, T, G4, K, ep, stream);
}

}  // namespace rlt

using namespace rlt;

extern "C" {

size_t rlt_bilstm_saved_bytes(const rlt_bilstm_desc* d) {
  if (check_lstm(d) != RLT_OK) return 0;
  return lstm_saved(*d).total * sizeof(float);
}
size_t rlt_bilstm_workspace_bytes(const rlt_bilstm_desc* d) {
  if (check_lstm(d) != RLT_OK) return 0;
  return lstm_ws_floats(*d) * sizeof(float);
}

int rlt_bilstm_fwd(const rlt_bilstm_desc* d, const rlt_bilstm_weights* w, const float* x, float* y, void* saved_,
                   size_t saved_bytes, void* workspace, size_t workspace_bytes, rlt_stream_t stream_) {
  RLT_TRY(check_lstm(d));
  RLT_REQUIRE(w && x && y && workspace, RLT_INVALID_ARG, "bilstm fwd: null pointer");
  RLT_REQUIRE(workspace_bytes >= lstm_ws_floats(*d) * sizeof(float), RLT_WORKSPACE_TOO_SMALL,
              "bilstm fwd: workspace has %zu bytes, needs %zu", workspace_bytes, lstm_ws_floats(*d) * sizeof(float));
  const LstmSaved sl = lstm_saved(*d);
  float* saved = static_cast<float*>(saved_);
  RLT_REQUIRE(saved == nullptr || saved_bytes >= sl.total * sizeof(float), RLT_WORKSPACE_TOO_SMALL,
              "bilstm fwd: saved buffer has %zu bytes, needs %zu", saved_bytes, sl.total * sizeof(float));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int B = d->n_lists, L = d->seq_len, T = B * L, F = d->input_size;
  float* ws = static_cast<float*>(workspace);
  float* P = ws;
  float* y0_scratch = ws + size_t(T) * 2 * G4;              // layer-0 output when nothing is saved
  float* WT = y0_scratch + size_t(T) * 2 * H;
  float* bias_scratch = WT + 4 * size_t(H) * G4;
  float* y0 = saved ? saved + sl.y0 : y0_scratch;
  const bool tc = lstm_backend() == 0;
  if (!tc) {
    for (int l = 0; l < 2; ++l)
      for (int dir = 0; dir < 2; ++dir) {
        transpose_whh_kernel<<<(G4 * H + 255) / 256, 256, 0, stream>>>(w->w_hh[l][dir], WT + (l * 2 + dir) * size_t(H) * G4);
        RLT_CHECK_LAUNCH();
      }
  }
  for (int l = 0; l < 2; ++l) {
    const float* in = l == 0 ? x : y0;
    const int K = l == 0 ? F : 2 * H;
    const bool fused_in = tc && l == 0 && F <= 4;      // projection evaluated inside the recurrence: no P tensor
    if (fused_in) {
      // nothing to launch
    } else if (tc && !small_k(K)) {
      // both directions in ONE GEMM: [W_ih_fwd ; W_ih_rev] (1024 x K) and the summed biases are gathered into the
      // workspace (the transposed-W_hh region is unused on this backend), so the input is read once
      float* wcat = WT;
      RLT_REQUIRE(size_t(2) * G4 * K <= 4 * size_t(H) * G4, RLT_UNSUPPORTED_SHAPE, "bilstm: input width %d too large", K);
      for (int dir = 0; dir < 2; ++dir) {
        RLT_CHECK_CUDA(cudaMemcpyAsync(wcat + size_t(dir) * G4 * K, w->w_ih[l][dir], size_t(G4) * K * sizeof(float),
                                       cudaMemcpyDeviceToDevice, stream));
        RLT_CHECK_CUDA(cudaMemcpyAsync(bias_scratch + dir * G4, w->b_ih[l][dir], G4 * sizeof(float), cudaMemcpyDeviceToDevice, stream));
        copy_add_kernel<<<(G4 + 255) / 256, 256, 0, stream>>>(w->b_hh[l][dir], bias_scratch + dir * G4, G4);
        RLT_CHECK_LAUNCH();
      }
      EpiParams ep{};
      ep.alpha = 1.f; ep.out = P; ep.ldo = 2 * G4; ep.bias = bias_scratch;
      RLT_TRY(gemm_tn(in, K, wcat, K, T, 2 * G4, K, ep, stream));
    } else {
      for (int dir = 0; dir < 2; ++dir)
        RLT_TRY(input_projection(in, K, w->w_ih[l][dir], w->b_ih[l][dir], w->b_hh[l][dir], P + dir * G4, bias_scratch + dir * G4,
                                 T, stream));
    }
    float* out = l == 0 ? y0 : y;
    float* sv = saved ? saved + (l == 0 ? sl.s0 : sl.s1) : nullptr;
    time_begin(TAG_LSTM, stream);
    if (tc) {
      static DeviceOnce attr;
      if (attr.first()) {
        RLT_CHECK_CUDA(cudaFuncSetAttribute(lstm_um_fwd_kernel<false, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(LstmUmFwdSmem<64>::TOTAL)));
        RLT_CHECK_CUDA(cudaFuncSetAttribute(lstm_um_fwd_kernel<true, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(LstmUmFwdSmem<64>::TOTAL)));
        RLT_CHECK_CUDA(cudaFuncSetAttribute(lstm_um_fwd_kernel<false, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(LstmUmFwdSmem<32>::TOTAL)));
        RLT_CHECK_CUDA(cudaFuncSetAttribute(lstm_um_fwd_kernel<true, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(LstmUmFwdSmem<32>::TOTAL)));
      }
      LstmInProj inp{};
      if (fused_in) {
        inp.x = x; inp.F = F;
        for (int dir = 0; dir < 2; ++dir) { inp.w_ih[dir] = w->w_ih[0][dir]; inp.b_ih[dir] = w->b_ih[0][dir]; inp.b_hh[dir] = w->b_hh[0][dir]; }
      }
      const float* Pin = fused_in ? nullptr : P;
#define RLT_LSTM_FWD(FUSED, TILE_)                                                                                       \
  lstm_um_fwd_kernel<FUSED, TILE_><<<dim3((B + TILE_ - 1) / TILE_, 2), U_THREADS, LstmUmFwdSmem<TILE_>::TOTAL, stream>>>( \
      Pin, inp, w->w_hh[l][0], w->w_hh[l][1], out, sv, B, L)
      if (lstm_tile(B) == 32) { if (fused_in) RLT_LSTM_FWD(true, 32); else RLT_LSTM_FWD(false, 32); }
      else { if (fused_in) RLT_LSTM_FWD(true, 64); else RLT_LSTM_FWD(false, 64); }
#undef RLT_LSTM_FWD
    } else {
      lstm_rec_fwd_kernel<<<dim3(B, 2), H, 0, stream>>>(P, WT + (l * 2) * size_t(H) * G4, WT + (l * 2 + 1) * size_t(H) * G4,
                                                        out, sv, L);
    }
    time_end(TAG_LSTM, stream);
    RLT_CHECK_LAUNCH();
  }
  return RLT_OK;
}

int rlt_bilstm_bwd(const rlt_bilstm_desc* d, const rlt_bilstm_weights* w, const rlt_bilstm_grads* g, const float* x,
                   const void* saved_, const float* dy, float* dx, void* workspace, size_t workspace_bytes,
                   rlt_stream_t stream_) {
  RLT_TRY(check_lstm(d));
  RLT_REQUIRE(w && g && x && saved_ && dy && workspace, RLT_INVALID_ARG, "bilstm bwd: null pointer");
  RLT_REQUIRE(workspace_bytes >= lstm_ws_floats(*d) * sizeof(float), RLT_WORKSPACE_TOO_SMALL,
              "bilstm bwd: workspace has %zu bytes, needs %zu", workspace_bytes, lstm_ws_floats(*d) * sizeof(float));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const LstmSaved sl = lstm_saved(*d);
  const float* saved = static_cast<const float*>(saved_);
  const int B = d->n_lists, L = d->seq_len, T = B * L, F = d->input_size;
  float* ws = static_cast<float*>(workspace);
  float* dA = ws;                                   // [T, 1024]
  float* dY0 = ws + size_t(T) * 2 * G4;             // [T, 256]
  float* colsum_scratch = dY0 + size_t(T) * 2 * H + 4 * size_t(H) * G4;  // [1024]
  for (int l = 1; l >= 0; --l) {
    const float* sv = saved + (l == 0 ? sl.s0 : sl.s1);
    const float* dout = l == 1 ? dy : dY0;
    time_begin(TAG_LSTM, stream);
    if (lstm_backend() == 0) {
      // power-of-two scale that brings the incoming gradient into fp16's normal range (unscaled on read-back)
      unsigned int* amax = reinterpret_cast<unsigned int*>(colsum_scratch + 2 * G4);
      float* scale = colsum_scratch + 2 * G4 + 4;
      RLT_TRY(grad_scale(dout, size_t(T) * 2 * H, amax, scale, 6, stream));
      // bias gradients (column sums of dA) are accumulated by the recurrence itself
      RLT_CHECK_CUDA(cudaMemsetAsync(colsum_scratch, 0, 2 * G4 * sizeof(float), stream));
      static DeviceOnce attr;
      if (attr.first()) {
        RLT_CHECK_CUDA(cudaFuncSetAttribute(lstm_um_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(LstmUmBwdSmem<64>::TOTAL)));
        RLT_CHECK_CUDA(cudaFuncSetAttribute(lstm_um_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(LstmUmBwdSmem<32>::TOTAL)));
      }
      if (lstm_tile(B) == 32)
        lstm_um_bwd_kernel<32><<<dim3((B + 31) / 32, 2), U_THREADS, LstmUmBwdSmem<32>::TOTAL, stream>>>(
            dout, sv, w->w_hh[l][0], w->w_hh[l][1], scale, dA, colsum_scratch, B, L);
      else
        lstm_um_bwd_kernel<64><<<dim3((B + 63) / 64, 2), U_THREADS, LstmUmBwdSmem<64>::TOTAL, stream>>>(
            dout, sv, w->w_hh[l][0], w->w_hh[l][1], scale, dA, colsum_scratch, B, L);
    } else {
      lstm_rec_bwd_kernel<<<dim3(B, 2), H, 0, stream>>>(dout, sv, w->w_hh[l][0], w->w_hh[l][1], dA, L);
    }
    time_end(TAG_LSTM, stream);
    RLT_CHECK_LAUNCH();
    // biases: db_ih = db_hh = column sums of dA
    if (lstm_backend() != 0) {
      RLT_CHECK_CUDA(cudaMemsetAsync(colsum_scratch, 0, 2 * G4 * sizeof(float), stream));
      RLT_TRY(colsum(dA, colsum_scratch, T, 2 * G4, stream));
    }
    for (int dir = 0; dir < 2; ++dir) {
      copy_add_kernel<<<(G4 + 255) / 256, 256, 0, stream>>>(colsum_scratch + dir * G4, g->b_ih[l][dir], G4);
      RLT_CHECK_LAUNCH();
      copy_add_kernel<<<(G4 + 255) / 256, 256, 0, stream>>>(colsum_scratch + dir * G4, g->b_hh[l][dir], G4);
      RLT_CHECK_LAUNCH();
    }
    const float* in = l == 0 ? x : saved + sl.y0;
    const int K = l == 0 ? F : 2 * H;
    for (int dir = 0; dir < 2; ++dir) {
      const float* dAd = dA + dir * G4;
      // dW_hh += dA^T Hprev   (Hprev plane of the saved record; the two backends lay the record out differently)
      if (lstm_backend() == 0)
        RLT_TRY(gemm_dw(dAd, 2 * G4, sv + dir * U_REC + U_REC_HP, 2 * U_REC, T, G4, H, g->w_hh[l][dir], H, 1.f, stream));
      else
        RLT_TRY(gemm_dw(dAd, 2 * G4, sv + dir * (SAVE * H) + 5 * H, 2 * SAVE * H, T, G4, H, g->w_hh[l][dir], H, 1.f, stream));
      if (small_k(K)) {
        const int chunk = 256;
        const int grid = (T + chunk - 1) / chunk;
        if (K <= 4) lstm_dwih_small_kernel<4><<<grid, G4, 0, stream>>>(dAd, 2 * G4, in, K, g->w_ih[l][dir], T, chunk);
        else if (K <= 32) lstm_dwih_small_kernel<32><<<grid, G4, 0, stream>>>(dAd, 2 * G4, in, K, g->w_ih[l][dir], T, chunk);
        else lstm_dwih_small_kernel<64><<<grid, G4, 0, stream>>>(dAd, 2 * G4, in, K, g->w_ih[l][dir], T, chunk);
        RLT_CHECK_LAUNCH();
      } else {
        RLT_TRY(gemm_dw(dAd, 2 * G4, in, K, T, G4, K, g->w_ih[l][dir], K, 1.f, stream));
      }
    }
    // gradient w.r.t. the layer input
    if (l == 1) {
      EpiParams ep{};
      ep.alpha = 1.f; ep.out = dY0; ep.ldo = 2 * H;
      RLT_TRY(gemm_nn(dA, 2 * G4, w->w_ih[1][0], 2 * H, T, 2 * H, G4, ep, stream));
      ep.accumulate = 1;
      RLT_TRY(gemm_nn(dA + G4, 2 * G4, w->w_ih[1][1], 2 * H, T, 2 * H, G4, ep, stream));
    } else if (dx != nullptr) {
      if (small_k(K)) {
        lstm_dx_small_kernel<<<T, 128, 0, stream>>>(dA, w->w_ih[0][0], w->w_ih[0][1], K, dx, T);
        RLT_CHECK_LAUNCH();
      } else {
        EpiParams ep{};
        ep.alpha = 1.f; ep.out = dx; ep.ldo = K;
        RLT_TRY(gemm_nn(dA, 2 * G4, w->w_ih[0][0], K, T, K, G4, ep, stream));
        ep.accumulate = 1;
        RLT_TRY(gemm_nn(dA + G4, 2 * G4, w->w_ih[0][1], K, T, K, G4, ep, stream));
      }
    }
  }
  return RLT_OK;
}

}  // extern "C"
