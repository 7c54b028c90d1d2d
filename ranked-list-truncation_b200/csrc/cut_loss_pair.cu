// K3, packed kernels (cut_loss_pair.cuh) in their own translation unit: 64 instantiations (4 list-length classes x 4
// criteria x 2 rewards x 2 label formats) compile next to heads.cu instead of inside it.
#include <stdint.h>

#include "common.h"
#include "warp_utils.cuh"
#include "cut_loss_pair.cuh"

namespace rlt {

// deterministic single-CTA sum of the per-list losses: the same reduction order as heads.cu's reduce_scale_kernel, so the
// two label formats give bit-identical batch losses
__global__ void __launch_bounds__(256) pair_reduce_scale_kernel(const float* __restrict__ v, int n, float scale,
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

// the packed kernel: one instantiation per (loss kind, metric, label format)
#define RLT_K3P(LK, MD)                                                                                                  \
  case (LK) * 2 + (MD):                                                                                                  \
    cut_loss_pair_kernel<NP, LK, MD, kBits><<<grid, 128, 0, stream>>>(in, labels, probs_out, grad, loss_per_list, B, L, tau, \
                                                                      gscale);                                           \
    break;
template <int NP, bool kBits>
static void launch_cut_loss_pair(int cfg, int grid, cudaStream_t stream, const float* in, const void* labels, float* probs_out,
                                 float* grad, float* loss_per_list, int B, int L, float tau, float gscale) {
  switch (cfg) {
    RLT_K3P(0, 0) RLT_K3P(0, 1) RLT_K3P(1, 0) RLT_K3P(1, 1) RLT_K3P(2, 0) RLT_K3P(2, 1) RLT_K3P(3, 0) RLT_K3P(3, 1)
    default: break;
  }
}
template <bool kBits>
static int cut_loss_pair(const rlt_cut_loss_desc* c, const float* in, const void* labels, float* probs_out, float* grad,
                         float* loss_per_list, cudaStream_t stream) {
  const int B = c->n_lists, L = c->seq_len, grid = (B + 3) / 4;
  const int cfg = c->loss_kind * 2 + (c->metric_dcg ? 1 : 0);
#define RLT_K3P_NP(NP_) launch_cut_loss_pair<NP_, kBits>(cfg, grid, stream, in, labels, probs_out, grad, loss_per_list, B, L, c->tau, c->grad_scale)
  if (L <= 64) RLT_K3P_NP(1);
  else if (L <= 320) RLT_K3P_NP(5);
  else if (L <= 512) RLT_K3P_NP(8);
  else if (L <= 1024) RLT_K3P_NP(16);
  else return set_error(RLT_UNSUPPORTED_SHAPE, "list length %d exceeds 1024", L);
#undef RLT_K3P_NP
  return RLT_OK;
}
static bool aligned8(const void* a, const void* b, const void* c, const void* d) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
           reinterpret_cast<uintptr_t>(d)) & 7u) == 0;
}


bool cut_loss_pair_ok(int L, const void* in, const void* labels, const void* probs_out, const void* grad) {
  return L % 2 == 0 && aligned8(in, labels, probs_out, grad);
}
int cut_loss_pair_launch(const rlt_cut_loss_desc* c, const float* in, const void* labels, bool bits, float* probs_out, float* grad,
                         float* loss_per_list, cudaStream_t stream) {
  return bits ? cut_loss_pair<true>(c, in, labels, probs_out, grad, loss_per_list, stream)
              : cut_loss_pair<false>(c, in, labels, probs_out, grad, loss_per_list, stream);
}
int cut_loss_pair_set_rcoef(const float* rc_host, int n) {
  RLT_CHECK_CUDA(cudaMemcpyToSymbol(g_pair_rcoef32, rc_host, sizeof(float) * n));
  return RLT_OK;
}

}  // namespace rlt

using namespace rlt;

extern "C" {

int rlt_cut_loss_bits(const rlt_cut_loss_desc* c, const float* in, const uint32_t* label_bits, float* probs_out, float* grad,
                      float* loss_per_list, float* loss_out, rlt_stream_t stream_) {
  RLT_REQUIRE(c && in && label_bits, RLT_INVALID_ARG, "rlt_cut_loss_bits: null pointer");
  RLT_REQUIRE(c->n_lists > 0 && c->seq_len > 0, RLT_INVALID_ARG, "rlt_cut_loss_bits: n_lists=%d seq_len=%d", c->n_lists, c->seq_len);
  RLT_REQUIRE(c->loss_kind >= 0 && c->loss_kind <= 3, RLT_INVALID_ARG, "rlt_cut_loss_bits: loss_kind=%d", c->loss_kind);
  RLT_REQUIRE(c->input_kind == 0, RLT_INVALID_ARG, "rlt_cut_loss_bits: logits in only (input_kind 0)");
  RLT_REQUIRE(c->loss_kind == 0 || c->tau > 0.f, RLT_INVALID_ARG, "rlt_cut_loss_bits: tau must be positive");
  RLT_REQUIRE(c->seq_len % 2 == 0 && aligned8(in, probs_out, grad, nullptr), RLT_UNSUPPORTED_SHAPE,
              "rlt_cut_loss_bits: seq_len %d must be even and the float arrays 8-byte aligned", c->seq_len);
  RLT_REQUIRE(loss_out == nullptr || loss_per_list != nullptr, RLT_INVALID_ARG,
              "rlt_cut_loss_bits: loss_out needs the loss_per_list scratch");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RLT_TRY(cut_loss_pair<true>(c, in, label_bits, probs_out, grad, loss_per_list, stream));
  RLT_CHECK_LAUNCH();
  if (loss_out != nullptr) {
    pair_reduce_scale_kernel<<<1, 256, 0, stream>>>(loss_per_list, c->n_lists, c->loss_scale, loss_out, c->accumulate_loss);
    RLT_CHECK_LAUNCH();
  }
  return RLT_OK;
}

}  // extern "C"
