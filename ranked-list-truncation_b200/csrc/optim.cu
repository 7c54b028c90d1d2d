// Row N1 of SURVEY.md section 8(f): the optimizer step that follows the hot path.
// run.py:104 builds `optim.Adam(model.parameters(), lr, weight_decay)` (L2 decay folded into the gradient, not AdamW;
// amsgrad off) and run.py:129 calls `optimizer.step()` once per batch.  torch walks ~100 small tensors with a dozen
// foreach launches; here ONE launch updates every parameter tensor of the model: the parameters stay where torch
// allocated them (pointer table), the gradients are read where the backward left them (the flat all-reduced bucket of
// the Engine, or the modules' .grad tensors), the two moment buffers are flat.
// Arithmetic follows torch/optim/adam.py `_single_tensor_adam` operation by operation in fp32:
//     g   = g * grad_scale                      (1 / world size after a SUM all-reduce; 1 otherwise)
//     g   = g + wd * p
//     m   = m + (g - m) * (1 - beta1)           (Tensor.lerp_)
//     v   = v * beta2 + (1 - beta2) * g * g     (mul_ + addcmul_)
//     den = sqrt(v) / sqrt(1 - beta2^t) + eps
//     p   = p - (lr / (1 - beta1^t)) * m / den  (addcdiv_)
// The two bias-correction scalars are computed by the host in double, as torch does, and passed as floats.
#include "common.h"

namespace rlt {

constexpr int kAdamChunk = 4096;   // elements per CTA (256 threads x 4 float4)

struct AdamScalars {
  float step_size, bc2_sqrt, one_minus_beta1, beta2, one_minus_beta2, eps, weight_decay, grad_scale;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamScalars& a) {
  g = __fmul_rn(g, a.grad_scale);
  g = __fadd_rn(g, __fmul_rn(a.weight_decay, p));
  m = __fadd_rn(m, __fmul_rn(a.one_minus_beta1, __fsub_rn(g, m)));
  v = __fadd_rn(__fmul_rn(v, a.beta2), __fmul_rn(__fmul_rn(a.one_minus_beta2, g), g));
  const float den = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), a.bc2_sqrt), a.eps);
  p = __fadd_rn(p, __fmul_rn(-a.step_size, __fdiv_rn(m, den)));
}

__global__ void __launch_bounds__(256)
adam_step_kernel(const unsigned long long* __restrict__ param_ptrs, const unsigned long long* __restrict__ grad_ptrs,
                 float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq, const long long* __restrict__ state_offset,
                 const int* __restrict__ chunk_tensor, const int* __restrict__ chunk_first,
                 const int* __restrict__ chunk_len, AdamScalars a) {
  const int c = blockIdx.x;
  const int t = chunk_tensor[c], first = chunk_first[c], len = chunk_len[c];
  float* p = reinterpret_cast<float*>(param_ptrs[t]) + first;
  const float* g = reinterpret_cast<const float*>(grad_ptrs[t]) + first;
  float* m = exp_avg + state_offset[t] + first;
  float* v = exp_avg_sq + state_offset[t] + first;
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  if (vec) {
    const int n4 = len >> 2;
    for (int i = threadIdx.x; i < n4; i += 256) {
      float4 pp = reinterpret_cast<float4*>(p)[i];
      const float4 gg = reinterpret_cast<const float4*>(g)[i];
      float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
      adam_one(pp.x, gg.x, mm.x, vv.x, a); adam_one(pp.y, gg.y, mm.y, vv.y, a);
      adam_one(pp.z, gg.z, mm.z, vv.z, a); adam_one(pp.w, gg.w, mm.w, vv.w, a);
      reinterpret_cast<float4*>(p)[i] = pp;
      reinterpret_cast<float4*>(m)[i] = mm;
      reinterpret_cast<float4*>(v)[i] = vv;
    }
    for (int i = (n4 << 2) + threadIdx.x; i < len; i += 256) adam_one(p[i], g[i], m[i], v[i], a);
  } else {
    for (int i = threadIdx.x; i < len; i += 256) adam_one(p[i], g[i], m[i], v[i], a);
  }
}

// ---- per-tensor bookkeeping on the device (rlt_adam_step_masked)
struct AdamHyper {
  double lr, beta1, beta2;
  float one_minus_beta1, beta2_f, one_minus_beta2, eps, weight_decay, grad_scale;
};

__global__ void __launch_bounds__(256)
adam_step_masked_kernel(const unsigned long long* __restrict__ param_ptrs, const unsigned long long* __restrict__ grad_ptrs,
                        float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq, const long long* __restrict__ state_offset,
                        const int* __restrict__ chunk_tensor, const int* __restrict__ chunk_first,
                        const int* __restrict__ chunk_len, const int* __restrict__ tensor_steps,
                        const int* __restrict__ tensor_skip, AdamHyper h) {
  const int c = blockIdx.x;
  const int t = chunk_tensor[c], first = chunk_first[c], len = chunk_len[c];
  if (tensor_skip != nullptr && tensor_skip[t] != 0) return;       // torch: `if p.grad is None: continue`
  __shared__ AdamScalars sa;
  if (threadIdx.x == 0) {
    // torch/optim/adam.py: bias_correction{1,2} = 1 - beta ** step in double; step_size = lr / bias_correction1
    const double step = double(tensor_steps[t] + 1);
    AdamScalars a;
    a.step_size = float(h.lr / (1.0 - pow(h.beta1, step)));
    a.bc2_sqrt = float(sqrt(1.0 - pow(h.beta2, step)));
    a.one_minus_beta1 = h.one_minus_beta1; a.beta2 = h.beta2_f; a.one_minus_beta2 = h.one_minus_beta2;
    a.eps = h.eps; a.weight_decay = h.weight_decay; a.grad_scale = h.grad_scale;
    sa = a;
  }
  __syncthreads();
  const AdamScalars a = sa;
  float* p = reinterpret_cast<float*>(param_ptrs[t]) + first;
  const float* g = reinterpret_cast<const float*>(grad_ptrs[t]) + first;
  float* m = exp_avg + state_offset[t] + first;
  float* v = exp_avg_sq + state_offset[t] + first;
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  if (vec) {
    const int n4 = len >> 2;
    for (int i = threadIdx.x; i < n4; i += 256) {
      float4 pp = reinterpret_cast<float4*>(p)[i];
      const float4 gg = reinterpret_cast<const float4*>(g)[i];
      float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
      adam_one(pp.x, gg.x, mm.x, vv.x, a); adam_one(pp.y, gg.y, mm.y, vv.y, a);
      adam_one(pp.z, gg.z, mm.z, vv.z, a); adam_one(pp.w, gg.w, mm.w, vv.w, a);
      reinterpret_cast<float4*>(p)[i] = pp;
      reinterpret_cast<float4*>(m)[i] = mm;
      reinterpret_cast<float4*>(v)[i] = vv;
    }
    for (int i = (n4 << 2) + threadIdx.x; i < len; i += 256) adam_one(p[i], g[i], m[i], v[i], a);
  } else {
    for (int i = threadIdx.x; i < len; i += 256) adam_one(p[i], g[i], m[i], v[i], a);
  }
}

// runs after the update (stream order): every tensor that was stepped counts one more step
__global__ void adam_commit_steps_kernel(int* __restrict__ tensor_steps, const int* __restrict__ tensor_skip, int n) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n && (tensor_skip == nullptr || tensor_skip[t] == 0)) tensor_steps[t] += 1;
}

__global__ void adam_skip_from_status_kernel(const int32_t* __restrict__ status, int n_groups, const int* __restrict__ tensor_ids,
                                             int n_ids, int* __restrict__ tensor_skip) {
  __shared__ int any_active;
  if (threadIdx.x == 0) any_active = 0;
  __syncthreads();
  int a = 0;
  for (int g = threadIdx.x; g < n_groups; g += blockDim.x) a |= (status[g] & 2);
  if (a) atomicOr(&any_active, 1);
  __syncthreads();
  for (int i = threadIdx.x; i < n_ids; i += blockDim.x) tensor_skip[tensor_ids[i]] = any_active ? 0 : 1;
}

}  // namespace rlt

using namespace rlt;

extern "C" {

int rlt_adam_chunk_elems(void) { return kAdamChunk; }

int rlt_adam_step(const unsigned long long* param_ptrs, const unsigned long long* grad_ptrs, float* exp_avg,
                  float* exp_avg_sq, const long long* state_offset, const int* chunk_tensor, const int* chunk_first,
                  const int* chunk_len, int n_chunks, double lr, double beta1, double beta2, double eps,
                  double weight_decay, long long step, double grad_scale, rlt_stream_t stream_) {
  RLT_REQUIRE(param_ptrs && grad_ptrs && exp_avg && exp_avg_sq && state_offset && chunk_tensor && chunk_first && chunk_len,
              RLT_INVALID_ARG, "rlt_adam_step: null pointer");
  RLT_REQUIRE(step >= 1, RLT_INVALID_ARG, "rlt_adam_step: step %lld must count from 1", step);
  RLT_REQUIRE(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.0 && lr >= 0.0 && weight_decay >= 0.0,
              RLT_INVALID_ARG, "rlt_adam_step: invalid hyper-parameter (torch.optim.Adam raises ValueError for the same)");
  if (n_chunks <= 0) return RLT_OK;
  // torch/optim/adam.py: bias corrections and step size as Python floats (double), applied as fp32 scalars
  const double bc1 = 1.0 - pow(beta1, double(step));
  const double bc2 = 1.0 - pow(beta2, double(step));
  AdamScalars a;
  a.step_size = float(lr / bc1);
  a.bc2_sqrt = float(sqrt(bc2));
  a.one_minus_beta1 = float(1.0 - beta1);   // Python evaluates `1 - beta1` in double before the scalar becomes fp32
  a.beta2 = float(beta2);
  a.one_minus_beta2 = float(1.0 - beta2);
  a.eps = float(eps);
  a.weight_decay = float(weight_decay);
  a.grad_scale = float(grad_scale);
  adam_step_kernel<<<n_chunks, 256, 0, static_cast<cudaStream_t>(stream_)>>>(param_ptrs, grad_ptrs, exp_avg, exp_avg_sq,
                                                                             state_offset, chunk_tensor, chunk_first,
                                                                             chunk_len, a);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

int rlt_adam_step_masked(const unsigned long long* param_ptrs, const unsigned long long* grad_ptrs, float* exp_avg,
                         float* exp_avg_sq, const long long* state_offset, const int* chunk_tensor,
                         const int* chunk_first, const int* chunk_len, int n_chunks, int n_tensors, int* tensor_steps,
                         const int* tensor_skip, double lr, double beta1, double beta2, double eps, double weight_decay,
                         double grad_scale, rlt_stream_t stream_) {
  RLT_REQUIRE(param_ptrs && grad_ptrs && exp_avg && exp_avg_sq && state_offset && chunk_tensor && chunk_first && chunk_len &&
                  tensor_steps, RLT_INVALID_ARG, "rlt_adam_step_masked: null pointer");
  RLT_REQUIRE(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.0 && lr >= 0.0 && weight_decay >= 0.0,
              RLT_INVALID_ARG, "rlt_adam_step_masked: invalid hyper-parameter (torch.optim.Adam raises ValueError for the same)");
  if (n_chunks <= 0 || n_tensors <= 0) return RLT_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  AdamHyper h;
  h.lr = lr; h.beta1 = beta1; h.beta2 = beta2;
  h.one_minus_beta1 = float(1.0 - beta1);
  h.beta2_f = float(beta2);
  h.one_minus_beta2 = float(1.0 - beta2);
  h.eps = float(eps);
  h.weight_decay = float(weight_decay);
  h.grad_scale = float(grad_scale);
  adam_step_masked_kernel<<<n_chunks, 256, 0, stream>>>(param_ptrs, grad_ptrs, exp_avg, exp_avg_sq, state_offset, chunk_tensor,
                                                        chunk_first, chunk_len, tensor_steps, tensor_skip, h);
  RLT_CHECK_LAUNCH();
  adam_commit_steps_kernel<<<(n_tensors + 255) / 256, 256, 0, stream>>>(tensor_steps, tensor_skip, n_tensors);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

int rlt_adam_skip_from_status(const int32_t* status, int n_groups, const int* tensor_ids, int n_ids, int* tensor_skip,
                              rlt_stream_t stream_) {
  RLT_REQUIRE(status && tensor_ids && tensor_skip && n_groups > 0 && n_ids > 0, RLT_INVALID_ARG,
              "rlt_adam_skip_from_status: bad arguments");
  adam_skip_from_status_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream_)>>>(status, n_groups, tensor_ids, n_ids, tensor_skip);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

}  // extern "C"
