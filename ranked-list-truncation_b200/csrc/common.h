// Library-internal declarations shared by the translation units of librlt_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/rlt_b200.h"
#include "dropout.cuh"

namespace rlt {

struct EpiParams;

// ---- error plumbing (thread-local message behind rlt_last_error) ----
int set_error(int code, const char* fmt, ...);
#define RLT_CHECK_CUDA(expr)                                                                    \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return ::rlt::set_error(RLT_CUDA_ERROR, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                              __FILE__, __LINE__);                                              \
  } while (0)
// every kernel launch site goes through this: counts the launch (rlt_launch_count) and checks for errors
void note_launch();
#define RLT_CHECK_LAUNCH()             \
  do {                                 \
    ::rlt::note_launch();              \
    RLT_CHECK_CUDA(cudaGetLastError()); \
  } while (0)
#define RLT_REQUIRE(cond, code, ...)                          \
  do {                                                        \
    if (!(cond)) return ::rlt::set_error((code), __VA_ARGS__); \
  } while (0)
#define RLT_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc != RLT_OK) return _rc; \
  } while (0)

int num_sms();
// Per-device one-time setup (cudaFuncSetAttribute / occupancy queries are per device): `static DeviceOnce once;
// if (once.first()) {...}` runs the body once for every device a process uses.
struct DeviceOnce {
  bool done[64] = {};
  int value[64] = {};
  int dev() const { int d = 0; return (cudaGetDevice(&d) == cudaSuccess && d >= 0 && d < 64) ? d : 0; }
  bool first() { const int d = dev(); if (done[d]) return false; done[d] = true; return true; }
};

// In-situ kernel timing: when the "time_tag" option equals `tag`, the launch is bracketed by CUDA events on
// its own stream; rlt_timing_read() sums the elapsed times.  Tags name the GEMM call sites of the encoder.
enum KernelTag : int {
  TAG_NONE = 0, TAG_QKV = 1, TAG_OUT_PROJ = 2, TAG_FFN1 = 3, TAG_FFN2 = 4, TAG_D_FFN2 = 5, TAG_D_FFN1 = 6,
  TAG_DW_FFN2 = 7, TAG_DW_FFN1 = 8, TAG_ATTN_FWD = 9, TAG_ATTN_BWD = 10, TAG_FFN_FUSED = 11, TAG_LSTM = 12
};
void time_begin(int tag, cudaStream_t stream);
void time_end(int tag, cudaStream_t stream);

// ---- GEMM front-ends (gemm.cu).  Operands are fp32 containers; on the tensor-core path they
//      must hold tf32-rounded values unless tma_rounds() is true. ----
// C[M,N] = A[M,K] * B[N,K]^T with the fused epilogue described by EpiParams.
int gemm_tn(const float* A, int lda, const float* B, int ldb, int M, int N, int K, const EpiParams& ep,
            cudaStream_t stream);
// C[M,N] = A[M,K] * B[K,N]   (B row-major [K,N]: dX = dY W with the nn.Linear weight used as stored)
int gemm_nn(const float* A, int lda, const float* B, int ldb, int M, int N, int K, const EpiParams& ep,
            cudaStream_t stream);
// C[M,N] += alpha * sum_t A[t,m] * B[t,n]   (A: [T,lda], B: [T,ldb]; C pre-initialised by the caller)
// colsum_out (optional): colsum_out[m] += alpha * sum_t A[t, m] (the bias gradient), folded into the same pass over A
int gemm_dw(const float* A, int lda, const float* B, int ldb, int T, int M, int N, float* C, int ldc, float alpha,
            cudaStream_t stream, int tag = 0, float* colsum_out = nullptr);
int colsum(const float* src, float* out, int T, int C, cudaStream_t stream);   // encoder.cu

// fp16-operand variants (tensor-core backend only): C = A[M,K] B[N,K]^T and C += alpha * alpha_ptr[0] * A[T,M]^T B[T,N]
int gemm_tn_h(const __half* A, int lda, const __half* B, int ldb, int M, int N, int K, const EpiParams& ep,
              cudaStream_t stream);
bool attention_fwd_tc_ok(int S, int dh);
int attention_fwd_tc(const float* qkv, float* o, float* lse, int G, int S, int L, int d, int n_head, float scale,
                     cudaStream_t stream, int tag);
bool ffn_fwd_fused_ok(int d, int f);
void ffn_fwd_set_timeline(long long* dev_buf);
int ffn_fwd_fused(const __half* y16, const float* y, const __half* w1h, const float* b1, const __half* w2h, const float* b2,
                  const float* gamma, const float* beta, float* out, float* u2, float* stats, __half* h_out, int T, int d,
                  int f, float eps, cudaStream_t stream, int tag);
bool ffn_bwd_fused_ok(int d, int f);
int ffn_bwd_fused(const __half* du16, const __half* w2th, const __half* hh, __half* dh16, int T, int d, int f, float alpha,
                  const float* scale, float* db1, float* dW2, cudaStream_t stream, int tag);
int gemm_dw_h(const __half* A, int lda, const __half* B, int ldb, int T, int M, int N, float* C, int ldc, float alpha,
              const float* alpha_ptr, cudaStream_t stream, int tag = 0);
// dst = half(src * scale[0]) (scale may be null); dst[c, r] = half(src[r, c])
int convert_f16(const float* src, __half* dst, size_t n, const float* scale, cudaStream_t stream,
                DropCfg drop = DropCfg{0, 0, 1.f}, uint32_t site = 0);
// dst = src * keep-and-scale factor of dropout site `site` (element index = linear index)
int dropout_apply(const float* src, float* dst, size_t n, DropCfg drop, uint32_t site, cudaStream_t stream);
int transpose_f16(const float* src, __half* dst, int rows, int cols, cudaStream_t stream);
// scale = {2^k, 2^-k} with max|x| * 2^k in [2^(target-1), 2^target); amax_scratch: one device word
int grad_scale(const float* x, size_t n, unsigned int* amax_scratch, float* scale, int target, cudaStream_t stream);
// the same from an already reduced amax word (float bits)
int pow2_scale(const unsigned int* amax_bits, float* scale, int target, cudaStream_t stream);

// 0 = tcgen05 tensor-core kernels (default), 1 = plain SIMT kernels (validation backend only)
int gemm_backend();
// 0 = persistent tcgen05 LSTM recurrence (default), 1 = plain validation kernels (also forced by gemm_backend 1)
int lstm_backend();
void set_lstm_tile(int v);
int get_lstm_tile();
void set_lstm_backend(int v);
// true when the TMA tensor maps are encoded as TFLOAT32 (TMA rounds fp32->tf32 while loading), so
// producers need not materialise rounded operand copies.
bool tma_rounds();

// cut_loss_pair.cu: the packed K3 kernels (logits in, even L, 8-byte aligned rows); labels as floats or rlt_pack_labels words
bool cut_loss_pair_ok(int L, const void* in, const void* labels, const void* probs_out, const void* grad);
int cut_loss_pair_launch(const rlt_cut_loss_desc* c, const float* in, const void* labels, bool bits, float* probs_out, float* grad,
                         float* loss_per_list, cudaStream_t stream);
int cut_loss_pair_set_rcoef(const float* rc_host, int n);

}  // namespace rlt
