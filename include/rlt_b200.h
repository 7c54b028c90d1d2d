/* rlt_b200 — C ABI of the B200-native hot path for ranked-list truncation.
 *
 * This is the drop-in boundary.  The reference (Woody5962/Ranked-List-Truncation) is pure
 * Python/PyTorch and has no FFI of its own; what it exposes for this path is its Python API
 * (models/*.py forward, utils/losses.py criteria, utils/metrics.py Metric.f1/dcg, consumed by
 * run.py:59-145).  Every entry point below names the reference code it replaces (file:line under
 * the reference tree).  The Python mirror (ranked-list-truncation_b200/models, utils) binds these
 * with ctypes from torch.autograd.Function wrappers; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch / C++ types.  All tensor pointers are DEVICE pointers to
 *     dense row-major fp32 unless stated otherwise; the caller owns every buffer (inputs, outputs,
 *     saved-for-backward and workspace).  The library never allocates device memory.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant, and does
 *     not synchronise with the host.
 *   - return value: RLT_OK (0) or a negative rlt_status; a human-readable message for the calling
 *     thread is available from rlt_last_error().
 *   - "lists" are ranked lists (queries); a "group" is the set of S lists that one reference
 *     forward call sees together (the reference attends ACROSS the lists of a call because its
 *     encoder layers are built without batch_first).  Tokens are laid out
 *     [G*S lists][L positions][d features], exactly the reference's [B, L, d].
 */
#ifndef RLT_B200_H_
#define RLT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum rlt_status {
  RLT_OK = 0,
  RLT_INVALID_ARG = -1,
  RLT_UNSUPPORTED_SHAPE = -2,
  RLT_WORKSPACE_TOO_SMALL = -3,
  RLT_CUDA_ERROR = -4
} rlt_status;

typedef void* rlt_stream_t; /* cudaStream_t */

/* ------------------------------------------------------------------------------------------ */
/* library                                                                                    */
/* ------------------------------------------------------------------------------------------ */
const char* rlt_version(void);
const char* rlt_last_error(void);
/* options: "gemm_backend" (0 = tcgen05 tensor cores [default], 1 = SIMT validation kernels),
 *          "tma_round"    (1 = encode tensor maps as TFLOAT32 so TMA rounds operands on load). */
int rlt_set_option(const char* key, int value);
int rlt_get_option(const char* key);

/* ------------------------------------------------------------------------------------------ */
/* building blocks exported for tests and probes                                              */
/* ------------------------------------------------------------------------------------------ */
/* dst = tf32-rounded (round-to-nearest, ties away) copy of src; n elements. */
int rlt_round_tf32(const float* src, float* dst, size_t n, rlt_stream_t stream);
/* dst[c, r] = tf32(src[r, c]) for a [rows, cols] matrix (pre-transposed weight operand). */
int rlt_transpose_round_tf32(const float* src, float* dst, int rows, int cols, rlt_stream_t stream);
/* C[M,N] = act(alpha * A[M,K] B[N,K]^T + bias[N]); act = relu when relu != 0. (torch F.linear) */
int rlt_linear(const float* A, const float* B, const float* bias, float* C, int M, int N, int K, float alpha,
               int relu, rlt_stream_t stream);
/* C[M,N] += alpha * A[T,M]^T B[T,N]  (weight-gradient contraction over tokens). */
int rlt_grad_weight(const float* A, const float* B, float* C, int T, int M, int N, float alpha,
                    rlt_stream_t stream);
/* Probe: TMA-load a [rows<=128, 32] fp32 tile of src through a TFLOAT32 tensor map and copy the
 * shared-memory image (de-swizzled) to dst.  Used once to learn whether TMA rounds or truncates. */
int rlt_probe_tma_tf32(const float* src, float* dst, int rows, rlt_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RLT_B200_H_ */
