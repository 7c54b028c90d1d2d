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
 *          "tma_round"    (1 = encode tensor maps as TFLOAT32 so TMA rounds operands on load),
 *          "lstm_backend" (0 = tcgen05 recurrence [default], 1 = plain validation kernels),
 *          A/B switches, all default 1: "b_resident", "f16out_tma", "ffn_bwd_fused" (one-pass dH + db1 + dW2 kernel,
 *          d_model 128; tag 5 then covers it and tag 7 is not emitted), "dw_colsum" (in-projection bias gradient as an
 *          extra MMA of the weight-gradient GEMM instead of a separate column-sum kernel). */
int rlt_set_option(const char* key, int value);
int rlt_get_option(const char* key);
/* Number of CUDA kernels this library has launched in this process (all entry points). */
unsigned long long rlt_launch_count(void);
/* In-situ timing of one kernel class: set option "time_tag" to a tag (1 qkv, 2 out-proj, 3 ffn1, 4 ffn2,
 * 5 d_ffn2, 6 d_ffn1, 7 dW_ffn2, 8 dW_ffn1, 9 attention fwd, 10 attention bwd, 11 fused ffn, 12 lstm);
 * matching launches are bracketed by CUDA events on their stream.  rlt_timing_read sums them. */
int rlt_timing_reset(void);
int rlt_timing_read(double* total_ms, int* count);
/* With time_tag = -1 every tagged call site is bracketed; this reads the sum for one site. */
int rlt_timing_read_tag(int tag, double* total_ms, int* count);

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
/* Half-precision building blocks of the FFN hidden path (fp16 keeps TF32's 11 significant bits; void* = __half*):
 * dst = half(src * scale[0]) (scale may be NULL); C = act(alpha A B^T + bias) with fp16 A [M,K], B [N,K];
 * C += alpha A^T B with fp16 A [T,M], B [T,N]; C(fp16) = act(A B^T + bias) with fp32 operands. */
int rlt_convert_f16(const float* src, void* dst, size_t n, const float* scale, rlt_stream_t stream);
int rlt_linear_f16(const void* A, const void* B, const float* bias, float* C, int M, int N, int K, float alpha, int relu,
                   rlt_stream_t stream);
int rlt_grad_weight_f16(const void* A, const void* B, float* C, int T, int M, int N, float alpha, rlt_stream_t stream);
int rlt_linear_out_f16(const float* A, const float* B, const float* bias, void* C, int M, int N, int K, int relu,
                       rlt_stream_t stream);
/* Test hook for train-mode dropout: out[i] = 0 or 1/(1-p) exactly as the kernels apply it at dropout site `site`
 * (1 attention probabilities, index (item*S + query)*S + key with S = group_size; 2 after the attention block;
 * 3 FFN hidden; 4 after the FFN; 5 BiCut logits; index = linear element index). */
int rlt_dropout_mask(uint64_t seed, int site, float p, size_t n, int group_size, float* out, rlt_stream_t stream);
/* out[c] += sum_t src[t, c]  (bias gradients). n_cols must be a multiple of 4. */
int rlt_colsum(const float* src, float* out, int n_rows, int n_cols, rlt_stream_t stream);

/* ------------------------------------------------------------------------------------------ */
/* optimizer step                                               */
/* ------------------------------------------------------------------------------------------ */
/* Replaces run.py:104,129  `optim.Adam(model.parameters(), lr, weight_decay)` / `optimizer.step()`
 * (torch/optim/adam.py _single_tensor_adam: L2 decay added to the gradient, no amsgrad, no maximize) with ONE launch
 * over all parameter tensors.  param_ptrs / grad_ptrs: DEVICE arrays [n_tensors] of device addresses (fp32 tensors);
 * exp_avg / exp_avg_sq: flat fp32 moment buffers, tensor t at state_offset[t]; the work list chunk_* (device int
 * arrays [n_chunks]) names for every CTA a tensor, its first element and element count (<= rlt_adam_chunk_elems()).
 * step counts from 1; grad_scale multiplies the gradient first (1/world after a SUM all-reduce). */
int rlt_adam_chunk_elems(void);
int rlt_adam_step(const unsigned long long* param_ptrs, const unsigned long long* grad_ptrs, float* exp_avg,
                  float* exp_avg_sq, const long long* state_offset, const int* chunk_tensor, const int* chunk_first,
                  const int* chunk_len, int n_chunks, double lr, double beta1, double beta2, double eps,
                  double weight_decay, long long step, double grad_scale, rlt_stream_t stream);
/* The same step with torch.optim.Adam's per-parameter bookkeeping on the DEVICE: tensor_steps [n_tensors] (int32,
 * zero-initialised by the caller) holds each tensor's own step count -- torch keeps `state['step']` per parameter and
 * does not touch a parameter whose `.grad` is None (no decay, no moment update, no step increment; run.py:121
 * `optimizer.zero_grad()` resets the gradients to None every batch, and the rerank head has none while its hinge is
 * inactive, utils/losses.py:141).  tensor_skip [n_tensors] (int32, may be NULL): non-zero = leave tensor t alone this
 * step.  Bias corrections are evaluated per tensor from its own count; no scalar depends on the host, so the launch
 * can be captured in a CUDA graph. */
int rlt_adam_step_masked(const unsigned long long* param_ptrs, const unsigned long long* grad_ptrs, float* exp_avg,
                         float* exp_avg_sq, const long long* state_offset, const int* chunk_tensor,
                         const int* chunk_first, const int* chunk_len, int n_chunks, int n_tensors, int* tensor_steps,
                         const int* tensor_skip, double lr, double beta1, double beta2, double eps, double weight_decay,
                         double grad_scale, rlt_stream_t stream);
/* tensor_skip[tensor_ids[i]] = 1 when NO group of `status` (rlt_aux_heads_loss) has its hinge bit (2) set, else 0:
 * the Engine's device-side replacement of "param.grad is None" for the rerank-head parameters. */
int rlt_adam_skip_from_status(const int32_t* status, int n_groups, const int* tensor_ids, int n_ids, int* tensor_skip,
                              rlt_stream_t stream);
/* Probe: TMA-load a [rows<=128, 32] fp32 tile of src through a TFLOAT32 tensor map and copy the
 * shared-memory image (de-swizzled) to dst.  Used once to learn whether TMA rounds or truncates. */
int rlt_probe_tma_tf32(const float* src, float* dst, int rows, rlt_stream_t stream);

/* C[M,N] = A[M,K] B[K,N] (B row-major [K,N]; the dX = dY W contraction of a Linear backward). */
int rlt_linear_nn(const float* A, const float* B, float* C, int M, int N, int K, rlt_stream_t stream);

/* ------------------------------------------------------------------------------------------ */
/* Transformer encoder layer (replaces torch.nn.TransformerEncoderLayer.forward/backward as built
 * at models/Choopy.py:11-12, AttnCut.py:9-10, MtChoopy.py:11-12, MtAttnCut.py:9-10,
 * MMOECut.py:9-10: post-norm, ReLU, no batch_first => attention ACROSS the lists of a group).    */
/* ------------------------------------------------------------------------------------------ */
typedef struct rlt_encoder_desc {
  int32_t n_groups;    /* G independent attention groups (one reference forward call each)          */
  int32_t group_size;  /* S lists per group (the reference's batch size B of one call)              */
  int32_t seq_len;     /* L positions per list                                                       */
  int32_t d_model;     /* 128 (Choopy family) or 256 (AttnCut family / experts)                      */
  int32_t n_head;
  int32_t d_ff;        /* 2048 (torch default dim_feedforward)                                       */
  int32_t attend_axis; /* 0 = across lists (reference); 1 = within a list (not implemented)          */
  int32_t accumulate_dx; /* backward: d_x += instead of d_x = (several experts share one input)       */
  float ln_eps;        /* 1e-5                                                                       */
  float dropout_p;     /* train-mode dropout probability (0 = eval / no dropout)                     */
  uint64_t dropout_seed;
  int32_t inference;   /* 1 = forward only (torch.no_grad / eval, run.py:166-167): nothing is kept for a backward --
                          the FFN hidden never leaves the chip                                             */
  int32_t reserved;
} rlt_encoder_desc;

/* nn.TransformerEncoderLayer parameters in state_dict order (all fp32 device pointers). */
typedef struct rlt_encoder_weights {
  const float* in_proj_w;  /* self_attn.in_proj_weight [3d, d]  */
  const float* in_proj_b;  /* self_attn.in_proj_bias   [3d]     */
  const float* out_proj_w; /* self_attn.out_proj.weight [d, d]  */
  const float* out_proj_b; /* self_attn.out_proj.bias   [d]     */
  const float* lin1_w;     /* linear1.weight [d_ff, d]          */
  const float* lin1_b;     /* linear1.bias   [d_ff]             */
  const float* lin2_w;     /* linear2.weight [d, d_ff]          */
  const float* lin2_b;     /* linear2.bias   [d]                */
  const float* norm1_w;    /* norm1.weight [d]                  */
  const float* norm1_b;
  const float* norm2_w;
  const float* norm2_b;
} rlt_encoder_weights;

/* Cross-list multi-head attention core of one encoder layer, forward (what rlt_encoder_layer_fwd runs between the QKV
 * projection and the output projection; torch functional.multi_head_attention_forward with the reference's layout, SURVEY
 * section 0): qkv [G*S*L, 3d] (q | k | v, token = (g*S + s)*L + l), o [G*S*L, d], lse [G*S*L, n_head] (natural-log
 * log-sum-exp of the scaled scores, may be NULL).  For every group g, position l and head h the S lists attend to each
 * other.  Head dim 16 and S <= 64 run on tcgen05 (csrc/attention_tc.cuh), other shapes on the mma.sync / generic kernels. */
int rlt_attention_lists_fwd(const float* qkv, float* o, float* lse, int n_groups, int group_size, int seq_len, int d_model,
                            int n_head, rlt_stream_t stream);
/* The feed-forward block of one encoder layer as ONE kernel (csrc/ffn_fwd_fused.cuh; what rlt_encoder_layer_fwd runs for
 * d_model 128 / 256 without dropout): out = LayerNorm(y + relu(y W1^T + b1) W2^T + b2), torch TransformerEncoderLayer._ff_block
 * + norm2 (models/Choopy.py:11 etc., dim_feedforward 2048).  y16 / w1_h / w2_h: fp16 copies of y [T, d], linear1.weight
 * [d_ff, d] and linear2.weight [d, d_ff] (row-major); y: the fp32 residual.  Optional outputs (NULL = forward only):
 * u2 [T, d] pre-norm sum, stats [T, 2] (mean, rstd), h_out [T, d_ff] fp16 hidden.  d must be 128 or 256, d_ff % 128 == 0,
 * d_ff <= 2048; RLT_UNSUPPORTED_SHAPE otherwise. */
/* Tools only: device buffer of 64 x 16 int64 that the next rlt_ffn_fused_fwd launches fill with clock64 stamps of the first
 * CTA pair's MMA thread and first epilogue warp (NULL switches the timeline off, the default). */
int rlt_ffn_fused_set_timeline(long long* device_buffer);
int rlt_ffn_fused_fwd(const void* y16, const float* y, const void* w1_h, const float* b1, const void* w2_h, const float* b2,
                      const float* gamma, const float* beta, float* out, float* u2, float* stats, void* h_out, int n_tokens,
                      int d_model, int d_ff, float ln_eps, rlt_stream_t stream);
/* gradient accumulators, same shapes; the backward ADDS into them (zero them or pass .grad). */
typedef struct rlt_encoder_grads {
  float* in_proj_w;
  float* in_proj_b;
  float* out_proj_w;
  float* out_proj_b;
  float* lin1_w;
  float* lin1_b;
  float* lin2_w;
  float* lin2_b;
  float* norm1_w;
  float* norm1_b;
  float* norm2_w;
  float* norm2_b;
} rlt_encoder_grads;

size_t rlt_encoder_layer_saved_bytes(const rlt_encoder_desc* desc);
size_t rlt_encoder_layer_workspace_bytes(const rlt_encoder_desc* desc);
/* x, out: [G*S*L, d].  `saved` receives the activations the backward needs. */
int rlt_encoder_layer_fwd(const rlt_encoder_desc* desc, const rlt_encoder_weights* w, const float* x, float* out,
                          void* saved, size_t saved_bytes, rlt_stream_t stream);
/* d_x must not alias d_out.  workspace: rlt_encoder_layer_workspace_bytes(). */
int rlt_encoder_layer_bwd(const rlt_encoder_desc* desc, const rlt_encoder_weights* w, const rlt_encoder_grads* gw,
                          const float* x, const void* saved, const float* d_out, float* d_x, void* workspace,
                          size_t workspace_bytes, rlt_stream_t stream);

/* ------------------------------------------------------------------------------------------ */
/* K1 — 2-layer bidirectional LSTM, H = 128 (replaces torch.nn.LSTM forward/backward at models/Bicut.py:8-9,19,
 * AttnCut.py:8,17, MtAttnCut.py:8,22, MMOECut.py:63,88).  Pointers are indexed [layer][direction].   */
/* ------------------------------------------------------------------------------------------ */
typedef struct rlt_bilstm_desc {
  int32_t n_lists;
  int32_t seq_len;
  int32_t input_size;  /* F: 3 (robust04), 25 / 47 (mq2007); <= 64 or a multiple of 4 */
  int32_t hidden;      /* 128 */
  int32_t num_layers;  /* 2 */
  int32_t training;    /* reserved */
} rlt_bilstm_desc;
typedef struct rlt_bilstm_weights {
  const float* w_ih[2][2]; /* weight_ih_l{k}[_reverse] [512, F or 256] */
  const float* w_hh[2][2]; /* weight_hh_l{k}[_reverse] [512, 128]      */
  const float* b_ih[2][2]; /* bias_ih_l{k}[_reverse]   [512]           */
  const float* b_hh[2][2];
} rlt_bilstm_weights;
typedef struct rlt_bilstm_grads { /* accumulated into (+=) */
  float* w_ih[2][2];
  float* w_hh[2][2];
  float* b_ih[2][2];
  float* b_hh[2][2];
} rlt_bilstm_grads;
size_t rlt_bilstm_saved_bytes(const rlt_bilstm_desc* desc);
size_t rlt_bilstm_workspace_bytes(const rlt_bilstm_desc* desc);
/* x [B, L, F] -> y [B, L, 256] ([.., :128] forward direction, [.., 128:] reverse).  saved may be NULL (inference). */
int rlt_bilstm_fwd(const rlt_bilstm_desc* desc, const rlt_bilstm_weights* w, const float* x, float* y, void* saved,
                   size_t saved_bytes, void* workspace, size_t workspace_bytes, rlt_stream_t stream);
/* dx may be NULL when the input needs no gradient. */
int rlt_bilstm_bwd(const rlt_bilstm_desc* desc, const rlt_bilstm_weights* w, const rlt_bilstm_grads* g, const float* x,
                   const void* saved, const float* dy, float* dx, void* workspace, size_t workspace_bytes,
                   rlt_stream_t stream);

/* ------------------------------------------------------------------------------------------ */
/* K5 — MMOECut gates + towers (models/MMOECut.py:90-105): per-task softmax gate over the flattened LSTM
 * output, gate-weighted mixture of the experts, Linear(d,1) towers.  Returns tower LOGITS z [Tk, B, L].   */
/* ------------------------------------------------------------------------------------------ */
typedef struct rlt_moe_desc {
  int32_t n_lists;
  int32_t seq_len;
  int32_t d_lstm;    /* 256: features per position of the LSTM output (gate input = seq_len * d_lstm) */
  int32_t d_model;   /* expert width (256) */
  int32_t n_experts; /* <= 4 */
  int32_t n_tasks;   /* <= 3 */
} rlt_moe_desc;
/* w_gates: [Tk, L*d_lstm, E] (the ParameterList stacked); experts: E device pointers to [B, L, d];
 * tower_w [Tk, d], tower_b [Tk]; outputs gates [Tk, B, E] and z [Tk, B, L]. */
int rlt_moe_heads_fwd(const rlt_moe_desc* desc, const float* h_lstm, const float* w_gates, const float* const* experts,
                      const float* tower_w, const float* tower_b, float* gates, float* z, rlt_stream_t stream);
/* d_experts: E pointers, written; d_tower_w / d_tower_b / d_w_gates: accumulated (+=); d_h_lstm: written or
 * accumulated (accumulate_dh); dgate_scratch: [Tk, B, E] floats. */
int rlt_moe_heads_bwd(const rlt_moe_desc* desc, const float* h_lstm, const float* w_gates, const float* const* experts,
                      const float* tower_w, const float* gates, const float* dz, float* const* d_experts,
                      float* d_tower_w, float* d_tower_b, float* d_w_gates, float* d_h_lstm, int accumulate_dh,
                      float* dgate_scratch, rlt_stream_t stream);

/* ------------------------------------------------------------------------------------------ */
/* Choopy input assembly (models/Choopy.py:18-20, MtChoopy.py:24-25) and its table gradient.     */
/* ------------------------------------------------------------------------------------------ */
int rlt_choopy_embed_fwd(const float* score /*[B,L]*/, const float* pe /*[L,127]*/, float* x /*[B,L,128]*/,
                         int n_lists, int seq_len, rlt_stream_t stream);
int rlt_choopy_embed_bwd(const float* dx /*[B,L,128]*/, float* dpe /*[L,127], +=*/, int n_lists, int seq_len,
                         rlt_stream_t stream);

/* ------------------------------------------------------------------------------------------ */
/* Linear(d, 1) heads (decison_layer.0 / classi.0 / rerank / tower layers): z[h, t] = x[t] . w[h] + b[h] */
/* ------------------------------------------------------------------------------------------ */
int rlt_head_dots_fwd(const float* x /*[T,d]*/, const float* w /*[H,d]*/, const float* bias /*[H]*/,
                      float* z /*[H,T]*/, int n_tokens, int d, int n_heads, rlt_stream_t stream);
/* dx (+)= sum_h dz[h] w[h]; dw += dz^T x; db += sum dz. */
/* relu_gate != 0: x is a ReLU output and dx is masked by (x > 0) (BiCut: fc -> ReLU -> Linear(256, 2)). */
int rlt_head_dots_bwd(const float* x, const float* w, const float* dz /*[H,T]*/, float* dx, float* dw, float* db,
                      int n_tokens, int d, int n_heads, int accumulate_dx, int relu_gate, rlt_stream_t stream);
/* BiCut's nn.Softmax(dim=2) over the two classes: logit planes z [2, T] <-> probabilities o [T, 2]. */
int rlt_pair_softmax_fwd(const float* z, float* o, size_t n_tokens, float dropout_p, uint64_t dropout_seed,
                         rlt_stream_t stream);
int rlt_pair_softmax_bwd(const float* o, const float* d_o, float* dz, size_t n_tokens, float dropout_p,
                         uint64_t dropout_seed, rlt_stream_t stream);

/* softmax over the L positions of every list (nn.Softmax(dim=1) of the cut heads) and its backward */
int rlt_softmax_lists(const float* z, float* p, int n_lists, int seq_len, rlt_stream_t stream);
int rlt_softmax_lists_bwd(const float* p, const float* dp, float* dz, int n_lists, int seq_len, rlt_stream_t stream);

/* ------------------------------------------------------------------------------------------ */
/* K3 — cut losses.  Replaces the B x L Python reward loops and the loss arithmetic of
 * utils/losses.py:48-68 (ChoopyLoss), :71-96 (AttnCutLoss / RAML), :194-233 (DivLoss kl / js), with
 * rewards Metric_for_Loss.f1 / .dcg (utils/metrics.py:85-101).                                   */
/* ------------------------------------------------------------------------------------------ */
typedef struct rlt_cut_loss_desc {
  int32_t n_lists;
  int32_t seq_len;         /* <= 1024 */
  int32_t input_kind;      /* 0: `in` = logits (softmax fused, grad = dL/dlogits); 1: `in` = probabilities (grad = dL/dp) */
  int32_t loss_kind;       /* 0 ChoopyLoss, 1 AttnCutLoss (RAML), 2 DivLoss 'kl', 3 DivLoss 'js' */
  int32_t metric_dcg;      /* 0: F1 reward, 1: DCG reward */
  int32_t accumulate_loss; /* loss_out += instead of = */
  float tau;               /* reward temperature (0.95 RAML, 0.85 DivLoss augmented, 1.0 otherwise) */
  float grad_scale;        /* multiplies the gradient (1/B for the reference's batch mean) */
  float loss_scale;        /* multiplies the summed per-list losses (1/B) */
} rlt_cut_loss_desc;
/* probs_out [B,L] (optional), grad [B,L] (optional), loss_per_list [B] (optional scratch/outputs),
 * loss_out: device scalar (optional; deterministic single-CTA reduction of loss_per_list). */
int rlt_cut_loss(const rlt_cut_loss_desc* desc, const float* in, const float* labels, float* probs_out, float* grad,
                 float* loss_per_list, float* loss_out, rlt_stream_t stream);
/* The same criteria with the labels as the bit masks of rlt_pack_labels ([n_lists, ceil(seq_len/32)] words): 8 L + L/8
 * bytes of traffic per list instead of 12 L.  Logits in (input_kind 0), even seq_len, 8-byte aligned float arrays. */
int rlt_cut_loss_bits(const rlt_cut_loss_desc* desc, const float* in, const uint32_t* label_bits, float* probs_out, float* grad,
                      float* loss_per_list, float* loss_out, rlt_stream_t stream);
/* Upload the DCG coefficient tables built on the host with math.log(j+2, 2) (utils/metrics.py:7). */
int rlt_set_dcg_tables(const float* coef32_host, const double* term64_host, int n);

/* ------------------------------------------------------------------------------------------ */
/* K4 — cut selection + metrics.  Replaces run.py:131-142 (argmax / BiCut rule) and Metric.f1 / Metric.dcg
 * (utils/metrics.py:15-38) per list; the Python wrapper takes np.mean on the host.              */
/* ------------------------------------------------------------------------------------------ */
/* mode 0: probs [B,L]; mode 1 (BiCut): probs [B,L,2].  Outputs optional: k, count of relevant in the cut,
 * number of relevant in the list, F1 and DCG as float64 with numpy's arithmetic. */
int rlt_eval_cut(const float* probs, const float* labels, int n_lists, int seq_len, int mode, int32_t* k_out,
                 int32_t* count_out, int32_t* nrel_out, double* f1_out, double* dcg_out, rlt_stream_t stream);
/* mode 0 with the labels as the bit masks of rlt_pack_labels ([n_lists, ceil(seq_len/32)] words): 4 L + L/8 bytes per
 * list instead of 8 L.  Even seq_len, probs 8-byte aligned. */
int rlt_eval_cut_bits(const float* probs, const uint32_t* label_bits, int n_lists, int seq_len, int32_t* k_out,
                      int32_t* count_out, int32_t* nrel_out, double* f1_out, double* dcg_out, rlt_stream_t stream);

/* Same metrics for caller-supplied cut positions (the Metric.f1 / Metric.dcg signature, utils/metrics.py:15,26).
 * pyint_in[b] = 1 marks a k that was a Python int in the reference (float32 precision, run.py:135). */
int rlt_eval_given_k(const float* labels, const int32_t* k_in, const int32_t* pyint_in, int n_lists, int seq_len,
                     int32_t* count_out, int32_t* nrel_out, double* f1_out, double* dcg_out, rlt_stream_t stream);
/* ------------------------------------------------------------------------------------------ */
/* Row N3 - device-resident data path.  Replaces the host-side collation of a shuffled batch by torch's DataLoader
 * (dataloader/attncut_dataloader.py:82-87) and the two host->device copies of run.py:123-124.         */
/* ------------------------------------------------------------------------------------------ */
/* x_out[o] = x[index[o]] ([n_src, seq_len, n_features] float32 rows), labels_out[o] = labels[index[o]] from EITHER float32
 * labels [n_src, seq_len] OR bit masks label_bits [n_src, ceil(seq_len/32)] (bit i%32 of word i/32 = document i is
 * relevant), expanded to the reference's float32 {0., 1.}.  index == NULL: identity (the first n_out lists).
 * labels_out may be NULL.  status (device int32, zeroed by the caller) |= 1 when an index is out of range. */
int rlt_gather_lists(const float* x, const float* labels, const uint32_t* label_bits, const int64_t* index, int n_src,
                     int n_out, int seq_len, int n_features, float* x_out, float* labels_out, int32_t* status,
                     rlt_stream_t stream);
/* float32 {0., 1.} labels -> bit masks (layout above).  status |= 2 when a label is neither 0. nor 1. */
int rlt_pack_labels(const float* labels, int n_lists, int seq_len, uint32_t* label_bits, int32_t* status, rlt_stream_t stream);

/* Full-list metrics of the verify scripts, one value per list:
 *   dcg_out[b]  = Metric.taskr_metric's DCG_sample (utils/metrics.py:51-57): documents ordered by descending score
 *                 (ties keep list order), +-inv_log2[i] added left to right in float64; inv_log2[i] = 1/math.log2(i+2)
 *                 is supplied by the caller (seq_len doubles, device memory) so the table is the host libm's;
 *   auc_out[b]  = sklearn.metrics.roc_auc_score(labels[b], scores[b]) (utils/metrics.py:73) as the Mann-Whitney
 *                 statistic, auc_valid[b] = 0 for the one-class lists utils/metrics.py:72 skips.
 * Either output may be NULL.  The means over lists (utils/metrics.py:58, :76) are taken by the caller. */
int rlt_rank_metrics(const float* scores, const float* labels, const double* inv_log2, int n_lists, int seq_len,
                     double* dcg_out, double* auc_out, int32_t* auc_valid, rlt_stream_t stream);
/* Reward matrix r[b, j] = Metric_for_Loss.f1 / .dcg (label_b, k = j+1)  (utils/metrics.py:85-101). */
int rlt_reward_matrix(const float* labels, float* rewards, int n_lists, int seq_len, int metric_dcg,
                      rlt_stream_t stream);

/* ------------------------------------------------------------------------------------------ */
/* Auxiliary heads of MtCutLoss (utils/losses.py:164-191): BCELoss on the class head and RerankLoss
 * (utils/losses.py:99-141) on the rerank head, per group (both are batch-global in the reference). */
/* ------------------------------------------------------------------------------------------ */
typedef struct rlt_aux_loss_desc {
  int32_t n_groups;
  int32_t group_size;
  int32_t seq_len;
  int32_t rerank_softmax;  /* 1: rerank head output is softmax over positions (MMOECut TowerRerank) */
  int32_t class_probs;     /* 1: zc already holds sigmoid outputs and dzc is the gradient w.r.t. them (API boundary) */
  int32_t accumulate_loss;
  float margin;            /* 5e-4 */
  float class_weight;
  float rerank_weight;
  float grad_scale;        /* 1 / n_groups when averaging group losses */
  float loss_scale;
} rlt_aux_loss_desc;
/* zc / zr: class / rerank head LOGITS [G*S, L] (either may be NULL).  probs_c: sigmoid(zc) (optional);
 * out_r: softmax(zr) when rerank_softmax (required then).  status[g]: bit 0 set when a group has no relevant or
 * no irrelevant document (the reference raises RuntimeError there; the wrapper does too); bit 1 set when the rerank
 * hinge of the group is active (utils/losses.py:141 returns a gradient-free constant otherwise: the rerank head then
 * has no gradient at all and torch.optim.Adam skips its parameters, see rlt_adam_step_masked). */
int rlt_aux_heads_loss(const rlt_aux_loss_desc* desc, const float* zc, const float* zr, const float* labels,
                       float* probs_c, float* out_r, float* dzc, float* dzr, float* loss_group, int32_t* status,
                       float* loss_out, rlt_stream_t stream);

/* ------------------------------------------------------------------------------------------ */
/* BiCut head loss (utils/losses.py:11-45).                                                       */
/* ------------------------------------------------------------------------------------------ */
typedef struct rlt_bicut_loss_desc {
  int32_t n_lists;
  int32_t seq_len;
  int32_t input_kind;      /* 0: u = 2-class logits, 1: u = probabilities (reference API boundary) */
  int32_t metric_nci;      /* 1: the 'nci' weights (losses.py:39), 0: the alpha / r weights (:41) */
  int32_t accumulate_loss;
  float alpha;             /* 0.65 */
  float r;                 /* 0.0971134020 */
  float grad_scale;
  float loss_scale;
} rlt_bicut_loss_desc;
int rlt_bicut_loss(const rlt_bicut_loss_desc* desc, const float* u /*[B,L,2]*/, const float* labels,
                   float* probs_out, float* grad, float* loss_per_list, float* loss_out, rlt_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RLT_B200_H_ */
