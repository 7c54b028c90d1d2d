// Library plumbing (errors, options, TMA tensor maps) and the GEMM front-ends.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <atomic>
#include <mutex>
#include <vector>

#include <cuda_fp16.h>

#include "common.h"
#include "gemm_tc.cuh"
#include "gemm_f16out.cuh"
#include "ffn_bwd_fused.cuh"
#include "ffn_fwd_fused.cuh"
#include "attention_tc.cuh"

namespace rlt {

// ------------------------------------------------------------------------------------------
// errors / options
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}
static int g_gemm_backend = env_int("RLT_GEMM_BACKEND", 0);
// Measured on B200 (tests/test_gemm_gpu.py::test_probe_tma_tfloat32_rounding): a TFLOAT32 tensor map makes
// TMA round fp32 -> tf32 to NEAREST while it fills shared memory (identical to cvt.rna.tf32.f32 except on
// exact ties), so operands stay exact fp32 in HBM and no rounded copies are materialised.  Without it the
// tensor core would truncate the low 13 mantissa bits (a systematic -2^-11 relative bias per operand).
static int g_tma_round = env_int("RLT_TMA_ROUND", 1);
static int g_b_resident = env_int("RLT_B_RESIDENT", 1);
static int g_ffn_bwd_fused = env_int("RLT_FFN_BWD_FUSED", 1);   // one-pass dH / dW2 / db1 kernel (d_model 128)
static int g_ffn_fwd_fused = env_int("RLT_FFN_FWD_FUSED", 1);   // fused FFN1 + ReLU + FFN2 + residual + LayerNorm2 (cta_group::2)
static int g_attn_tc = env_int("RLT_ATTN_TC", 1);               // tcgen05 attention forward (head dim 16, groups <= 64 lists)
static int g_dw_colsum = env_int("RLT_DW_COLSUM", 1);         // bias-gradient column sums as an extra MMA of the weight-gradient GEMM
static int g_f16out_tma = env_int("RLT_F16OUT_TMA", 1);     // copy-engine epilogue kernel for the fp16-output GEMMs   // gemm_tn: keep the CTA's B slice in shared memory when it fits

static std::atomic<unsigned long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static int g_time_tag = 0;
static std::mutex g_time_mu;
static std::vector<cudaEvent_t> g_time_events;  // begin/end pairs
static std::vector<int> g_time_tags;            // tag of every event
// time_tag option: 0 = off, t > 0 = bracket the launches of call site t, -1 = bracket every tagged call site
void time_begin(int tag, cudaStream_t stream) {
  if (tag == 0 || (tag != g_time_tag && g_time_tag != -1)) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, stream);
  std::lock_guard<std::mutex> lk(g_time_mu);
  g_time_events.push_back(e);
  g_time_tags.push_back(tag);
}
void time_end(int tag, cudaStream_t stream) { time_begin(tag, stream); }

int gemm_backend() { return g_gemm_backend; }
bool tma_rounds() { return g_tma_round != 0 && g_gemm_backend == 0; }

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// ------------------------------------------------------------------------------------------
// TMA tensor maps.  cuTensorMapEncodeTiled is fetched through the runtime so the library has no
// link-time dependency on libcuda.so.
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D row-major matrix [rows, cols] with leading dimension ld (elements); box = box_rows x (128 bytes of columns:
// 32 fp32 or 64 fp16), SWIZZLE_128B (or the 32-byte-atom variant for MN-major fp32 operands), zero fill outside.
static int make_tmap_any(CUtensorMap* out, const void* base, int elem_bytes, CUtensorMapDataType dtype, uint64_t rows,
                         uint64_t cols, uint64_t ld, uint32_t box_rows, bool atom32, uint32_t box_cols = 0,
                         CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = encode_fn();
  RLT_REQUIRE(fn != nullptr, RLT_CUDA_ERROR, "cuTensorMapEncodeTiled is not available from the driver");
  RLT_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * elem_bytes) % 16 == 0, RLT_INVALID_ARG,
              "TMA operand must be 16-byte aligned with a row pitch that is a multiple of 16 bytes "
              "(ptr=%p ld=%llu)", base, (unsigned long long)ld);
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {ld * elem_bytes};
  const cuuint32_t box[2] = {box_cols ? box_cols : cuuint32_t(128 / elem_bytes), box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = fn(out, dtype, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : swz,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RLT_REQUIRE(r == CUDA_SUCCESS, RLT_CUDA_ERROR, "cuTensorMapEncodeTiled failed with CUresult %d", int(r));
  return RLT_OK;
}
static int make_tmap(CUtensorMap* out, const float* base, uint64_t rows, uint64_t cols, uint64_t ld,
                     uint32_t box_rows, bool round_tf32, bool atom32 = false) {
  return make_tmap_any(out, base, 4, round_tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rows,
                       cols, ld, box_rows, atom32);
}
static int make_tmap_h(CUtensorMap* out, const __half* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  return make_tmap_any(out, base, 2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rows, cols, ld, box_rows, false);
}

// ------------------------------------------------------------------------------------------
// SIMT validation backend (same contracts, no tensor cores, exact fp32 FMA)
// ------------------------------------------------------------------------------------------
__global__ void gemm_tn_simt_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                    int M, int N, int K, EpiParams ep, int b_major_n) {
  __shared__ float sA[16][65];
  __shared__ float sB[16][65];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int r = i >> 4, c = i & 15;
      sA[c][r] = (m0 + r < M && k0 + c < K) ? A[size_t(m0 + r) * lda + k0 + c] : 0.f;
      sB[c][r] = (n0 + r < N && k0 + c < K)
                     ? (b_major_n ? B[size_t(k0 + c) * ldb + n0 + r] : B[size_t(n0 + r) * ldb + k0 + c])
                     : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sA[kk][ty * 4 + i]; b[i] = sB[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      const int r = m0 + ty * 4 + i, c = n0 + tx * 4 + j;
      if (r >= M || c >= N) continue;
      float alpha = ep.alpha;
      if (ep.scale_mode == 1 || ep.scale_mode == 2) alpha *= ep.scale_ptr[ep.scale_mode == 1 ? 0 : 1];
      float v = acc[i][j] * alpha;
      if (ep.bias) v += ep.bias[c];
      if (ep.relu) v = fmaxf(v, 0.f);
      const size_t off = size_t(r) * ep.ldo + c;
      if (ep.drop.thr) v *= drop_factor(drop_bits(ep.drop.seed, ep.drop_site, off >> 2), int(off & 3), ep.drop.thr, ep.drop.scale);
      if (ep.gate_src) v = ep.gate_src[off] > 0.f ? v : 0.f;
      if (ep.gate_h) v = __half2float(ep.gate_h[off]) > 0.f ? v : 0.f;
      if (ep.residual) v += ep.residual[off];
      if (ep.out_h) {
        ep.out_h[off] = __float2half_rn(v);
      } else {
        if (ep.accumulate) v += ep.out[off];
        ep.out[off] = v;
      }
      if (ep.colsum) atomicAdd(ep.colsum + c, (ep.scale_mode == 1 || ep.scale_mode == 3) ? v * ep.scale_ptr[1] : v);
    }
}

__global__ void gemm_dw_simt_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                    int T, int M, int N, float* __restrict__ C, int ldc, float alpha, int tchunk) {
  // one thread per (m, n) output, looping over a chunk of tokens; blockIdx.z selects the chunk
  const int n = blockIdx.x * 16 + (threadIdx.x & 15);
  const int m = blockIdx.y * 16 + (threadIdx.x >> 4);
  if (m >= M || n >= N) return;
  const int t0 = blockIdx.z * tchunk, t1 = min(T, t0 + tchunk);
  float acc = 0.f;
  for (int t = t0; t < t1; ++t) acc = fmaf(A[size_t(t) * lda + m], B[size_t(t) * ldb + n], acc);
  atomicAdd(C + size_t(m) * ldc + n, alpha * acc);
}

// ------------------------------------------------------------------------------------------
// front-ends
// ------------------------------------------------------------------------------------------
template <int BN, int OP, int EF>
static int launch_tn_ef(const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int K, const EpiParams& ep, int grid,
                        cudaStream_t stream) {
  using Cfg = GemmTnCfg<BN, OP>;
  // B-resident mode: every k-block of the CTA's B slice fits next to an A ring, and there is more than one row block
  const int num_kb = (K + (OP == OP_F16_K ? 64 : 32) - 1) / (OP == OP_F16_K ? 64 : 32);
  const int tiles_m = (M + Cfg::BM - 1) / Cfg::BM, tiles_n = N / BN;
  const int b_res = (g_b_resident != 0 && BN >= 128 && size_t(num_kb) * Cfg::B_BYTES <= size_t(Cfg::RES_B_BYTES) &&
                     tiles_m >= 2 * (grid / tiles_n)) ? 1 : 0;
  static DeviceOnce attr_set;
  if (attr_set.first()) {
    RLT_CHECK_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<BN, OP, EF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        int(Cfg::SMEM_BYTES)));
  }
  time_begin(ep.tag, stream);
  gemm_tn_kernel<BN, OP, EF><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, M, N, K, ep, b_res);
  time_end(ep.tag, stream);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

template <int BN, int OP>
static int launch_tn(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const EpiParams& ep,
                     cudaStream_t stream) {
  using Cfg = GemmTnCfg<BN, OP>;
  RLT_REQUIRE(ep.out != nullptr || ep.out_h != nullptr, RLT_INVALID_ARG, "gemm: null output");
  RLT_REQUIRE(ep.scale_mode == 0 || ep.scale_ptr != nullptr, RLT_INVALID_ARG, "gemm: scale_mode without scale_ptr");
  CUtensorMap tmA, tmB;
  if (OP == OP_F16_K) {
    RLT_TRY(make_tmap_h(&tmA, static_cast<const __half*>(A), M, K, lda, Cfg::BM));
    RLT_TRY(make_tmap_h(&tmB, static_cast<const __half*>(B), N, K, ldb, BN));
  } else {
    RLT_TRY(make_tmap(&tmA, static_cast<const float*>(A), M, K, lda, Cfg::BM, tma_rounds()));
    if (OP == OP_TF32_N) RLT_TRY(make_tmap(&tmB, static_cast<const float*>(B), K, N, ldb, Cfg::BK, tma_rounds(), true));
    else RLT_TRY(make_tmap(&tmB, static_cast<const float*>(B), N, K, ldb, BN, tma_rounds()));
  }
  // grid = a multiple of the number of column blocks (each CTA keeps one column block, see the kernel), at most
  // one CTA per SM and no more row-block walkers than there are row blocks
  const int tiles_m = (M + Cfg::BM - 1) / Cfg::BM, tiles_n = N / BN;
  RLT_REQUIRE(tiles_n <= num_sms(), RLT_UNSUPPORTED_SHAPE, "gemm: N=%d needs more column blocks than there are SMs", N);
  int walkers = num_sms() / tiles_n;
  if (walkers > tiles_m) walkers = tiles_m;
  const int grid = walkers * tiles_n;
  // epilogue specialisations of the hot call sites (encoder / BiLSTM); anything else takes the run-time epilogue
  if (BN >= 128) {
    switch (epi_mask(ep)) {
#define RLT_EF_CASE(mask) case (mask): return launch_tn_ef<BN, OP, (mask)>(tmA, tmB, M, N, K, ep, grid, stream)
      RLT_EF_CASE(0);
      RLT_EF_CASE(EF_BIAS);
      RLT_EF_CASE(EF_BIAS | EF_RELU);
      RLT_EF_CASE(EF_BIAS | EF_RES);
      RLT_EF_CASE(EF_GATE | EF_COLSUM);
      RLT_EF_CASE(EF_RES);
      RLT_EF_CASE(EF_ACC);
      RLT_EF_CASE(EF_RES | EF_ACC);                                 // dX of an expert that shares its input (MMOECut)
      RLT_EF_CASE(EF_BIAS | EF_RELU | EF_OUT_H);                    // FFN1 -> fp16 hidden
      RLT_EF_CASE(EF_GATE_H | EF_COLSUM | EF_OUT_H | EF_SCALE);     // dH = s (dU W2) [h > 0] -> fp16
      RLT_EF_CASE(EF_RES | EF_SCALE);                               // dY = dU + (dH W1) / s
      RLT_EF_CASE(EF_BIAS | EF_RELU | EF_OUT_H | EF_DROP);          // train-mode dropout variants
      RLT_EF_CASE(EF_BIAS | EF_RES | EF_DROP);
#undef RLT_EF_CASE
      default: break;
    }
  }
  return launch_tn_ef<BN, OP, EF_RUNTIME>(tmA, tmB, M, N, K, ep, grid, stream);
}

template <int OP>
static int gemm_any(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const EpiParams& ep,
                    cudaStream_t stream) {
  RLT_REQUIRE(M > 0 && N > 0 && K > 0, RLT_INVALID_ARG, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  RLT_REQUIRE(N % 32 == 0, RLT_UNSUPPORTED_SHAPE, "gemm: N=%d must be a multiple of 32", N);
  RLT_REQUIRE(ep.ldo % 4 == 0, RLT_UNSUPPORTED_SHAPE, "gemm: output leading dimension %d must be a multiple of 4",
              ep.ldo);
  if (N % 256 == 0) return launch_tn<256, OP>(A, lda, B, ldb, M, N, K, ep, stream);
  if (N % 128 == 0) return launch_tn<128, OP>(A, lda, B, ldb, M, N, K, ep, stream);
  if (OP == OP_F16_K) return set_error(RLT_UNSUPPORTED_SHAPE, "gemm (fp16 operands): N=%d must be a multiple of 128", N);
  if (N % 64 == 0) return launch_tn<64, OP == OP_F16_K ? OP_TF32_K : OP>(A, lda, B, ldb, M, N, K, ep, stream);
  return launch_tn<32, OP == OP_F16_K ? OP_TF32_K : OP>(A, lda, B, ldb, M, N, K, ep, stream);
}

static int gemm_simt(const float* A, int lda, const float* B, int ldb, int M, int N, int K, const EpiParams& ep,
                     cudaStream_t stream, int b_major_n) {
  RLT_REQUIRE(M > 0 && N > 0 && K > 0, RLT_INVALID_ARG, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  dim3 grid((N + 63) / 64, (M + 63) / 64);
  gemm_tn_simt_kernel<<<grid, 256, 0, stream>>>(A, lda, B, ldb, M, N, K, ep, b_major_n);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

int gemm_tn(const float* A, int lda, const float* B, int ldb, int M, int N, int K, const EpiParams& ep,
            cudaStream_t stream) {
  if (gemm_backend() == 1) return gemm_simt(A, lda, B, ldb, M, N, K, ep, stream, 0);
  return gemm_any<OP_TF32_K>(A, lda, B, ldb, M, N, K, ep, stream);
}
int gemm_nn(const float* A, int lda, const float* B, int ldb, int M, int N, int K, const EpiParams& ep,
            cudaStream_t stream) {
  if (gemm_backend() == 1) return gemm_simt(A, lda, B, ldb, M, N, K, ep, stream, 1);
  return gemm_any<OP_TF32_N>(A, lda, B, ldb, M, N, K, ep, stream);
}
// fp16 operands AND fp16 output, K <= 128, N % 256 == 0: the copy-engine epilogue kernel (gemm_f16out.cuh)
template <int BN, int EF>
static int launch_f16out(const __half* A, int lda, const __half* B, int ldb, int M, int N, int K, const EpiParams& ep,
                         cudaStream_t stream) {
  using Cfg = GemmF16OutCfg<BN>;
  CUtensorMap tmA, tmB, tmOut, tmGate;
  RLT_TRY(make_tmap_h(&tmA, A, M, K, lda, Cfg::BM));
  RLT_TRY(make_tmap_h(&tmB, B, N, K, ldb, Cfg::BN));
  RLT_TRY(make_tmap_any(&tmOut, ep.out_h, 2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, M, N, ep.ldo, 32, false, 32, CU_TENSOR_MAP_SWIZZLE_64B));
  tmGate = tmOut;
  if (ep.gate_h != nullptr)
    RLT_TRY(make_tmap_any(&tmGate, ep.gate_h, 2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, M, N, ep.ldo, 32, false, 32, CU_TENSOR_MAP_SWIZZLE_64B));
  static DeviceOnce attr_set;
  if (attr_set.first()) {
    RLT_CHECK_CUDA(cudaFuncSetAttribute(gemm_f16out_kernel<BN, EF>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg::SMEM_BYTES)));
  }
  const int tiles_m = (M + Cfg::BM - 1) / Cfg::BM, tiles_n = N / Cfg::BN;
  RLT_REQUIRE(tiles_n <= num_sms(), RLT_UNSUPPORTED_SHAPE, "gemm: N=%d needs more column blocks than there are SMs", N);
  int walkers = num_sms() / tiles_n;
  if (walkers > tiles_m) walkers = tiles_m;
  time_begin(ep.tag, stream);
  gemm_f16out_kernel<BN, EF><<<walkers * tiles_n, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, tmOut, tmGate, M, N, K, ep);
  time_end(ep.tag, stream);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

int gemm_tn_h(const __half* A, int lda, const __half* B, int ldb, int M, int N, int K, const EpiParams& ep,
              cudaStream_t stream) {
  RLT_REQUIRE(gemm_backend() == 0, RLT_INVALID_ARG, "gemm_tn_h: fp16 operands exist only on the tensor-core backend");
  RLT_REQUIRE(K % 8 == 0, RLT_UNSUPPORTED_SHAPE, "gemm_tn_h: K=%d must be a multiple of 8", K);
  if (g_f16out_tma != 0 && ep.out_h != nullptr && K <= 256 && N % 256 == 0 && ep.ldo % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(ep.out_h) & 15) == 0) {
#define RLT_F16OUT(mask)                                                                                    \
  case (mask):                                                                                              \
    return K <= 128 ? launch_f16out<256, (mask)>(A, lda, B, ldb, M, N, K, ep, stream)                       \
                    : launch_f16out<128, (mask)>(A, lda, B, ldb, M, N, K, ep, stream)
    // measured at d_model 256 (K = 256, BN = 128 tiles): the gated dH GEMM gains (3.36 -> 2.63 ms), the plain FFN1 store
    // does not (1.75 -> 2.13 ms: 32 KB tiles, per-tile overheads) and stays on gemm_tn_kernel<256>
    const int mask = epi_mask(ep);
    const bool plain = (mask & EF_GATE_H) == 0;
    switch ((plain && K > 128) ? -1 : mask) {
      RLT_F16OUT(EF_BIAS | EF_RELU | EF_OUT_H);
      RLT_F16OUT(EF_BIAS | EF_RELU | EF_OUT_H | EF_DROP);
      RLT_F16OUT(EF_GATE_H | EF_COLSUM | EF_OUT_H | EF_SCALE);
      default: break;
    }
#undef RLT_F16OUT
  }
  return gemm_any<OP_F16_K>(A, lda, B, ldb, M, N, K, ep, stream);
}

template <int BN, bool kF16, bool kColsum>
static int launch_dw_impl(const void* A, int lda, const void* B, int ldb, int T, int M, int N, float* C, int ldc,
                          float alpha, const float* alpha_ptr, cudaStream_t stream, float* colsum_out) {
  using Cfg = GemmDwCfg<BN, kF16, kColsum>;
  CUtensorMap tmA, tmB;
  if (kF16) {
    RLT_TRY(make_tmap_h(&tmA, static_cast<const __half*>(A), T, M, lda, Cfg::BT));
    RLT_TRY(make_tmap_h(&tmB, static_cast<const __half*>(B), T, N, ldb, Cfg::BT));
  } else {
    RLT_TRY(make_tmap(&tmA, static_cast<const float*>(A), T, M, lda, Cfg::BT, tma_rounds(), true));
    RLT_TRY(make_tmap(&tmB, static_cast<const float*>(B), T, N, ldb, Cfg::BT, tma_rounds(), true));
  }
  static DeviceOnce attr_set;
  if (attr_set.first()) {
    RLT_CHECK_CUDA(cudaFuncSetAttribute(gemm_dw_kernel<BN, kF16, kColsum>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        int(Cfg::SMEM_BYTES)));
  }
  const int tiles = ((M + Cfg::BM - 1) / Cfg::BM) * (N / BN);
  const int num_tb = (T + Cfg::BT - 1) / Cfg::BT;
  // one wave of CTAs: split the token axis so that tiles * splits ~ #SMs, at least 8 token blocks per split
  int splits = num_sms() / tiles;
  if (splits < 1) splits = 1;
  const int max_splits = (num_tb + 7) / 8;
  if (splits > max_splits) splits = max_splits;
  gemm_dw_kernel<BN, kF16, kColsum><<<dim3(tiles, splits), 192, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, T, M, N, C, ldc, alpha, alpha_ptr, colsum_out);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}
template <int BN, bool kF16>
static int launch_dw(const void* A, int lda, const void* B, int ldb, int T, int M, int N, float* C, int ldc,
                     float alpha, const float* alpha_ptr, cudaStream_t stream, float* colsum_out = nullptr) {
  if constexpr (!kF16 && BN == 128) {      // the one call site with a fused bias gradient: dW_in (N = d_model 128)
    if (colsum_out != nullptr)
      return launch_dw_impl<BN, kF16, true>(A, lda, B, ldb, T, M, N, C, ldc, alpha, alpha_ptr, stream, colsum_out);
  }
  RLT_REQUIRE(colsum_out == nullptr, RLT_UNSUPPORTED_SHAPE, "gemm_dw: fused column sums are built for N %% 256 == 128 only (N=%d)", N);
  return launch_dw_impl<BN, kF16, false>(A, lda, B, ldb, T, M, N, C, ldc, alpha, alpha_ptr, stream, nullptr);
}

struct TimeScope {
  int tag; cudaStream_t s;
  TimeScope(int t, cudaStream_t st) : tag(t), s(st) { time_begin(tag, s); }
  ~TimeScope() { time_end(tag, s); }
};

int gemm_dw(const float* A, int lda, const float* B, int ldb, int T, int M, int N, float* C, int ldc, float alpha,
            cudaStream_t stream, int tag, float* colsum_out) {
  TimeScope scope(tag, stream);
  RLT_REQUIRE(T > 0 && M > 0 && N > 0, RLT_INVALID_ARG, "gemm_dw: empty problem T=%d M=%d N=%d", T, M, N);
  if (colsum_out != nullptr && (gemm_backend() == 1 || g_dw_colsum == 0 || N % 256 != 128)) {
    RLT_REQUIRE(lda == M, RLT_INVALID_ARG, "gemm_dw: column sums need a dense A (lda=%d, M=%d)", lda, M);
    RLT_TRY(colsum(A, colsum_out, T, M, stream));
    colsum_out = nullptr;
  }
  if (gemm_backend() == 1) {
    const int tchunk = 2048;
    dim3 grid((N + 15) / 16, (M + 15) / 16, (T + tchunk - 1) / tchunk);
    gemm_dw_simt_kernel<<<grid, 256, 0, stream>>>(A, lda, B, ldb, T, M, N, C, ldc, alpha, tchunk);
    RLT_CHECK_LAUNCH();
    return RLT_OK;
  }
  RLT_REQUIRE(N % 32 == 0, RLT_UNSUPPORTED_SHAPE, "gemm_dw: N=%d must be a multiple of 32", N);
  if (N % 256 == 0) return launch_dw<256, false>(A, lda, B, ldb, T, M, N, C, ldc, alpha, nullptr, stream, colsum_out);
  if (N % 128 == 0) return launch_dw<128, false>(A, lda, B, ldb, T, M, N, C, ldc, alpha, nullptr, stream, colsum_out);
  if (N % 64 == 0) return launch_dw<64, false>(A, lda, B, ldb, T, M, N, C, ldc, alpha, nullptr, stream, colsum_out);
  return launch_dw<32, false>(A, lda, B, ldb, T, M, N, C, ldc, alpha, nullptr, stream, colsum_out);
}

// Fused FFN backward (ffn_bwd_fused.cuh): dH (fp16), db1 and dW2 in one pass over h.
bool ffn_bwd_fused_ok(int d, int f) {
  return g_ffn_bwd_fused != 0 && gemm_backend() == 0 && d == FfnBwdCfg::D && f % FfnBwdCfg::BN == 0 && f / FfnBwdCfg::BN <= num_sms();
}
int ffn_bwd_fused(const __half* du16, const __half* w2th, const __half* hh, __half* dh16, int T, int d, int f, float alpha,
                  const float* scale, float* db1, float* dW2, cudaStream_t stream, int tag) {
  using Cfg = FfnBwdCfg;
  RLT_REQUIRE(ffn_bwd_fused_ok(d, f), RLT_UNSUPPORTED_SHAPE, "ffn_bwd_fused: d=%d f=%d unsupported", d, f);
  CUtensorMap tmDU, tmW2, tmH, tmDH;
  RLT_TRY(make_tmap_h(&tmDU, du16, T, d, d, Cfg::BM));
  RLT_TRY(make_tmap_h(&tmW2, w2th, f, d, d, Cfg::BN));
  RLT_TRY(make_tmap_h(&tmH, hh, T, f, f, Cfg::BM));
  RLT_TRY(make_tmap_any(&tmDH, dh16, 2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, T, f, f, 32, false, 32, CU_TENSOR_MAP_SWIZZLE_64B));
  static DeviceOnce attr_set;
  if (attr_set.first()) {
    RLT_CHECK_CUDA(cudaFuncSetAttribute(ffn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg::SMEM_BYTES)));
  }
  const int tiles_m = (T + Cfg::BM - 1) / Cfg::BM, tiles_n = f / Cfg::BN;
  int walkers = num_sms() / tiles_n;
  if (walkers > tiles_m) walkers = tiles_m;
  TimeScope scope(tag, stream);
  ffn_bwd_kernel<<<walkers * tiles_n, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tmDU, tmW2, tmH, tmDH, T, f, alpha, scale, db1, dW2);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

static long long* g_ffn_dbg = nullptr;      // device buffer for a kernel timeline (rlt_ffn_fused_set_timeline; tools only)
void ffn_fwd_set_timeline(long long* dev_buf) { g_ffn_dbg = dev_buf; }
// ------------------------------------------------------------------------------------------
// tcgen05 attention forward (attention_tc.cuh)
// ------------------------------------------------------------------------------------------
// 3-D fp32 tensor map over a [n_lists, L, width] row-major tensor: dims (width, L, n_lists), box (box_w, 1, 64 lists)
static int make_tmap_lists3d(CUtensorMap* out, const float* base, int width, int L, long long n_lists, int box_w) {
  EncodeTiledFn fn = encode_fn();
  RLT_REQUIRE(fn != nullptr, RLT_CUDA_ERROR, "cuTensorMapEncodeTiled is not available from the driver");
  RLT_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (width * 4) % 16 == 0, RLT_INVALID_ARG,
              "TMA operand must be 16-byte aligned with a row pitch that is a multiple of 16 bytes");
  const cuuint64_t dims[3] = {cuuint64_t(width), cuuint64_t(L), cuuint64_t(n_lists)};
  const cuuint64_t strides[2] = {cuuint64_t(width) * 4, cuuint64_t(L) * cuuint64_t(width) * 4};
  const cuuint32_t box[3] = {cuuint32_t(box_w), 1u, 64u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RLT_REQUIRE(r == CUDA_SUCCESS, RLT_CUDA_ERROR, "cuTensorMapEncodeTiled (3-D) failed with CUresult %d", int(r));
  return RLT_OK;
}
bool attention_fwd_tc_ok(int S, int dh) { return g_attn_tc != 0 && gemm_backend() == 0 && dh == 16 && S <= 64; }
int attention_fwd_tc(const float* qkv, float* o, float* lse, int G, int S, int L, int d, int n_head, float scale,
                     cudaStream_t stream, int tag) {
  using Cfg = AttnTcCfg<16>;
  RLT_REQUIRE(attention_fwd_tc_ok(S, d / n_head), RLT_UNSUPPORTED_SHAPE, "attention_fwd_tc: S=%d dh=%d unsupported", S, d / n_head);
  CUtensorMap tmQKV, tmO;
  RLT_TRY(make_tmap_lists3d(&tmQKV, qkv, 3 * d, L, (long long)G * S, 16));
  RLT_TRY(make_tmap_lists3d(&tmO, o, d, L, (long long)G * S, 16));
  static DeviceOnce once;
  if (once.first())
    RLT_CHECK_CUDA(cudaFuncSetAttribute(attn_lists_fwd_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg::SMEM_BYTES)));
  const long long n_super = (long long)G * ((L + 1) / 2);
  long long grid = num_sms();
  if (grid > n_super) grid = n_super;
  TimeScope scope(tag, stream);
  attn_lists_fwd_tc_kernel<16><<<int(grid), Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tmQKV, tmO, o, lse, G, S, L, d, n_head,
                                                                                    scale * 1.4426950408889634f, g_ffn_dbg);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

// Fused FFN forward (ffn_fwd_fused.cuh): out = LN2(y + relu(y W1^T + b1) W2^T + b2), hidden on chip.
bool ffn_fwd_fused_ok(int d, int f) {
  return g_ffn_fwd_fused != 0 && gemm_backend() == 0 && (d == 128 || d == 256) && f % 128 == 0 && f <= FfnFwdCfg<128>::MAX_F;
}
template <int D>
static int ffn_fwd_fused_launch(const __half* y16, const float* y, const __half* w1h, const float* b1, const __half* w2h,
                                const float* b2, const float* gamma, const float* beta, float* out, float* u2, float* stats,
                                __half* h_out, int T, int f, float eps, cudaStream_t stream, int tag) {
  using Cfg = FfnFwdCfg<D>;
  CUtensorMap tmY, tmW1, tmW2, tmH, tmOut, tmU2;
  RLT_TRY(make_tmap_h(&tmY, y16, T, D, D, Cfg::BM));
  // out / pre-norm sums: per-warp [32 rows x 16 columns] fp32 boxes (64-byte rows, SWIZZLE_64B) stored by the copy engine
  RLT_TRY(make_tmap_any(&tmOut, out, 4, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, T, D, D, 32, false, 16, CU_TENSOR_MAP_SWIZZLE_64B));
  tmU2 = tmOut;
  if (u2 != nullptr)
    RLT_TRY(make_tmap_any(&tmU2, u2, 4, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, T, D, D, 32, false, 16, CU_TENSOR_MAP_SWIZZLE_64B));
  RLT_TRY(make_tmap_h(&tmW1, w1h, f, D, D, Cfg::CH / 2));
  RLT_TRY(make_tmap_h(&tmW2, w2h, D, f, f, D / 2));
  tmH = tmY;       // placeholder when the hidden is not saved (never dereferenced then)
  if (h_out != nullptr)      // per-warp [32 rows x CH/4 columns] fp16 tiles: 64-byte rows swizzled, 32-byte rows plain
    RLT_TRY(make_tmap_any(&tmH, h_out, 2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, T, f, f, 32, false, Cfg::CH / 4,
                          Cfg::CH / 4 == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE));
  static DeviceOnce once;
  if (once.first()) {
    RLT_CHECK_CUDA(cudaFuncSetAttribute(ffn_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg::SMEM_BYTES)));
    // how many CTA pairs the device can hold at once (a GPC with an odd number of free SMs leaves one unpaired)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(unsigned(num_sms() / 2 * 2));
    cfg.blockDim = dim3(Cfg::THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int n_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&n_clusters, ffn_fwd_kernel<D>, &cfg) != cudaSuccess || n_clusters < 1) {
      (void)cudaGetLastError();
      n_clusters = num_sms() / 2;
    }
    once.value[once.dev()] = n_clusters;
  }
  const int n_tiles = (T + 2 * Cfg::BM - 1) / (2 * Cfg::BM);
  int pairs = once.value[once.dev()];
  if (pairs > n_tiles) pairs = n_tiles;
  FfnFwdParams prm;
  prm.y = y; prm.b1 = b1; prm.b2 = b2; prm.gamma = gamma; prm.beta = beta; prm.out = out; prm.u2 = u2; prm.stats = stats;
  prm.h_out = h_out; prm.T = T; prm.F = f; prm.eps = eps;
  prm.dbg = g_ffn_dbg;
  TimeScope scope(tag, stream);
  ffn_fwd_kernel<D><<<2 * pairs, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tmY, tmW1, tmW2, tmH, tmOut, tmU2, prm);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}
int ffn_fwd_fused(const __half* y16, const float* y, const __half* w1h, const float* b1, const __half* w2h, const float* b2,
                  const float* gamma, const float* beta, float* out, float* u2, float* stats, __half* h_out, int T, int d,
                  int f, float eps, cudaStream_t stream, int tag) {
  RLT_REQUIRE(ffn_fwd_fused_ok(d, f), RLT_UNSUPPORTED_SHAPE, "ffn_fwd_fused: d=%d f=%d unsupported", d, f);
  if (d == 128)
    return ffn_fwd_fused_launch<128>(y16, y, w1h, b1, w2h, b2, gamma, beta, out, u2, stats, h_out, T, f, eps, stream, tag);
  return ffn_fwd_fused_launch<256>(y16, y, w1h, b1, w2h, b2, gamma, beta, out, u2, stats, h_out, T, f, eps, stream, tag);
}

int gemm_dw_h(const __half* A, int lda, const __half* B, int ldb, int T, int M, int N, float* C, int ldc, float alpha,
              const float* alpha_ptr, cudaStream_t stream, int tag) {
  TimeScope scope(tag, stream);
  RLT_REQUIRE(T > 0 && M > 0 && N > 0, RLT_INVALID_ARG, "gemm_dw_h: empty problem T=%d M=%d N=%d", T, M, N);
  RLT_REQUIRE(gemm_backend() == 0, RLT_INVALID_ARG, "gemm_dw_h: fp16 operands exist only on the tensor-core backend");
  RLT_REQUIRE(N % 128 == 0 && M % 64 == 0, RLT_UNSUPPORTED_SHAPE, "gemm_dw_h: M=%d N=%d must be multiples of 64 / 128", M, N);
  if (N % 256 == 0) return launch_dw<256, true>(A, lda, B, ldb, T, M, N, C, ldc, alpha, alpha_ptr, stream);
  return launch_dw<128, true>(A, lda, B, ldb, T, M, N, C, ldc, alpha, alpha_ptr, stream);
}

// ------------------------------------------------------------------------------------------
// fp16 operand copies
// ------------------------------------------------------------------------------------------
// dst = half(src * (scale ? scale[0] : 1) * dropout factor), n a multiple of 4
__global__ void __launch_bounds__(256) convert_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, size_t n4,
                                                          const float* __restrict__ scale, DropCfg drop, uint32_t site) {
  const float s = scale != nullptr ? scale[0] : 1.f;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = reinterpret_cast<const float4*>(src)[i];
    if (drop.thr) {
      const uint64_t bits = drop_bits(drop.seed, site, i);
      v.x *= drop_factor(bits, 0, drop.thr, drop.scale); v.y *= drop_factor(bits, 1, drop.thr, drop.scale);
      v.z *= drop_factor(bits, 2, drop.thr, drop.scale); v.w *= drop_factor(bits, 3, drop.thr, drop.scale);
    }
    const __half2 lo = __floats2half2_rn(v.x * s, v.y * s), hi = __floats2half2_rn(v.z * s, v.w * s);
    reinterpret_cast<uint2*>(dst)[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
  }
}
int convert_f16(const float* src, __half* dst, size_t n, const float* scale, cudaStream_t stream, DropCfg drop,
                uint32_t site) {
  RLT_REQUIRE(n % 4 == 0, RLT_INVALID_ARG, "convert_f16: n must be a multiple of 4");
  size_t blocks = (n / 4 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > size_t(num_sms()) * 16) blocks = size_t(num_sms()) * 16;
  convert_f16_kernel<<<unsigned(blocks), 256, 0, stream>>>(src, dst, n / 4, scale, drop, site);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}
// dst = src * dropout factor (fp32; the masked copy of a gradient that feeds the GEMMs of a dropped branch)
__global__ void __launch_bounds__(256) dropout_apply_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t n4,
                                                            DropCfg drop, uint32_t site) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = reinterpret_cast<const float4*>(src)[i];
    const uint64_t bits = drop_bits(drop.seed, site, i);
    v.x *= drop_factor(bits, 0, drop.thr, drop.scale); v.y *= drop_factor(bits, 1, drop.thr, drop.scale);
    v.z *= drop_factor(bits, 2, drop.thr, drop.scale); v.w *= drop_factor(bits, 3, drop.thr, drop.scale);
    reinterpret_cast<float4*>(dst)[i] = v;
  }
}
int dropout_apply(const float* src, float* dst, size_t n, DropCfg drop, uint32_t site, cudaStream_t stream) {
  RLT_REQUIRE(n % 4 == 0, RLT_INVALID_ARG, "dropout_apply: n must be a multiple of 4");
  size_t blocks = (n / 4 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > size_t(num_sms()) * 16) blocks = size_t(num_sms()) * 16;
  dropout_apply_kernel<<<unsigned(blocks), 256, 0, stream>>>(src, dst, n / 4, drop, site);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}
// Test hook: out[i] = keep-and-scale factor of element i of a linear site (groups of 4), or, for the attention site,
// out[(item * S + query) * S + key] with the pair indexing of the attention kernels.
__global__ void dropout_mask_kernel(float* __restrict__ out, size_t n, DropCfg drop, uint32_t site, int S) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (site == DROP_ATTN) {
      const size_t key = i % S;
      out[i] = drop_factor(drop_bits(drop.seed, site, i - (key & 1)), int(key & 1), drop.thr, drop.scale);
    } else {
      out[i] = drop_factor(drop_bits(drop.seed, site, i >> 2), int(i & 3), drop.thr, drop.scale);
    }
  }
}
// dst[c, r] = half(src[r, c])   ([rows, cols] -> [cols, rows])
__global__ void transpose_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, int rows, int cols) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? src[size_t(r) * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[size_t(c) * rows + r] = __float2half_rn(tile[threadIdx.x][i]);
  }
}
int transpose_f16(const float* src, __half* dst, int rows, int cols, cudaStream_t stream) {
  dim3 grid((cols + 31) / 32, (rows + 31) / 32);
  transpose_f16_kernel<<<grid, dim3(32, 8), 0, stream>>>(src, dst, rows, cols);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}
// amax |x| -> bits of a non-negative float in *out (zero-initialised by the caller); n4 = n / 4
__global__ void __launch_bounds__(256) amax_abs4_kernel(const float4* __restrict__ x, size_t n4, unsigned int* __restrict__ out) {
  float m = 0.f;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = x[i];
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}
// scale[0] = 2^k with amax * 2^k in [2^(target-1), 2^target), scale[1] = 2^-k  (1, 1 when amax is 0 or not finite)
__global__ void pow2_scale_kernel(const unsigned int* __restrict__ amax_bits, float* __restrict__ scale, int target) {
  const float a = __uint_as_float(*amax_bits);
  float s = 1.f;
  if (a > 0.f && a < 3.0e38f) {
    int e;
    frexpf(a, &e);                   // a = m * 2^e, m in [0.5, 1)
    s = ldexpf(1.f, target - e);
  }
  scale[0] = s;
  scale[1] = 1.f / s;
}
int pow2_scale(const unsigned int* amax_bits, float* scale, int target, cudaStream_t stream) {
  pow2_scale_kernel<<<1, 1, 0, stream>>>(amax_bits, scale, target);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}
int grad_scale(const float* x, size_t n, unsigned int* amax_scratch, float* scale, int target, cudaStream_t stream) {
  RLT_REQUIRE(n % 4 == 0, RLT_INVALID_ARG, "grad_scale: n must be a multiple of 4");
  RLT_CHECK_CUDA(cudaMemsetAsync(amax_scratch, 0, sizeof(unsigned int), stream));
  size_t blocks = (n / 4 + 255) / 256;
  if (blocks > size_t(num_sms()) * 16) blocks = size_t(num_sms()) * 16;
  if (blocks < 1) blocks = 1;
  amax_abs4_kernel<<<unsigned(blocks), 256, 0, stream>>>(reinterpret_cast<const float4*>(x), n / 4, amax_scratch);
  RLT_CHECK_LAUNCH();
  pow2_scale_kernel<<<1, 1, 0, stream>>>(amax_scratch, scale, target);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

// ------------------------------------------------------------------------------------------
// operand preparation
// ------------------------------------------------------------------------------------------
__global__ void round_tf32_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t n) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  const size_t n4 = n / 4;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    reinterpret_cast<float4*>(dst)[i] = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
  }
  for (size_t i = n4 * 4 + size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = to_tf32(src[i]);
}

__global__ void transpose_round_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? src[size_t(r) * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[size_t(c) * rows + r] = to_tf32(tile[threadIdx.x][i]);
  }
}

// ------------------------------------------------------------------------------------------
// probe: what does a TFLOAT32 tensor map do to fp32 data?
// ------------------------------------------------------------------------------------------
__global__ void probe_tma_kernel(const __grid_constant__ CUtensorMap tm, float* __restrict__ dst, int rows) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, 128 * 32 * 4);
    tma_load_2d(smem, &tm, &bar, 0, 0);
  }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < rows * 32; i += blockDim.x) {
    const int r = i >> 5, c = i & 31;
    dst[i] = *reinterpret_cast<const float*>(smem + sw128_offset(r, c >> 2) + (c & 3) * 4);
  }
}

}  // namespace rlt

using namespace rlt;

extern "C" {

const char* rlt_version(void) { return "rlt_b200 0.1.0 (sm_100a)"; }
const char* rlt_last_error(void) { return g_err; }

unsigned long long rlt_launch_count(void) { return g_launches.load(); }

int rlt_timing_reset(void) {
  std::lock_guard<std::mutex> lk(g_time_mu);
  for (cudaEvent_t e : g_time_events) cudaEventDestroy(e);
  g_time_events.clear();
  g_time_tags.clear();
  return RLT_OK;
}

/* Sum of the bracketed kernel durations since the last reset (synchronises on the recorded events). */
int rlt_timing_read(double* total_ms, int* count) {
  RLT_REQUIRE(total_ms && count, RLT_INVALID_ARG, "rlt_timing_read: null pointer");
  std::lock_guard<std::mutex> lk(g_time_mu);
  double tot = 0.0;
  int n = 0;
  for (size_t i = 0; i + 1 < g_time_events.size(); i += 2) {
    RLT_CHECK_CUDA(cudaEventSynchronize(g_time_events[i + 1]));
    float ms = 0.f;
    RLT_CHECK_CUDA(cudaEventElapsedTime(&ms, g_time_events[i], g_time_events[i + 1]));
    tot += ms;
    ++n;
  }
  *total_ms = tot;
  *count = n;
  return RLT_OK;
}

/* Same for one call site (KernelTag) when every site was bracketed (time_tag = -1). */
int rlt_timing_read_tag(int tag, double* total_ms, int* count) {
  RLT_REQUIRE(total_ms && count, RLT_INVALID_ARG, "rlt_timing_read_tag: null pointer");
  std::lock_guard<std::mutex> lk(g_time_mu);
  double tot = 0.0;
  int n = 0;
  for (size_t i = 0; i + 1 < g_time_events.size(); i += 2) {
    if (g_time_tags[i] != tag) continue;
    RLT_CHECK_CUDA(cudaEventSynchronize(g_time_events[i + 1]));
    float ms = 0.f;
    RLT_CHECK_CUDA(cudaEventElapsedTime(&ms, g_time_events[i], g_time_events[i + 1]));
    tot += ms;
    ++n;
  }
  *total_ms = tot;
  *count = n;
  return RLT_OK;
}

int rlt_set_option(const char* key, int value) {
  if (key == nullptr) return set_error(RLT_INVALID_ARG, "rlt_set_option: null key");
  if (strcmp(key, "time_tag") == 0) { g_time_tag = value; return RLT_OK; }
  if (strcmp(key, "lstm_backend") == 0) { set_lstm_backend(value); return RLT_OK; }
  if (strcmp(key, "lstm_tile") == 0) { set_lstm_tile(value); return RLT_OK; }
  if (strcmp(key, "gemm_backend") == 0) { g_gemm_backend = value; return RLT_OK; }
  if (strcmp(key, "tma_round") == 0) { g_tma_round = value; return RLT_OK; }
  if (strcmp(key, "b_resident") == 0) { g_b_resident = value; return RLT_OK; }
  if (strcmp(key, "f16out_tma") == 0) { g_f16out_tma = value; return RLT_OK; }
  if (strcmp(key, "ffn_bwd_fused") == 0) { g_ffn_bwd_fused = value; return RLT_OK; }
  if (strcmp(key, "ffn_fwd_fused") == 0) { g_ffn_fwd_fused = value; return RLT_OK; }
  if (strcmp(key, "attn_tc") == 0) { g_attn_tc = value; return RLT_OK; }
  if (strcmp(key, "dw_colsum") == 0) { g_dw_colsum = value; return RLT_OK; }
  return set_error(RLT_INVALID_ARG, "rlt_set_option: unknown option '%s'", key);
}
int rlt_get_option(const char* key) {
  if (key == nullptr) return set_error(RLT_INVALID_ARG, "rlt_get_option: null key");
  if (strcmp(key, "gemm_backend") == 0) return g_gemm_backend;
  if (strcmp(key, "tma_round") == 0) return g_tma_round;
  if (strcmp(key, "b_resident") == 0) return g_b_resident;
  if (strcmp(key, "f16out_tma") == 0) return g_f16out_tma;
  if (strcmp(key, "ffn_bwd_fused") == 0) return g_ffn_bwd_fused;
  if (strcmp(key, "ffn_fwd_fused") == 0) return g_ffn_fwd_fused;
  if (strcmp(key, "attn_tc") == 0) return g_attn_tc;
  if (strcmp(key, "dw_colsum") == 0) return g_dw_colsum;
  if (strcmp(key, "time_tag") == 0) return g_time_tag;
  if (strcmp(key, "lstm_backend") == 0) return lstm_backend();
  if (strcmp(key, "lstm_tile") == 0) return get_lstm_tile();
  return set_error(RLT_INVALID_ARG, "rlt_get_option: unknown option '%s'", key);
}

int rlt_round_tf32(const float* src, float* dst, size_t n, rlt_stream_t stream) {
  RLT_REQUIRE(src && dst, RLT_INVALID_ARG, "rlt_round_tf32: null pointer");
  if (n == 0) return RLT_OK;
  RLT_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0, RLT_INVALID_ARG,
              "rlt_round_tf32: pointers must be 16-byte aligned");
  size_t blocks = (n / 4 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > size_t(num_sms()) * 8) blocks = size_t(num_sms()) * 8;
  round_tf32_kernel<<<unsigned(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, n);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

int rlt_transpose_round_tf32(const float* src, float* dst, int rows, int cols, rlt_stream_t stream) {
  RLT_REQUIRE(src && dst && rows > 0 && cols > 0, RLT_INVALID_ARG, "rlt_transpose_round_tf32: bad arguments");
  dim3 grid((cols + 31) / 32, (rows + 31) / 32);
  transpose_round_kernel<<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(src, dst, rows, cols);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

int rlt_linear(const float* A, const float* B, const float* bias, float* C, int M, int N, int K, float alpha,
               int relu, rlt_stream_t stream) {
  RLT_REQUIRE(A && B && C, RLT_INVALID_ARG, "rlt_linear: null pointer");
  EpiParams ep{};
  ep.out = C;
  ep.ldo = N;
  ep.bias = bias;
  ep.relu = relu;
  ep.alpha = alpha;
  return gemm_tn(A, K, B, K, M, N, K, ep, static_cast<cudaStream_t>(stream));
}

int rlt_linear_nn(const float* A, const float* B, float* C, int M, int N, int K, rlt_stream_t stream) {
  RLT_REQUIRE(A && B && C, RLT_INVALID_ARG, "rlt_linear_nn: null pointer");
  EpiParams ep{};
  ep.out = C;
  ep.ldo = N;
  ep.alpha = 1.f;
  return gemm_nn(A, K, B, N, M, N, K, ep, static_cast<cudaStream_t>(stream));
}

int rlt_grad_weight(const float* A, const float* B, float* C, int T, int M, int N, float alpha,
                    rlt_stream_t stream) {
  RLT_REQUIRE(A && B && C, RLT_INVALID_ARG, "rlt_grad_weight: null pointer");
  return gemm_dw(A, M, B, N, T, M, N, C, N, alpha, static_cast<cudaStream_t>(stream));
}

int rlt_convert_f16(const float* src, void* dst, size_t n, const float* scale, rlt_stream_t stream) {
  RLT_REQUIRE(src && dst, RLT_INVALID_ARG, "rlt_convert_f16: null pointer");
  return convert_f16(src, static_cast<__half*>(dst), n, scale, static_cast<cudaStream_t>(stream));
}

int rlt_dropout_mask(uint64_t seed, int site, float p, size_t n, int group_size, float* out, rlt_stream_t stream) {
  RLT_REQUIRE(out && n > 0 && site >= 1 && site <= 5 && p >= 0.f && p < 1.f, RLT_INVALID_ARG, "rlt_dropout_mask: bad arguments");
  RLT_REQUIRE(site != DROP_ATTN || group_size > 0, RLT_INVALID_ARG, "rlt_dropout_mask: the attention site needs group_size");
  dropout_mask_kernel<<<num_sms() * 4, 256, 0, static_cast<cudaStream_t>(stream)>>>(out, n, make_drop(p, seed), uint32_t(site),
                                                                                  group_size);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

int rlt_linear_f16(const void* A, const void* B, const float* bias, float* C, int M, int N, int K, float alpha, int relu,
                   rlt_stream_t stream) {
  RLT_REQUIRE(A && B && C, RLT_INVALID_ARG, "rlt_linear_f16: null pointer");
  EpiParams ep{};
  ep.out = C;
  ep.ldo = N;
  ep.bias = bias;
  ep.relu = relu;
  ep.alpha = alpha;
  return gemm_tn_h(static_cast<const __half*>(A), K, static_cast<const __half*>(B), K, M, N, K, ep,
                   static_cast<cudaStream_t>(stream));
}

int rlt_grad_weight_f16(const void* A, const void* B, float* C, int T, int M, int N, float alpha, rlt_stream_t stream) {
  RLT_REQUIRE(A && B && C, RLT_INVALID_ARG, "rlt_grad_weight_f16: null pointer");
  return gemm_dw_h(static_cast<const __half*>(A), M, static_cast<const __half*>(B), N, T, M, N, C, N, alpha, nullptr,
                   static_cast<cudaStream_t>(stream), 0);
}

/* C (fp16) = relu(A B^T + bias): the FFN1 form (fp32 operands, half-precision result). */
int rlt_linear_out_f16(const float* A, const float* B, const float* bias, void* C, int M, int N, int K, int relu,
                       rlt_stream_t stream) {
  RLT_REQUIRE(A && B && C, RLT_INVALID_ARG, "rlt_linear_out_f16: null pointer");
  EpiParams ep{};
  ep.out_h = static_cast<__half*>(C);
  ep.ldo = N;
  ep.bias = bias;
  ep.relu = relu;
  ep.alpha = 1.f;
  return gemm_tn(A, K, B, K, M, N, K, ep, static_cast<cudaStream_t>(stream));
}

int rlt_probe_tma_tf32(const float* src, float* dst, int rows, rlt_stream_t stream) {
  RLT_REQUIRE(src && dst && rows > 0 && rows <= 128, RLT_INVALID_ARG, "rlt_probe_tma_tf32: rows must be in [1,128]");
  CUtensorMap tm;
  RLT_TRY(make_tmap(&tm, src, rows, 32, 32, 128, true));
  probe_tma_kernel<<<1, 128, 128 * 32 * 4 + 1024, static_cast<cudaStream_t>(stream)>>>(tm, dst, rows);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

}  // extern "C"
