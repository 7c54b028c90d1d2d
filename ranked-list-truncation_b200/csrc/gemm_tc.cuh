// Persistent, warp-specialised TF32 GEMMs on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
//
//   gemm_tn_kernel : C[M,N]  = A[M,K] * B[N,K]^T         (both operands K-major; Linear fwd, dX)
//   gemm_dw_kernel : C[M,N] += sum_t A[t,M] * B[t,N]     (both operands MN-major; weight grads,
//                                                         contraction over the token axis, split
//                                                         over CTAs, fp32 red.add into C)
//
// Roles per CTA (gemm_tn: 320 threads): warp 0 = TMA producer, warp 1 = single-thread MMA issuer,
// warps 2..9 = epilogue.  An epilogue warp reads its TMEM lane quarter (tcgen05.ld: one output row
// per thread), transposes the 32x32 chunk through a private swizzled 4 KB shared-memory buffer and
// then touches global memory with whole 128-byte row segments per 8 lanes (4 L1 wavefronts per
// 128-bit warp access instead of 32), for the output store as well as for the gate / residual /
// accumulate reads.  Accumulators are double-buffered in TMEM so the epilogue of tile i overlaps the
// main loop of tile i+1.  Operands are fp32 in HBM; the TFLOAT32 tensor maps round them on load.
#pragma once
#include <cuda_fp16.h>

#include "dropout.cuh"
#include "sm100.cuh"

namespace rlt {

struct EpiParams {
  float* out;             // [M, ldo] fp32 result (null when out_h is used)
  int ldo;
  const float* bias;      // [N] added to every row (may be null)
  const float* gate_src;  // [M, ldo]: result *= (gate_src > 0)   (ReLU backward; may be null)
  const float* residual;  // [M, ldo]: result += residual          (skip connections; may be null)
  float* colsum;          // [N]: atomicAdd of the column sums of the final values (bias grads; may be null)
  int relu;               // max(x, 0)
  int accumulate;         // out += result instead of out = result
  float alpha;            // result scale applied first
  int tag;                // KernelTag of the call site (in-situ timing; 0 = none)
  // ---- half-precision activations of the FFN hidden layer (fp16 = the 11 significant bits of TF32)
  __half* out_h;          // [M, ldo] fp16 result instead of `out`
  const __half* gate_h;   // [M, ldo] fp16 gate source instead of `gate_src`
  // device-side power-of-two scale {s, 1/s} (gradients are ~1e-6: fp16 needs them scaled into its normal range):
  //   scale_mode 1: result *= s and the column sums are multiplied by 1/s (they stay unscaled)
  //   scale_mode 2: result *= 1/s (unscale a contraction over scaled operands)
  //   scale_mode 3: the result is left as it is (an operand already carries s); column sums are multiplied by 1/s
  const float* scale_ptr;
  int scale_mode;
  // ---- train-mode dropout applied to (alpha acc + bias) after the ReLU and BEFORE the residual is added; the element
  //      index of the hash is row * ldo + col (dropout.cuh); drop.thr == 0 disables it
  DropCfg drop;
  uint32_t drop_site;
};

// Sum v[j] over the 32 lanes of a warp for 32 different j: on return lane j holds the column-j
// total.  31 shuffles instead of 32x5.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int j = 0; j < half; ++j) {
      const float send = upper ? v[j] : v[j + half];
      const float keep = upper ? v[j + half] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  return v[0];
}

// Operand kinds (template parameter OP):
//   OP_TF32_K : A [M, K] fp32, B [N, K] fp32 row-major (K-major operands, nn.Linear weight used as-is: x W^T)
//   OP_TF32_N : A [M, K] fp32, B [K, N] fp32 row-major (MN-major B: x W with W stored [K, N])
//   OP_F16_K  : A [M, K] fp16, B [N, K] fp16 (K-major, kind::f16: 64 elements per 128-byte swizzle row)
enum : int { OP_TF32_K = 0, OP_TF32_N = 1, OP_F16_K = 2 };

template <int BN, int OP>
struct GemmTnCfg {
  static constexpr int BM = 128;
  static constexpr int BK = 32;  // 32 tf32 = 128 B = one swizzle row (64 fp16)
  static constexpr int A_BYTES = BM * BK * 4;
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // The epilogue of these store-bound GEMMs is a latency chain per warp (TMEM read -> smem transpose -> global), so
  // the fp16-operand kernels (FFN hidden path) trade one pipeline stage for 16 epilogue warps instead of 8.
  static constexpr int EPI_WARPS = (OP == OP_F16_K && BN >= 128) ? 16 : 8;   // warps per TMEM lane quarter: 4 or 2
  static constexpr int STAGES = (BN >= 256) ? (EPI_WARPS == 16 ? 3 : 4) : (BN >= 128 ? (EPI_WARPS == 16 ? 5 : 6) : 8);
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  static constexpr int RES_B_BYTES = BN < 128 ? 0 : (EPI_WARPS == 16 ? 64 * 1024 : 128 * 1024);   // B-resident mode: the CTA's B slice ...
  static constexpr int RES_STAGES = BN >= 128 ? (STAGES * STAGE_BYTES - RES_B_BYTES) / A_BYTES : 1;   // ... and an A-only ring
  static_assert(BN < 128 || RES_STAGES >= 4, "A ring of the B-resident mode");
  static constexpr int NBAR = STAGES > RES_STAGES ? STAGES : RES_STAGES;   // full / empty barrier pairs
  static_assert(2 * NBAR + 6 <= 31, "barrier block");
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
  static constexpr int EPI_STAGE_BYTES = 32 * 32 * 4;  // one 32x32 fp32 chunk per epilogue warp
  static constexpr size_t SMEM_BYTES = 1024 /*align slack*/ + size_t(STAGES) * STAGE_BYTES + EPI_WARPS * EPI_STAGE_BYTES +
                                       BN * 4 /*column sums*/ + 256 /*barriers*/;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

// Epilogue feature mask (template parameter EF of the kernels): call sites with a known combination get a kernel in
// which the unused branches do not exist (the fully unrolled epilogue is otherwise ~3700 instructions and the eight
// epilogue warps thrash the instruction cache: measured +30-45% on the store-bound GEMMs); EF_RUNTIME keeps every
// branch and tests the EpiParams fields at run time.
enum : int { EF_BIAS = 1, EF_RELU = 2, EF_GATE = 4, EF_RES = 8, EF_COLSUM = 16, EF_ACC = 32, EF_RUNTIME = 64,
              EF_OUT_H = 128, EF_GATE_H = 256, EF_SCALE = 512, EF_DROP = 1024 };
__host__ __device__ inline int epi_mask(const EpiParams& ep) {
  return (ep.bias ? EF_BIAS : 0) | (ep.relu ? EF_RELU : 0) | (ep.gate_src ? EF_GATE : 0) | (ep.residual ? EF_RES : 0) |
         (ep.colsum ? EF_COLSUM : 0) | (ep.accumulate ? EF_ACC : 0) | (ep.out_h ? EF_OUT_H : 0) |
         (ep.gate_h ? EF_GATE_H : 0) | (ep.scale_mode ? EF_SCALE : 0) | (ep.drop.thr ? EF_DROP : 0);
}
template <int EF> __device__ __forceinline__ bool ef_bias(const EpiParams& ep) { return EF == EF_RUNTIME ? ep.bias != nullptr : (EF & EF_BIAS) != 0; }
template <int EF> __device__ __forceinline__ bool ef_relu(const EpiParams& ep) { return EF == EF_RUNTIME ? ep.relu != 0 : (EF & EF_RELU) != 0; }
template <int EF> __device__ __forceinline__ bool ef_gate(const EpiParams& ep) { return EF == EF_RUNTIME ? ep.gate_src != nullptr : (EF & EF_GATE) != 0; }
template <int EF> __device__ __forceinline__ bool ef_res(const EpiParams& ep) { return EF == EF_RUNTIME ? ep.residual != nullptr : (EF & EF_RES) != 0; }
template <int EF> __device__ __forceinline__ bool ef_colsum(const EpiParams& ep) { return EF == EF_RUNTIME ? ep.colsum != nullptr : (EF & EF_COLSUM) != 0; }
template <int EF> __device__ __forceinline__ bool ef_acc(const EpiParams& ep) { return EF == EF_RUNTIME ? ep.accumulate != 0 : (EF & EF_ACC) != 0; }
template <int EF> __device__ __forceinline__ bool ef_out_h(const EpiParams& ep) { return EF == EF_RUNTIME ? ep.out_h != nullptr : (EF & EF_OUT_H) != 0; }
template <int EF> __device__ __forceinline__ bool ef_gate_h(const EpiParams& ep) { return EF == EF_RUNTIME ? ep.gate_h != nullptr : (EF & EF_GATE_H) != 0; }
template <int EF> __device__ __forceinline__ bool ef_drop(const EpiParams& ep) { return EF == EF_RUNTIME ? ep.drop.thr != 0 : (EF & EF_DROP) != 0; }
template <int EF> __device__ __forceinline__ bool ef_scale(const EpiParams& ep) { return EF == EF_RUNTIME ? ep.scale_mode != 0 : (EF & EF_SCALE) != 0; }

// Operand of the epilogue that comes from global memory, fetched for a whole 32x32 chunk BEFORE the accumulator is
// read so that the 8 independent 128-bit loads per lane are in flight together instead of one DRAM round trip per
// row group (measured: 25 us per tile for the serial form of dH = (dU W2) * [h > 0]).  Only ONE operand is
// prefetched - the gate source, else the residual, else the running output of an accumulating store; a second one
// (no call site has it on the hot path) is read inside the loop.
struct EpiAux {
  float4 a[8];
};
template <int EF>
__device__ __forceinline__ const float* epi_aux_src(const EpiParams& ep) {
  if (ef_gate<EF>(ep)) return ep.gate_src;
  if (ef_res<EF>(ep)) return ep.residual;
  if (ef_acc<EF>(ep)) return ep.out;
  return nullptr;
}
template <int EF>
__device__ __forceinline__ void epilogue_fetch_aux(const EpiParams& ep, EpiAux& aux, int row0, int M, int col0, int lane) {
  const int col = col0 + (lane & 7) * 4;
  if (ef_gate_h<EF>(ep)) {   // fp16 gate source: 4 halves = 8 bytes per lane, kept in the .x/.y words
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int row = row0 + it * 4 + (lane >> 3);
      uint2 g = make_uint2(0u, 0u);
      if (row < M) g = *reinterpret_cast<const uint2*>(ep.gate_h + size_t(row) * ep.ldo + col);
      aux.a[it].x = __uint_as_float(g.x);
      aux.a[it].y = __uint_as_float(g.y);
    }
    return;
  }
  const float* src = epi_aux_src<EF>(ep);
  if (src == nullptr) return;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int row = row0 + it * 4 + (lane >> 3);
    aux.a[it] = row < M ? *reinterpret_cast<const float4*>(src + size_t(row) * ep.ldo + col) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// One 32 (rows) x 32 (columns) accumulator chunk: v[] holds this lane's ROW (lane = row inside the warp's TMEM lane
// quarter).  The chunk is transposed through `stage` (4 KB, private to the warp, 16-byte units XOR-swizzled by the
// row so that both phases are bank-conflict free) and then processed with lane = (row % 4 group, 16-byte column
// unit): every global access of the warp covers 4 rows x 128 contiguous bytes.  Column sums of the final values
// (bias gradients) are reduced in the warp and added to the CTA's shared-memory accumulator.
template <int EF>
__device__ __forceinline__ void epilogue_store_chunk(const EpiParams& ep, float (&v)[32], const EpiAux& aux, uint8_t* stage,
                                                     int row0, int M, int col0, int lane,
                                                     float* s_colsum /* [32] of this chunk */, float alpha, float4 bias) {
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4)
    *reinterpret_cast<float4*>(stage + lane * 128 + (((j4 ^ lane) & 7) << 4)) =
        make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
  __syncwarp();
  const int cu = lane & 7;          // 16-byte unit (4 columns) inside the 128-byte row segment
  const int rsub = lane >> 3;       // row inside a group of 4
  const int col = col0 + cu * 4;
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
  // which optional operand travelled through aux (same priority as epi_aux_src)
  const bool gate_in_aux = ef_gate<EF>(ep) || ef_gate_h<EF>(ep);
  const bool res_in_aux = !gate_in_aux && ef_res<EF>(ep);
  const bool acc_in_aux = !gate_in_aux && !res_in_aux && ef_acc<EF>(ep);
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + rsub;
    float4 x = *reinterpret_cast<const float4*>(stage + r * 128 + (((cu ^ r) & 7) << 4));
    const int row = row0 + r;
    if (row < M) {
      x.x = fmaf(x.x, alpha, bias.x); x.y = fmaf(x.y, alpha, bias.y);
      x.z = fmaf(x.z, alpha, bias.z); x.w = fmaf(x.w, alpha, bias.w);
      if (ef_relu<EF>(ep)) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
      const size_t off = size_t(row) * ep.ldo + col;
      if (ef_drop<EF>(ep)) {
        const uint64_t bits = drop_bits(ep.drop.seed, ep.drop_site, off >> 2);
        x.x *= drop_factor(bits, 0, ep.drop.thr, ep.drop.scale); x.y *= drop_factor(bits, 1, ep.drop.thr, ep.drop.scale);
        x.z *= drop_factor(bits, 2, ep.drop.thr, ep.drop.scale); x.w *= drop_factor(bits, 3, ep.drop.thr, ep.drop.scale);
      }
      if (ef_gate_h<EF>(ep)) {
        // fp16 values are >= +0 after the ReLU: "> 0" is "any bit set" of the 16-bit pattern
        const uint32_t g0 = __float_as_uint(aux.a[it].x), g1 = __float_as_uint(aux.a[it].y);
        x.x = (g0 & 0xffffu) ? x.x : 0.f; x.y = (g0 >> 16) ? x.y : 0.f;
        x.z = (g1 & 0xffffu) ? x.z : 0.f; x.w = (g1 >> 16) ? x.w : 0.f;
      } else if (ef_gate<EF>(ep)) {
        const float4 g = aux.a[it];
        x.x = g.x > 0.f ? x.x : 0.f; x.y = g.y > 0.f ? x.y : 0.f;
        x.z = g.z > 0.f ? x.z : 0.f; x.w = g.w > 0.f ? x.w : 0.f;
      }
      if (ef_res<EF>(ep)) {
        const float4 q = res_in_aux ? aux.a[it] : *reinterpret_cast<const float4*>(ep.residual + off);
        x.x += q.x; x.y += q.y; x.z += q.z; x.w += q.w;
      }
      if (ef_out_h<EF>(ep)) {
        const __half2 lo = __floats2half2_rn(x.x, x.y), hi = __floats2half2_rn(x.z, x.w);
        *reinterpret_cast<uint2*>(ep.out_h + off) =
            make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
      } else {
        float4* o = reinterpret_cast<float4*>(ep.out + off);
        if (ef_acc<EF>(ep)) {
          const float4 q = acc_in_aux ? aux.a[it] : *o;
          x.x += q.x; x.y += q.y; x.z += q.z; x.w += q.w;
        }
        *o = x;
      }
      if (ef_colsum<EF>(ep)) { cs.x += x.x; cs.y += x.y; cs.z += x.z; cs.w += x.w; }
    }
  }
  if (ef_colsum<EF>(ep)) {
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
      cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
    }
    if (lane < 8) {
      atomicAdd(s_colsum + cu * 4 + 0, cs.x); atomicAdd(s_colsum + cu * 4 + 1, cs.y);
      atomicAdd(s_colsum + cu * 4 + 2, cs.z); atomicAdd(s_colsum + cu * 4 + 3, cs.w);
    }
  }
  __syncwarp();   // the staging buffer is rewritten by the next chunk
}

// MN-major B: its stage is
//                   BN/32 boxes of 32 k-rows x 32 columns in the SWIZZLE_128B_BASE32B layout.
template <int BN, int OP, int EF>
__global__ void __launch_bounds__(GemmTnCfg<BN, OP>::THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N,
               int K, EpiParams ep, int b_res) {
  using Cfg = GemmTnCfg<BN, OP>;
  constexpr bool kBMajorN = OP == OP_TF32_N;
  constexpr bool kF16 = OP == OP_F16_K;
  constexpr int BKE = kF16 ? 64 : 32;     // elements per k-block (always 128 bytes)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* epi_stage = smem + size_t(Cfg::STAGES) * Cfg::STAGE_BYTES;
  float* s_colsum = reinterpret_cast<float*>(epi_stage + Cfg::EPI_WARPS * Cfg::EPI_STAGE_BYTES);   // [BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_colsum + BN);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::NBAR;
  uint64_t* tfull = bars + 2 * Cfg::NBAR;
  uint64_t* tempty = tfull + 2;
  uint64_t* bres = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres + 1);
  // b_res (B-resident mode, chosen by the host when all k-blocks of the CTA's B slice fit in RES_B_BYTES): the slice is
  // loaded ONCE - the column-stationary schedule never changes it - and only A streams through a ring of RES_STAGES
  // stages behind it.  For the K = 128 GEMMs (FFN1, dH, QKV) this removes 2/3 of the L2 -> SM operand traffic.
  uint8_t* a_ring = smem + Cfg::RES_B_BYTES;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Tile schedule: gridDim.x is a multiple of tiles_n (launch_tn), CTA b owns the column block n = b % tiles_n for
  // its whole life and walks the row blocks m = b / tiles_n, + gridDim.x / tiles_n, ...  The CTAs that run side by
  // side share their A row block through L2, every CTA re-reads the same B slice (L2-resident weights), and the
  // column sums of a CTA stay in registers until the kernel ends.
  const int tiles_m = (M + Cfg::BM - 1) / Cfg::BM;
  const int tiles_n = N / BN;
  const int num_kb = (K + BKE - 1) / BKE;
  const int n0 = (int(blockIdx.x) % tiles_n) * BN;
  const int m_first = int(blockIdx.x) / tiles_n;
  const int m_step = int(gridDim.x) / tiles_n;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      for (int s = 0; s < Cfg::NBAR; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
      for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], Cfg::EPI_WARPS); }
      mbar_init(bres, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  for (int j = threadIdx.x; j < BN; j += blockDim.x) s_colsum[j] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      uint32_t it = 0;
      if (b_res) {
        mbar_expect_tx(bres, uint32_t(num_kb) * Cfg::B_BYTES);
        for (int kb = 0; kb < num_kb; ++kb) {
          uint8_t* sb = smem + size_t(kb) * Cfg::B_BYTES;
          if constexpr (!kBMajorN) {
            tma_load_2d(sb, &tmB, bres, kb * BKE, n0);
          } else {
#pragma unroll
            for (int b = 0; b < BN / 32; ++b) tma_load_2d(sb + b * 4096, &tmB, bres, n0 + b * 32, kb * BKE);
          }
        }
        for (int mt = m_first; mt < tiles_m; mt += m_step) {
          const int m0 = mt * Cfg::BM;
          for (int kb = 0; kb < num_kb; ++kb, ++it) {
            const uint32_t s = it % Cfg::RES_STAGES, ph = (it / Cfg::RES_STAGES) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            mbar_expect_tx(&full[s], Cfg::A_BYTES);
            tma_load_2d(a_ring + size_t(s) * Cfg::A_BYTES, &tmA, &full[s], kb * BKE, m0);
          }
        }
      } else {
        for (int mt = m_first; mt < tiles_m; mt += m_step) {
          const int m0 = mt * Cfg::BM;
          for (int kb = 0; kb < num_kb; ++kb, ++it) {
            const uint32_t s = it % Cfg::STAGES, ph = (it / Cfg::STAGES) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
            uint8_t* sa = smem + size_t(s) * Cfg::STAGE_BYTES;
            tma_load_2d(sa, &tmA, &full[s], kb * BKE, m0);
            if constexpr (!kBMajorN) {
              tma_load_2d(sa + Cfg::A_BYTES, &tmB, &full[s], kb * BKE, n0);
            } else {
#pragma unroll
              for (int b = 0; b < BN / 32; ++b)
                tma_load_2d(sa + Cfg::A_BYTES + b * 4096, &tmB, &full[s], n0 + b * 32, kb * BKE);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kF16 ? kFmtF16 : kFmtTF32, Cfg::BM, BN, false, kBMajorN);
      uint32_t it = 0, lt = 0;
      if (b_res) mbar_wait(bres, 0);
      for (int mt = m_first; mt < tiles_m; mt += m_step, ++lt) {
        const uint32_t buf = lt & 1, bph = (lt >> 1) & 1;
        mbar_wait(&tempty[buf], bph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t nst = b_res ? Cfg::RES_STAGES : Cfg::STAGES;
          const uint32_t s = it % nst, ph = (it / nst) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = b_res ? smem_u32(a_ring + size_t(s) * Cfg::A_BYTES) : smem_u32(smem + size_t(s) * Cfg::STAGE_BYTES);
          const uint32_t b_addr = b_res ? smem_u32(smem + size_t(kb) * Cfg::B_BYTES) : a_addr + Cfg::A_BYTES;
          const uint64_t da = make_smem_desc_sw128(a_addr, 16, 1024);
          const uint64_t db = kBMajorN ? make_smem_desc_sw128(b_addr, 4096, 512, kLayoutSw128Base32)
                                       : make_smem_desc_sw128(b_addr, 16, 1024);
          // per MMA (K = 8): A advances 32 B inside its 128 B row (+2); a K-major B likewise, an
          // MN-major B advances 8 k-rows = 1024 B (+64)
          constexpr uint64_t kBStep = kBMajorN ? 64 : 2;
#pragma unroll
          for (int k = 0; k < 4; ++k) {   // 32 bytes of K per MMA: 8 tf32 or 16 fp16
            if constexpr (kF16) umma_f16(d_tmem, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
            else umma_tf32(d_tmem, da + uint64_t(2 * k), db + kBStep * uint64_t(k), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&tfull[buf]);
      }
    }
  } else {
    // ------------------------------ epilogue ------------------------------
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, +32) are the only ones this warp may read
    const int ew = warp - 2;       // 0 .. EPI_WARPS-1
    constexpr int kColSplit = Cfg::EPI_WARPS / 4;       // warps sharing a lane quarter interleave the 32-column chunks
    const int csub = ew >> 2;
    uint8_t* stage = epi_stage + ew * Cfg::EPI_STAGE_BYTES;
    // result scale: the host value times the device-side power-of-two factor (read once, not per chunk)
    float alpha = ep.alpha;
    if (ef_scale<EF>(ep) && ep.scale_mode != 3) alpha *= ep.scale_ptr[ep.scale_mode == 1 ? 0 : 1];
    uint32_t lt = 0;
    for (int mt = m_first; mt < tiles_m; mt += m_step, ++lt) {
      const uint32_t buf = lt & 1, bph = (lt >> 1) & 1;
      const int m0 = mt * Cfg::BM;
      // the gate / residual operands of the first chunk do not depend on the accumulator: fetch them while the MMAs run
      EpiAux aux;
      epilogue_fetch_aux<EF>(ep, aux, m0 + quarter * 32, M, n0 + csub * 32, lane);
      mbar_wait(&tfull[buf], bph);
      tc_fence_after();
#pragma unroll 1
      for (int c = csub; c < BN / 32; c += kColSplit) {
        float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);   // this lane's 4 columns; in flight during the TMEM read
        if (ef_bias<EF>(ep)) bias = __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + c * 32 + (lane & 7) * 4));
        float v[32];
        tmem_ld32(tmem_base + (uint32_t(quarter * 32) << 16) + buf * BN + c * 32, v);
        const EpiAux cur = aux;
        if (c + kColSplit < BN / 32)   // next chunk's operands: in flight during this chunk's stores
          epilogue_fetch_aux<EF>(ep, aux, m0 + quarter * 32, M, n0 + (c + kColSplit) * 32, lane);
        epilogue_store_chunk<EF>(ep, v, cur, stage, m0 + quarter * 32, M, n0 + c * 32, lane, s_colsum + c * 32, alpha, bias);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);
    }
    if (ef_colsum<EF>(ep)) {
      // all epilogue warps have added their partial column sums into s_colsum: one global atomic per column and CTA
      asm volatile("bar.sync 1, %0;" ::"n"(Cfg::EPI_WARPS * 32) : "memory");
      const float cscale = (ef_scale<EF>(ep) && ep.scale_mode != 2) ? ep.scale_ptr[1] : 1.f;
      for (int j = threadIdx.x - 64; j < BN; j += Cfg::EPI_WARPS * 32) atomicAdd(ep.colsum + n0 + j, s_colsum[j] * cscale);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

// -----------------------------------------------------------------------------------------------
// Weight-gradient GEMM:  C[M,N] += sum over tokens t of A[t, m] * B[t, n].
// A is [T, lda] row-major (M contiguous), B is [T, ldb] row-major (N contiguous): both MN-major.
// One smem "box" = 32 tokens x 32 columns (128 B rows, TMA SWIZZLE_128B_ATOM_32B): the UMMA atom
// (layout SWIZZLE_128B_BASE32B, the only MN-major layout for 32-bit operands) is 4 tokens x 128 B,
// SBO = 512 B (next 4 tokens), LBO = 4096 B (next 32-column block); one K=8 MMA spans two atoms.
// grid = (tiles_m * tiles_n, splits); every CTA reduces a contiguous token range and red.adds its
// partial tile into C (C must be initialised by the caller: zeros or the running gradient).
// -----------------------------------------------------------------------------------------------
template <int BN, bool kF16, bool kColsum = false>
struct GemmDwCfg {
  static constexpr int BM = 128;
  static constexpr int BT = kF16 ? 64 : 32;           // tokens per stage (4 MMAs of K = 8 tf32 / 16 fp16)
  static constexpr int BOX_COLS = kF16 ? 64 : 32;     // 128-byte rows
  static constexpr int BOX_BYTES = 128 * BT;          // 4 KB (tf32) / 8 KB (fp16)
  static constexpr int A_BYTES = (BM / BOX_COLS) * BOX_BYTES;
  static constexpr int B_BYTES = (BN / BOX_COLS) * BOX_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (STAGE_BYTES >= 48 * 1024) ? 4 : 6;
  // accumulator columns [0, BN) + 32 columns for the optional column-sum product (A^T . ones), as a power of two
  static constexpr int ACC_COLS = kColsum ? BN + 32 : BN;
  static constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : ACC_COLS <= 256 ? 256 : 512;
  static constexpr int ONES_BYTES = kColsum ? BOX_BYTES : 0;   // one box filled with 1.0: the B operand of the column-sum product
  static constexpr size_t SMEM_BYTES = 1024 + size_t(STAGES) * STAGE_BYTES + ONES_BYTES + 256;
};

// kF16: both operands fp16, MN-major with the plain SWIZZLE_128B layout (an atom is 8 tokens x 128 B = 64 columns;
// SBO = 1024 B to the next 8 tokens, LBO = one box to the next 64 columns; one K = 16 MMA spans two atoms).
// alpha_ptr (optional): the partial tile is multiplied by alpha * alpha_ptr[0] (device-side unscale of fp16 gradients).
// colsum (optional): colsum[m] += alpha * sum_t A[t, m] -- the bias gradient that goes with the weight gradient.  It is
// one more (N = 16) MMA per k-step against a shared-memory box of ones, so A is not read a second time by a separate
// column-sum kernel; only the CTAs of the first column block (n0 == 0) add it.  kColsum is a template flag: the plain
// instantiation keeps its smaller TMEM allocation and issue loop.
template <int BN, bool kF16, bool kColsum>
__global__ void __launch_bounds__(192, 1)
gemm_dw_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int T, int M,
               int N, float* __restrict__ C, int ldc, float alpha, const float* __restrict__ alpha_ptr,
               float* __restrict__ colsum) {
  using Cfg = GemmDwCfg<BN, kF16, kColsum>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* ones = smem + size_t(Cfg::STAGES) * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ones + Cfg::ONES_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::STAGES;
  uint64_t* tfull = bars + 2 * Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_n = N / BN;
  const int m0 = (blockIdx.x / tiles_n) * Cfg::BM;
  const int n0 = (blockIdx.x % tiles_n) * BN;
  // token range of this split, in units of BT-token blocks
  const int num_tb = (T + Cfg::BT - 1) / Cfg::BT;
  const int per = (num_tb + gridDim.y - 1) / gridDim.y;
  const int tb0 = blockIdx.y * per;
  const int tb1 = min(num_tb, tb0 + per);
  const int nblk = tb1 - tb0;  // may be <= 0 for trailing splits
  const bool do_colsum = kColsum && colsum != nullptr && n0 == 0;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
      mbar_init(&tfull[0], 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  if (do_colsum) {
    const uint32_t one = kF16 ? 0x3c003c00u : 0x3f800000u;
    for (int i = threadIdx.x; i < Cfg::ONES_BYTES / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(ones)[i] = one;
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (nblk > 0) {
    if (warp == 0) {
      if (lane == 0) {
        for (int i = 0; i < nblk; ++i) {
          const uint32_t s = i % Cfg::STAGES, ph = (i / Cfg::STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
          uint8_t* sa = smem + size_t(s) * Cfg::STAGE_BYTES;
          const int t0 = (tb0 + i) * Cfg::BT;
#pragma unroll
          for (int b = 0; b < Cfg::BM / Cfg::BOX_COLS; ++b)
            tma_load_2d(sa + b * Cfg::BOX_BYTES, &tmA, &full[s], m0 + b * Cfg::BOX_COLS, t0);
#pragma unroll
          for (int b = 0; b < BN / Cfg::BOX_COLS; ++b)
            tma_load_2d(sa + Cfg::A_BYTES + b * Cfg::BOX_BYTES, &tmB, &full[s], n0 + b * Cfg::BOX_COLS, t0);
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        constexpr uint32_t idesc = make_idesc(kF16 ? kFmtF16 : kFmtTF32, Cfg::BM, BN, true, true);
        constexpr uint32_t idesc_cs = make_idesc(kF16 ? kFmtF16 : kFmtTF32, Cfg::BM, 16, true, true);
        constexpr uint32_t kLayout = kF16 ? kLayoutSw128 : kLayoutSw128Base32;
        constexpr uint32_t kSbo = kF16 ? 1024 : 512;
        // next MMA: 8 tf32 tokens = 1024 B (+64 in the address field), 16 fp16 tokens = 2048 B (+128)
        constexpr uint64_t kStep = kF16 ? 128 : 64;
        for (int i = 0; i < nblk; ++i) {
          const uint32_t s = i % Cfg::STAGES, ph = (i / Cfg::STAGES) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + size_t(s) * Cfg::STAGE_BYTES);
          const uint64_t da = make_smem_desc_sw128(a_addr, Cfg::BOX_BYTES, kSbo, kLayout);
          const uint64_t db = make_smem_desc_sw128(a_addr + Cfg::A_BYTES, Cfg::BOX_BYTES, kSbo, kLayout);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if constexpr (kF16) umma_f16(tmem_base, da + kStep * uint64_t(k), db + kStep * uint64_t(k), idesc, (i | k) != 0 ? 1u : 0u);
            else umma_tf32(tmem_base, da + kStep * uint64_t(k), db + kStep * uint64_t(k), idesc, (i | k) != 0 ? 1u : 0u);
          }
          if (do_colsum) {                     // [128, 16] += A^T . ones: every column is the column sum of A's rows
            const uint64_t d1 = make_smem_desc_sw128(smem_u32(ones), Cfg::BOX_BYTES, kSbo, kLayout);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if constexpr (kF16) umma_f16(tmem_base + BN, da + kStep * uint64_t(k), d1, idesc_cs, (i | k) != 0 ? 1u : 0u);
              else umma_tf32(tmem_base + BN, da + kStep * uint64_t(k), d1, idesc_cs, (i | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&tfull[0]);
      }
    } else {
      const int quarter = warp & 3;
      const float a = alpha_ptr != nullptr ? alpha * alpha_ptr[0] : alpha;
      mbar_wait(&tfull[0], 0);
      tc_fence_after();
      const int row = m0 + quarter * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        float v[32];
        tmem_ld32(tmem_base + (uint32_t(quarter * 32) << 16) + c * 32, v);
        if (row < M) {
          float* dst = C + size_t(row) * ldc + n0 + c * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(dst + j, a * v[j]);
        }
      }
      if (do_colsum) {
        float v[32];
        tmem_ld32(tmem_base + (uint32_t(quarter * 32) << 16) + BN, v);
        if (row < M) atomicAdd(colsum + row, a * v[0]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

}  // namespace rlt
