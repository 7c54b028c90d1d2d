// Persistent, warp-specialised TF32 GEMMs on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
//
//   gemm_tn_kernel : C[M,N]  = A[M,K] * B[N,K]^T         (both operands K-major; Linear fwd, dX)
//   gemm_dw_kernel : C[M,N] += sum_t A[t,M] * B[t,N]     (both operands MN-major; weight grads,
//                                                         contraction over the token axis, split
//                                                         over CTAs, fp32 red.add into C)
//
// Roles per CTA (192 threads): warp 0 = TMA producer, warp 1 = single-thread MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> global), one TMEM lane (= output row) per thread.
// Accumulators are double-buffered in TMEM so the epilogue of tile i overlaps the main loop of
// tile i+1.  Operands are fp32 containers holding TF32-rounded values (see to_tf32()).
#pragma once
#include "sm100.cuh"

namespace rlt {

struct EpiParams {
  float* out;             // [M, ldo] fp32 result (may be null)
  float* out_tf32;        // same values rounded to tf32 (operand copy for the next GEMM; may be null)
  int ldo;
  const float* bias;      // [N] added to every row (may be null)
  const float* gate_src;  // [M, ldo]: result *= (gate_src > 0)   (ReLU backward; may be null)
  const float* residual;  // [M, ldo]: result += residual          (skip connections; may be null)
  float* colsum;          // [N]: atomicAdd of the column sums of the final values (bias grads; may be null)
  int relu;               // max(x, 0)
  int accumulate;         // out += result instead of out = result
  float alpha;            // result scale applied first
  int tag;                // KernelTag of the call site (in-situ timing; 0 = none)
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

template <int BN>
struct GemmTnCfg {
  static constexpr int BM = 128;
  static constexpr int BK = 32;  // 32 tf32 = 128 B = one swizzle row
  static constexpr int A_BYTES = BM * BK * 4;
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 128 ? 6 : 8);
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  static constexpr size_t SMEM_BYTES = 1024 /*align slack*/ + size_t(STAGES) * STAGE_BYTES + 256 /*barriers*/;
};

__device__ __forceinline__ void epilogue_store_chunk(const EpiParams& ep, float (&v)[32], int row, int M,
                                                     int col0, int lane, float* s_colsum /*per warp [32] or null*/) {
  const bool row_ok = row < M;
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] *= ep.alpha;
  if (ep.bias != nullptr) {
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + col0) + j4);
      v[4 * j4 + 0] += b.x; v[4 * j4 + 1] += b.y; v[4 * j4 + 2] += b.z; v[4 * j4 + 3] += b.w;
    }
  }
  if (ep.relu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  const size_t off = size_t(row) * ep.ldo + col0;
  if (ep.gate_src != nullptr && row_ok) {
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(ep.gate_src + off) + j4);
      v[4 * j4 + 0] = g.x > 0.f ? v[4 * j4 + 0] : 0.f;
      v[4 * j4 + 1] = g.y > 0.f ? v[4 * j4 + 1] : 0.f;
      v[4 * j4 + 2] = g.z > 0.f ? v[4 * j4 + 2] : 0.f;
      v[4 * j4 + 3] = g.w > 0.f ? v[4 * j4 + 3] : 0.f;
    }
  }
  if (ep.residual != nullptr && row_ok) {
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      const float4 r = __ldg(reinterpret_cast<const float4*>(ep.residual + off) + j4);
      v[4 * j4 + 0] += r.x; v[4 * j4 + 1] += r.y; v[4 * j4 + 2] += r.z; v[4 * j4 + 3] += r.w;
    }
  }
  if (row_ok) {
    if (ep.out != nullptr) {
      float4* o = reinterpret_cast<float4*>(ep.out + off);
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        float4 r = make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
        if (ep.accumulate) {
          const float4 p = o[j4];
          r.x += p.x; r.y += p.y; r.z += p.z; r.w += p.w;
          v[4 * j4] = r.x; v[4 * j4 + 1] = r.y; v[4 * j4 + 2] = r.z; v[4 * j4 + 3] = r.w;
        }
        o[j4] = r;
      }
    }
    if (ep.out_tf32 != nullptr) {
      float4* o = reinterpret_cast<float4*>(ep.out_tf32 + off);
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4)
        o[j4] = make_float4(to_tf32(v[4 * j4]), to_tf32(v[4 * j4 + 1]), to_tf32(v[4 * j4 + 2]),
                            to_tf32(v[4 * j4 + 3]));
    }
  }
  if (ep.colsum != nullptr) {
    if (!row_ok) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0.f;
    }
    const float s = warp_colsum32(v, lane);
    atomicAdd(ep.colsum + col0 + lane, s);
  }
}

// kBMajorN = false: B is [N, K] row-major (K-major operand, nn.Linear weight used as-is: x W^T).
// kBMajorN = true : B is [K, N] row-major (MN-major operand: x W with W stored [K, N]); its stage is
//                   BN/32 boxes of 32 k-rows x 32 columns in the SWIZZLE_128B_BASE32B layout.
template <int BN, bool kBMajorN>
__global__ void __launch_bounds__(192, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N,
               int K, EpiParams ep) {
  using Cfg = GemmTnCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + size_t(Cfg::STAGES) * Cfg::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::STAGES;
  uint64_t* tfull = bars + 2 * Cfg::STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_m = (M + Cfg::BM - 1) / Cfg::BM;
  const int tiles_n = N / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (K + Cfg::BK - 1) / Cfg::BK;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
      for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 4); }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * Cfg::BM;
        const int n0 = (tile % tiles_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % Cfg::STAGES, ph = (it / Cfg::STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
          uint8_t* sa = smem + size_t(s) * Cfg::STAGE_BYTES;
          tma_load_2d(sa, &tmA, &full[s], kb * Cfg::BK, m0);
          if constexpr (!kBMajorN) {
            tma_load_2d(sa + Cfg::A_BYTES, &tmB, &full[s], kb * Cfg::BK, n0);
          } else {
#pragma unroll
            for (int b = 0; b < BN / 32; ++b)
              tma_load_2d(sa + Cfg::A_BYTES + b * 4096, &tmB, &full[s], n0 + b * 32, kb * Cfg::BK);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kFmtTF32, Cfg::BM, BN, false, kBMajorN);
      uint32_t it = 0, lt = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
        const uint32_t buf = lt & 1, bph = (lt >> 1) & 1;
        mbar_wait(&tempty[buf], bph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % Cfg::STAGES, ph = (it / Cfg::STAGES) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + size_t(s) * Cfg::STAGE_BYTES);
          const uint64_t da = make_smem_desc_sw128(a_addr, 16, 1024);
          const uint64_t db = kBMajorN ? make_smem_desc_sw128(a_addr + Cfg::A_BYTES, 4096, 512, kLayoutSw128Base32)
                                       : make_smem_desc_sw128(a_addr + Cfg::A_BYTES, 16, 1024);
          // per MMA (K = 8): A advances 32 B inside its 128 B row (+2); a K-major B likewise, an
          // MN-major B advances 8 k-rows = 1024 B (+64)
          constexpr uint64_t kBStep = kBMajorN ? 64 : 2;
#pragma unroll
          for (int k = 0; k < Cfg::BK / 8; ++k)
            umma_tf32(d_tmem, da + uint64_t(2 * k), db + kBStep * uint64_t(k), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty[s]);
        }
        umma_commit(&tfull[buf]);
      }
    }
  } else {
    // ------------------------------ epilogue ------------------------------
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, +32) are the only ones this warp may read
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const uint32_t buf = lt & 1, bph = (lt >> 1) & 1;
      const int m0 = (tile / tiles_n) * Cfg::BM;
      const int n0 = (tile % tiles_n) * BN;
      mbar_wait(&tfull[buf], bph);
      tc_fence_after();
      const int row = m0 + quarter * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        float v[32];
        tmem_ld32(tmem_base + (uint32_t(quarter * 32) << 16) + buf * BN + c * 32, v);
        epilogue_store_chunk(ep, v, row, M, n0 + c * 32, lane, nullptr);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);
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
template <int BN>
struct GemmDwCfg {
  static constexpr int BM = 128;
  static constexpr int BT = 32;  // tokens per stage (4 MMAs of K = 8)
  static constexpr int BOX_BYTES = 32 * BT * 4;  // 4 KB
  static constexpr int A_BYTES = (BM / 32) * BOX_BYTES;
  static constexpr int B_BYTES = (BN / 32) * BOX_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 4 : 6;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  static constexpr size_t SMEM_BYTES = 1024 + size_t(STAGES) * STAGE_BYTES + 256;
};

template <int BN>
__global__ void __launch_bounds__(192, 1)
gemm_dw_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int T, int M,
               int N, float* __restrict__ C, int ldc, float alpha) {
  using Cfg = GemmDwCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + size_t(Cfg::STAGES) * Cfg::STAGE_BYTES);
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
          for (int b = 0; b < Cfg::BM / 32; ++b) tma_load_2d(sa + b * Cfg::BOX_BYTES, &tmA, &full[s], m0 + b * 32, t0);
#pragma unroll
          for (int b = 0; b < BN / 32; ++b)
            tma_load_2d(sa + Cfg::A_BYTES + b * Cfg::BOX_BYTES, &tmB, &full[s], n0 + b * 32, t0);
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        constexpr uint32_t idesc = make_idesc(kFmtTF32, Cfg::BM, BN, true, true);
        for (int i = 0; i < nblk; ++i) {
          const uint32_t s = i % Cfg::STAGES, ph = (i / Cfg::STAGES) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + size_t(s) * Cfg::STAGE_BYTES);
          const uint64_t da = make_smem_desc_sw128(a_addr, Cfg::BOX_BYTES, 512, kLayoutSw128Base32);
          const uint64_t db = make_smem_desc_sw128(a_addr + Cfg::A_BYTES, Cfg::BOX_BYTES, 512, kLayoutSw128Base32);
#pragma unroll
          for (int k = 0; k < Cfg::BT / 8; ++k)  // next 8 tokens = +1024 B: +64 in the address field
            umma_tf32(tmem_base, da + uint64_t(64 * k), db + uint64_t(64 * k), idesc, (i | k) != 0 ? 1u : 0u);
          umma_commit(&empty[s]);
        }
        umma_commit(&tfull[0]);
      }
    } else {
      const int quarter = warp & 3;
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
          for (int j = 0; j < 32; ++j) atomicAdd(dst + j, alpha * v[j]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

}  // namespace rlt
