// C(fp16)[M, N] = epilogue( A(fp16)[M, K] * B(fp16)[N, K]^T ) for the two hidden-sized, store-bound GEMMs of the
// encoder layer (K = d_model <= 128, N = d_ff):
//     FFN1 : h  = dropout(relu(y W1^T + b1))                       (EF_BIAS | EF_RELU [| EF_DROP])
//     dH   : dH = alpha (s dU2) W2 * [h > 0], db1 += colsum / s     (EF_GATE_H | EF_COLSUM | EF_SCALE)
// ncu on the generic gemm_tn kernel (profiles/r01_ncu_ffn1.txt, r01_ncu_dh.txt): the SM's L1/shared-memory pipe is 93 %
// busy in FFN1 (fp32 staging STS + LDS + 64-byte STG per row segment + tensor-core operand reads), and the dH
// epilogue waits on its register-prefetched gate loads.  Here ALL global traffic of the epilogue is done by the copy
// engine: the thread that owns an accumulator row (tcgen05.ld 32x32b) finishes the row in registers, packs it to
// fp16, writes 64 bytes into a SWIZZLE_64B staging tile and one lane issues `cp.async.bulk.tensor` shared -> global;
// the gate tile arrives the same way (TMA load, two chunks ahead, per-warp mbarriers).  No LDS/STG transposition,
// half the staging bytes, no per-thread global addresses.
// Same persistent structure as gemm_tn_kernel: column-stationary CTAs (BN = 256), the B slice resident in shared
// memory, A streaming through a 3-stage ring, double-buffered TMEM accumulators, 16 epilogue warps.
#pragma once
#include "gemm_tc.cuh"

namespace rlt {

// BN = 256 for K <= 128 (d_model 128), BN = 128 for K <= 256 (d_model 256): the resident B slice is 64 KB either way.
template <int BN_>
struct GemmF16OutCfg {
  static constexpr int BM = 128, BN = BN_, BKE = 64;
  static constexpr int A_BYTES = BM * 128;                 // one k-block: 128 rows x 128 B
  static constexpr int B_BYTES = BN * 128;
  static constexpr int MAX_KB = 64 * 1024 / B_BYTES;       // 2 (K <= 128) or 4 (K <= 256)
  static constexpr int A_STAGES = 3;
  static constexpr int EPI_WARPS = 16;
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
  static constexpr int TILE_BYTES = 32 * 64;               // 32 rows x 32 fp16 columns
  static constexpr int EPI_WARP_BYTES = 3 * TILE_BYTES;    // out staging + 2 gate buffers
  static constexpr int OFF_A = MAX_KB * B_BYTES;
  static constexpr int OFF_EPI = OFF_A + A_STAGES * A_BYTES;
  static constexpr int OFF_BIAS = OFF_EPI + EPI_WARPS * EPI_WARP_BYTES;
  static constexpr int OFF_COLSUM = OFF_BIAS + BN * 4;
  static constexpr int OFF_BARS = OFF_COLSUM + BN * 4;
  static constexpr int N_BARS = 2 * A_STAGES + 2 + 2 + 1 + 2 * EPI_WARPS;
  static constexpr size_t SMEM_BYTES = 1024 + OFF_BARS + N_BARS * 8 + 16;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

template <int BN_, int EF>
__global__ void __launch_bounds__(GemmF16OutCfg<BN_>::THREADS, 1)
gemm_f16out_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmGate, int M, int N,
                   int K, EpiParams ep) {
  using Cfg = GemmF16OutCfg<BN_>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* a_ring = smem + Cfg::OFF_A;
  float* s_bias = reinterpret_cast<float*>(smem + Cfg::OFF_BIAS);
  float* s_colsum = reinterpret_cast<float*>(smem + Cfg::OFF_COLSUM);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BARS);
  uint64_t* full = bars;
  uint64_t* empty = full + Cfg::A_STAGES;
  uint64_t* tfull = empty + Cfg::A_STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* bres = tempty + 2;
  uint64_t* gbar = bres + 1;                               // [EPI_WARPS][2] gate tile landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbar + 2 * Cfg::EPI_WARPS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (M + Cfg::BM - 1) / Cfg::BM;
  const int tiles_n = N / Cfg::BN;
  const int num_kb = (K + Cfg::BKE - 1) / Cfg::BKE;
  const int n0 = (int(blockIdx.x) % tiles_n) * Cfg::BN;
  const int m_first = int(blockIdx.x) / tiles_n;
  const int m_step = int(gridDim.x) / tiles_n;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      tma_prefetch_desc(&tmOut);
      if (ef_gate_h<EF>(ep)) tma_prefetch_desc(&tmGate);
      for (int s = 0; s < Cfg::A_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
      for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], Cfg::EPI_WARPS); }
      mbar_init(bres, 1);
      for (int i = 0; i < 2 * Cfg::EPI_WARPS; ++i) mbar_init(&gbar[i], 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<2 * Cfg::BN>(tmem_slot);
  }
  for (int j = threadIdx.x; j < Cfg::BN; j += blockDim.x) {
    s_colsum[j] = 0.f;
    s_bias[j] = ef_bias<EF>(ep) ? ep.bias[n0 + j] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      mbar_expect_tx(bres, uint32_t(num_kb) * Cfg::B_BYTES);
      for (int kb = 0; kb < num_kb; ++kb) tma_load_2d(smem + size_t(kb) * Cfg::B_BYTES, &tmB, bres, kb * Cfg::BKE, n0);
      uint32_t it = 0;
      for (int mt = m_first; mt < tiles_m; mt += m_step) {
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % Cfg::A_STAGES, ph = (it / Cfg::A_STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], Cfg::A_BYTES);
          tma_load_2d(a_ring + size_t(s) * Cfg::A_BYTES, &tmA, &full[s], kb * Cfg::BKE, mt * Cfg::BM);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kFmtF16, Cfg::BM, Cfg::BN, false, false);
      uint32_t it = 0, lt = 0;
      mbar_wait(bres, 0);
      for (int mt = m_first; mt < tiles_m; mt += m_step, ++lt) {
        const uint32_t buf = lt & 1, bph = (lt >> 1) & 1;
        mbar_wait(&tempty[buf], bph ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % Cfg::A_STAGES, ph = (it / Cfg::A_STAGES) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint64_t da = make_smem_desc_sw128(smem_u32(a_ring + size_t(s) * Cfg::A_BYTES), 16, 1024);
          const uint64_t db = make_smem_desc_sw128(smem_u32(smem + size_t(kb) * Cfg::B_BYTES), 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base + buf * Cfg::BN, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty[s]);
        }
        umma_commit(&tfull[buf]);
      }
    }
  } else {
    // ------------------------------ epilogue ------------------------------
    const int quarter = warp & 3;        // TMEM lane quarter
    const int ew = warp - 2;             // 0..15
    const int csub = ew >> 2;            // this warp takes the 32-column chunks c = csub, csub + 4
    constexpr int kChunks = Cfg::BN / 32 / 4;
    uint8_t* st_out = smem + Cfg::OFF_EPI + ew * Cfg::EPI_WARP_BYTES;
    uint8_t* st_gate = st_out + Cfg::TILE_BYTES;          // two buffers
    uint64_t* my_gbar = gbar + 2 * ew;
    const int sw = (lane >> 1) & 3;                       // SWIZZLE_64B: 16-byte unit index ^= (row >> 1) & 3
    float alpha = ep.alpha;
    if (ef_scale<EF>(ep) && ep.scale_mode != 3) alpha *= ep.scale_ptr[ep.scale_mode == 1 ? 0 : 1];

    // gate tiles are independent of the accumulator: requested two chunks ahead of their use
    int n_my_tiles = 0;
    for (int mt = m_first; mt < tiles_m; mt += m_step) ++n_my_tiles;
    const uint32_t total_chunks = uint32_t(n_my_tiles) * kChunks;
    auto request_gate = [&](uint32_t n) {                 // chunk n of this warp: tile n / kChunks, column chunk csub + 4 (n % kChunks)
      if (!ef_gate_h<EF>(ep) || n >= total_chunks) return;
      if (lane == 0) {
        const int mt = m_first + int(n / kChunks) * m_step;
        const int c = csub + 4 * int(n % kChunks);
        mbar_expect_tx(&my_gbar[n & 1], Cfg::TILE_BYTES);
        tma_load_2d(st_gate + (n & 1) * Cfg::TILE_BYTES, &tmGate, &my_gbar[n & 1], n0 + c * 32, mt * Cfg::BM + quarter * 32);
      }
    };
    request_gate(0);
    request_gate(1);

    uint32_t lt = 0, n = 0;
    for (int mt = m_first; mt < tiles_m; mt += m_step, ++lt) {
      const uint32_t buf = lt & 1, bph = (lt >> 1) & 1;
      const int row = mt * Cfg::BM + quarter * 32 + lane;
      const bool row_ok = row < M;
      mbar_wait(&tfull[buf], bph);
      tc_fence_after();
#pragma unroll 1
      for (int i = 0; i < kChunks; ++i, ++n) {
        const int c = csub + 4 * i;
        float v[32];
        tmem_ld32(tmem_base + (uint32_t(quarter * 32) << 16) + buf * Cfg::BN + c * 32, v);
        // ---- row math in registers
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 b = *reinterpret_cast<const float4*>(s_bias + c * 32 + j4 * 4);   // same address in every lane: broadcast
          v[4 * j4 + 0] = fmaf(v[4 * j4 + 0], alpha, b.x); v[4 * j4 + 1] = fmaf(v[4 * j4 + 1], alpha, b.y);
          v[4 * j4 + 2] = fmaf(v[4 * j4 + 2], alpha, b.z); v[4 * j4 + 3] = fmaf(v[4 * j4 + 3], alpha, b.w);
        }
        if (ef_relu<EF>(ep)) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (ef_drop<EF>(ep)) {
          const size_t e0 = size_t(row) * ep.ldo + n0 + c * 32;
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const uint64_t bits = drop_bits(ep.drop.seed, ep.drop_site, (e0 >> 2) + j4);
#pragma unroll
            for (int e = 0; e < 4; ++e) v[4 * j4 + e] *= drop_factor(bits, e, ep.drop.thr, ep.drop.scale);
          }
        }
        if (ef_gate_h<EF>(ep)) {
          mbar_wait(&my_gbar[n & 1], (n >> 1) & 1);
          const uint8_t* g = st_gate + (n & 1) * Cfg::TILE_BYTES + lane * 64;
#pragma unroll
          for (int u = 0; u < 4; ++u) {                  // 16-byte unit u of this row = columns 8u .. 8u+7
            const uint4 q = *reinterpret_cast<const uint4*>(g + ((u ^ sw) << 4));
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {                // fp16 values are >= +0 after the ReLU: "> 0" is "any bit set"
              v[8 * u + 2 * e] = (w[e] & 0xffffu) ? v[8 * u + 2 * e] : 0.f;
              v[8 * u + 2 * e + 1] = (w[e] >> 16) ? v[8 * u + 2 * e + 1] : 0.f;
            }
          }
          // every lane has read its gate row: the buffer may be refilled with the tile two chunks ahead
          fence_proxy_async_smem();
          __syncwarp();
          request_gate(n + 2);
        }
        if (!row_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        // ---- pack to fp16 and stage (SWIZZLE_64B: conflict-free 16-byte units), then let the copy engine store
        uint32_t h[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const __half2 p2 = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
          h[k] = *reinterpret_cast<const uint32_t*>(&p2);
        }
        if (lane == 0) tma_store_wait_read<0>();          // the previous store of this warp has drained the staging tile
        __syncwarp();
#pragma unroll
        for (int u = 0; u < 4; ++u)
          *reinterpret_cast<uint4*>(st_out + lane * 64 + ((u ^ sw) << 4)) = make_uint4(h[4 * u], h[4 * u + 1], h[4 * u + 2], h[4 * u + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmOut, st_out, n0 + c * 32, mt * Cfg::BM + quarter * 32);
          tma_store_commit();
        }
        if (ef_colsum<EF>(ep)) {
          const float tot = warp_colsum32(v, lane);       // lane j <- sum over the 32 rows of column j
          atomicAdd(s_colsum + c * 32 + lane, tot);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);
    }
    if (lane == 0) tma_store_wait<0>();                   // global writes performed before the CTA exits
    if (ef_colsum<EF>(ep)) {
      asm volatile("bar.sync 1, %0;" ::"n"(Cfg::EPI_WARPS * 32) : "memory");
      const float cscale = (ef_scale<EF>(ep) && ep.scale_mode != 2) ? ep.scale_ptr[1] : 1.f;
      for (int j = threadIdx.x - 64; j < Cfg::BN; j += Cfg::EPI_WARPS * 32) atomicAdd(ep.colsum + n0 + j, s_colsum[j] * cscale);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<2 * Cfg::BN>(tmem_base);
}

}  // namespace rlt
