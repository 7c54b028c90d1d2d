// Fused feed-forward backward of the encoder layer (d_model = 128, fp16 hidden path): ONE pass over the hidden
// activation h [T, d_ff] produces
//     dH  = alpha (s dU2) W2 * [h > 0]          (fp16, written once; the dW1 and dY GEMMs read it)
//     db1 += colsum(dH) / s
//     dW2 += (s dU2)^T h / s                     (lin2 weight gradient, [d, d_ff])
// dW2 needs exactly the two operands the dH GEMM already has in shared memory (dU2 tile, h tile), so the separate
// weight-gradient GEMM and its second 5 GB read of h disappear.  (Folding dW1 = dH^T y in as well was measured:
// the y tiles cost a pipeline stage of the h ring and the kernel became latency-bound, 3.1 ms against 1.8 + 0.76.)
//
// Column-stationary persistent CTAs: a CTA owns 128 hidden units (n0) and walks token tiles of 128.  Per tile:
//     MMA1  S[tok, hid]      = dU2[tok, :] . W2^T[hid, :]        K = d      operands K-major
//     MMA3  dW2acc[d, hid]  += dU2[tok, d]^T . h[tok, hid]       K = tokens operands MN-major (the SAME smem tiles)
//     epilogue: S -> gate by h (read from the TMA-loaded h tile) -> fp16 -> per-warp SWIZZLE_64B staging -> TMA store
// A [128 tok x 128 col] fp16 tile is two TMA boxes of [128 rows x 64 columns] (SWIZZLE_128B): read as K-major
// (k-block = box, SBO 1024) for MMA1 and as MN-major (LBO = box, SBO 1024, 16 tokens per MMA) for the transposed
// product, so no operand is transposed or loaded twice.
// TMEM: S double buffer 2 x 128 columns | dW2acc 128 columns; the weight-gradient accumulator stays in TMEM for the
// CTA's whole token range and is red.added to global once at the end.
// Shared memory: W2 slice 32 KB | dU2 ring 2 x 32 KB (L2-resident operand) | h ring 3 x 32 KB (the HBM stream) |
// dH staging 32 KB.
// Warp roles: 0 TMA producer, 1 MMA issuer, 2..17 epilogue (TMEM lane quarter x 32-column chunk: one chunk per warp
// and tile).
#pragma once
#include "gemm_tc.cuh"

namespace rlt {

struct FfnBwdCfg {
  static constexpr int BM = 128, BN = 128, D = 128;
  static constexpr int KB_BYTES = 128 * 128;               // one box: 128 rows x 64 fp16 columns
  static constexpr int TILE_BYTES = 2 * KB_BYTES;
  static constexpr int DU_STAGES = 2, H_STAGES = 3;
  static constexpr int EPI_WARPS = 16;
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
  static constexpr int OFF_W2 = 0;                                  // resident W2^T slice [128 hid x 128 d]
  static constexpr int OFF_DU = OFF_W2 + TILE_BYTES;
  static constexpr int OFF_H = OFF_DU + DU_STAGES * TILE_BYTES;
  static constexpr int OFF_DH = OFF_H + H_STAGES * TILE_BYTES;      // dH staging tile
  static constexpr int OFF_COLSUM = OFF_DH + TILE_BYTES;
  static constexpr int OFF_BARS = OFF_COLSUM + BN * 4;
  static constexpr int N_BARS = 2 * DU_STAGES + 2 * H_STAGES + 4 + 2;
  static constexpr size_t SMEM_BYTES = 1024 + OFF_BARS + N_BARS * 8 + 16;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

__global__ void __launch_bounds__(FfnBwdCfg::THREADS, 1)
ffn_bwd_kernel(const __grid_constant__ CUtensorMap tmDU, const __grid_constant__ CUtensorMap tmW2,
               const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmDH, int T, int F, float alpha,
               const float* __restrict__ scale_ptr /* {s, 1/s} */, float* __restrict__ db1, float* __restrict__ dW2) {
  using Cfg = FfnBwdCfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sW2 = smem + Cfg::OFF_W2;
  uint8_t* sDU = smem + Cfg::OFF_DU;
  uint8_t* sH = smem + Cfg::OFF_H;
  uint8_t* sDH = smem + Cfg::OFF_DH;
  float* s_colsum = reinterpret_cast<float*>(smem + Cfg::OFF_COLSUM);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BARS);
  uint64_t* full_du = bars;                          // TMA landed
  uint64_t* empty_du = full_du + Cfg::DU_STAGES;     // MMA3 of the tile retired
  uint64_t* full_h = empty_du + Cfg::DU_STAGES;
  uint64_t* empty_h = full_h + Cfg::H_STAGES;        // MMA3 retired + every epilogue warp has read its gate rows (count 17)
  uint64_t* tfull = empty_h + Cfg::H_STAGES;         // [2] S accumulator ready
  uint64_t* tempty = tfull + 2;                      // [2] S accumulator drained (count 16)
  uint64_t* bres = tempty + 2;                       // W2 slice landed
  uint64_t* acc_done = bres + 1;                     // every MMA of the CTA retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + Cfg::N_BARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (T + Cfg::BM - 1) / Cfg::BM;
  const int tiles_n = F / Cfg::BN;
  const int n0 = (int(blockIdx.x) % tiles_n) * Cfg::BN;
  const int m_first = int(blockIdx.x) / tiles_n;
  const int m_step = int(gridDim.x) / tiles_n;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmDU);
      tma_prefetch_desc(&tmW2);
      tma_prefetch_desc(&tmH);
      tma_prefetch_desc(&tmDH);
      for (int s = 0; s < Cfg::DU_STAGES; ++s) { mbar_init(&full_du[s], 1); mbar_init(&empty_du[s], 1); }
      for (int s = 0; s < Cfg::H_STAGES; ++s) { mbar_init(&full_h[s], 1); mbar_init(&empty_h[s], 1 + Cfg::EPI_WARPS); }
      for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], Cfg::EPI_WARPS); }
      mbar_init(bres, 1);
      mbar_init(acc_done, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  for (int j = threadIdx.x; j < Cfg::BN; j += blockDim.x) s_colsum[j] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t kColW2 = 256;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      mbar_expect_tx(bres, Cfg::TILE_BYTES);
      for (int kb = 0; kb < 2; ++kb) tma_load_2d(sW2 + kb * Cfg::KB_BYTES, &tmW2, bres, kb * 64, n0);
      uint32_t lt = 0;
      for (int mt = m_first; mt < tiles_m; mt += m_step, ++lt) {
        const uint32_t sd = lt % Cfg::DU_STAGES, pd = (lt / Cfg::DU_STAGES) & 1;
        const uint32_t sh = lt % Cfg::H_STAGES, phh = (lt / Cfg::H_STAGES) & 1;
        mbar_wait(&empty_h[sh], phh ^ 1);
        mbar_expect_tx(&full_h[sh], Cfg::TILE_BYTES);
        for (int kb = 0; kb < 2; ++kb)
          tma_load_2d(sH + sh * Cfg::TILE_BYTES + kb * Cfg::KB_BYTES, &tmH, &full_h[sh], n0 + kb * 64, mt * Cfg::BM);
        mbar_wait(&empty_du[sd], pd ^ 1);
        mbar_expect_tx(&full_du[sd], Cfg::TILE_BYTES);
        for (int kb = 0; kb < 2; ++kb)
          tma_load_2d(sDU + sd * Cfg::TILE_BYTES + kb * Cfg::KB_BYTES, &tmDU, &full_du[sd], kb * 64, mt * Cfg::BM);
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc_k = make_idesc(kFmtF16, 128, 128, false, false);
      constexpr uint32_t idesc_mn = make_idesc(kFmtF16, 128, 128, true, true);
      const uint32_t w2_addr = smem_u32(sW2), du_addr = smem_u32(sDU), h_addr = smem_u32(sH);
      uint32_t lt = 0;
      mbar_wait(bres, 0);
      for (int mt = m_first; mt < tiles_m; mt += m_step, ++lt) {
        const uint32_t sd = lt % Cfg::DU_STAGES, pd = (lt / Cfg::DU_STAGES) & 1;
        const uint32_t sh = lt % Cfg::H_STAGES, phh = (lt / Cfg::H_STAGES) & 1;
        const uint32_t buf = lt & 1, bph = (lt >> 1) & 1;
        mbar_wait(&full_du[sd], pd);
        mbar_wait(&tempty[buf], bph ^ 1);
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t da = make_smem_desc_sw128(du_addr + sd * Cfg::TILE_BYTES + kb * Cfg::KB_BYTES, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(w2_addr + kb * Cfg::KB_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base + buf * Cfg::BN, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc_k, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&tfull[buf]);
        // dW2acc[d, hid] += dU2^T h
        mbar_wait(&full_h[sh], phh);
        tc_fence_after();
        {
          const uint64_t da = make_smem_desc_sw128(du_addr + sd * Cfg::TILE_BYTES, Cfg::KB_BYTES, 1024);
          const uint64_t db = make_smem_desc_sw128(h_addr + sh * Cfg::TILE_BYTES, Cfg::KB_BYTES, 1024);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_f16(tmem_base + kColW2, da + uint64_t(128 * k), db + uint64_t(128 * k), idesc_mn, (lt | uint32_t(k)) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_du[sd]);
        umma_commit(&empty_h[sh]);
      }
      umma_commit(acc_done);
    }
  } else {
    // ------------------------------ epilogue ------------------------------
    const int quarter = warp & 3;        // TMEM lane quarter
    const int ew = warp - 2;             // 0..15
    const int csub = ew >> 2;            // 32-column chunk of the 128-column tile
    const int r = quarter * 32 + lane;   // row inside the tile
    // this thread's 64 bytes inside a [128 x 128] fp16 tile: box csub/2, row r, 16-byte units 4 (csub%2) + u, swizzled
    const uint32_t tile_off = uint32_t(csub >> 1) * Cfg::KB_BYTES + uint32_t(r) * 128;
    const uint32_t ub = uint32_t(csub & 1) * 4, rx = uint32_t(r & 7);
    const uint32_t lane_tmem = tmem_base + (uint32_t(quarter * 32) << 16) + csub * 32;
    uint8_t* st_out = sDH + ew * 2048;   // this warp's [32 rows x 32 columns] fp16 staging tile (SWIZZLE_64B)
    const int sw = (lane >> 1) & 3;      // SWIZZLE_64B: 16-byte unit index ^= (row >> 1) & 3

    uint32_t lt = 0;
    for (int mt = m_first; mt < tiles_m; mt += m_step, ++lt) {
      const uint32_t sh = lt % Cfg::H_STAGES, phh = (lt / Cfg::H_STAGES) & 1;
      const uint32_t buf = lt & 1, bph = (lt >> 1) & 1;
      mbar_wait(&tfull[buf], bph);
      tc_fence_after();
      float v[32];
      tmem_ld32(lane_tmem + buf * Cfg::BN, v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);            // the accumulator is in registers
      // ---- gate: the relu / keep mask is "h != 0" (fp16 h >= +0)
      mbar_wait(&full_h[sh], phh);
      const uint8_t* g = sH + sh * Cfg::TILE_BYTES + tile_off;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint4 q = *reinterpret_cast<const uint4*>(g + (((ub + u) ^ rx) << 4));
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          v[8 * u + 2 * e] = (w[e] & 0xffffu) ? v[8 * u + 2 * e] * alpha : 0.f;
          v[8 * u + 2 * e + 1] = (w[e] >> 16) ? v[8 * u + 2 * e + 1] * alpha : 0.f;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_h[sh]);
      // ---- pack, stage, hand over to the copy engine
      uint32_t h[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const __half2 p2 = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
        h[k] = *reinterpret_cast<const uint32_t*>(&p2);
      }
      if (lane == 0) tma_store_wait_read<0>();             // the previous store of this warp has drained its staging tile
      __syncwarp();
#pragma unroll
      for (int u = 0; u < 4; ++u)
        *reinterpret_cast<uint4*>(st_out + lane * 64 + ((u ^ sw) << 4)) = make_uint4(h[4 * u], h[4 * u + 1], h[4 * u + 2], h[4 * u + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&tmDH, st_out, n0 + csub * 32, mt * Cfg::BM + quarter * 32);
        tma_store_commit();
      }
      // ---- bias gradient: column sums of the fp32 values
      const float tot = warp_colsum32(v, lane);
      atomicAdd(s_colsum + csub * 32 + lane, tot);
    }
    if (lane == 0) tma_store_wait<0>();
    if (lt > 0) {
      // ---- weight-gradient accumulator: TMEM -> red.add into the running gradient
      const float inv_s = scale_ptr[1];
      mbar_wait(acc_done, 0);
      tc_fence_after();
      float v[32];
      tmem_ld32(tmem_base + (uint32_t(quarter * 32) << 16) + kColW2 + csub * 32, v);
      float* dst = dW2 + size_t(r) * F + n0 + csub * 32;               // row = d, columns = hidden units
#pragma unroll
      for (int j = 0; j < 32; ++j) atomicAdd(dst + j, v[j] * inv_s);
      asm volatile("bar.sync 1, %0;" ::"n"(Cfg::EPI_WARPS * 32) : "memory");
      for (int j = threadIdx.x - 64; j < Cfg::BN; j += Cfg::EPI_WARPS * 32) atomicAdd(db1 + n0 + j, s_colsum[j] * inv_s);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}

}  // namespace rlt
