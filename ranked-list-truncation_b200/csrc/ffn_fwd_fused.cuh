// Fused feed-forward block of the encoder layer (torch TransformerEncoderLayer._ff_block + norm2, post-norm):
//
//     out = LayerNorm2( y + relu(y W1^T + b1) W2^T + b2 )          y [T, d], W1 [f, d], W2 [d, f], f = 2048
//
// in ONE kernel: the hidden activation h = relu(y W1^T + b1) (T x 2048, 4 KB per token in fp16) never goes to HBM in
// inference and is written exactly once -- never read back -- in training (the backward kernels consume it).  It
// replaces three launches (FFN1 GEMM -> h, FFN2 GEMM <- h, LayerNorm2) whose h round trip bounded them at the HBM
// roofline; fused, the block is bound by the tensor pipe.
//
// Work decomposition.  A CTA PAIR (cluster of 2, `tcgen05 ... cta_group::2`) owns a tile of 256 tokens, 128 rows per
// CTA, and walks the hidden dimension in chunks of CH units:
//     MMA1  S_j [256 x CH]  = y16 [256 x d] . W1_j^T              A: shared memory (this CTA's 128 rows of y, fp16)
//                                                                 B: W1 rows of chunk j, HALF per CTA
//     epilogue (16 warps per CTA, thread = row): S_j + b1 -> relu -> fp16 -> H_j written to TENSOR MEMORY
//                                                (train: the same registers go to HBM as the saved hidden)
//     MMA2  Z [256 x d]    += H_j [256 x CH] . W2_j^T             A: tensor memory (tcgen05.mma with a TMEM A operand)
//                                                                 B: W2 columns of chunk j, HALF of the d rows per CTA
// and after the last chunk: Z + b2 + y (fp32 residual from HBM) -> LayerNorm over d -> out (+ u2, statistics when saved).
// The weights stream through two TMA rings; with the pair each CTA loads and feeds the tensor core HALF of every
// weight chunk, so a 256-token tile costs 1 MB of L2 -> SM traffic (2 MB for two independent CTAs -- more than the
// measured L2 throughput at full tensor rate) and B-operand shared-memory reads are halved.  Putting H in tensor memory
// takes the whole A operand of MMA2 (4 KB per MMA) off the shared-memory pipe.
//
// Tensor memory (512 columns per CTA): d 256: S double buffer 2 x 64 | Z 256 | H double buffer 2 x 32 (448);
// d 128: S TRIPLE buffer 3 x 128 with H_g written in place over the score columns its warp has read | Z 128 (512).
// Shared memory: y16 tile(s) 128 x d fp16 | W1 ring 3 x 16 KB (half-chunks) | W2 ring 4 x 16 KB | b1 | b2, gamma, beta |
// LayerNorm partials | 16 x 2 KB staging tiles of the saved hidden (TMA store) | barriers.
// Warp roles per CTA: 0 TMA producer (both CTAs load their own halves; the bytes are counted on the LEADER's barriers),
// 1 MMA1 issuer and 18 MMA2 issuer (leader CTA only), 2..17 epilogue (TMEM lane quarter = warp % 4, column quarter =
// (warp - 2) / 4).
// Barriers consumed by the MMA thread live in the leader CTA; the peer's epilogue warps arrive through the cluster
// window (mapa); barriers consumed by producers / epilogues are local and signalled by multicast tcgen05.commit.
#pragma once
#include <cuda_fp16.h>

#include "gemm_tc.cuh"

namespace rlt {

// packed fp32 add (sm_100 FADD2): two IEEE additions per instruction
__device__ __forceinline__ float2 add_f32x2(float2 a, float2 b) {
  unsigned long long x, y, r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(b.x), "f"(b.y));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(x), "l"(y));
  float2 o;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(r));
  return o;
}

template <int D>
struct FfnFwdCfg {
  static_assert(D == 128 || D == 256, "d_model");
  static constexpr int BM = 128;                       // rows per CTA (256 per pair)
  static constexpr int CH = 128 * 128 / D;             // hidden units per chunk
  static constexpr int NS1 = 3;                        // W1 ring stages (freed by MMA1 of the chunk)
  static constexpr int NS2 = 4;                        // W2 ring stages (freed by MMA2, which trails MMA1)
  static constexpr int YB = D == 128 ? 2 : 1;          // y16 tile buffers
  static constexpr int EPI_WARPS = 16;
  static constexpr int THREADS = 64 + 32 * EPI_WARPS + 32;     // + the MMA2 issuer warp
  static constexpr int Y_BOX_BYTES = BM * 128;         // [128 rows x 64 fp16]
  static constexpr int Y_BYTES = (D / 64) * Y_BOX_BYTES;
  static constexpr int W1_BOX_BYTES = (CH / 2) * 128;  // [CH/2 rows x 64 fp16]
  static constexpr int W1_BYTES = (D / 64) * W1_BOX_BYTES;
  static constexpr int W2_BOX_BYTES = (D / 2) * 128;   // [D/2 rows x 64 fp16]
  static constexpr int W2_BYTES = (CH / 64) * W2_BOX_BYTES;
  static constexpr int MAX_F = 2048;
  static constexpr int OFF_Y = 0;
  static constexpr int OFF_W1 = OFF_Y + YB * Y_BYTES;
  static constexpr int OFF_W2 = OFF_W1 + NS1 * W1_BYTES;
  static constexpr int OFF_B1 = OFF_W2 + NS2 * W2_BYTES;
  static constexpr int OFF_VEC = OFF_B1 + MAX_F * 4;             // b2 | gamma | beta
  static constexpr int OFF_RED = OFF_VEC + 3 * D * 4;            // [2][4 quarters][4 column quarters][32 lanes]
  static constexpr int OFF_STAGE = ((OFF_RED + 2 * 4 * 4 * 32 * 4 + 1023) / 1024) * 1024;   // per-warp [32 x 32] fp16 tiles of the saved hidden
  static constexpr int OFF_BARS = OFF_STAGE + EPI_WARPS * 2048;
  // d_model 128: H_g is written IN PLACE over the S_g columns its own warp has just read (the attention kernel does the
  // same with P), which frees 128 tensor-memory columns for a THIRD S / H buffer: the per-buffer dependency loop
  // MMA1_g -> bias / ReLU / pack -> MMA2_g -> MMA1_{g+NSB} (~2 900 cycles, measured with two buffers: 1 450 per chunk) is
  // then shared by three chunks in flight instead of two.
  static constexpr bool ALIAS_H = D == 128;
  static constexpr int NSB = ALIAS_H ? 3 : 2;          // S (and H) buffers
  static constexpr int N_BARS = 2 * NS1 + 2 * NS2 + 2 * YB + 4 * NSB + 2;
  static constexpr size_t SMEM_BYTES = 1024 + OFF_BARS + N_BARS * 8 + 16;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
  // tensor memory columns
  static constexpr uint32_t COL_S = 0;                 // NSB x CH
  static constexpr uint32_t COL_Z = NSB * CH;          // D
  static constexpr uint32_t COL_H = 2 * CH + D;        // 2 x CH/2 (not ALIAS_H)
  static_assert(ALIAS_H ? (COL_Z + D <= 512) : (COL_H + CH <= 512), "tensor memory budget");
};

struct FfnFwdParams {
  const float* y;        // [T, D] fp32 residual
  const float* b1;       // [F]
  const float* b2;       // [D]
  const float* gamma;    // [D]
  const float* beta;     // [D]
  float* out;            // [T, D]
  float* u2;             // [T, D] pre-norm sum (optional)
  float* stats;          // [T, 2] mean, rstd (optional)
  __half* h_out;         // [T, F] hidden (optional: training)
  int T, F;
  float eps;
  long long* dbg;        // optional timeline of pair 0's leader (tools/bench_ffn.py --timeline): [64 chunks][16] clock64 stamps
};

template <int D>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FfnFwdCfg<D>::THREADS, 1)
ffn_fwd_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmW1,
               const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmH,
               const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmU2, const FfnFwdParams p) {
  using Cfg = FfnFwdCfg<D>;
  constexpr int CH = Cfg::CH, NS1 = Cfg::NS1, NS2 = Cfg::NS2, YB = Cfg::YB;
  constexpr int SC = CH / 4;       // S columns per epilogue thread
  constexpr int ZC = D / 4;        // Z columns per epilogue thread
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sY = smem + Cfg::OFF_Y;
  uint8_t* sW1 = smem + Cfg::OFF_W1;
  uint8_t* sW2 = smem + Cfg::OFF_W2;
  float* sB1 = reinterpret_cast<float*>(smem + Cfg::OFF_B1);
  float* sVec = reinterpret_cast<float*>(smem + Cfg::OFF_VEC);
  float* sRed = reinterpret_cast<float*>(smem + Cfg::OFF_RED);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BARS);
  uint64_t* w1_full = bars;                // [NS1]  leader: both CTAs' halves of a W1 chunk landed
  uint64_t* w1_empty = w1_full + NS1;      // [NS1]  local: MMA1 of the chunk retired
  uint64_t* w2_full = w1_empty + NS1;      // [NS2]  leader
  uint64_t* w2_empty = w2_full + NS2;      // [NS2]  local: MMA2 of the chunk retired
  uint64_t* y_full = w2_empty + NS2;       // [YB]   leader
  uint64_t* y_empty = y_full + YB;         // [YB]   local: last MMA1 of the tile retired
  constexpr int NSB = Cfg::NSB;
  uint64_t* s_full = y_empty + YB;         // [NSB]  local: S_j ready
  uint64_t* s_empty = s_full + NSB;        // [NSB]  leader: both CTAs' epilogues have S_j in registers (count 32; unused with ALIAS_H)
  uint64_t* h_full = s_empty + NSB;        // [NSB]  leader: both CTAs' epilogues wrote H_j (count 32)
  uint64_t* h_empty = h_full + NSB;        // [NSB]  local: MMA2_j retired (ALIAS_H: the S buffer may be rewritten)
  uint64_t* z_full = h_empty + NSB;        // local
  uint64_t* z_empty = z_full + 1;          // leader (count 32)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + Cfg::N_BARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int n_tiles = (p.T + 2 * Cfg::BM - 1) / (2 * Cfg::BM);
  const int n_chunks = p.F / CH;
  const int pair = int(blockIdx.x) >> 1, n_pairs = int(gridDim.x) >> 1;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmY);
      tma_prefetch_desc(&tmW1);
      tma_prefetch_desc(&tmW2);
      for (int s = 0; s < NS1; ++s) { mbar_init(&w1_full[s], 1); mbar_init(&w1_empty[s], 1); }
      for (int s = 0; s < NS2; ++s) { mbar_init(&w2_full[s], 1); mbar_init(&w2_empty[s], 1); }
      for (int s = 0; s < YB; ++s) { mbar_init(&y_full[s], 1); mbar_init(&y_empty[s], 1); }
      for (int s = 0; s < NSB; ++s) {
        mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 2 * Cfg::EPI_WARPS);
        mbar_init(&h_full[s], 2 * Cfg::EPI_WARPS); mbar_init(&h_empty[s], 1);
      }
      mbar_init(z_full, 1);
      mbar_init(z_empty, 2 * Cfg::EPI_WARPS);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc_pair<512>(tmem_slot);
  }
  for (int j = threadIdx.x; j < p.F; j += blockDim.x) sB1[j] = p.b1[j];
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    sVec[j] = p.b2[j];
    sVec[D + j] = p.gamma[j];
    sVec[2 * D + j] = p.beta[j];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // barriers of both CTAs initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer (both CTAs) ------------------------------
    // One flat walk over this pair's chunks (all its tiles): W1 of chunk g, then W2 of chunk g - 2 -- the order in which
    // the MMA thread consumes them (MMA2 trails MMA1 by two chunks).  All ring / tile cursors advance incrementally: a
    // runtime integer division costs this single thread ~100 cycles.
    if (lane == 0) {
      const int my_tiles = pair < n_tiles ? (n_tiles - pair + n_pairs - 1) / n_pairs : 0;
      const uint32_t total = uint32_t(my_tiles) * uint32_t(n_chunks);
      uint32_t st2 = 0, ph2 = 0;
      int j2 = 0;                                   // chunk-in-tile of the next W2 load
      auto load_w2 = [&]() {
        mbar_wait(&w2_empty[st2], ph2 ^ 1);
        if (leader) mbar_expect_tx(&w2_full[st2], 2 * Cfg::W2_BYTES);
        const uint32_t bar = mapa_u32(smem_u32(&w2_full[st2]), 0);
        // W2 [D, F]: this CTA's half of the d rows, columns of chunk j2
        for (int kb = 0; kb < CH / 64; ++kb)
          tma_load_2d_pair(sW2 + st2 * Cfg::W2_BYTES + kb * Cfg::W2_BOX_BYTES, &tmW2, bar, j2 * CH + kb * 64, int(rank) * (D / 2));
        if (++st2 == NS2) { st2 = 0; ph2 ^= 1; }
        if (++j2 == n_chunks) j2 = 0;
      };
      uint32_t st1 = 0, ph1 = 0, yb = 0, yph = 0;
      int j = 0, tile = pair;
      // The y tile of the NEXT tile is requested in the middle of the current one (two buffers): it is the only operand
      // that comes from HBM, and at the tile boundary its latency under the hidden-store write stream held up the W1 / W2
      // loads queued behind it by ~2 500 cycles (tools/bench_ffn.py --timeline train).  With one buffer (d = 256) it can
      // only be requested at the boundary.
      auto load_y = [&]() {
        const int row0 = tile * 2 * Cfg::BM + int(rank) * Cfg::BM;
        mbar_wait(&y_empty[yb], yph ^ 1);
        if (leader) mbar_expect_tx(&y_full[yb], 2 * Cfg::Y_BYTES);
        const uint32_t bar = mapa_u32(smem_u32(&y_full[yb]), 0);
        for (int kb = 0; kb < D / 64; ++kb)
          tma_load_2d_pair(sY + yb * Cfg::Y_BYTES + kb * Cfg::Y_BOX_BYTES, &tmY, bar, kb * 64, row0);
        if (++yb == YB) { yb = 0; yph ^= 1; }
        tile += n_pairs;
      };
      constexpr bool kEarlyY = YB > 1;
      if (kEarlyY && total > 0) load_y();
      for (uint32_t g = 0; g < total; ++g) {
        if (kEarlyY ? (j == n_chunks / 2 && g + n_chunks / 2 < total) : (j == 0)) load_y();
        {
          mbar_wait(&w1_empty[st1], ph1 ^ 1);
          if (leader) mbar_expect_tx(&w1_full[st1], 2 * Cfg::W1_BYTES);
          const uint32_t bar = mapa_u32(smem_u32(&w1_full[st1]), 0);
          // W1 [F, D]: this CTA's half of chunk j = rows j CH + rank CH/2 ... (+ CH/2)
          for (int kb = 0; kb < D / 64; ++kb)
            tma_load_2d_pair(sW1 + st1 * Cfg::W1_BYTES + kb * Cfg::W1_BOX_BYTES, &tmW1, bar, kb * 64, j * CH + int(rank) * (CH / 2));
          if (++st1 == NS1) { st1 = 0; ph1 ^= 1; }
        }
        if (g >= 2) load_w2();
        if (++j == n_chunks) j = 0;
      }
      if (total >= 2) load_w2();
      if (total >= 1) load_w2();
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------ MMA1 issuer (leader CTA): S_g = y . W1_g^T ------------------------------
    // Two issuing threads (this warp: MMA1, warp 18: MMA2).  A tcgen05.mma holds its issuing thread for about the
    // duration of the instruction (measured ~96 cycles at M 256 / N 128), so ONE thread that also polls five barriers per
    // chunk leaves the tensor core idle a third of the time; two threads with independent barrier sets keep two
    // accumulation chains in flight, and tcgen05.commit only covers the committing thread's own MMAs -- which is exactly
    // the dependency each barrier needs.
    if (leader && lane == 0) {
      constexpr uint32_t idesc1 = make_idesc(kFmtF16, 256, CH, false, false);
      const uint32_t y_addr = smem_u32(sY), w1_addr = smem_u32(sW1);
      const int my_tiles = pair < n_tiles ? (n_tiles - pair + n_pairs - 1) / n_pairs : 0;
      const uint32_t total = uint32_t(my_tiles) * uint32_t(n_chunks);
      uint32_t st1 = 0, ph1 = 0, yb = 0, yph = 0, sb1 = 0, sp1 = 0;
      int j = 0;
      for (uint32_t g = 0; g < total; ++g) {
        const uint32_t buf1 = sb1, u1 = sp1;                       // S buffer of chunk g (g % NSB, parity of g / NSB)
        const bool dbg = p.dbg != nullptr && pair == 0 && g < 64;
        if (dbg) p.dbg[g * 24 + 0] = clock64();
        if (j == 0) mbar_wait_cluster(&y_full[yb], yph);
        mbar_wait_cluster(&w1_full[st1], ph1);
        if (dbg) p.dbg[g * 24 + 18] = clock64();
        if constexpr (Cfg::ALIAS_H) mbar_wait_cluster(&h_empty[buf1], u1 ^ 1);   // H_{g-NSB} (inside this buffer) consumed by MMA2
        else mbar_wait_cluster(&s_empty[buf1], u1 ^ 1);
        if (dbg) p.dbg[g * 24 + 1] = clock64();
        tc_fence_after();
        const uint32_t a1 = y_addr + yb * Cfg::Y_BYTES, b1a = w1_addr + st1 * Cfg::W1_BYTES;
#pragma unroll
        for (int kb = 0; kb < D / 64; ++kb) {
          const uint64_t da = make_smem_desc_sw128(a1 + kb * Cfg::Y_BOX_BYTES, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(b1a + kb * Cfg::W1_BOX_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_pair(tmem_base + Cfg::COL_S + buf1 * CH, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc1, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit_pair(&s_full[buf1]);
        umma_commit_pair(&w1_empty[st1]);
        if (++sb1 == uint32_t(NSB)) { sb1 = 0; sp1 ^= 1; }
        if (j == n_chunks - 1) {
          umma_commit_pair(&y_empty[yb]);
          if (++yb == YB) { yb = 0; yph ^= 1; }
          j = 0;
        } else {
          ++j;
        }
        if (++st1 == NS1) { st1 = 0; ph1 ^= 1; }
        if (dbg) p.dbg[g * 24 + 2] = clock64();
      }
    }
    __syncwarp();
  } else if (warp == 2 + Cfg::EPI_WARPS) {
    // ------------------------------ MMA2 issuer (leader CTA): Z += H_c . W2_c^T ------------------------------
    if (leader && lane == 0) {
      constexpr uint32_t idesc2 = make_idesc(kFmtF16, 256, D, false, false);
      const uint32_t w2_addr = smem_u32(sW2);
      const int my_tiles = pair < n_tiles ? (n_tiles - pair + n_pairs - 1) / n_pairs : 0;
      const uint32_t total = uint32_t(my_tiles) * uint32_t(n_chunks);
      uint32_t st2 = 0, ph2 = 0, ztile = 0, sb2 = 0, sp2 = 0;
      int cj = 0;
      for (uint32_t c = 0; c < total; ++c) {
        const uint32_t buf2 = sb2, u2 = sp2;                       // H buffer of chunk c
        const bool dbg = p.dbg != nullptr && pair == 0 && c < 64;
        if (dbg) p.dbg[c * 24 + 4] = clock64();
        mbar_wait_cluster(&w2_full[st2], ph2);
        if (dbg) p.dbg[c * 24 + 16] = clock64();
        mbar_wait_cluster(&h_full[buf2], u2);
        if (dbg) p.dbg[c * 24 + 17] = clock64();
        if (cj == 0) mbar_wait_cluster(z_empty, (ztile & 1) ^ 1);
        if (dbg) p.dbg[c * 24 + 5] = clock64();
        tc_fence_after();
        const uint32_t b2a = w2_addr + st2 * Cfg::W2_BYTES;
        // A operand in tensor memory: H_c contiguous behind Z, or (ALIAS_H) the 16 packed columns at the start of every
        // 32-column quarter of the S buffer -- K step ks = hidden units [16 ks, 16 ks + 16) = quarter ks / 2, half ks % 2
        const uint32_t a_t = Cfg::ALIAS_H ? tmem_base + Cfg::COL_S + buf2 * CH : tmem_base + Cfg::COL_H + buf2 * (CH / 2);
#pragma unroll
        for (int kb = 0; kb < CH / 64; ++kb) {
          const uint64_t db = make_smem_desc_sw128(b2a + kb * Cfg::W2_BOX_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int ks = kb * 4 + k;
            const uint32_t a_k = Cfg::ALIAS_H ? a_t + uint32_t((ks >> 1) * 32 + (ks & 1) * 8) : a_t + uint32_t(kb * 32 + k * 8);
            umma_f16_ts_pair(tmem_base + Cfg::COL_Z, a_k, db + uint64_t(2 * k), idesc2, (cj | kb | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit_pair(&h_empty[buf2]);
        umma_commit_pair(&w2_empty[st2]);
        if (++sb2 == uint32_t(NSB)) { sb2 = 0; sp2 ^= 1; }
        if (cj == n_chunks - 1) {
          umma_commit_pair(z_full);
          cj = 0;
          ++ztile;
        } else {
          ++cj;
        }
        if (++st2 == NS2) { st2 = 0; ph2 ^= 1; }
        if (dbg) p.dbg[c * 24 + 6] = clock64();
      }
    }
    __syncwarp();
  } else {
    // ------------------------------ epilogue (both CTAs) ------------------------------
    const int quarter = warp & 3;          // TMEM lane quarter this warp may access
    const int ew = warp - 2;
    const int cq = ew >> 2;                // column quarter
    const int r = quarter * 32 + lane;     // row inside this CTA's 128-row tile
    const uint32_t lane_tmem = tmem_base + (uint32_t(quarter * 32) << 16);
    uint32_t s_empty_r[NSB], h_full_r[NSB];
#pragma unroll
    for (int b = 0; b < NSB; ++b) {
      s_empty_r[b] = mapa_u32(smem_u32(&s_empty[b]), 0);
      h_full_r[b] = mapa_u32(smem_u32(&h_full[b]), 0);
    }
    const uint32_t z_empty_r = mapa_u32(smem_u32(z_empty), 0);
    uint8_t* st_out = smem + Cfg::OFF_STAGE + ew * 2048;     // this warp's [32 rows x 32 columns] fp16 staging tile (SWIZZLE_64B)
    const int sw = (lane >> 1) & 3;                          // SWIZZLE_64B: 16-byte unit index ^= (row >> 1) & 3
    float* red0 = sRed + (quarter * 4) * 32;                 // [cq][lane] partial sums of this lane quarter
    float* red1 = sRed + 4 * 4 * 32 + (quarter * 4) * 32;
    uint32_t g = 0, tt = 0, sbe = 0, spe = 0;
    for (int tile = pair; tile < n_tiles; tile += n_pairs, ++tt) {
      const long long row = (long long)tile * 2 * Cfg::BM + (long long)rank * Cfg::BM + r;
      const bool live = row < p.T;
      for (int j = 0; j < n_chunks; ++j, ++g) {
        const uint32_t buf = sbe, u = spe;
        if (++sbe == uint32_t(NSB)) { sbe = 0; spe ^= 1; }
        const bool dbg = p.dbg != nullptr && pair == 0 && leader && ew == 0 && lane == 0 && g < 64;
        if (dbg) p.dbg[g * 24 + 8] = clock64();
        mbar_wait(&s_full[buf], u);
        if (dbg) p.dbg[g * 24 + 9] = clock64();
        tc_fence_after();
        float v[SC];
        if constexpr (SC == 32) tmem_ld32(lane_tmem + Cfg::COL_S + buf * CH + cq * SC, v);
        else tmem_ld16(lane_tmem + Cfg::COL_S + buf * CH + cq * SC, v);
        tc_fence_before();
        __syncwarp();
        if (!Cfg::ALIAS_H && lane == 0) mbar_arrive_cluster(s_empty_r[buf]);          // S_j is in registers
        if (dbg) p.dbg[g * 24 + 10] = clock64();
        // ---- bias + relu + fp16 pack (two hidden units per 32-bit word = the TMEM A-operand layout): packed fp32x2
        // adds, one pack per pair, relu on the packed halves (max commutes with the rounding)
        const float4* bias = reinterpret_cast<const float4*>(sB1 + j * CH + cq * SC);
        uint32_t hp[SC / 2];
#pragma unroll
        for (int q4 = 0; q4 < SC / 4; ++q4) {
          const float4 b = bias[q4];
          const float2 s0 = add_f32x2(make_float2(v[4 * q4], v[4 * q4 + 1]), make_float2(b.x, b.y));
          const float2 s1 = add_f32x2(make_float2(v[4 * q4 + 2], v[4 * q4 + 3]), make_float2(b.z, b.w));
          const __half2 zero = __float2half2_rn(0.f);
          const __half2 lo = __hmax2(__floats2half2_rn(s0.x, s0.y), zero);
          const __half2 hi = __hmax2(__floats2half2_rn(s1.x, s1.y), zero);
          hp[2 * q4] = *reinterpret_cast<const uint32_t*>(&lo);
          hp[2 * q4 + 1] = *reinterpret_cast<const uint32_t*>(&hi);
        }
        if (p.h_out != nullptr) {
          // training: the hidden is saved -- written once by the copy engine (per-warp staging tile -> TMA store; per-thread
          // global stores of 32 rows 4 KB apart cost the LSU one cycle per row and instruction: measured +1 ms per launch);
          // rows beyond T are clipped by the tensor map
          if (lane == 0) tma_store_wait_read<0>();             // the previous store of this warp has drained its staging tile
          __syncwarp();
          if constexpr (SC == 32) {            // [32 rows x 32 columns] fp16, 64-byte rows, SWIZZLE_64B
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4)
              *reinterpret_cast<uint4*>(st_out + lane * 64 + ((q4 ^ sw) << 4)) = make_uint4(hp[4 * q4], hp[4 * q4 + 1], hp[4 * q4 + 2], hp[4 * q4 + 3]);
          } else {                             // [32 rows x 16 columns] fp16, 32-byte rows, no swizzle
#pragma unroll
            for (int q4 = 0; q4 < 2; ++q4)
              *reinterpret_cast<uint4*>(st_out + lane * 32 + (q4 << 4)) = make_uint4(hp[4 * q4], hp[4 * q4 + 1], hp[4 * q4 + 2], hp[4 * q4 + 3]);
          }
          // (the proxy fence and the TMA store are issued AFTER the H hand-over below: by then the staging writes have
          // landed and the fence costs nothing on the path MMA2 waits for)
        }
        if (dbg) p.dbg[g * 24 + 11] = clock64();
        // separate H buffers: MMA2 of chunk g - 2 has consumed this one.  ALIAS_H: H_g goes into the 16 packed columns at
        // the start of this warp's own 32 score columns (it has them in registers; MMA1 rewrites the buffer only after
        // MMA2_g has retired, which s_full of this chunk already implies for chunk g - NSB)
        if constexpr (!Cfg::ALIAS_H) mbar_wait(&h_empty[buf], u ^ 1);
        if (dbg) p.dbg[g * 24 + 12] = clock64();
        tc_fence_after();
        if constexpr (SC == 32) {
          const uint32_t(&h16)[16] = hp;
          tmem_st16(Cfg::ALIAS_H ? lane_tmem + Cfg::COL_S + buf * CH + cq * SC : lane_tmem + Cfg::COL_H + buf * (CH / 2) + cq * (SC / 2), h16);
        } else {
          const uint32_t(&h8)[8] = hp;
          tmem_st8(lane_tmem + Cfg::COL_H + buf * (CH / 2) + cq * (SC / 2), h8);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(h_full_r[buf]);
        if (p.h_out != nullptr) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmH, st_out, j * CH + cq * SC, tile * 2 * Cfg::BM + int(rank) * Cfg::BM + quarter * 32);
            tma_store_commit();
          }
        }
        if (dbg) p.dbg[g * 24 + 13] = clock64();
      }
      // ---------------- Z epilogue: + b2 + residual -> LayerNorm -> out ----------------
      // This warp owns a [32 rows x 32 columns] fp32 block of the tile (thread = row in tensor memory).  Row-per-thread
      // global accesses touch 32 different 128-byte lines per instruction and serialise in the LSU (measured: 12 000
      // cycles per tile for residual + out, a third of the kernel); instead global memory is accessed with 8 rows x 64
      // contiguous bytes per instruction and the block is turned between the two layouts through the warp's 2 KB staging
      // tile, 16 columns at a time.
      // d = 256: a warp owns 64 columns = two such blocks; the pre-norm sums go back to tensor memory (the Z columns are
      // free once read) instead of staying in 64 registers.
      constexpr int NB = ZC / 32;
      const long long row_base = (long long)tile * 2 * Cfg::BM + (long long)rank * Cfg::BM + quarter * 32;
      const int crow = lane >> 2, cu = lane & 3;                   // coalesced side: row 8 i + crow, 16-byte unit cu
      // coalesced-side addresses inside the staging tile (64-byte rows, 16-byte units XOR-swizzled)
      uint32_t c_off[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = 8 * i + crow;
        c_off[i] = uint32_t(rr * 64 + ((cu ^ ((rr >> 1) & 3)) << 4));
      }
      float4 rq[2][4];
      auto load_residual = [&](int zb) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const long long rg = row_base + 8 * i + crow;
            rq[h][i] = rg < p.T ? __ldg(reinterpret_cast<const float4*>(p.y + size_t(rg) * D + cq * ZC + 32 * zb + 16 * h + 4 * cu))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
          }
      };
      // a [32 x 32] block of this warp from registers (thread = row) to global memory: two [32 rows x 16 columns] boxes
      // through the staging tile, written by the copy engine (TMA store, SWIZZLE_64B).  Per-thread stores -- even the
      // coalesced 8 rows x 64 bytes per instruction of the first version -- held the warp for 2 300 cycles per 64 KB
      // tile and tensor (tools/bench_ffn.py --timeline); rows beyond T are clipped by the tensor map.
      auto store_block = [&](const CUtensorMap* tm, const float (&val)[32], int zb) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (lane == 0) tma_store_wait_read<0>();               // the previous box has left the staging tile
          __syncwarp();
#pragma unroll
          for (int u4 = 0; u4 < 4; ++u4)
            *reinterpret_cast<float4*>(st_out + lane * 64 + ((u4 ^ sw) << 4)) =
                make_float4(val[16 * h + 4 * u4], val[16 * h + 4 * u4 + 1], val[16 * h + 4 * u4 + 2], val[16 * h + 4 * u4 + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(tm, st_out, cq * ZC + 32 * zb + 16 * h, int(row_base));
            tma_store_commit();
          }
        }
      };
      load_residual(0);                                            // issued before waiting for the last MMA2
      const bool zdbg = p.dbg != nullptr && pair == 0 && leader && ew == 0 && lane == 0 && g - 1 < 64;
      mbar_wait(z_full, tt & 1);
      tc_fence_after();
      if (zdbg) p.dbg[(g - 1) * 24 + 3] = clock64();
      float uv[32];
      float s = 0.f;
#pragma unroll
      for (int zb = 0; zb < NB; ++zb) {
        if (zb > 0) load_residual(zb);
        tmem_ld32(lane_tmem + Cfg::COL_Z + cq * ZC + 32 * zb, uv);
        if (zb == 0) {
          if constexpr (NB == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(z_empty_r);         // Z is in registers: the next tile may accumulate
          }
        }
        if (lane == 0) tma_store_wait_read<0>();                   // the last store (hidden / previous block) has left the staging tile
        __syncwarp();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(st_out + c_off[i]) = rq[h][i];
          __syncwarp();
#pragma unroll
          for (int u4 = 0; u4 < 4; ++u4) {
            const float4 t4 = *reinterpret_cast<const float4*>(st_out + lane * 64 + ((u4 ^ sw) << 4));
            const float* b2v = sVec + cq * ZC + 32 * zb + 16 * h + 4 * u4;
            float* dstv = uv + 16 * h + 4 * u4;
            dstv[0] += b2v[0] + t4.x; dstv[1] += b2v[1] + t4.y; dstv[2] += b2v[2] + t4.z; dstv[3] += b2v[3] + t4.w;
            s += (dstv[0] + dstv[1]) + (dstv[2] + dstv[3]);
          }
          __syncwarp();
        }
        if (zdbg) p.dbg[(g - 1) * 24 + 7] = clock64();
        if (p.u2 != nullptr) store_block(&tmU2, uv, zb);
        if constexpr (NB > 1) tmem_st32f(lane_tmem + Cfg::COL_Z + cq * ZC + 32 * zb, uv);
      }
      if (zdbg) p.dbg[(g - 1) * 24 + 14] = clock64();
      if constexpr (NB > 1) tmem_st_wait();
      // row statistics across the four column quarters (warps quarter, quarter + 4, ...): two-pass like torch
      red0[cq * 32 + lane] = s;
      asm volatile("bar.sync %0, 128;" ::"r"(1 + quarter) : "memory");
      const float mu = (red0[lane] + red0[32 + lane] + red0[64 + lane] + red0[96 + lane]) * (1.f / D);
      float q = 0.f;
#pragma unroll
      for (int zb = 0; zb < NB; ++zb) {
        if constexpr (NB > 1) tmem_ld32(lane_tmem + Cfg::COL_Z + cq * ZC + 32 * zb, uv);
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float dlt = uv[c] - mu;
          q += dlt * dlt;
        }
      }
      red1[cq * 32 + lane] = q;
      asm volatile("bar.sync %0, 128;" ::"r"(1 + quarter) : "memory");
      const float rstd = rsqrtf((red1[lane] + red1[32 + lane] + red1[64 + lane] + red1[96 + lane]) * (1.f / D) + p.eps);
      if (zdbg) p.dbg[(g - 1) * 24 + 15] = clock64();
#pragma unroll
      for (int zb = 0; zb < NB; ++zb) {
        if constexpr (NB > 1) tmem_ld32(lane_tmem + Cfg::COL_Z + cq * ZC + 32 * zb, uv);
        const float* gm = sVec + D + cq * ZC + 32 * zb;
        const float* bt = sVec + 2 * D + cq * ZC + 32 * zb;
#pragma unroll
        for (int c = 0; c < 32; ++c) uv[c] = (uv[c] - mu) * rstd * gm[c] + bt[c];
        if constexpr (NB > 1) {
          if (zb == NB - 1) {                                      // the scratch use of the Z columns is over
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(z_empty_r);
          }
        }
        store_block(&tmOut, uv, zb);
      }
      if (p.stats != nullptr && cq == 0 && live) {
        p.stats[2 * size_t(row)] = mu;
        p.stats[2 * size_t(row) + 1] = rstd;
      }
    }
  }
  if (warp >= 2 && warp < 2 + Cfg::EPI_WARPS && lane == 0) tma_store_wait<0>();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // both CTAs are done with tensor memory and with each other's barriers
  if (warp == 0) tmem_dealloc_pair<512>(tmem_base);
}

}  // namespace rlt
