// Cross-list attention FORWARD on the 5th-generation tensor core (tcgen05.mma, accumulators in tensor memory).
//
// Reference semantics: nn.TransformerEncoderLayer without batch_first attends over dim 0 --
// for every position l and head h the S lists of a group attend to each other: S x S x dh problems with S = 64 (the
// reference's batch), dh = 16 (Choopy family, 8 heads).  Far below a 128-row MMA tile on their own, so one work item is
// TWO positions of one (group, head): 128 query rows = 2 positions x 64 lists, and the 128 x 128 score tile is
// block-diagonal (a query only sees the 64 keys of its own position).  Half of the score tile is wasted arithmetic
// (attention is ~3 % of the model's flops); what matters is that the tensor core replaces the warp-level fragment
// traffic that bound the mma.sync kernels (ncu: shared-memory pipe 66 %, MMA 9 % of the instructions).
//
//   TMA      q, k, v rows of the item (fp32, 64 bytes per token and operand) -> raw staging tiles, 6 box loads
//   convert  (4 warps, thread = token): fp16 operands in the canonical K-major SWIZZLE_128B layout:
//              QQ row = [q_hi | q_lo | q_hi], KK row = [k_hi | k_hi | k_lo]  (hi = fp16(x), lo = fp16(x - hi): 22 bits)
//              VT = v^T, [dh rows x 128 keys]
//   MMA      S = QQ . KK^T as three K = 16 steps = q_hi k_hi + q_lo k_hi + q_hi k_lo  (scores at ~fp32 accuracy, like the
//            3xTF32 products of the mma.sync kernels)                                   -> tensor memory, 128 columns
//   softmax  (4 warps, thread = query row): its 64 valid columns -> max, exp2, sum -> P as fp16 pairs written back to
//            TENSOR MEMORY (the invalid half of the row as zeros)
//   MMA      O' = P' . [V_pos0 | V_pos1] with P' as the TMEM A operand: a row keeps only the 64 probabilities of its own
//            position (K = 64: four K = 16 steps instead of eight over a zero-padded 128-key row) against BOTH positions'
//            values side by side (N = 2 dh); rows of position p read columns [p dh, p dh + dh)   -> tensor memory, 2 dh columns
//   epilogue the softmax warps scale O by 1 / sum and store it through a staging transposition (8 rows x 64 B per store
//            instruction); the log-sum-exp of the pair's heads goes to lse[token, :] for the backward.
// Warps: 0 TMA, 1 MMA issuer (scores), 2-5 convert, 6-9 softmax, 10-13 epilogue (softmax statistics handed over in shared
// memory), 14 MMA issuer (P . V).
// Two items are in flight (double-buffered staging / operand / TMEM slots); the MMA thread issues the score product of
// item i + 1 before the P . V of item i.
// Handles S <= 64 (keys >= S masked, rows >= S not stored), odd L (the second position of the last pair is empty).
#pragma once
#include <cuda_fp16.h>

#include "sm100.cuh"

namespace rlt {

template <int DH>
struct AttnTcCfg {
  static_assert(DH == 16, "attention_tc: head dim 16");
  static constexpr int ROWS = 128;                                   // 2 positions x 64 lists
  static constexpr int RAW_BYTES = 3 * ROWS * DH * 4;                // q | k | v fp32 rows
  static constexpr int QQ_BYTES = ROWS * 128;                        // [128 rows x 64 fp16]  SWIZZLE_128B K-major
  static constexpr int VT_BYTES = 2 * DH * 128;                      // two k-blocks of [DH rows x 64 keys]
  static constexpr int OP_BYTES = 2 * QQ_BYTES + VT_BYTES;
  static constexpr int OUT_BYTES = ROWS * DH * 4;                    // 4 x 2 KB store-staging tiles | [128 rows][8] log-sum-exp
  static constexpr int NR = 4;                                       // raw staging ring: the copy engine runs up to 4 items ahead
  static constexpr int NS = 3;                                       // operand / tensor-memory stages (items in flight)
  static constexpr int OFF_RAW = 0;
  static constexpr int OFF_OP = OFF_RAW + NR * RAW_BYTES;            // (1024-aligned: RAW_BYTES = 24576)
  static constexpr int OFF_OUT = OFF_OP + NS * ((OP_BYTES + 1023) / 1024 * 1024);
  static constexpr int OFF_STAT = OFF_OUT + 2 * OUT_BYTES;           // [NS stages][128 rows] (max, sum) of the softmax
  static constexpr int OFF_BARS = OFF_STAT + NS * ROWS * 8;
  static constexpr int N_BARS = 7 * NS + 2 * NR;

  static constexpr size_t SMEM_BYTES = 1024 + OFF_BARS + N_BARS * 8 + 16;
  static constexpr int THREADS = 32 * 15;                            // TMA | MMA(S) | 4 convert | 4 softmax | 4 epilogue | MMA(PV)
  // tensor memory columns: NS x 128 for S, with P (64 packed columns) written IN PLACE over the first half of the stage's
  // score columns (a lane only ever touches its own row: it has its scores in registers before it writes P) | NS x 2 DH
  // for O' (both positions' value columns; a row uses its own position's half)
  static constexpr uint32_t COL_S = 0, COL_P = 0, COL_O = NS * 128;
  static_assert(NS * 128 + NS * 2 * DH <= 512, "tensor memory budget");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

__device__ __forceinline__ float ex2_ftz(float x) {       // 2^x on the SFU (arguments <= 0 here)
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// hi = fp16(x), lo = fp16(x - hi) for two values, packed (x0 in the low half)
__device__ __forceinline__ void split_f16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

template <int DH>
__global__ void __launch_bounds__(AttnTcCfg<DH>::THREADS, 1)
attn_lists_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO,
                         float* __restrict__ o, float* __restrict__ lse, int G, int S, int L, int d, int n_head,
                         float scale_log2e, long long* __restrict__ dbg) {
  using Cfg = AttnTcCfg<DH>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sRaw = smem + Cfg::OFF_RAW;
  uint8_t* sOp = smem + Cfg::OFF_OP;
  uint8_t* sOut = smem + Cfg::OFF_OUT;
  constexpr int OP_STRIDE = (Cfg::OP_BYTES + 1023) / 1024 * 1024;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BARS);
  constexpr int NR = Cfg::NR;
  constexpr int NS = Cfg::NS;
  uint64_t* op_full = bars;                 // [NS] operand tiles written (count 128)
  uint64_t* op_empty = op_full + NS;        // [NS] the P . V MMAs of the item retired (commit)
  uint64_t* s_full = op_empty + NS;         // [NS] scores in tensor memory (commit)
  uint64_t* p_full = s_full + NS;           // [NS] P written to tensor memory (count 128)
  uint64_t* o_full = p_full + NS;           // [NS] O in tensor memory (commit)
  uint64_t* t_empty = o_full + NS;          // [NS] S / P / O slots of the stage read by the epilogue (count 128)
  uint64_t* stat_full = t_empty + NS;       // [NS] softmax statistics of the item in shared memory (count 128)
  uint64_t* raw_full = stat_full + NS;      // [NR] TMA landed
  uint64_t* raw_empty = raw_full + NR;  // [NR] convert warps have consumed the raw tile (count 128)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + Cfg::N_BARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_pairs = (L + 1) >> 1;
  const long long n_super = (long long)G * n_pairs;          // (group, position pair); every CTA walks all heads of its pairs
  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQKV);
      tma_prefetch_desc(&tmO);
      for (int s = 0; s < NR; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], 128); }
      for (int s = 0; s < NS; ++s) {
        mbar_init(&op_full[s], 128); mbar_init(&op_empty[s], 1);
        mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 128);
        mbar_init(&o_full[s], 1); mbar_init(&t_empty[s], 128);
        mbar_init(&stat_full[s], 128);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      uint32_t it = 0;
      for (long long sp = blockIdx.x; sp < n_super; sp += gridDim.x) {
        const int g = int(sp / n_pairs), l0 = int(sp % n_pairs) * 2;
        for (int h = 0; h < n_head; ++h, ++it) {
          const uint32_t st = it % NR, ph = (it / NR) & 1;
          mbar_wait(&raw_empty[st], ph ^ 1);
          mbar_expect_tx(&raw_full[st], Cfg::RAW_BYTES);
          uint8_t* dst = sRaw + st * Cfg::RAW_BYTES;
          for (int part = 0; part < 3; ++part)          // q, k, v columns of head h
            for (int pos = 0; pos < 2; ++pos)           // rows [64 pos, 64 pos + 64) = lists of the group at position l0 + pos
              tma_load_3d(dst + part * (Cfg::ROWS * DH * 4) + pos * (64 * DH * 4), &tmQKV, &raw_full[st], part * d + h * DH,
                          l0 + pos, g * S);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------ MMA issuer 1: S = QQ . KK^T ------------------------------
    // Two issuing threads (this one and warp 14 for P . V): a tcgen05.mma holds its issuing thread ~70 - 100 cycles and
    // every barrier poll ~150; one thread doing 11 MMAs, 3 commits and 4 waits per item was the bottleneck (2 200 cycles).
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc(kFmtF16, 128, 128, false, false);
      const uint32_t op_addr = smem_u32(sOp);
      long long my_super = 0;
      for (long long sp = blockIdx.x; sp < n_super; sp += gridDim.x) ++my_super;
      const uint32_t total = uint32_t(my_super) * uint32_t(n_head);
      uint32_t st = 0, ph = 0;
      for (uint32_t i = 0; i < total; ++i) {
        const bool d_ = dbg != nullptr && blockIdx.x == 0 && i < 48;
        if (d_) dbg[i * 16 + 0] = clock64();
        mbar_wait(&op_full[st], ph);
        mbar_wait(&t_empty[st], ph ^ 1);
        if (d_) dbg[i * 16 + 1] = clock64();
        tc_fence_after();
        const uint64_t da = make_smem_desc_sw128(op_addr + st * OP_STRIDE, 16, 1024);
        const uint64_t db = make_smem_desc_sw128(op_addr + st * OP_STRIDE + Cfg::QQ_BYTES, 16, 1024);
#pragma unroll
        for (int t = 0; t < 3; ++t)
          umma_f16(tmem_base + Cfg::COL_S + st * 128, da + uint64_t(2 * t), db + uint64_t(2 * t), idesc_s, t != 0 ? 1u : 0u);
        umma_commit(&s_full[st]);
        if (d_) dbg[i * 16 + 2] = clock64();
        if (++st == NS) { st = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 14) {
    // ------------------------------ MMA issuer 2: O = P . V ------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc_o = make_idesc(kFmtF16, 128, 2 * DH, false, false);
      const uint32_t op_addr = smem_u32(sOp);
      long long my_super = 0;
      for (long long sp = blockIdx.x; sp < n_super; sp += gridDim.x) ++my_super;
      const uint32_t total = uint32_t(my_super) * uint32_t(n_head);
      uint32_t st = 0, ph = 0;
      for (uint32_t i = 0; i < total; ++i) {
        const bool d_ = dbg != nullptr && blockIdx.x == 0 && i < 48;
        if (d_) dbg[i * 16 + 3] = clock64();
        mbar_wait(&p_full[st], ph);
        if (d_) dbg[i * 16 + 4] = clock64();
        tc_fence_after();
        // B = the two positions' V^T blocks as ONE [2 DH rows x 64 keys] K-major tile (they are adjacent 8-row-swizzled
        // blocks of the stage); A = P' (32 packed columns); a tcgen05.mma costs its issuer the same ~100+ cycles at N = 16
        // and N = 32, so halving the number of K steps halves this thread's time per item
        const uint64_t db = make_smem_desc_sw128(op_addr + st * OP_STRIDE + 2 * Cfg::QQ_BYTES, 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ts(tmem_base + Cfg::COL_O + st * (2 * DH), tmem_base + Cfg::COL_P + st * 128 + uint32_t(k * 8),
                      db + uint64_t(2 * k), idesc_o, k != 0 ? 1u : 0u);
        umma_commit(&o_full[st]);
        umma_commit(&op_empty[st]);
        if (d_) dbg[i * 16 + 5] = clock64();
        if (++st == NS) { st = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp < 6) {
    // ------------------------------ convert warps: thread = token row ------------------------------
    const int r = (warp - 2) * 32 + lane;                 // row of the item: position r / 64, list r % 64
    uint32_t it = 0;
    for (long long sp = blockIdx.x; sp < n_super; sp += gridDim.x) {
      for (int h = 0; h < n_head; ++h, ++it) {
        const uint32_t st = it % NS, ph = (it / NS) & 1;
        const uint32_t rs = it % NR, rph = (it / NR) & 1;
        const bool d_ = dbg != nullptr && blockIdx.x == 0 && it < 48 && warp == 2 && lane == 0;
        if (d_) dbg[it * 16 + 6] = clock64();
        mbar_wait(&raw_full[rs], rph);
        if (d_) dbg[it * 16 + 7] = clock64();
        const float4* raw = reinterpret_cast<const float4*>(sRaw + rs * Cfg::RAW_BYTES);
        float4 q4[DH / 4], k4[DH / 4], v4[DH / 4];
#pragma unroll
        for (int c = 0; c < DH / 4; ++c) {
          q4[c] = raw[r * (DH / 4) + c];
          k4[c] = raw[Cfg::ROWS * (DH / 4) + r * (DH / 4) + c];
          v4[c] = raw[2 * Cfg::ROWS * (DH / 4) + r * (DH / 4) + c];
        }
        mbar_wait(&op_empty[st], ph ^ 1);                  // the MMAs of the previous item on this stage are done with it
        if (d_) dbg[it * 16 + 8] = clock64();
        uint8_t* qq = sOp + st * OP_STRIDE;
        uint8_t* kk = qq + Cfg::QQ_BYTES;
        uint8_t* vt = kk + Cfg::QQ_BYTES;
        uint32_t qh[DH / 2], ql[DH / 2], kh[DH / 2], kl[DH / 2];
#pragma unroll
        for (int c = 0; c < DH / 4; ++c) {
          split_f16x2(q4[c].x, q4[c].y, qh[2 * c], ql[2 * c]);
          split_f16x2(q4[c].z, q4[c].w, qh[2 * c + 1], ql[2 * c + 1]);
          split_f16x2(k4[c].x, k4[c].y, kh[2 * c], kl[2 * c]);
          split_f16x2(k4[c].z, k4[c].w, kh[2 * c + 1], kl[2 * c + 1]);
        }
        // rows of 128 bytes = eight 16-byte chunks: [hi hi | lo lo | hi hi | - -] (Q) and [hi hi | hi hi | lo lo | - -] (K)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint4 qh4 = make_uint4(qh[4 * c], qh[4 * c + 1], qh[4 * c + 2], qh[4 * c + 3]);
          const uint4 ql4 = make_uint4(ql[4 * c], ql[4 * c + 1], ql[4 * c + 2], ql[4 * c + 3]);
          const uint4 kh4 = make_uint4(kh[4 * c], kh[4 * c + 1], kh[4 * c + 2], kh[4 * c + 3]);
          const uint4 kl4 = make_uint4(kl[4 * c], kl[4 * c + 1], kl[4 * c + 2], kl[4 * c + 3]);
          *reinterpret_cast<uint4*>(qq + sw128_offset(r, c)) = qh4;
          *reinterpret_cast<uint4*>(qq + sw128_offset(r, 2 + c)) = ql4;
          *reinterpret_cast<uint4*>(qq + sw128_offset(r, 4 + c)) = qh4;
          *reinterpret_cast<uint4*>(kk + sw128_offset(r, c)) = kh4;
          *reinterpret_cast<uint4*>(kk + sw128_offset(r, 2 + c)) = kh4;
          *reinterpret_cast<uint4*>(kk + sw128_offset(r, 4 + c)) = kl4;
        }
        // v^T: element (key r, column c) -> k-block r / 64, row c, key r % 64
        {
          uint8_t* vb = vt + (r >> 6) * (DH * 128);
          const int key = r & 63;
          const float vv[DH] = {v4[0].x, v4[0].y, v4[0].z, v4[0].w, v4[1].x, v4[1].y, v4[1].z, v4[1].w,
                                v4[2].x, v4[2].y, v4[2].z, v4[2].w, v4[3].x, v4[3].y, v4[3].z, v4[3].w};
#pragma unroll
          for (int c = 0; c < DH; ++c)
            *reinterpret_cast<__half*>(vb + sw128_offset(c, key >> 3) + (key & 7) * 2) = __float2half_rn(vv[c]);
        }
        fence_proxy_async_smem();
        mbar_arrive(&op_full[st]);
        // The raw tile is released only HERE, after every value read from it has been consumed: an arrive placed right
        // after the loads let the copy engine overwrite the tile while the LDS were still in flight (measured: garbage in
        // 5 of 8 heads; the arrive does not wait for the thread's outstanding shared-memory loads).
        mbar_arrive(&raw_empty[rs]);
        if (d_) dbg[it * 16 + 9] = clock64();
      }
    }
  } else if (warp < 10) {
    // ------------------------------ softmax warps: thread = query row ------------------------------
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int pos = r >> 6;
    const uint32_t lane_tmem = tmem_base + (uint32_t(quarter * 32) << 16);
    float2* s_stat = reinterpret_cast<float2*>(smem + Cfg::OFF_STAT);
    uint32_t it = 0;
    for (long long sp = blockIdx.x; sp < n_super; sp += gridDim.x) {
      for (int h = 0; h < n_head; ++h, ++it) {
        const uint32_t st = it % NS, ph = (it / NS) & 1;
        const bool d_ = dbg != nullptr && blockIdx.x == 0 && it < 48 && warp == 6 && lane == 0;
        if (d_) dbg[it * 16 + 10] = clock64();
        mbar_wait(&s_full[st], ph);
        if (d_) dbg[it * 16 + 11] = clock64();
        tc_fence_after();
        // this row's 64 valid score columns (the keys of its own position)
        float sc[64];
        {
          float t32[32];
          tmem_ld32(lane_tmem + Cfg::COL_S + st * 128 + pos * 64, t32);
#pragma unroll
          for (int c = 0; c < 32; ++c) sc[c] = t32[c];
          tmem_ld32(lane_tmem + Cfg::COL_S + st * 128 + pos * 64 + 32, t32);
#pragma unroll
          for (int c = 0; c < 32; ++c) sc[32 + c] = t32[c];
        }
        // max -> one FFMA + one ex2 per key -> fp16 pairs; four partial sums (fp32, unrounded: the denominators differ from
        // the rounded numerators by an unbiased 2^-11 / sqrt(64))
        float m0 = -INFINITY, m1 = -INFINITY;
        if (S < 64) {
#pragma unroll
          for (int c = 0; c < 64; ++c) sc[c] = c < S ? sc[c] : -INFINITY;
        }
#pragma unroll
        for (int c = 0; c < 64; c += 2) { m0 = fmaxf(m0, sc[c]); m1 = fmaxf(m1, sc[c + 1]); }
        const float m = fmaxf(m0, m1) * scale_log2e;               // scale_log2e > 0: max commutes with the scaling
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        uint32_t pk[32];
#pragma unroll
        for (int c = 0; c < 64; c += 4) {
          const float p0 = ex2_ftz(fmaf(sc[c], scale_log2e, -m)), p1 = ex2_ftz(fmaf(sc[c + 1], scale_log2e, -m));
          const float p2 = ex2_ftz(fmaf(sc[c + 2], scale_log2e, -m)), p3 = ex2_ftz(fmaf(sc[c + 3], scale_log2e, -m));
          a0 += p0; a1 += p1; a2 += p2; a3 += p3;
          const __half2 h01 = __floats2half2_rn(p0, p1), h23 = __floats2half2_rn(p2, p3);
          pk[c >> 1] = *reinterpret_cast<const uint32_t*>(&h01);
          pk[(c >> 1) + 1] = *reinterpret_cast<const uint32_t*>(&h23);
        }
        const float sum = (a0 + a1) + (a2 + a3);
        // (the statistics slot of this stage was read by the epilogue of item it - 2 before it released t_empty, which
        // the score product of this item waited for)
        s_stat[st * Cfg::ROWS + r] = make_float2(m, sum);
        mbar_arrive(&stat_full[st]);
        // P' row: the 64 keys of this row's own position = 32 packed columns at the start of the stage's score columns
        {
          const uint32_t pbase = lane_tmem + Cfg::COL_P + st * 128;
          uint32_t a16[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) a16[c] = pk[c];
          tmem_st16(pbase, a16);
#pragma unroll
          for (int c = 0; c < 16; ++c) a16[c] = pk[16 + c];
          tmem_st16(pbase + 16, a16);
          tmem_st_wait();
        }
        tc_fence_before();
        mbar_arrive(&p_full[st]);
        if (d_) dbg[it * 16 + 12] = clock64();
      }
    }
  } else if (warp < 14) {
    // ------------------------------ epilogue warps: thread = query row ------------------------------
    // O leaves through a per-warp [32 rows x 64 B] staging tile: written row-per-thread, read back 8 rows x 64 B per
    // instruction, so a global store touches 8 lines instead of 32 (rows are L d 4 bytes apart in HBM).  The log-sum-exp
    // values of the pair's heads are collected in shared memory and written once per (row, 8 heads).
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int pos = r >> 6, list = r & 63;
    const uint32_t lane_tmem = tmem_base + (uint32_t(quarter * 32) << 16);
    constexpr float kLn2 = 0.6931471805599453f;
    const float2* s_stat = reinterpret_cast<const float2*>(smem + Cfg::OFF_STAT);
    uint8_t* stg = sOut + (warp - 10) * 2048;
    float* s_lse = reinterpret_cast<float*>(sOut + 4 * 2048) + r * 8;      // [128 rows][8 heads]
    const int crow = lane >> 2, cu = lane & 3;
    uint32_t it = 0;
    for (long long sp = blockIdx.x; sp < n_super; sp += gridDim.x) {
      const int g = int(sp / n_pairs), l0 = int(sp % n_pairs) * 2;
      const bool row_live = list < S && l0 + pos < L;
      const size_t tok = (size_t(g) * S + list) * L + l0 + pos;
      for (int h = 0; h < n_head; ++h, ++it) {
        const uint32_t st = it % NS, ph = (it / NS) & 1;
        const bool d_ = dbg != nullptr && blockIdx.x == 0 && it < 48 && warp == 10 && lane == 0;
        mbar_wait(&stat_full[st], ph);
        const float2 ms = s_stat[st * Cfg::ROWS + r];
        if (d_) dbg[it * 16 + 13] = clock64();
        mbar_wait(&o_full[st], ph);
        tc_fence_after();
        float ov[DH];
        tmem_ld16(lane_tmem + Cfg::COL_O + st * (2 * DH) + pos * DH, ov);     // this row's position: its half of O'
        tc_fence_before();
        mbar_arrive(&t_empty[st]);                                   // S / P / O (and the statistics slot) of this stage are free
        if (d_) dbg[it * 16 + 15] = clock64();
        const float inv = __fdividef(1.f, ms.y);
#pragma unroll
        for (int c = 0; c < DH / 4; ++c)
          *reinterpret_cast<float4*>(stg + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)) =
              make_float4(ov[4 * c] * inv, ov[4 * c + 1] * inv, ov[4 * c + 2] * inv, ov[4 * c + 3] * inv);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = 8 * i + crow;                               // row inside this warp's 32
          const int row = quarter * 32 + rr, ps = row >> 6, ls = row & 63;
          const float4 t4 = *reinterpret_cast<const float4*>(stg + rr * 64 + ((cu ^ ((rr >> 1) & 3)) << 4));
          if (ls < S && l0 + ps < L)
            *reinterpret_cast<float4*>(o + ((size_t(g) * S + ls) * L + l0 + ps) * d + h * DH + 4 * cu) = t4;
        }
        __syncwarp();
        if (d_) dbg[it * 16 + 6] = clock64();
        if (lse != nullptr) {
          const float v = (ms.x + __log2f(ms.y)) * kLn2;
          if (n_head <= 8) {
            s_lse[h] = v;
            if (h == n_head - 1 && row_live) {
              float4* dst = reinterpret_cast<float4*>(lse + tok * n_head);
              for (int c = 0; c < n_head / 4; ++c) dst[c] = reinterpret_cast<const float4*>(s_lse)[c];
            }
          } else if (row_live) {
            lse[tok * n_head + h] = v;
          }
        }
        if (d_) dbg[it * 16 + 14] = clock64();
      }
    }
  }
  (void)sOut;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}

}  // namespace rlt
