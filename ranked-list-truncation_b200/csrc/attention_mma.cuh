// Cross-list attention on the warp-level tensor-core path (mma.sync.m16n8k8 TF32, fp32 accumulate).
//
// One work item per (group g, position l, head h); the problem is S x S x dh with S = lists per group (<= 128) and
// dh in {16, 32, 64}: far below a 128-row tcgen05 tile, so each warp owns a 16-row tile and keeps everything in
// registers.  Score products use the 3xTF32 split (a_hi b_hi + a_lo b_hi + a_hi b_lo) so that the
// softmax sees ~fp32-accurate logits; P V and the gradient contractions use single TF32.
// The contraction index of the second GEMM is permuted so that accumulator fragments (cols 2t, 2t+1) feed the
// next MMA's A fragment (cols t, t+4) without any shuffle: key 8j+2t -> k-slot t, key 8j+2t+1 -> k-slot t+4, and
// the B fragment rows are read from shared memory with the same permutation.
// Backward recomputes P from the saved log-sum-exp (no S x S tensor in HBM): phase A per query tile (dQ; P and dS stay
// in shared memory, column-swizzled), phase B per pair of key tiles contracts them transposed (dK, dV); no atomics.
#pragma once
#include "dropout.cuh"
#include "sm100.cuh"

namespace rlt {

// fp32 -> tf32 operand bits, round to nearest (ties away): the tensor core ignores the low 13 mantissa bits, so adding
// half a tf32 ulp to the bit pattern and letting the MMA truncate IS cvt.rna.tf32.f32 (except for inf / NaN inputs) -
// in ONE integer add.  (ncu: `cvt.rna.tf32.f32` is emulated on sm_100a with an FSETP / IMAD / LOP3 / SEL sequence; at
// ~500 conversions per warp and item it was more than half of the attention kernels' instructions.)
__device__ __forceinline__ uint32_t tf32_bits(float x) { return __float_as_uint(x) + 0x1000u; }
// the same value with the low bits cleared, for the hi / lo split of the 3xTF32 products (lo = x - hi must be exact)
__device__ __forceinline__ uint32_t tf32_hi(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
// 2^x on the SFU (ex2.approx.ftz: ~2 ulp; arguments here are <= 0, results in [0, 1]); exp2f() adds a denormal-range
// rescaling sequence (FSETP / FMUL) per call that the softmax does not need
__device__ __forceinline__ float ex2_fast(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// acc[j] (16 x 8 tiles, j < NT) += A[r0 : r0+16, 0:DH] * B[8j : 8j+8, 0:DH]^T ; rows of A and B live in shared memory
// with pitch DH + 4 floats (conflict-free fragment loads).
template <int DH, int NT, bool X3>
__device__ __forceinline__ void tile_abT(const float* __restrict__ sA, int r0, const float* __restrict__ sB,
                                         float (&acc)[NT][4], int lane) {
  constexpr int P = DH + 4;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int ks = 0; ks < DH / 8; ++ks) {
    const float fa0 = sA[(r0 + g) * P + ks * 8 + t], fa1 = sA[(r0 + g + 8) * P + ks * 8 + t];
    const float fa2 = sA[(r0 + g) * P + ks * 8 + t + 4], fa3 = sA[(r0 + g + 8) * P + ks * 8 + t + 4];
    const uint32_t a0 = X3 ? tf32_hi(fa0) : tf32_bits(fa0), a1 = X3 ? tf32_hi(fa1) : tf32_bits(fa1);
    const uint32_t a2 = X3 ? tf32_hi(fa2) : tf32_bits(fa2), a3 = X3 ? tf32_hi(fa3) : tf32_bits(fa3);
    uint32_t l0 = 0, l1 = 0, l2 = 0, l3 = 0;
    if (X3) {
      l0 = tf32_bits(fa0 - __uint_as_float(a0)); l1 = tf32_bits(fa1 - __uint_as_float(a1));
      l2 = tf32_bits(fa2 - __uint_as_float(a2)); l3 = tf32_bits(fa3 - __uint_as_float(a3));
    }
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const float fb0 = sB[(j * 8 + g) * P + ks * 8 + t], fb1 = sB[(j * 8 + g) * P + ks * 8 + t + 4];
      const uint32_t b0 = X3 ? tf32_hi(fb0) : tf32_bits(fb0), b1 = X3 ? tf32_hi(fb1) : tf32_bits(fb1);
      if (X3) {
        const uint32_t m0 = tf32_bits(fb0 - __uint_as_float(b0)), m1 = tf32_bits(fb1 - __uint_as_float(b1));
        mma_tf32(acc[j], l0, l1, l2, l3, b0, b1);   // small terms first
        mma_tf32(acc[j], a0, a1, a2, a3, m0, m1);
      }
      mma_tf32(acc[j], a0, a1, a2, a3, b0, b1);
    }
  }
}

// o[n] (16 x 8 tiles, n < DH/8) += Pfrag[16, 8*NT] * B[0 : 8*NT, 0:DH], Pfrag given as accumulator fragments.
template <int DH, int NT>
__device__ __forceinline__ void tile_pB(const float (&p)[NT][4], const float* __restrict__ sB, float (&o)[DH / 8][4],
                                        int lane) {
  constexpr int P = DH + 4;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const uint32_t a0 = tf32_bits(p[j][0]), a1 = tf32_bits(p[j][2]), a2 = tf32_bits(p[j][1]), a3 = tf32_bits(p[j][3]);
#pragma unroll
    for (int n = 0; n < DH / 8; ++n) {
      const uint32_t b0 = tf32_bits(sB[(j * 8 + 2 * t) * P + n * 8 + g]);
      const uint32_t b1 = tf32_bits(sB[(j * 8 + 2 * t + 1) * P + n * 8 + g]);
      mma_tf32(o[n], a0, a1, a2, a3, b0, b1);
    }
  }
}

// Column swizzle of the [ROWS][ROWS] probability / dS matrices of the backward kernel (no row padding): element
// (row, col) lives at row * ROWS + (col ^ (8 * pswz(row))).  pswz takes four distinct values on rows {0..3}, {4..7},
// {0,2,4,6} and {1,3,5,7} (mod 8), which makes BOTH access patterns conflict-free: the 64-bit accumulator-fragment
// stores of phase A (a half-warp = 4 rows x 8 words) and the transposed fragment loads of phase B (rows 2t / 2t+1 x 8
// columns).  With the padded pitch 68 the stores were 2-way conflicted: 128 of ~370 wavefronts per warp and item.
__device__ __forceinline__ int pswz(int row) { return (row & 3) ^ ((row >> 2) & 1); }

// Two 16-column tiles (c0, c0 + 16; c0 % 32 == 0) of the same transposed, swizzled operand against the same B rows: the
// B fragments are loaded once per k-step and feed both tiles (the kernel is bound by shared-memory wavefronts).  The
// contraction index is permuted (slot t <-> row 8ks+2t, slot t+4 <-> row 8ks+2t+1) so that the B loads are conflict-free.
template <int DH, int NT>
__device__ __forceinline__ void tile_tAB2(const float* __restrict__ sT, int c0, const float* __restrict__ sB,
                                          float (&acc0)[DH / 8][4], float (&acc1)[DH / 8][4], int lane) {
  constexpr int P = DH + 4;
  constexpr int PP = NT * 8;
  const int g = lane >> 2, t = lane & 3;
  const int fe = pswz(2 * t), fo = pswz(2 * t + 1);
  const float* pe = sT + (2 * t) * PP + c0 + g;
  const float* po = pe + PP;
  const int e0 = (0 ^ fe) * 8, e1 = (1 ^ fe) * 8, e2 = (2 ^ fe) * 8, e3 = (3 ^ fe) * 8;
  const int o0 = (0 ^ fo) * 8, o1 = (1 ^ fo) * 8, o2 = (2 ^ fo) * 8, o3 = (3 ^ fo) * 8;
#pragma unroll
  for (int ks = 0; ks < NT; ++ks) {
    uint32_t b0[DH / 8], b1[DH / 8];
#pragma unroll
    for (int n = 0; n < DH / 8; ++n) {
      b0[n] = tf32_bits(sB[(ks * 8 + 2 * t) * P + n * 8 + g]);
      b1[n] = tf32_bits(sB[(ks * 8 + 2 * t + 1) * P + n * 8 + g]);
    }
    const float* re = pe + ks * 8 * PP;
    const float* ro = po + ks * 8 * PP;
    {
      const uint32_t a0 = tf32_bits(re[e0]), a1 = tf32_bits(re[e1]), a2 = tf32_bits(ro[o0]), a3 = tf32_bits(ro[o1]);
#pragma unroll
      for (int n = 0; n < DH / 8; ++n) mma_tf32(acc0[n], a0, a1, a2, a3, b0[n], b1[n]);
    }
    {
      const uint32_t a0 = tf32_bits(re[e2]), a1 = tf32_bits(re[e3]), a2 = tf32_bits(ro[o2]), a3 = tf32_bits(ro[o3]);
#pragma unroll
      for (int n = 0; n < DH / 8; ++n) mma_tf32(acc1[n], a0, a1, a2, a3, b0[n], b1[n]);
    }
  }
}

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// ------------------------------------------------------------------------------------------------------------
// Persistent, software-pipelined kernels: a CTA walks work items (group, position, head) with the head index
// fastest (neighbouring heads share 128-byte lines of the qkv rows) and stages the next item's rows into the other
// half of a double buffer with cp.async while the tensor cores work on the current one.  Memory latency is hidden by
// the pipeline instead of by occupancy (the first, one-CTA-per-item kernels were register-limited to 2 CTAs per SM and
// spent 60% of their samples on the staging loads).  The softmax scale is applied to the accumulator (raw rows are copied).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Work item (group g, position l, head h), head fastest.  A CTA visits items blockIdx.x, + gridDim.x, ...: the
// (g, l, h) digits are advanced with carries instead of being re-derived by division (the 64-bit divisions and the
// 64-bit address products of every staged 16-byte chunk were ~25 % of the kernels' instructions).
struct AttnPos {
  int g, l, h;
};
struct AttnWalk {
  int dg, dl, dh, L, n_head;
  __device__ __forceinline__ AttnWalk(int step, int L_, int n_head_) : L(L_), n_head(n_head_) {
    dh = step % n_head_;
    const int r = step / n_head_;
    dl = r % L_;
    dg = r / L_;
  }
  __device__ __forceinline__ AttnPos first(int item) const {
    AttnPos p;
    p.h = item % n_head;
    const int r = item / n_head;
    p.l = r % L;
    p.g = r / L;
    return p;
  }
  __device__ __forceinline__ AttnPos next(AttnPos p) const {
    p.h += dh;
    int c = p.h >= n_head ? 1 : 0;
    p.h -= c ? n_head : 0;
    p.l += dl + c;
    c = p.l >= L ? 1 : 0;
    p.l -= c ? L : 0;
    p.g += dg + c;
    return p;
  }
};
// token index of list 0 of the item's group at the item's position
__device__ __forceinline__ size_t attn_tok0(const AttnPos& p, int S, int L) { return size_t(p.g) * S * L + p.l; }

// kDrop: train-mode dropout on the attention probabilities (a separate instantiation: the hash code costs registers)
// kFull: S == 8 NT (no padded rows / columns: the bounds predicates compile away)
template <int DH, int NT, bool kDrop, bool kFull>
__global__ void __launch_bounds__(128) attn_lists_fwd_pipe_kernel(const float* __restrict__ qkv, float* __restrict__ o,
                                                                  float* __restrict__ lse, int S, int L, int d, int n_head,
                                                                  float scale, int n_items, DropCfg drop) {
  constexpr int P = DH + 4;
  constexpr int ROWS = NT * 8;
  constexpr int BUF = 3 * ROWS * P;
  constexpr int CH = DH / 4, RS = 128 / CH;     // 16-byte chunks per row; rows covered by one pass of the 128 threads
  extern __shared__ float sm[];
  const int ld = 3 * d;
  for (int i = threadIdx.x; i < 2 * BUF; i += 128) sm[i] = 0.f;   // padded rows stay zero for the whole kernel
  __syncthreads();
  // staging: thread -> (row s0 + k RS, chunk c0) of every matrix; 32-bit element offsets from the item's base pointer
  const int s0 = threadIdx.x / CH, c0 = (threadIdx.x % CH) * 4;
  const uint32_t goff0 = uint32_t(s0) * uint32_t(L) * uint32_t(ld) + uint32_t(c0), gstep = uint32_t(RS) * uint32_t(L) * uint32_t(ld);
  const int soff0 = s0 * P + c0;
  const AttnWalk walk(int(gridDim.x), L, n_head);
  auto stage = [&](const AttnPos& it, int b) {
    float* base = sm + b * BUF + soff0;
    const float* src = qkv + attn_tok0(it, S, L) * ld + it.h * DH;
    uint32_t go = goff0;
    int so = 0;
    for (int s_ = s0; s_ < S; s_ += RS, go += gstep, so += RS * P) {
      cp_async16(base + so, src + go);
      cp_async16(base + ROWS * P + so, src + d + go);
      cp_async16(base + 2 * ROWS * P + so, src + 2 * d + go);
    }
    cp_async_commit();
  };
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gq = lane >> 2, t = lane & 3;
  const float sc = scale * kLog2e;
  int item = blockIdx.x;
  AttnPos cur = walk.first(item);
  if (item < n_items) stage(cur, 0);
  int b = 0;
  for (; item < n_items; item += gridDim.x, b ^= 1) {
    const AttnPos nxt = walk.next(cur);
    if (item + int(gridDim.x) < n_items) { stage(nxt, b ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const size_t tok0 = attn_tok0(cur, S, L);
    float* o_item = o + tok0 * d + cur.h * DH;
    float* lse_item = lse != nullptr ? lse + tok0 * n_head + cur.h : nullptr;
    const float* sQ = sm + b * BUF;
    const float* sK = sQ + ROWS * P;
    const float* sV = sK + ROWS * P;
    for (int r0 = warp * 16; r0 < S; r0 += 64) {
      float acc[NT][4];
#pragma unroll
      for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
      tile_abT<DH, NT, true>(sQ, r0, sK, acc, lane);
      float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const int c = j * 8 + 2 * t;
        const bool ok0 = kFull || c < S, ok1 = kFull || c + 1 < S;
        acc[j][0] = ok0 ? acc[j][0] * sc : -INFINITY; acc[j][2] = ok0 ? acc[j][2] * sc : -INFINITY;
        acc[j][1] = ok1 ? acc[j][1] * sc : -INFINITY; acc[j][3] = ok1 ? acc[j][3] * sc : -INFINITY;
        m0 = fmaxf(m0, fmaxf(acc[j][0], acc[j][1]));
        m1 = fmaxf(m1, fmaxf(acc[j][2], acc[j][3]));
      }
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
      m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
      float s0_ = 0.f, s1_ = 0.f;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        acc[j][0] = ex2_fast(acc[j][0] - m0); acc[j][1] = ex2_fast(acc[j][1] - m0);
        acc[j][2] = ex2_fast(acc[j][2] - m1); acc[j][3] = ex2_fast(acc[j][3] - m1);
        s0_ += acc[j][0] + acc[j][1];
        s1_ += acc[j][2] + acc[j][3];
      }
      s0_ += __shfl_xor_sync(0xffffffffu, s0_, 1); s0_ += __shfl_xor_sync(0xffffffffu, s0_, 2);
      s1_ += __shfl_xor_sync(0xffffffffu, s1_, 1); s1_ += __shfl_xor_sync(0xffffffffu, s1_, 2);
      const int ra = r0 + gq, rb = r0 + gq + 8;
      if (kDrop) {   // dropout on the attention probabilities (the row sums above stay undropped)
        const uint64_t ea = (uint64_t(item) * S + ra) * S, eb = (uint64_t(item) * S + rb) * S;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const uint64_t ba = drop_bits(drop.seed, DROP_ATTN, ea + j * 8 + 2 * t), bb = drop_bits(drop.seed, DROP_ATTN, eb + j * 8 + 2 * t);
          acc[j][0] *= drop_factor(ba, 0, drop.thr, drop.scale); acc[j][1] *= drop_factor(ba, 1, drop.thr, drop.scale);
          acc[j][2] *= drop_factor(bb, 0, drop.thr, drop.scale); acc[j][3] *= drop_factor(bb, 1, drop.thr, drop.scale);
        }
      }
      float oacc[DH / 8][4];
#pragma unroll
      for (int n = 0; n < DH / 8; ++n) oacc[n][0] = oacc[n][1] = oacc[n][2] = oacc[n][3] = 0.f;
      tile_pB<DH, NT>(acc, sV, oacc, lane);
      const float i0 = 1.f / s0_, i1 = 1.f / s1_;
      const uint32_t rowa = uint32_t(ra) * uint32_t(L), rowb = uint32_t(rb) * uint32_t(L);   // token offsets of the two rows
      if (kFull || ra < S) {
        float* op = o_item + rowa * uint32_t(d);
#pragma unroll
        for (int n = 0; n < DH / 8; ++n)
          *reinterpret_cast<float2*>(op + n * 8 + 2 * t) = make_float2(oacc[n][0] * i0, oacc[n][1] * i0);
        if (t == 0 && lse_item != nullptr) lse_item[rowa * uint32_t(n_head)] = (m0 + log2f(s0_)) * kLn2;
      }
      if (kFull || rb < S) {
        float* op = o_item + rowb * uint32_t(d);
#pragma unroll
        for (int n = 0; n < DH / 8; ++n)
          *reinterpret_cast<float2*>(op + n * 8 + 2 * t) = make_float2(oacc[n][2] * i1, oacc[n][3] * i1);
        if (t == 0 && lse_item != nullptr) lse_item[rowb * uint32_t(n_head)] = (m1 + log2f(s1_)) * kLn2;
      }
    }
    __syncthreads();   // every warp is done with buffer b before the next iteration refills it
    cur = nxt;
  }
}

// Backward.  D_i = sum_j P_ij dP_ij (= dO_i . O_i) is taken from the phase-A fragments, so the attention output is not
// read at all.  Phase A leaves P (dropped) and dS in shared memory; phase B contracts them TRANSPOSED with Q / dO for
// dK / dV instead of recomputing the scores a second time (the kernel is instruction-issue bound).
// kBufs: 2 = cp.async double buffer over items; 1 = single staging buffer (large head dims: two buffers would leave room
// for ONE CTA per SM; a second resident CTA hides the staging latency instead)
template <int DH, int NT, bool kDrop, int kBufs, bool kFull>
__global__ void __launch_bounds__(128) attn_lists_bwd_pipe_kernel(const float* __restrict__ qkv, const float* __restrict__ lse,
                                                                  const float* __restrict__ d_o, float* __restrict__ dqkv,
                                                                  int S, int L, int d, int n_head, float scale,
                                                                  int n_items, DropCfg drop) {
  constexpr int P = DH + 4;
  constexpr int ROWS = NT * 8;
  constexpr int BUF = 4 * ROWS * P + ROWS;      // q, k, v, dO rows + lse
  constexpr int PP = ROWS;                      // probability / dS matrices: unpadded rows, columns swizzled by pswz(row)
  extern __shared__ float sm[];
  float* sPm = sm + kBufs * BUF;                // [ROWS][PP] (dropped) probabilities of the current item
  float* sS = sPm + ROWS * PP;                  // [ROWS][PP] dS of the current item
  const int ld = 3 * d;
  for (int i = threadIdx.x; i < kBufs * BUF + 2 * ROWS * PP; i += 128) sm[i] = 0.f;   // padded rows stay zero
  __syncthreads();
  for (int b = 0; b < kBufs; ++b)
    for (int s_ = S + threadIdx.x; s_ < ROWS; s_ += 128) sm[b * BUF + 4 * ROWS * P + s_] = INFINITY;  // padded queries: P = 0
  __syncthreads();
  // staging: thread -> (row s0 + k RS, chunk c0) of every matrix; 32-bit element offsets from the item's base pointers
  constexpr int CH = DH / 4, RS = 128 / CH;
  const int s0 = threadIdx.x / CH, c0 = (threadIdx.x % CH) * 4;
  const uint32_t rowtok = uint32_t(s0) * uint32_t(L), rowstep = uint32_t(RS) * uint32_t(L);
  const int soff0 = s0 * P + c0;
  const AttnWalk walk(int(gridDim.x), L, n_head);
  auto stage = [&](const AttnPos& it, int b) {
    float* base = sm + b * BUF;
    const size_t tok0 = attn_tok0(it, S, L);
    const float* src = qkv + tok0 * ld + it.h * DH + c0;
    const float* srcg = d_o + tok0 * d + it.h * DH + c0;
    uint32_t rt = rowtok;
    int so = soff0;
    for (int s_ = s0; s_ < S; s_ += RS, rt += rowstep, so += RS * P) {
      const uint32_t go = rt * uint32_t(ld);
      cp_async16(base + so, src + go);
      cp_async16(base + ROWS * P + so, src + d + go);
      cp_async16(base + 2 * ROWS * P + so, src + 2 * d + go);
      cp_async16(base + 3 * ROWS * P + so, srcg + rt * uint32_t(d));
    }
    const float* srcl = lse + tok0 * n_head + it.h;
    for (int s_ = threadIdx.x; s_ < S; s_ += 128)
      cp_async4(base + 4 * ROWS * P + s_, srcl + uint32_t(s_) * uint32_t(L) * uint32_t(n_head));
    cp_async_commit();
  };
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gq = lane >> 2, t = lane & 3;
  const float sc = scale * kLog2e;
  int item = blockIdx.x;
  AttnPos cur = walk.first(item);
  if (kBufs == 2 && item < n_items) stage(cur, 0);
  int b = 0;
  for (; item < n_items; item += gridDim.x, b ^= (kBufs - 1)) {
    const AttnPos nxt = walk.next(cur);
    if (kBufs == 2) {
      if (item + int(gridDim.x) < n_items) { stage(nxt, b ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    } else {
      stage(cur, 0);
      cp_async_wait<0>();
    }
    __syncthreads();
    float* dq_item = dqkv + attn_tok0(cur, S, L) * ld + cur.h * DH;
    const float* sQ = sm + b * BUF;
    const float* sK = sQ + ROWS * P;
    const float* sV = sK + ROWS * P;
    const float* sG = sV + ROWS * P;
    const float* sL = sG + ROWS * P;            // natural-log lse
    // ---------------- phase A: query tiles -> P, dS (kept in shared memory for phase B), dQ
    for (int r0 = warp * 16; r0 < S; r0 += 64) {
      float p[NT][4], dp[NT][4];
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        p[j][0] = p[j][1] = p[j][2] = p[j][3] = 0.f;
        dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
      }
      tile_abT<DH, NT, true>(sQ, r0, sK, p, lane);
      tile_abT<DH, NT, false>(sG, r0, sV, dp, lane);
      const float la = sL[r0 + gq] * kLog2e, lb = sL[r0 + gq + 8] * kLog2e;
      float da = 0.f, db = 0.f;
      const int fsw = pswz(gq);                 // rows r0 + gq and r0 + gq + 8 share the swizzle (r0 % 16 == 0)
      float* pa_row = sPm + (r0 + gq) * PP + 2 * t;
      float* pb_row = pa_row + 8 * PP;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const int c = j * 8 + 2 * t;
        const bool ok0 = kFull || c < S, ok1 = kFull || c + 1 < S;
        p[j][0] = ok0 ? ex2_fast(fmaf(p[j][0], sc, -la)) : 0.f;
        p[j][1] = ok1 ? ex2_fast(fmaf(p[j][1], sc, -la)) : 0.f;
        p[j][2] = ok0 ? ex2_fast(fmaf(p[j][2], sc, -lb)) : 0.f;
        p[j][3] = ok1 ? ex2_fast(fmaf(p[j][3], sc, -lb)) : 0.f;
        float m0 = 1.f, m1 = 1.f, m2 = 1.f, m3 = 1.f;
        if (kDrop) {   // d(attn) = mask/(1-p) * d(dropped attn); dV sees the dropped probabilities
          const uint64_t ba = drop_bits(drop.seed, DROP_ATTN, (uint64_t(item) * S + r0 + gq) * S + c);
          const uint64_t bb = drop_bits(drop.seed, DROP_ATTN, (uint64_t(item) * S + r0 + gq + 8) * S + c);
          m0 = drop_factor(ba, 0, drop.thr, drop.scale); m1 = drop_factor(ba, 1, drop.thr, drop.scale);
          m2 = drop_factor(bb, 0, drop.thr, drop.scale); m3 = drop_factor(bb, 1, drop.thr, drop.scale);
          dp[j][0] *= m0; dp[j][1] *= m1; dp[j][2] *= m2; dp[j][3] *= m3;
        }
        *reinterpret_cast<float2*>(pa_row + ((j ^ fsw) << 3)) = make_float2(p[j][0] * m0, p[j][1] * m1);
        *reinterpret_cast<float2*>(pb_row + ((j ^ fsw) << 3)) = make_float2(p[j][2] * m2, p[j][3] * m3);
        da = fmaf(p[j][0], dp[j][0], fmaf(p[j][1], dp[j][1], da));
        db = fmaf(p[j][2], dp[j][2], fmaf(p[j][3], dp[j][3], db));
      }
      da += __shfl_xor_sync(0xffffffffu, da, 1); da += __shfl_xor_sync(0xffffffffu, da, 2);
      db += __shfl_xor_sync(0xffffffffu, db, 1); db += __shfl_xor_sync(0xffffffffu, db, 2);
      float* sa_row = sS + (r0 + gq) * PP + 2 * t;
      float* sb_row = sa_row + 8 * PP;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        p[j][0] *= dp[j][0] - da; p[j][1] *= dp[j][1] - da;
        p[j][2] *= dp[j][2] - db; p[j][3] *= dp[j][3] - db;
        *reinterpret_cast<float2*>(sa_row + ((j ^ fsw) << 3)) = make_float2(p[j][0], p[j][1]);
        *reinterpret_cast<float2*>(sb_row + ((j ^ fsw) << 3)) = make_float2(p[j][2], p[j][3]);
      }
      float acc[DH / 8][4];
#pragma unroll
      for (int n = 0; n < DH / 8; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
      tile_pB<DH, NT>(p, sK, acc, lane);
      const int ra = r0 + gq, rb = r0 + gq + 8;
      if (kFull || ra < S) {
        float* out = dq_item + uint32_t(ra) * uint32_t(L) * uint32_t(ld);
#pragma unroll
        for (int n = 0; n < DH / 8; ++n)
          *reinterpret_cast<float2*>(out + n * 8 + 2 * t) = make_float2(acc[n][0] * scale, acc[n][1] * scale);
      }
      if (kFull || rb < S) {
        float* out = dq_item + uint32_t(rb) * uint32_t(L) * uint32_t(ld);
#pragma unroll
        for (int n = 0; n < DH / 8; ++n)
          *reinterpret_cast<float2*>(out + n * 8 + 2 * t) = make_float2(acc[n][2] * scale, acc[n][3] * scale);
      }
    }
    __syncthreads();   // P and dS complete
    // ---------------- phase B: key tiles -> dK = dS^T Q (warps 0, 1), dV = (dropped P)^T dO (warps 2, 3); operands are
    // read transposed from smem, each warp takes pairs of 16-key tiles so that the Q / dO fragments are loaded once per pair
    {
      const bool is_v = warp >= 2;
      const float* sT = is_v ? sPm : sS;
      const float* sB = is_v ? sG : sQ;
      const float osc = is_v ? 1.f : scale;
      float* out_m = dq_item + (is_v ? 2 * d : d);
      for (int c0 = (warp & 1) * 32; c0 < S; c0 += 64) {
        float a0[DH / 8][4], a1[DH / 8][4];
#pragma unroll
        for (int n = 0; n < DH / 8; ++n) {
          a0[n][0] = a0[n][1] = a0[n][2] = a0[n][3] = 0.f;
          a1[n][0] = a1[n][1] = a1[n][2] = a1[n][3] = 0.f;
        }
        tile_tAB2<DH, NT>(sT, c0, sB, a0, a1, lane);
#pragma unroll
        for (int hrow = 0; hrow < 4; ++hrow) {            // key rows c0 + gq + {0, 8, 16, 24}
          const int kr = c0 + gq + 8 * hrow;
          if (kFull || kr < S) {
            float* outk = out_m + uint32_t(kr) * uint32_t(L) * uint32_t(ld);
#pragma unroll
            for (int n = 0; n < DH / 8; ++n) {
              const float v0 = hrow < 2 ? a0[n][2 * (hrow & 1)] : a1[n][2 * (hrow & 1)];
              const float v1 = hrow < 2 ? a0[n][2 * (hrow & 1) + 1] : a1[n][2 * (hrow & 1) + 1];
              *reinterpret_cast<float2*>(outk + n * 8 + 2 * t) = make_float2(v0 * osc, v1 * osc);
            }
          }
        }
      }
    }
    __syncthreads();   // every warp is done with buffer b (and sD) before the next iteration refills them
    cur = nxt;
  }
}

}  // namespace rlt
