// K1 — persistent BiLSTM recurrence on tcgen05 (one CTA per tile of 128 lists x direction).
//
// Forward, per time step:   a[128 lists, 512 gates] = P_t + h_{t-1} W_hh^T
//   * W_hh lives in shared memory for the whole scan as an fp16 K-major SWIZZLE_128B operand (128 KB); h_{t-1} is the
//     fp16 A operand (32 KB) that the gate warps rewrite every step.  fp16 keeps the 10-bit mantissa of TF32 and both
//     operands are bounded (|h| < 1, |W_hh| small), so the contraction has TF32-equivalent accuracy at twice the rate
//     and half the bytes; accumulation is fp32 in TMEM.
//   * gate rows are permuted so that each N = 256 MMA group yields i,f,g,o of 64 hidden units: half 0 -> TMEM columns
//     [0,256), half 1 -> [256,512).  Gate warps 1-4 consume half 0 while the tensor core computes half 1; warps 5-8
//     consume half 1.  One thread owns one list row of its half: c_t stays in registers (64 floats), the
//     nonlinearities are evaluated in registers (ex2 + rcp: sigmoid(x) = 1/(1+2^(-x log2 e)), tanh(x) = 2 sigmoid(2x) - 1).
//   * P_t (input projection + both biases, fp32, written by the tcgen05 GEMM) is added in the gate math.
// Backward (BPTT), per time step:  dh_rec[128, 128] = da[128, 512] W_hh, with da in fp16 scaled by a power of two
// chosen from max|dy| (gradients are ~1e-6; the scale keeps them in fp16's normal range; D is unscaled on read).
//   * W_hh^T is resident as the B operand (128 KB); da is produced in K-blocks of (16 units x 4 gates) = 64 columns,
//     each pushed through a 4-slot ring (mbarrier full/empty) and consumed by the MMA as soon as it is complete.
#pragma once
#include "sm100.cuh"

namespace rlt {

constexpr int LH = 128;          // hidden
constexpr int LG4 = 512;         // gate rows
constexpr int LSAVE = 6;         // saved planes

__device__ __forceinline__ float fast_sigmoid(float x) {
  // 1 / (1 + 2^(-x log2e)); ex2.approx + rcp.approx: ~1e-7 relative error
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return r;
}
__device__ __forceinline__ float fast_tanh(float x) { return 2.f * fast_sigmoid(2.f * x) - 1.f; }

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %2, %1;" : "=r"(r) : "f"(a), "f"(b));  // low half <- a
  return r;
}

struct LstmFwdSmem {
  static constexpr int B_BYTES = 2 * LG4 * 128;   // two k-blocks of [512 rows x 128 B]
  static constexpr int A_BYTES = 2 * 128 * 128;   // two k-blocks of [128 rows x 128 B]; TWO such buffers (ping-pong)
  static constexpr size_t TOTAL = 1024 + B_BYTES + 2 * A_BYTES + 256;
};

// B-operand row (0..511) -> W_hh gate row: half = n / 256, gate = (n % 256) / 64, unit = half*64 + n % 64
__device__ __forceinline__ int fwd_brow_to_wrow(int n) { return ((n & 255) >> 6) * LH + (n >> 8) * 64 + (n & 63); }

__global__ void __launch_bounds__(288, 1)
lstm_rec_fwd_tc_kernel(const float* __restrict__ P, const float* __restrict__ whh_f, const float* __restrict__ whh_r,
                       float* __restrict__ y, float* __restrict__ saved, int B, int L) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sB = smem;
  uint8_t* sA = smem + LstmFwdSmem::B_BYTES;
  // h ping-pong: step s reads buffer s&1 (h_{s-1}) and the gate warps write h_s into buffer (s+1)&1, because the
  // half-1 MMAs of step s are still reading h_{s-1} while the half-0 warps already produce h_s.
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + 2 * LstmFwdSmem::A_BYTES);
  uint64_t* bar_h = bars;          // h_{t} complete in smem (count 256)
  uint64_t* bar_acc = bars + 1;    // [2] accumulator half ready (tcgen05.commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, dir = blockIdx.y;
  const float* whh = dir ? whh_r : whh_f;

  // ---- one-time: stage W_hh (fp32 [512,128]) as the permuted fp16 K-major operand; zero h_0
  for (int i = threadIdx.x; i < LG4 * 16; i += blockDim.x) {
    const int n = i >> 4, c = i & 15;                 // c: 8-element chunk along k (0..15)
    const float* src = whh + size_t(fwd_brow_to_wrow(n)) * LH + c * 8;
    const float4 v0 = *reinterpret_cast<const float4*>(src), v1 = *reinterpret_cast<const float4*>(src + 4);
    uint4 pk = make_uint4(pack_half2(v0.x, v0.y), pack_half2(v0.z, v0.w), pack_half2(v1.x, v1.y), pack_half2(v1.z, v1.w));
    *reinterpret_cast<uint4*>(sB + (c >> 3) * (LG4 * 128) + sw128_offset(n, c & 7)) = pk;
  }
  for (int i = threadIdx.x; i < 2 * LstmFwdSmem::A_BYTES / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(sA)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == 0) {
    if (lane == 0) {
      mbar_init(bar_h, 256);
      mbar_init(&bar_acc[0], 1);
      mbar_init(&bar_acc[1], 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  fence_proxy_async_smem();   // generic-proxy writes of sA / sB -> visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kFmtF16, 128, 256, false, false);
      const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
      for (int step = 0; step < L; ++step) {
        if (step > 0) {
          mbar_wait(bar_h, (step - 1) & 1);
          tc_fence_after();
        }
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint64_t da = make_smem_desc_sw128(a_addr + (step & 1) * LstmFwdSmem::A_BYTES + kb * (128 * 128), 16, 1024);
            const uint64_t db = make_smem_desc_sw128(b_addr + kb * (LG4 * 128) + hf * (256 * 128), 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)   // K = 16 fp16 = 32 B per MMA
              umma_f16(tmem_base + hf * 256, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&bar_acc[hf]);
        }
      }
    } else {
      // lanes 1..31: pull the P rows (2 KB per list, contiguous) of the step that is kPrefetch ahead into L2, paced by
      // the accumulator barrier, so that the gate warps' dependent loads hit L2 instead of DRAM
      constexpr int kPrefetch = 3;
      for (int step = 0; step < L; ++step) {
        if (step >= kPrefetch) mbar_wait(&bar_acc[0], (step - kPrefetch) & 1);
        const int t = dir ? (L - 1 - step) : step;
        for (int r = lane - 1; r < 128; r += 31) {
          const int bb = tile * 128 + r;
          if (bb < B) prefetch_l2_bulk(P + (size_t(bb) * L + t) * (2 * LG4) + dir * LG4, LG4 * 4);
        }
      }
    }
  } else {
    // ------------------------------ gate warps ------------------------------
    const int hf = (warp - 1) >> 2;            // which half of the hidden units
    const int quarter = warp & 3;              // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;       // list row inside the tile
    const int b = tile * 128 + row;
    const bool live = b < B;
    float c[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) c[i] = 0.f;
    const uint32_t t_row = tmem_base + (uint32_t(quarter * 32) << 16) + hf * 256;
    for (int step = 0; step < L; ++step) {
      const int t = dir ? (L - 1 - step) : step;
      const size_t tok = size_t(live ? b : 0) * L + t;
      const float* p = P + tok * (2 * LG4) + dir * LG4 + hf * 64;
      float* yo = y + tok * (2 * LH) + dir * LH + hf * 64;
      float* sv = saved ? saved + (tok * 2 + dir) * (LSAVE * LH) + hf * 64 : nullptr;
      // h_{t-1} of this thread's units, for the saved plane (read back from y of the previous step)
      mbar_wait(&bar_acc[hf], step & 1);
      tc_fence_after();
#pragma unroll
      for (int sc = 0; sc < 4; ++sc) {          // 16 units at a time
        float ai[16], af[16], ag[16], ao[16];
        tmem_ld16(t_row + 0 * 64 + sc * 16, ai);
        tmem_ld16(t_row + 1 * 64 + sc * 16, af);
        tmem_ld16(t_row + 2 * 64 + sc * 16, ag);
        tmem_ld16(t_row + 3 * 64 + sc * 16, ao);
        uint32_t hp[8];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          float4 pi = make_float4(0.f, 0.f, 0.f, 0.f), pf = pi, pg = pi, po = pi;
          if (live) {
            pi = *reinterpret_cast<const float4*>(p + 0 * LH + sc * 16 + q4 * 4);
            pf = *reinterpret_cast<const float4*>(p + 1 * LH + sc * 16 + q4 * 4);
            pg = *reinterpret_cast<const float4*>(p + 2 * LH + sc * 16 + q4 * 4);
            po = *reinterpret_cast<const float4*>(p + 3 * LH + sc * 16 + q4 * 4);
          }
          const float pis[4] = {pi.x, pi.y, pi.z, pi.w}, pfs[4] = {pf.x, pf.y, pf.z, pf.w};
          const float pgs[4] = {pg.x, pg.y, pg.z, pg.w}, pos[4] = {po.x, po.y, po.z, po.w};
          float hv[4], gi4[4], gf4[4], gg4[4], go4[4], cv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int u = q4 * 4 + e;
            const float gi = fast_sigmoid(ai[u] + pis[e]), gf = fast_sigmoid(af[u] + pfs[e]);
            const float gg = fast_tanh(ag[u] + pgs[e]), go = fast_sigmoid(ao[u] + pos[e]);
            const float cn = gf * c[sc * 16 + u] + gi * gg;
            c[sc * 16 + u] = cn;
            hv[e] = go * fast_tanh(cn);
            gi4[e] = gi; gf4[e] = gf; gg4[e] = gg; go4[e] = go; cv[e] = cn;
          }
          hp[q4 * 2] = pack_half2(hv[0], hv[1]);
          hp[q4 * 2 + 1] = pack_half2(hv[2], hv[3]);
          if (live) {
            const int off = sc * 16 + q4 * 4;
            if (sv != nullptr) {
              *reinterpret_cast<float4*>(sv + 0 * LH + off) = make_float4(gi4[0], gi4[1], gi4[2], gi4[3]);
              *reinterpret_cast<float4*>(sv + 1 * LH + off) = make_float4(gf4[0], gf4[1], gf4[2], gf4[3]);
              *reinterpret_cast<float4*>(sv + 2 * LH + off) = make_float4(gg4[0], gg4[1], gg4[2], gg4[3]);
              *reinterpret_cast<float4*>(sv + 3 * LH + off) = make_float4(go4[0], go4[1], go4[2], go4[3]);
              *reinterpret_cast<float4*>(sv + 4 * LH + off) = make_float4(cv[0], cv[1], cv[2], cv[3]);
              // h_{t-1}: what this thread wrote to y one step ago (zero at the first step)
              float4 hprev = make_float4(0.f, 0.f, 0.f, 0.f);
              if (step > 0) {
                const int tp = dir ? (t + 1) : (t - 1);
                hprev = *reinterpret_cast<const float4*>(y + (size_t(b) * L + tp) * (2 * LH) + dir * LH + hf * 64 + off);
              }
              *reinterpret_cast<float4*>(sv + 5 * LH + off) = hprev;
            }
            *reinterpret_cast<float4*>(yo + off) = make_float4(hv[0], hv[1], hv[2], hv[3]);
          }
        }
        // 16 fp16 values = two 16-byte chunks of this row's 128-byte k-block row (k-block = hf)
        uint8_t* arow = sA + ((step + 1) & 1) * LstmFwdSmem::A_BYTES + hf * (128 * 128);
        *reinterpret_cast<uint4*>(arow + sw128_offset(row, sc * 2)) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
        *reinterpret_cast<uint4*>(arow + sw128_offset(row, sc * 2 + 1)) = make_uint4(hp[4], hp[5], hp[6], hp[7]);
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(bar_h);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------------------------
// Backward recurrence.
// K ordering of the contraction dh_rec = da W_hh: K-block kb = hf*4 + sc holds (gate q, unit hf*64 + sc*16 + uu) at
// column j = q*16 + uu.  B operand row n = hidden index of dh_rec.
// ------------------------------------------------------------------------------------------------------------
struct LstmBwdSmem {
  static constexpr int KB_BYTES = 128 * 128;            // one k-block: 128 rows x 128 B
  static constexpr int B_BYTES = 8 * KB_BYTES;          // W_hh^T, 8 k-blocks
  static constexpr int RING = 4;                        // da k-block slots
  static constexpr size_t TOTAL = 1024 + B_BYTES + RING * KB_BYTES + 256;
};

// amax |x| over n floats -> *out (uint bits of a non-negative float; zero-initialised by the caller)
__global__ void __launch_bounds__(256) amax_abs_kernel(const float* __restrict__ x, size_t n, unsigned int* __restrict__ out) {
  float m = 0.f;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}
// scale[0] = 2^k with amax * 2^k in [32, 64]; scale[1] = 2^-k   (1, 1 when amax is 0 or not finite)
__global__ void grad_scale_kernel(const unsigned int* __restrict__ amax_bits, float* __restrict__ scale) {
  const float a = __uint_as_float(*amax_bits);
  float s = 1.f;
  if (a > 0.f && a < 3.0e38f) {
    int e;
    frexpf(a, &e);              // a = m * 2^e, m in [0.5, 1)
    s = ldexpf(1.f, 6 - e);     // a * s in [32, 64)
  }
  scale[0] = s;
  scale[1] = 1.f / s;
}

__global__ void __launch_bounds__(288, 1)
lstm_rec_bwd_tc_kernel(const float* __restrict__ dy, const float* __restrict__ saved, const float* __restrict__ whh_f,
                       const float* __restrict__ whh_r, const float* __restrict__ scale_ptr, float* __restrict__ dA,
                       int B, int L) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sB = smem;
  uint8_t* sRing = smem + LstmBwdSmem::B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRing + LstmBwdSmem::RING * LstmBwdSmem::KB_BYTES);
  uint64_t* blk_full = bars;        // [4] count 128
  uint64_t* blk_empty = bars + 4;   // [4] count 1 (tcgen05.commit)
  uint64_t* bar_d = bars + 8;       // dh_rec of the step complete (tcgen05.commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, dir = blockIdx.y;
  const float* whh = dir ? whh_r : whh_f;

  // ---- one-time: W_hh^T as fp16 K-major operand with the K ordering above
  for (int i = threadIdx.x; i < 8 * 8 * 128; i += blockDim.x) {
    const int n = i & 127, c = (i >> 7) & 7, kb = i >> 10;    // n fastest: coalesced reads of W_hh rows
    const int hf = kb >> 2, sc = kb & 3, q = c >> 1, uu0 = (c & 1) * 8;
    const float* src = whh + size_t(q * LH + hf * 64 + sc * 16 + uu0) * LH + n;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = src[size_t(e) * LH];
    *reinterpret_cast<uint4*>(sB + kb * LstmBwdSmem::KB_BYTES + sw128_offset(n, c)) =
        make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]), pack_half2(v[6], v[7]));
  }
  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < 4; ++i) { mbar_init(&blk_full[i], 128); mbar_init(&blk_empty[i], 1); }
      mbar_init(bar_d, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<256>(tmem_slot);   // two dh_rec accumulators of 128 columns (ping-pong over steps)
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kFmtF16, 128, 128, false, false);
      const uint32_t ring_addr = smem_u32(sRing), b_addr = smem_u32(sB);
      for (int it = 0; it < L; ++it) {                 // it-th processed step (reverse time order)
        const uint32_t d_tmem = tmem_base + ((it + 1) & 1) * 128;   // produces dh_rec for the NEXT processed step
#pragma unroll
        for (int sc = 0; sc < 4; ++sc) {
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int slot = hf * 2 + (sc & 1);
            const uint32_t uses = uint32_t(it) * 2 + (sc >> 1);
            mbar_wait(&blk_full[slot], uses & 1);
            tc_fence_after();
            const uint64_t da = make_smem_desc_sw128(ring_addr + slot * LstmBwdSmem::KB_BYTES, 16, 1024);
            const uint64_t db = make_smem_desc_sw128(b_addr + (hf * 4 + sc) * LstmBwdSmem::KB_BYTES, 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(d_tmem, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, (sc | hf | k) != 0 ? 1u : 0u);
            umma_commit(&blk_empty[slot]);
          }
        }
        umma_commit(bar_d);
      }
    } else {
      // lanes 1..31: L2 prefetch of the saved record (3 KB) and dy row (512 B) of the step kPrefetch ahead
      constexpr int kPrefetch = 3;
      for (int it = 0; it < L; ++it) {
        if (it >= kPrefetch) mbar_wait(bar_d, (it - kPrefetch) & 1);
        const int step = L - 1 - it;
        const int t = dir ? (L - 1 - step) : step;
        for (int r = lane - 1; r < 128; r += 31) {
          const int bb = tile * 128 + r;
          if (bb < B) {
            const size_t tok = size_t(bb) * L + t;
            prefetch_l2_bulk(saved + (tok * 2 + dir) * (LSAVE * LH), LSAVE * LH * 4);
            prefetch_l2_bulk(dy + tok * (2 * LH) + dir * LH, LH * 4);
          }
        }
      }
    }
  } else {
    // ------------------------------ gate warps ------------------------------
    const int hf = (warp - 1) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int b = tile * 128 + row;
    const bool live = b < B;
    const float scale = scale_ptr[0], inv_scale = scale_ptr[1];
    float dc_rec[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) dc_rec[i] = 0.f;
    for (int it = 0; it < L; ++it) {
      const int step = L - 1 - it;                      // forward step index being differentiated
      const int t = dir ? (L - 1 - step) : step;
      const size_t tok = size_t(live ? b : 0) * L + t;
      const float* sv = saved + (tok * 2 + dir) * (LSAVE * LH) + hf * 64;
      const float* svp = nullptr;                       // saved record of the previous forward step (for c_{t-1})
      if (step > 0) {
        const int tp = dir ? (t + 1) : (t - 1);
        svp = saved + ((size_t(live ? b : 0) * L + tp) * 2 + dir) * (LSAVE * LH) + hf * 64;
      }
      const float* dyr = dy + tok * (2 * LH) + dir * LH + hf * 64;
      float* dar = dA + tok * (2 * LG4) + dir * LG4 + hf * 64;
      const uint32_t d_tmem = tmem_base + (uint32_t(quarter * 32) << 16) + (it & 1) * 128 + hf * 64;
      if (it > 0) {
        mbar_wait(bar_d, (it - 1) & 1);
        tc_fence_after();
      }
#pragma unroll
      for (int sc = 0; sc < 4; ++sc) {
        float dhr[16];
        if (it > 0) {
          tmem_ld16(d_tmem + sc * 16, dhr);
        } else {
#pragma unroll
          for (int u = 0; u < 16; ++u) dhr[u] = 0.f;
        }
        uint32_t pk[4][8];   // [gate][8 packed half2] : 16 units per gate
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int off = sc * 16 + q4 * 4;
          float4 vi = make_float4(0.f, 0.f, 0.f, 0.f), vf = vi, vg = vi, vo = vi, vc = vi, vcp = vi, vdy = vi;
          if (live) {
            vi = *reinterpret_cast<const float4*>(sv + 0 * LH + off);
            vf = *reinterpret_cast<const float4*>(sv + 1 * LH + off);
            vg = *reinterpret_cast<const float4*>(sv + 2 * LH + off);
            vo = *reinterpret_cast<const float4*>(sv + 3 * LH + off);
            vc = *reinterpret_cast<const float4*>(sv + 4 * LH + off);
            if (svp != nullptr) vcp = *reinterpret_cast<const float4*>(svp + 4 * LH + off);
            vdy = *reinterpret_cast<const float4*>(dyr + off);
          }
          const float gi[4] = {vi.x, vi.y, vi.z, vi.w}, gf[4] = {vf.x, vf.y, vf.z, vf.w};
          const float gg[4] = {vg.x, vg.y, vg.z, vg.w}, go[4] = {vo.x, vo.y, vo.z, vo.w};
          const float cc[4] = {vc.x, vc.y, vc.z, vc.w}, cp[4] = {vcp.x, vcp.y, vcp.z, vcp.w};
          const float dyv[4] = {vdy.x, vdy.y, vdy.z, vdy.w};
          float dai[4], daf[4], dag[4], dao[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int u = q4 * 4 + e;
            const float dh = dyv[e] + dhr[u] * inv_scale;
            const float tc = fast_tanh(cc[e]);
            const float d_o = dh * tc;
            const float dc = dc_rec[sc * 16 + u] + dh * go[e] * (1.f - tc * tc);
            dc_rec[sc * 16 + u] = dc * gf[e];
            dai[e] = dc * gg[e] * gi[e] * (1.f - gi[e]);
            daf[e] = dc * cp[e] * gf[e] * (1.f - gf[e]);
            dag[e] = dc * gi[e] * (1.f - gg[e] * gg[e]);
            dao[e] = d_o * go[e] * (1.f - go[e]);
          }
          if (live) {
            *reinterpret_cast<float4*>(dar + 0 * LH + off) = make_float4(dai[0], dai[1], dai[2], dai[3]);
            *reinterpret_cast<float4*>(dar + 1 * LH + off) = make_float4(daf[0], daf[1], daf[2], daf[3]);
            *reinterpret_cast<float4*>(dar + 2 * LH + off) = make_float4(dag[0], dag[1], dag[2], dag[3]);
            *reinterpret_cast<float4*>(dar + 3 * LH + off) = make_float4(dao[0], dao[1], dao[2], dao[3]);
          }
          pk[0][q4 * 2] = pack_half2(dai[0] * scale, dai[1] * scale); pk[0][q4 * 2 + 1] = pack_half2(dai[2] * scale, dai[3] * scale);
          pk[1][q4 * 2] = pack_half2(daf[0] * scale, daf[1] * scale); pk[1][q4 * 2 + 1] = pack_half2(daf[2] * scale, daf[3] * scale);
          pk[2][q4 * 2] = pack_half2(dag[0] * scale, dag[1] * scale); pk[2][q4 * 2 + 1] = pack_half2(dag[2] * scale, dag[3] * scale);
          pk[3][q4 * 2] = pack_half2(dao[0] * scale, dao[1] * scale); pk[3][q4 * 2 + 1] = pack_half2(dao[2] * scale, dao[3] * scale);
        }
        // publish this thread's 128-byte row of K-block (hf, sc) into its ring slot
        const int slot = hf * 2 + (sc & 1);
        const uint32_t uses = uint32_t(it) * 2 + (sc >> 1);
        mbar_wait(&blk_empty[slot], (uses & 1) ^ 1);
        uint8_t* dst = sRing + slot * LstmBwdSmem::KB_BYTES;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          *reinterpret_cast<uint4*>(dst + sw128_offset(row, q * 2)) = make_uint4(pk[q][0], pk[q][1], pk[q][2], pk[q][3]);
          *reinterpret_cast<uint4*>(dst + sw128_offset(row, q * 2 + 1)) = make_uint4(pk[q][4], pk[q][5], pk[q][6], pk[q][7]);
        }
        tc_fence_before();
        fence_proxy_async_smem();
        mbar_arrive(&blk_full[slot]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem_base);
}

}  // namespace rlt
