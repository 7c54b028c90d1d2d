// K1 — persistent BiLSTM recurrence on tcgen05, "unit-major" mapping.
//
// The per-step contraction is computed TRANSPOSED:  a^T[512 gate rows, lists] = W_hh[512, 128] . h^T[128, lists]
//   * A operand = W_hh as stored by nn.LSTM ([4H, H], gate blocks i,f,g,o), fp16, RESIDENT IN TENSOR MEMORY for the
//     whole scan (tcgen05.mma with a TMEM A operand: lane = row of the 128-row gate block, one 32-bit column = two
//     consecutive k; 4 blocks x 64 columns = 256 columns, written once with tcgen05.st).  The other 256 columns are the
//     accumulators: the kernel owns all of TMEM and no shared memory is spent on the weights.
//   * B operand = h_{t-1} of the CTA's lists ([lists, 128] fp16, K-major, 8 KB per 32 lists), rewritten every step
//     by the gate warps.
//   * D lives in TMEM with lane = hidden unit and column = list, so a gate thread owns ONE hidden unit of a few
//     lists: consecutive lanes are consecutive hidden units and every global access of the recurrence (P, the saved
//     gates, y, dy, dA) is a 128-byte contiguous warp access on the plain [token, feature] layouts.  (The earlier
//     list-per-lane mapping touched 32 cache lines per warp instruction and ran at ~50 us per step.)
//   * A CTA carries TWO independent halves of 32 lists each (MMA N = 32).  While the gate warps of one half evaluate
//     the nonlinearities, the tensor core runs the other half's contraction; the TMEM-resident W_hh serves both.
//   * fp16 operands keep the 10-bit mantissa of TF32 (|h| < 1, |W_hh| small), fp32 accumulation in TMEM; the cell
//     state c stays in fp32 (shared memory, one private slot per thread and list).
//   * Nonlinearities: 4 ex2 + ONE rcp for the four gates (the four denominators share a reciprocal), ex2 + rcp for
//     tanh(c): 7 MUFU operations per cell instead of 10.
// Backward (BPTT) uses the same mapping: dh_rec^T[128, lists] = W_hh^T[128, 512] . da^T[512, lists] with da in fp16
// scaled by a power of two taken from max|dy| (unscaled when the accumulator is read).
//
// Roles per CTA (576 threads): warp 0 lane 0 issues the MMAs; warps 1..16 are gate warps (half = (warp-1)/8, TMEM
// lane quarter = warp % 4, list sub-range = ((warp-1)/4) % 2); warp 17 pulls the rows of the coming steps into L2.
#pragma once
#include "sm100.cuh"

namespace rlt {

constexpr int UH = 128;          // hidden units
constexpr int UG4 = 512;         // gate rows
// Saved record per (token, direction), U_REC 32-bit words: [0,128) half2(i, f) | [128,256) half2(g, o) | [256,384) c fp32 |
// [384,512) h_{t-1} fp32.  The gate activations only feed the backward's products (11-bit significand is what the
// TF32 GEMMs consuming dA keep anyway); c stays fp32 because tanh(c) and c_{t-1} multiply the recurrent gradient.
constexpr int U_REC = 512;
constexpr int U_REC_GO = 128, U_REC_C = 256, U_REC_HP = 384;
// Tile shape (template parameter TILE = lists per CTA): two pipeline halves of TILE / 2 lists (= MMA N), TILE / 4 lists per
// gate thread.  TILE 64 for large batches; TILE 32 when that still fills no more than one wave of CTAs -- the reference's
// own batch of 63 lists then runs on 4 CTAs instead of 2 with half the gate work per step and thread.
constexpr int U_CHUNK = 4;       // lists per register chunk
constexpr int U_THREADS = 32 + 16 * 32 + 32;   // MMA warp, 16 gate warps, L2 prefetch warp

__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// Measured and NOT adopted (round 2, -DRLT_LSTM_HW_TANH): gate nonlinearities on the hardware tanh (MUFU.TANH, relative
// error 2^-11; sigmoid(x) = 0.5 tanh(0.5 x) + 0.5: 5 MUFU and ~9 FP32 operations per cell instead of 7 MUFU and ~25).
// Forward recurrence 10 % faster, forward + backward 2.5 %; model outputs stay at 3e-4 of the reference, but one
// near-tied cut position of the B = 16 goldens moves (297 -> 296), which the parity contract does not allow.
#ifdef RLT_LSTM_HW_TANH
__device__ __forceinline__ float tanh_hw(float x) {
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float tanh_2mufu(float x) { return tanh_hw(x); }
__device__ __forceinline__ void lstm_gate_values(float ai, float af, float ag, float ao, float& gi, float& gf, float& gg,
                                                 float& go) {
  gi = fmaf(tanh_hw(0.5f * ai), 0.5f, 0.5f);
  gf = fmaf(tanh_hw(0.5f * af), 0.5f, 0.5f);
  go = fmaf(tanh_hw(0.5f * ao), 0.5f, 0.5f);
  gg = tanh_hw(ag);
}
#else
__device__ __forceinline__ float tanh_2mufu(float x) {
  // 2 / (1 + e^(-2x)) - 1 ; saturates correctly when the exponential overflows / underflows
  return fmaf(2.f, rcp_approx(1.f + ex2_approx(-2.f * 1.4426950408889634f * x)), -1.f);
}
// sigmoid(ai), sigmoid(af), tanh(ag), sigmoid(ao) with one shared reciprocal.  Pre-activations are clamped from below so
// that the product of the four denominators stays finite: sigmoid(-20) and tanh(-10) are exact to 2e-9.
__device__ __forceinline__ void lstm_gate_values(float ai, float af, float ag, float ao, float& gi, float& gf, float& gg,
                                                 float& go) {
  constexpr float kL = 1.4426950408889634f;
  // only the negative side can overflow (e^{-x} for x -> -inf); for x -> +inf the exponential underflows to 0
  ai = fmaxf(ai, -20.f);
  af = fmaxf(af, -20.f);
  ao = fmaxf(ao, -20.f);
  ag = fmaxf(ag, -10.f);
  const float di = 1.f + ex2_approx(-kL * ai), df = 1.f + ex2_approx(-kL * af);
  const float dg = 1.f + ex2_approx(-2.f * kL * ag), dO = 1.f + ex2_approx(-kL * ao);
  const float pa = di * df, pb = dg * dO;
  const float r = rcp_approx(pa * pb);
  const float ra = r * pb, rb = r * pa;     // 1 / (di df), 1 / (dg do)
  gi = ra * df;
  gf = ra * di;
  go = rb * dg;
  gg = fmaf(2.f, rb * dO, -1.f);
}
#endif

__device__ __forceinline__ unsigned short f32_to_f16_bits(float x) {
  unsigned short r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(x));
  return r;
}

// two fp32 values as one 32-bit word of fp16 (lo, hi), and back
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %2, %1;" : "=r"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f16x2(uint32_t w, float& lo, float& hi) {
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
  lo = f.x;
  hi = f.y;
}

template <int TILE>
struct LstmUmFwdSmem {
  static constexpr int U_HALF = TILE / 2, U_CELLS = TILE / 4;
  static constexpr int W_BYTES = 0;                        // W_hh lives in tensor memory (A operand), not in shared memory
  static constexpr int H_BYTES = 2 * U_HALF * 128;         // per half: two k-blocks of [32 rows x 128 B]
  static constexpr int C_BYTES = U_CELLS * 512 * 4;        // cell state, [cell][gate thread]
  // the kernel owns ALL 512 TMEM columns: requesting more than half of the SM's shared memory keeps a second CTA (whose
  // tcgen05.alloc would block until this one exits) off the SM
  static constexpr size_t USED = 1024 + W_BYTES + 2 * H_BYTES + C_BYTES + 256;
  static constexpr size_t TOTAL = USED > 120 * 1024 ? USED : 120 * 1024;
};

// Layer-0 input projection fused into the recurrence (kFusedIn): with F <= 4 input features (robust04: 3) the
// pre-activation P_t = x_t W_ih^T + b_ih + b_hh is 12 FMAs per cell on values the thread keeps in registers, so the
// [T, 1024] fp32 tensor P (4 KB per token written by a projection kernel and read back here) does not exist at all.
struct LstmInProj {
  const float* x;          // [B, L, F]
  int F;                   // 1..4
  const float* w_ih[2];    // [512, F] per direction
  const float* b_ih[2];    // [512]
  const float* b_hh[2];
};

template <bool kFusedIn, int TILE>
__global__ void __launch_bounds__(U_THREADS, 1)
lstm_um_fwd_kernel(const float* __restrict__ P, LstmInProj inp, const float* __restrict__ whh_f,
                   const float* __restrict__ whh_r, float* __restrict__ y, float* __restrict__ saved, int B, int L) {
  constexpr int U_TILE = TILE, U_HALF = TILE / 2, U_CELLS = TILE / 4;
  using Smem = LstmUmFwdSmem<TILE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sW = smem;
  uint8_t* sH = sW + Smem::W_BYTES;
  float* sC = reinterpret_cast<float*>(sH + 2 * Smem::H_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sC) + Smem::C_BYTES);
  uint64_t* bar_h = bars;          // [2] h of the half complete in smem (count 256)
  uint64_t* bar_acc = bars + 2;    // [2] accumulator of the half ready (tcgen05.commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  volatile int* s_progress = reinterpret_cast<volatile int*>(tmem_slot + 1);   // steps finished by the gate warps (prefetch pacing)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, dir = blockIdx.y;
  const float* whh = dir ? whh_r : whh_f;

  // ---- one-time: zero h_0 and c_0 (W_hh goes to tensor memory below)
  for (int i = threadIdx.x; i < (2 * Smem::H_BYTES + Smem::C_BYTES) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(sH)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == 0) {
    if (lane == 0) {
      for (int h = 0; h < 2; ++h) { mbar_init(&bar_h[h], 256); mbar_init(&bar_acc[h], 1); }
      *s_progress = 0;
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  fence_proxy_async_smem();   // generic-proxy writes of sH -> visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // W_hh as the TMEM-resident A operand: columns [256, 512), gate block q at 256 + 64 q, lane = row of the block,
  // column c = fp16 pair (k = 2c, 2c + 1).  Written once by the 16 gate warps (TMEM lane quarter x gate block).
  if (warp >= 1 && warp <= 16) {
    const int quarter = warp & 3, qblk = (warp - 1) >> 2;
    const float* src = whh + size_t(qblk * 128 + quarter * 32 + lane) * UH;
#pragma unroll 1
    for (int c0 = 0; c0 < 64; c0 += 16) {
      uint32_t r[16];
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const float4 v = *reinterpret_cast<const float4*>(src + 2 * (c0 + j));
        asm("cvt.rn.satfinite.f16x2.f32 %0, %2, %1;" : "=r"(r[j]) : "f"(v.x), "f"(v.y));
        asm("cvt.rn.satfinite.f16x2.f32 %0, %2, %1;" : "=r"(r[j + 1]) : "f"(v.z), "f"(v.w));
      }
      tmem_st16(tmem_base + (uint32_t(quarter * 32) << 16) + 256 + qblk * 64 + c0, r);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------ MMA issuer ------------------------------
      // (measured and rejected in round 2: one issuing thread per half -- 4 % slower, the other half's gate phase already
      // hides the issue time; k-outer / gate-block-inner instruction order -- no change, the ~70 cycles per instruction
      // are issue cost, not accumulator dependency)
      constexpr uint32_t idesc = make_idesc(kFmtF16, 128, U_HALF, false, false);
      const uint32_t h_addr = smem_u32(sH);
      for (int step = 0; step < L; ++step) {
#pragma unroll 1
        for (int hf = 0; hf < 2; ++hf) {
          if (step > 0) {
            mbar_wait(&bar_h[hf], (step - 1) & 1);
            tc_fence_after();
          }
          // k outer, gate block inner: consecutive MMAs write four DIFFERENT accumulators (a chain of dependent
          // tcgen05.mma retires one instruction per ~70 cycles whatever its size)
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint64_t db = make_smem_desc_sw128(h_addr + hf * Smem::H_BYTES + kb * (U_HALF * 128), 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)     // K = 16 fp16: 8 TMEM columns of A, 32 B of B per MMA
#pragma unroll
              for (int q = 0; q < 4; ++q)
                umma_f16_ts(tmem_base + hf * (4 * U_HALF) + q * U_HALF, tmem_base + 256 + q * 64 + (kb * 4 + k) * 8,
                            db + uint64_t(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&bar_acc[hf]);
        }
      }
    }
  } else if (warp == 17) {
    // pull the P rows (2 KB per list and step, contiguous) of the step that is kAhead ahead into L2, paced by a step
    // counter the gate warps publish (an mbarrier parity wait aliases when the waiter is two phases behind: the
    // backward version of this loop dead-locked on its last iterations)
    constexpr int kAhead = 1;
    for (int step = 0; step < L && !kFusedIn; ++step) {
      while (*s_progress < step - kAhead) __nanosleep(200);
      const int t = dir ? (L - 1 - step) : step;
      for (int r = lane; r < U_TILE; r += 32) {
        const int bb = tile * U_TILE + r;
        if (bb < B) prefetch_l2_bulk(P + (size_t(bb) * L + t) * (2 * UG4) + dir * UG4, UG4 * 4);
      }
    }
  } else {
    // ------------------------------ gate warps ------------------------------
    const int gw = warp - 1;
    const int hf = gw >> 3;                    // pipeline half
    const int sub = (gw >> 2) & 1;             // which 16 lists of the half
    const int quarter = warp & 3;              // TMEM lane quarter this warp may access
    const int u = quarter * 32 + lane;         // hidden unit
    const int gt = threadIdx.x - 32;           // gate thread index 0..511 (cell-state slot)
    const int row0 = sub * U_CELLS;            // first list row inside the half
    const int list0 = tile * U_TILE + hf * U_HALF + row0;
    const uint32_t t_lane = tmem_base + (uint32_t(quarter * 32) << 16) + hf * (4 * U_HALF) + row0;
    // this thread's 2-byte slot inside a list's h row: k-block u/64, 16-byte unit (u%64)/8
    uint8_t* hbase = sH + hf * Smem::H_BYTES + (u >> 6) * (U_HALF * 128) + (u & 7) * 2;
    const int hunit = (u & 63) >> 3;
    // All global addressing below uses 32-bit ELEMENT offsets from the tensor bases (token count * 1536 < 2^32 is
    // checked by the host): one IMAD + one IMAD.WIDE per access instead of the 64-bit multiply chains the size_t
    // form compiled to (ncu: 22 % of the kernel's instructions were IMAD, 36 per cell).
    const uint32_t Lu = uint32_t(L);
    const uint32_t tok_list0 = uint32_t(list0) * Lu;      // token index of (list0, t = 0)
    const uint32_t cP = uint32_t(dir * UG4 + u), cY = uint32_t(dir * UH + u), cS = uint32_t(dir * U_REC + u);
    // fused input projection: this unit's four W_ih rows (F <= 4 columns, zero padded) and summed biases
    float wi[4][4], bsum[4];
    if (kFusedIn) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        bsum[q] = inp.b_ih[dir][q * UH + u] + inp.b_hh[dir][q * UH + u];
#pragma unroll
        for (int f = 0; f < 4; ++f) wi[q][f] = f < inp.F ? inp.w_ih[dir][size_t(q * UH + u) * inp.F + f] : 0.f;
      }
    }
    // pre-activations of U_CHUNK lists whose first token index is tok: from P, or from x (the same 3 floats for every
    // lane: broadcast loads)
    auto load_pre = [&](float (&dst)[4][U_CHUNK], uint32_t tok, int c0) {
#pragma unroll
      for (int li = 0; li < U_CHUNK; ++li) {
        const bool live = list0 + c0 + li < B;
        const uint32_t tk = tok + uint32_t(li) * Lu;
        if (kFusedIn) {
          const float* xr = inp.x + size_t(tk) * inp.F;
          float xv[4];
#pragma unroll
          for (int f = 0; f < 4; ++f) xv[f] = (live && f < inp.F) ? __ldg(xr + f) : 0.f;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            dst[q][li] = fmaf(xv[3], wi[q][3], fmaf(xv[2], wi[q][2], fmaf(xv[1], wi[q][1], fmaf(xv[0], wi[q][0], bsum[q]))));
        } else {
          const float* pr = P + (tk * uint32_t(2 * UG4) + cP);
#pragma unroll
          for (int q = 0; q < 4; ++q) dst[q][li] = live ? __ldg(pr + q * UH) : 0.f;
        }
      }
    };

    for (int step = 0; step < L; ++step) {
      const int t = dir ? (L - 1 - step) : step;
      const uint32_t tokb = tok_list0 + uint32_t(t);       // token of this thread's first list at this step
      // h_t is the "previous h" plane of the NEXT step's record: +-1 token
      const uint32_t dnext = (dir ? 0u - uint32_t(2 * U_REC) : uint32_t(2 * U_REC)) + uint32_t(U_REC_HP);
      const bool has_next = step + 1 < L;
      float pc[4][U_CHUNK];
      load_pre(pc, tokb, 0);
      mbar_wait(&bar_acc[hf], step & 1);
      tc_fence_after();
      uint32_t tok_ch = tokb;                              // token of the first list of the current chunk
#pragma unroll 1
      for (int ch = 0; ch < U_CELLS / U_CHUNK; ++ch, tok_ch += U_CHUNK * Lu) {
        float a[4][U_CHUNK];   // the four gate blocks of this chunk: four TMEM loads in flight, one wait
        tmem_ld4x4(t_lane + 0 * U_HALF + ch * U_CHUNK, t_lane + 1 * U_HALF + ch * U_CHUNK, t_lane + 2 * U_HALF + ch * U_CHUNK,
                   t_lane + 3 * U_HALF + ch * U_CHUNK, a);
        float pn[4][U_CHUNK];
        if (ch + 1 < U_CELLS / U_CHUNK) load_pre(pn, tok_ch + U_CHUNK * Lu, (ch + 1) * U_CHUNK);
        // Pure math for the chunk's four cells first, in ONE branch-free block: the per-cell chains are four MUFU
        // latencies deep (ex2 -> rcp -> ex2 -> rcp) and the `if (live)` store blocks used to sit between the cells as
        // control-flow barriers, so the compiler evaluated the cells one after the other.  Stores follow below.
        float gi[U_CHUNK], gf[U_CHUNK], gg[U_CHUNK], go[U_CHUNK], cn[U_CHUNK], hv[U_CHUNK];
#pragma unroll
        for (int li = 0; li < U_CHUNK; ++li) {
          const int cell = ch * U_CHUNK + li;
          lstm_gate_values(a[0][li] + pc[0][li], a[1][li] + pc[1][li], a[2][li] + pc[2][li], a[3][li] + pc[3][li], gi[li],
                           gf[li], gg[li], go[li]);
          cn[li] = fmaf(gf[li], sC[cell * 512 + gt], gi[li] * gg[li]);
          sC[cell * 512 + gt] = cn[li];
          hv[li] = go[li] * tanh_2mufu(cn[li]);
        }
#pragma unroll
        for (int li = 0; li < U_CHUNK; ++li) {
          const int cell = ch * U_CHUNK + li;
          const bool live = list0 + cell < B;
          const uint32_t tk = tok_ch + uint32_t(li) * Lu;
          if (live) {
            if (saved != nullptr) {
              const uint32_t so = tk * uint32_t(2 * U_REC) + cS;
              float* sv = saved + so;
              sv[0] = __uint_as_float(pack_f16x2(gi[li], gf[li]));
              sv[U_REC_GO] = __uint_as_float(pack_f16x2(gg[li], go[li]));
              sv[U_REC_C] = cn[li];
              if (step == 0) sv[U_REC_HP] = 0.f;
              if (has_next) saved[uint32_t(so + dnext)] = hv[li];
            }
            y[tk * uint32_t(2 * UH) + cY] = hv[li];
          }
          const int r = row0 + cell;
          *reinterpret_cast<unsigned short*>(hbase + (r >> 3) * 1024 + (r & 7) * 128 + (((hunit ^ r) & 7) << 4)) =
              f32_to_f16_bits(hv[li]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int li = 0; li < U_CHUNK; ++li) pc[q][li] = pn[q][li];
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(&bar_h[hf]);
      if (gt == 511) *s_progress = step + 1;     // last gate thread of the second half
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------------------------
// Backward recurrence:  dh_rec^T[128 units k, lists] = W_hh^T[128, 512] . da^T[512, lists]
//   A operand: W_hh^T, fp16 K-major, 8 k-blocks of [128 rows (k) x 64 gate rows n] (128 KB, resident)
//   B operand: da of the half's 32 lists, [32 rows x 512 n] fp16 K-major (8 k-blocks x 4 KB), scaled by 2^s
//   D: TMEM lane = hidden unit k, column = list (32 columns per half)
// ------------------------------------------------------------------------------------------------------------
template <int TILE>
struct LstmUmBwdSmem {
  static constexpr int U_HALF = TILE / 2, U_CELLS = TILE / 4;
  static constexpr int KB_BYTES = 128 * 128;               // one k-block of W_hh^T
  static constexpr int W_BYTES = 0;                        // W_hh^T lives in tensor memory (A operand)
  static constexpr int DA_KB = U_HALF * 128;               // one k-block of da (32 rows x 128 B)
  static constexpr int DA_BYTES = 8 * DA_KB;               // per half
  static constexpr int C_BYTES = U_CELLS * 512 * 4;        // dc_rec, [cell][gate thread]
  static constexpr size_t USED = 1024 + W_BYTES + 2 * DA_BYTES + C_BYTES + 256;
  static constexpr size_t TOTAL = USED > 120 * 1024 ? USED : 120 * 1024;   // one CTA per SM (it owns all of TMEM)
};

template <int TILE>
__global__ void __launch_bounds__(U_THREADS, 1)
lstm_um_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ saved, const float* __restrict__ whh_f,
                   const float* __restrict__ whh_r, const float* __restrict__ scale_ptr, float* __restrict__ dA,
                   float* __restrict__ db /* [2][512] += column sums of dA (bias gradients) */, int B, int L) {
  constexpr int U_TILE = TILE, U_HALF = TILE / 2, U_CELLS = TILE / 4;
  using Smem = LstmUmBwdSmem<TILE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sW = smem;
  uint8_t* sDA = sW + Smem::W_BYTES;
  float* sDC = reinterpret_cast<float*>(sDA + 2 * Smem::DA_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sDC) + Smem::C_BYTES);
  uint64_t* bar_da = bars;         // [2] da of the half complete in smem (count 256)
  uint64_t* bar_d = bars + 2;      // [2] dh_rec of the half ready (tcgen05.commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  volatile int* s_progress = reinterpret_cast<volatile int*>(tmem_slot + 1);   // steps finished by the gate warps (prefetch pacing)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, dir = blockIdx.y;
  const float* whh = dir ? whh_r : whh_f;

  // ---- one-time: zero da and dc_rec (W_hh^T goes to tensor memory below)
  for (int i = threadIdx.x; i < (2 * Smem::DA_BYTES + Smem::C_BYTES) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(sDA)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == 0) {
    if (lane == 0) {
      for (int h = 0; h < 2; ++h) { mbar_init(&bar_da[h], 256); mbar_init(&bar_d[h], 1); }
      *s_progress = 0;
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // W_hh^T as the TMEM-resident A operand: columns [256, 512): lane = hidden unit k, column 256 + c = fp16 pair of gate
  // rows n = 2c, 2c + 1 (A[k][n] = W_hh[n][k]).  Written once by the 16 gate warps: lane quarter x 128-row gate block.
  if (warp >= 1 && warp <= 16) {
    const int quarter = warp & 3, qblk = (warp - 1) >> 2;
    const float* src = whh + size_t(qblk * 128) * UH + quarter * 32 + lane;
#pragma unroll 1
    for (int c0 = 0; c0 < 64; c0 += 16) {
      uint32_t r[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float lo = src[size_t(2 * (c0 + j)) * UH], hi = src[size_t(2 * (c0 + j) + 1) * UH];
        asm("cvt.rn.satfinite.f16x2.f32 %0, %2, %1;" : "=r"(r[j]) : "f"(lo), "f"(hi));
      }
      tmem_st16(tmem_base + (uint32_t(quarter * 32) << 16) + 256 + qblk * 64 + c0, r);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------ MMA issuer ------------------------------
      constexpr uint32_t idesc = make_idesc(kFmtF16, 128, U_HALF, false, false);
      const uint32_t da_addr = smem_u32(sDA);
      for (int it = 0; it + 1 < L; ++it) {             // the last processed step has no consumer for dh_rec
#pragma unroll 1
        for (int hf = 0; hf < 2; ++hf) {
          mbar_wait(&bar_da[hf], it & 1);
          tc_fence_after();
#pragma unroll
          for (int kb = 0; kb < 8; ++kb) {
            const uint64_t db = make_smem_desc_sw128(da_addr + hf * Smem::DA_BYTES + kb * Smem::DA_KB, 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_ts(tmem_base + hf * U_HALF, tmem_base + 256 + (kb * 4 + k) * 8, db + uint64_t(2 * k), idesc,
                          (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&bar_d[hf]);
        }
      }
    }
  } else if (warp == 17) {
    // L2 prefetch of the saved record (gates + c: 1.5 KB) and dy row (512 B) of the step kAhead ahead
    constexpr int kAhead = 1;
    for (int it = 0; it < L; ++it) {
      while (*s_progress < it - kAhead) __nanosleep(200);
      const int step = L - 1 - it;
      const int t = dir ? (L - 1 - step) : step;
      for (int r = lane; r < U_TILE; r += 32) {
        const int bb = tile * U_TILE + r;
        if (bb < B) {
          const size_t tok = size_t(bb) * L + t;
          prefetch_l2_bulk(saved + (tok * 2 + dir) * U_REC, U_REC_HP * 4);
          prefetch_l2_bulk(dy + tok * (2 * UH) + dir * UH, UH * 4);
        }
      }
    }
  } else {
    // ------------------------------ gate warps ------------------------------
    const int gw = warp - 1;
    const int hf = gw >> 3;
    const int sub = (gw >> 2) & 1;
    const int quarter = warp & 3;
    const int u = quarter * 32 + lane;
    const int gt = threadIdx.x - 32;
    const int row0 = sub * U_CELLS;
    const int list0 = tile * U_TILE + hf * U_HALF + row0;
    const uint32_t t_lane = tmem_base + (uint32_t(quarter * 32) << 16) + hf * U_HALF + row0;
    const float scale = scale_ptr[0], inv_scale = scale_ptr[1];
    // da[list row r][n = q*128 + u]: k-block q*2 + u/64, 16-byte unit (u%64)/8, 2 bytes at (u%8)*2
    uint8_t* dabase = sDA + hf * Smem::DA_BYTES + (u >> 6) * Smem::DA_KB + (u & 7) * 2;
    const int dunit = (u & 63) >> 3;
    // 32-bit element offsets from the tensor bases, as in the forward kernel
    const uint32_t Lu = uint32_t(L);
    const uint32_t tok_list0 = uint32_t(list0) * Lu;
    const uint32_t cS = uint32_t(dir * U_REC + u), cY = uint32_t(dir * UH + u), cA = uint32_t(dir * UG4 + u);
    float dbs[4] = {0.f, 0.f, 0.f, 0.f};                // this unit's share of db = sum over (list, t) of da
    // (i, f), (g, o), c of the step, c of the previous step, dy: U_CHUNK lists whose first token index is tok
    auto load_rec = [&](float (&dst)[5][U_CHUNK], uint32_t tok, int c0, bool has_prev, uint32_t dprev) {
#pragma unroll
      for (int li = 0; li < U_CHUNK; ++li) {
        const bool live = list0 + c0 + li < B;
        const uint32_t tk = tok + uint32_t(li) * Lu;
        const uint32_t so = tk * uint32_t(2 * U_REC) + cS;
        const float* sr = saved + so;
        dst[0][li] = live ? __ldg(sr) : 0.f;
        dst[1][li] = live ? __ldg(sr + U_REC_GO) : 0.f;
        dst[2][li] = live ? __ldg(sr + U_REC_C) : 0.f;
        dst[3][li] = (live && has_prev) ? __ldg(saved + uint32_t(so + dprev)) : 0.f;   // offset sum wraps in 32 bits
        dst[4][li] = live ? __ldg(dy + (tk * uint32_t(2 * UH) + cY)) : 0.f;
      }
    };

    for (int it = 0; it < L; ++it) {
      const int step = L - 1 - it;                      // forward step index being differentiated
      const int t = dir ? (L - 1 - step) : step;
      const uint32_t tokb = tok_list0 + uint32_t(t);
      // c_{t-1} lives in the record of the previous forward step: -+1 token
      const uint32_t dprev = (dir ? uint32_t(2 * U_REC) : 0u - uint32_t(2 * U_REC)) + uint32_t(U_REC_C);
      const bool has_prev = step > 0;
      // operands of the first chunk: in flight while the tensor core finishes dh_rec
      float vc[5][U_CHUNK];   // (i, f), (g, o), c, c_prev, dy
      load_rec(vc, tokb, 0, has_prev, dprev);
      if (it > 0) {
        mbar_wait(&bar_d[hf], (it - 1) & 1);
        tc_fence_after();
      }
      uint32_t tok_ch = tokb;
#pragma unroll 1
      for (int ch = 0; ch < U_CELLS / U_CHUNK; ++ch, tok_ch += U_CHUNK * Lu) {
        float dhr[U_CHUNK];
        if (it > 0) {
          tmem_ld4(t_lane + ch * U_CHUNK, dhr);
        } else {
#pragma unroll
          for (int li = 0; li < U_CHUNK; ++li) dhr[li] = 0.f;
        }
        float vn[5][U_CHUNK];
        if (ch + 1 < U_CELLS / U_CHUNK) load_rec(vn, tok_ch + U_CHUNK * Lu, (ch + 1) * U_CHUNK, has_prev, dprev);
        // math of the chunk's four cells in one branch-free block (see the forward kernel), stores afterwards
        float dai[U_CHUNK], daf[U_CHUNK], dag[U_CHUNK], dao[U_CHUNK];
#pragma unroll
        for (int li = 0; li < U_CHUNK; ++li) {
          const int cell = ch * U_CHUNK + li;
          float gi, gf, gg, go;
          unpack_f16x2(__float_as_uint(vc[0][li]), gi, gf);
          unpack_f16x2(__float_as_uint(vc[1][li]), gg, go);
          const float cc = vc[2][li], cp = vc[3][li];
          const float dh = fmaf(dhr[li], inv_scale, vc[4][li]);
          const float tc = tanh_2mufu(cc);
          const float d_o = dh * tc;
          const float dc = fmaf(dh * go, 1.f - tc * tc, sDC[cell * 512 + gt]);
          sDC[cell * 512 + gt] = dc * gf;
          dai[li] = dc * gg * gi * (1.f - gi);
          daf[li] = dc * cp * gf * (1.f - gf);
          dag[li] = dc * gi * (1.f - gg * gg);
          dao[li] = d_o * go * (1.f - go);
        }
#pragma unroll
        for (int li = 0; li < U_CHUNK; ++li) {
          const int cell = ch * U_CHUNK + li;
          const bool live = list0 + cell < B;
          if (live) {
            float* o = dA + ((tok_ch + uint32_t(li) * Lu) * uint32_t(2 * UG4) + cA);
            o[0 * UH] = dai[li]; o[1 * UH] = daf[li]; o[2 * UH] = dag[li]; o[3 * UH] = dao[li];
            dbs[0] += dai[li]; dbs[1] += daf[li]; dbs[2] += dag[li]; dbs[3] += dao[li];
          }
          const int r = row0 + cell;
          uint8_t* dst = dabase + (r >> 3) * 1024 + (r & 7) * 128 + (((dunit ^ r) & 7) << 4);
          *reinterpret_cast<unsigned short*>(dst + 0 * 2 * Smem::DA_KB) = f32_to_f16_bits(dai[li] * scale);
          *reinterpret_cast<unsigned short*>(dst + 1 * 2 * Smem::DA_KB) = f32_to_f16_bits(daf[li] * scale);
          *reinterpret_cast<unsigned short*>(dst + 2 * 2 * Smem::DA_KB) = f32_to_f16_bits(dag[li] * scale);
          *reinterpret_cast<unsigned short*>(dst + 3 * 2 * Smem::DA_KB) = f32_to_f16_bits(dao[li] * scale);
        }
#pragma unroll
        for (int pl = 0; pl < 5; ++pl)
#pragma unroll
          for (int li = 0; li < U_CHUNK; ++li) vc[pl][li] = vn[pl][li];
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(&bar_da[hf]);
      if (gt == 511) *s_progress = it + 1;
    }
    if (db != nullptr) {
#pragma unroll
      for (int q = 0; q < 4; ++q) atomicAdd(db + dir * UG4 + q * UH + u, dbs[q]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}

}  // namespace rlt
