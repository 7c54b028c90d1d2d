// Packed fp32 arithmetic of sm_100 (FADD2 / FMUL2 / FFMA2: two IEEE operations per instruction, each component rounded
// exactly like the scalar instruction) on float2 values.  Used where a kernel is bound by instruction issue: the K3 loss
// kernels (cut_loss_pair.cuh).  (Measured and rejected for the BiLSTM gate math: two cells per packed instruction made
// the forward recurrence 6-7 % SLOWER -- that phase is bound by its dependent MUFU chains, and packing couples two cells'
// chains and adds register-pair moves.)
#pragma once

namespace rlt {

__device__ __forceinline__ float2 f2_fma(float2 a, float2 b, float2 c) {
  unsigned long long x, y, z, r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(z) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(x), "l"(y), "l"(z));
  float2 o;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(r));
  return o;
}
__device__ __forceinline__ float2 f2_mul(float2 a, float2 b) {
  unsigned long long x, y, r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(b.x), "f"(b.y));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(x), "l"(y));
  float2 o;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(r));
  return o;
}
__device__ __forceinline__ float2 f2_add(float2 a, float2 b) {
  unsigned long long x, y, r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(b.x), "f"(b.y));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(x), "l"(y));
  float2 o;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(r));
  return o;
}
__device__ __forceinline__ float2 f2_dup(float v) { return make_float2(v, v); }

}  // namespace rlt
