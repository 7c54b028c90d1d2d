// Train-mode dropout (reference: nn.TransformerEncoderLayer(dropout=p) at models/Choopy.py:11, AttnCut.py:9, ...;
// nn.Dropout(p) on the BiCut logits, models/Bicut.py:14).  torch's Philox stream cannot be reproduced bit for bit, so
// the masks come from a stateless counter hash: keep(seed, site, element) is a pure function, evaluated again in the
// backward kernels instead of storing masks.  p is quantised to 1/65536: a 64-bit hash yields four 16-bit uniforms
// (one group of 4 consecutive elements), an element is dropped when its uniform is < thr = round(p * 65536), kept
// values are scaled by 65536 / (65536 - thr).
#pragma once
#include <stdint.h>

namespace rlt {

enum DropSite : uint32_t { DROP_ATTN = 1, DROP_AFTER_ATTN = 2, DROP_FFN = 3, DROP_AFTER_FFN = 4, DROP_LOGITS = 5 };

struct DropCfg {
  uint64_t seed;
  uint32_t thr;     // 0 = no dropout
  float scale;      // 1 / (1 - p_quantised)
};

__host__ __device__ inline DropCfg make_drop(float p, uint64_t seed) {
  DropCfg c;
  c.seed = seed;
  int t = int(p * 65536.f + 0.5f);
  if (t < 0) t = 0;
  if (t > 65535) t = 65535;
  c.thr = uint32_t(t);
  c.scale = 65536.f / float(65536 - t);
  return c;
}

__host__ __device__ __forceinline__ uint64_t drop_mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// four 16-bit uniforms for group `g` (4 consecutive elements) of dropout site `site`
__host__ __device__ __forceinline__ uint64_t drop_bits(uint64_t seed, uint32_t site, uint64_t g) {
  return drop_mix64((seed + uint64_t(site) * 0xD1B54A32D192ED03ull) ^ (g * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull));
}
// keep-and-scale factor of element e (0..3) of a group
__host__ __device__ __forceinline__ float drop_factor(uint64_t bits, int e, uint32_t thr, float scale) {
  return (uint32_t(bits >> (16 * e)) & 0xFFFFu) >= thr ? scale : 0.f;
}

}  // namespace rlt
