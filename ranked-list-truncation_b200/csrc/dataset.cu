// Row N3 of SURVEY.md section 8(f): the step before the hot path.
// The reference keeps the whole split as two host tensors (dataloader/attncut_dataloader.py:59: X [N, L, F] float32,
// y [N, L] float32 in {0, 1}) and lets a torch DataLoader collate a shuffled batch per step on the host
// (attncut_dataloader.py:86-87), followed by two host->device copies in run.py:123-124.  Here the split lives in HBM
// once; a batch is ONE launch that gathers the selected lists into the contiguous [b, L, F] / [b, L] tensors the
// kernels read.  Labels may be stored as bit masks (one uint32 per 32 documents: 32x less resident memory and label
// read traffic); the gather expands them back to the float32 {0., 1.} layout of the reference.
#include "common.h"

namespace rlt {

constexpr int kGatherThreads = 256;

// One WARP per output list (grid-stride over warps): the X row (L*F floats) and the label row (L floats or
// ceil(L/32) words) of list index[o] are copied with 128-bit accesses when the row length allows it, four loads in
// flight per lane before the first store.  (The first version gave a list to a whole CTA: at L = 300, F = 1 a row is 75
// float4, so 181 of 256 threads idled behind one dependent load each -- 0.21-0.47 of the copy peak,
// profiles/r02_bench_data.txt.)
template <typename T>
__device__ __forceinline__ void warp_copy(const T* __restrict__ src, T* __restrict__ dst, int n, int lane) {
  int i = lane;
  for (; i + 96 < n; i += 128) {
    const T a = __ldg(src + i), b = __ldg(src + i + 32), c = __ldg(src + i + 64), d = __ldg(src + i + 96);
    dst[i] = a; dst[i + 32] = b; dst[i + 64] = c; dst[i + 96] = d;
  }
  for (; i < n; i += 32) dst[i] = __ldg(src + i);
}

template <bool kVec4>
__global__ void __launch_bounds__(kGatherThreads) gather_lists_kernel(const float* __restrict__ x,
                                                                      const float* __restrict__ y,
                                                                      const uint32_t* __restrict__ y_bits,
                                                                      const int64_t* __restrict__ index, int n_src, int n_out,
                                                                      int seq_len, int n_features, float* __restrict__ x_out,
                                                                      float* __restrict__ y_out, int32_t* __restrict__ status) {
  const int row_x = seq_len * n_features;
  const int words = (seq_len + 31) >> 5;
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * kGatherThreads) >> 5;
  for (int o = (blockIdx.x * kGatherThreads + threadIdx.x) >> 5; o < n_out; o += warps) {
    const int64_t src = index != nullptr ? index[o] : int64_t(o);
    if (src < 0 || src >= n_src) {                       // IndexError on the host side
      if (lane == 0) atomicOr(status, 1);
      continue;
    }
    const float* xs = x + size_t(src) * row_x;
    float* xd = x_out + size_t(o) * row_x;
    if (kVec4) warp_copy(reinterpret_cast<const float4*>(xs), reinterpret_cast<float4*>(xd), row_x >> 2, lane);
    else warp_copy(xs, xd, row_x, lane);
    if (y_out == nullptr) continue;
    float* yd = y_out + size_t(o) * seq_len;
    if (y_bits != nullptr) {
      const uint32_t* ws = y_bits + size_t(src) * words;
      for (int w = 0; w < words; ++w) {                  // one word serves the 32 lanes (broadcast load)
        const int i = w * 32 + lane;
        const uint32_t m = __ldg(ws + w);
        if (i < seq_len) yd[i] = float((m >> lane) & 1u);
      }
    } else {
      const float* ys = y + size_t(src) * seq_len;
      if (kVec4) warp_copy(reinterpret_cast<const float4*>(ys), reinterpret_cast<float4*>(yd), seq_len >> 2, lane);
      else warp_copy(ys, yd, seq_len, lane);
    }
  }
}

// One warp per list: 32 labels -> one ballot -> one word.  A label other than 0. or 1. cannot be a bit: status |= 2.
__global__ void __launch_bounds__(256) pack_labels_kernel(const float* __restrict__ y, int n_lists, int seq_len,
                                                          uint32_t* __restrict__ bits, int32_t* __restrict__ status) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int words = (seq_len + 31) >> 5;
  bool bad = false;
  for (int b = warp; b < n_lists; b += nwarps) {
    const float* row = y + size_t(b) * seq_len;
    for (int w = 0; w < words; ++w) {
      const int i = w * 32 + lane;
      const float v = i < seq_len ? row[i] : 0.f;
      bad |= !(v == 0.f || v == 1.f);
      const uint32_t m = __ballot_sync(0xffffffffu, v == 1.f);
      if (lane == 0) bits[size_t(b) * words + w] = m;
    }
  }
  if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(status, 2);
}

}  // namespace rlt

using namespace rlt;

extern "C" {

int rlt_gather_lists(const float* x, const float* labels, const uint32_t* label_bits, const int64_t* index, int n_src,
                     int n_out, int seq_len, int n_features, float* x_out, float* labels_out, int32_t* status,
                     rlt_stream_t stream_) {
  RLT_REQUIRE(x && x_out && status && n_src > 0 && n_out > 0 && seq_len > 0 && n_features > 0, RLT_INVALID_ARG,
              "rlt_gather_lists: bad arguments");
  RLT_REQUIRE(labels_out == nullptr || (labels != nullptr) != (label_bits != nullptr), RLT_INVALID_ARG,
              "rlt_gather_lists: labels_out needs exactly one of labels / label_bits");
  const size_t row_x = size_t(seq_len) * n_features;
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(x_out) | reinterpret_cast<uintptr_t>(labels) |
                         reinterpret_cast<uintptr_t>(labels_out)) & 15u) == 0;
  const bool vec4 = aligned && row_x % 4 == 0 && seq_len % 4 == 0;
  int grid = num_sms() * 8;                              // 8 CTAs x 8 warps per SM, one list per warp and pass
  if (grid > (n_out + 7) / 8) grid = (n_out + 7) / 8;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (vec4)
    gather_lists_kernel<true><<<grid, kGatherThreads, 0, stream>>>(x, labels, label_bits, index, n_src, n_out, seq_len, n_features,
                                                                   x_out, labels_out, status);
  else
    gather_lists_kernel<false><<<grid, kGatherThreads, 0, stream>>>(x, labels, label_bits, index, n_src, n_out, seq_len, n_features,
                                                                    x_out, labels_out, status);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

int rlt_pack_labels(const float* labels, int n_lists, int seq_len, uint32_t* label_bits, int32_t* status, rlt_stream_t stream_) {
  RLT_REQUIRE(labels && label_bits && status && n_lists > 0 && seq_len > 0, RLT_INVALID_ARG, "rlt_pack_labels: bad arguments");
  int grid = (n_lists + 7) / 8;
  const int cap = num_sms() * 8;
  if (grid > cap) grid = cap;
  pack_labels_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(labels, n_lists, seq_len, label_bits, status);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

}  // extern "C"
