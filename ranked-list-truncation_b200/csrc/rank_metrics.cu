// Row N4 of SURVEY.md section 8(f): the two full-list metrics of the verify scripts.
//   Metric.taskr_metric (utils/metrics.py:40-58): per list, sort the documents by descending prediction and add
//       +1/log2(i+2) for a relevant document at sorted position i, -1/log2(i+2) otherwise, left to right in float64.
//   Metric.taskc_metric (utils/metrics.py:60-76): per list, sklearn.metrics.roc_auc_score(labels, predictions), lists
//       whose labels are all 0 or all 1 skipped.
// Neither needs a sort: the sorted position of document i is its RANK
//       rank_i = #{j : s_j > s_i} + #{j < i : s_j == s_i}            (ties keep their list order: a stable argsort)
// and the area under the ROC curve is the Mann-Whitney statistic
//       AUC = sum_{i relevant} (#{j irrelevant : s_j < s_i} + 1/2 #{j irrelevant : s_j == s_i}) / (n_rel * n_irr),
// kept as the exact integer 2U until the one final division.  One CTA per list: scores and labels are read ONCE from
// HBM (coalesced, 8 bytes per document) into shared memory; every thread then walks all L documents for its own ones
// (all lanes read the same shared word: broadcast, no bank conflicts).  The DCG terms are scattered to their rank and
// added by one thread in rank order, which reproduces the reference's Python-float accumulation bit for bit.
#include "common.h"

namespace rlt {

constexpr int kRankMaxLen = 1024;
constexpr int kRankThreads = 128;

__global__ void __launch_bounds__(kRankThreads) rank_metrics_kernel(const float* __restrict__ scores,
                                                                    const float* __restrict__ labels,
                                                                    const double* __restrict__ inv_log2, int seq_len,
                                                                    double* __restrict__ dcg_out,
                                                                    double* __restrict__ auc_out,
                                                                    int32_t* __restrict__ auc_valid) {
  __shared__ float s_score[kRankMaxLen];
  __shared__ float s_label[kRankMaxLen];
  __shared__ double s_term[kRankMaxLen];
  __shared__ int s_npos;
  __shared__ unsigned long long s_u2;
  const int L = seq_len;
  const size_t base = size_t(blockIdx.x) * L;
  if (threadIdx.x == 0) { s_npos = 0; s_u2 = 0ull; }
  for (int i = threadIdx.x; i < L; i += kRankThreads) {
    s_score[i] = scores[base + i];
    s_label[i] = labels[base + i];
  }
  __syncthreads();
  int my_pos = 0;
  unsigned long long my_u2 = 0ull;
  for (int i = threadIdx.x; i < L; i += kRankThreads) {
    const float si = s_score[i];
    const bool truthy = s_label[i] != 0.f;     // `if sample_label[origin_index]` (metrics.py:56)
    const bool positive = s_label[i] == 1.f;   // sklearn: pos_label = 1 for {0, 1} labels
    int rank = 0, neg_less = 0, neg_eq = 0;
    for (int j = 0; j < L; ++j) {
      const float sj = s_score[j];
      const bool eq = sj == si;
      rank += int(sj > si) + int(eq && j < i);
      const bool negj = s_label[j] != 1.f;
      neg_less += int(negj && sj < si);
      neg_eq += int(negj && eq);
    }
    const double c = inv_log2[rank];
    s_term[rank] = truthy ? c : -c;
    if (positive) {
      ++my_pos;
      my_u2 += 2ull * unsigned(neg_less) + unsigned(neg_eq);
    }
  }
  if (my_pos) {
    atomicAdd(&s_npos, my_pos);
    atomicAdd(&s_u2, my_u2);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (dcg_out != nullptr) {
      double acc = 0.0;                          // DCG_sample = 0; DCG_sample += term  (metrics.py:53-56)
      for (int r = 0; r < L; ++r) acc = __dadd_rn(acc, s_term[r]);
      dcg_out[blockIdx.x] = acc;
    }
    if (auc_out != nullptr) {
      const int npos = s_npos, nneg = L - npos;
      const bool valid = npos > 0 && nneg > 0;   // metrics.py:72 skips one-class lists
      auc_out[blockIdx.x] = valid ? double(s_u2) / (2.0 * double(npos) * double(nneg)) : 0.0;
      if (auc_valid != nullptr) auc_valid[blockIdx.x] = valid ? 1 : 0;
    }
  }
}

}  // namespace rlt

using namespace rlt;

extern "C" {

int rlt_rank_metrics(const float* scores, const float* labels, const double* inv_log2, int n_lists, int seq_len,
                     double* dcg_out, double* auc_out, int32_t* auc_valid, rlt_stream_t stream_) {
  RLT_REQUIRE(scores && labels && inv_log2 && n_lists > 0 && seq_len > 0, RLT_INVALID_ARG, "rlt_rank_metrics: bad arguments");
  RLT_REQUIRE(dcg_out || auc_out, RLT_INVALID_ARG, "rlt_rank_metrics: no output requested");
  RLT_REQUIRE(seq_len <= kRankMaxLen, RLT_UNSUPPORTED_SHAPE, "rlt_rank_metrics: seq_len %d exceeds %d", seq_len, kRankMaxLen);
  rank_metrics_kernel<<<n_lists, kRankThreads, 0, static_cast<cudaStream_t>(stream_)>>>(scores, labels, inv_log2, seq_len, dcg_out,
                                                                                       auc_out, auc_valid);
  RLT_CHECK_LAUNCH();
  return RLT_OK;
}

}  // extern "C"
