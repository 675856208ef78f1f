// Evaluation tail on the device (reference Main.py:390-448): train-mask, top-K ranking and Recall / NDCG / Precision per
// test user without a host round trip per batch.
//
//   predict = score * (1 - mask) - mask * 1e8      (Main.py:410)  -> dmm_eval_mask_scores: the train items of each test
//                                                      user (CSR row) are overwritten with -1e8 in the score block
//   torch.topk(predict, K)                          (Main.py:411)  -> dmm_topk_edges with k = K for every row
//   calcRes                                         (Main.py:422-448) -> dmm_eval_metrics: one warp per user ranks its K
//                                                      selected columns by (score desc, column asc) and walks the user's
//                                                      test items in their stored order with the reference's float64
//                                                      arithmetic (the 1 / log2(pos + 2) and max-DCG tables come from
//                                                      the host's numpy, so every term is bit-identical to calcRes).
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) eval_mask_kernel(const int64_t* __restrict__ indptr,
                                                        const int32_t* __restrict__ indices,
                                                        const int64_t* __restrict__ row_ids, int64_t n_rows,
                                                        int64_t n_cols, float* __restrict__ scores, int64_t ld,
                                                        float fill) {
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n_rows) return;
  const int64_t u = row_ids[r];
  const int64_t b = indptr[u], e = indptr[u + 1];
  for (int64_t k = b + lane; k < e; k += 32) {
    const int32_t c = indices[k];
    if (c >= 0 && c < n_cols) scores[r * ld + c] = fill;
  }
}

__global__ void __launch_bounds__(256) eval_metrics_kernel(const float* __restrict__ scores, int64_t ld, int64_t n_rows,
                                                           const int32_t* __restrict__ top_items, int K,
                                                           const int64_t* __restrict__ row_ids,
                                                           const int64_t* __restrict__ test_ptr,
                                                           const int32_t* __restrict__ test_items,
                                                           const double* __restrict__ inv_log2,
                                                           const double* __restrict__ max_dcg,
                                                           double* __restrict__ out) {
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n_rows) return;
  // lane j < K holds the j-th selected column (ascending) and its score; rank = position in (score desc, column asc)
  const int col = lane < K ? top_items[r * K + lane] : 0x7FFFFFFF;
  const float s = lane < K ? scores[r * ld + col] : 0.f;
  int rank = 0;
  for (int j = 0; j < K; ++j) {
    const float sj = __shfl_sync(0xffffffffu, s, j);
    const int cj = __shfl_sync(0xffffffffu, col, j);
    rank += (sj > s || (sj == s && cj < col)) ? 1 : 0;
  }
  const int64_t u = row_ids[r];
  const int64_t b = test_ptr[u], e = test_ptr[u + 1];
  const int64_t tst = e - b;
  double dcg = 0.0;
  int64_t hits = 0;
  for (int64_t k = b; k < e; ++k) {           // the user's test items in their stored order (calcRes' loop order)
    const int32_t item = test_items[k];
    const uint32_t m = __ballot_sync(0xffffffffu, lane < K && col == item);
    if (m) {
      const int pos = __shfl_sync(0xffffffffu, rank, __ffs(m) - 1);
      dcg += inv_log2[pos];
      ++hits;
    }
  }
  if (lane == 0) {
    const double h = (double)hits;
    out[3 * r + 0] = tst > 0 ? h / (double)tst : 0.0;
    out[3 * r + 1] = tst > 0 ? dcg / max_dcg[tst < K ? tst : K] : 0.0;
    out[3 * r + 2] = h / (double)K;
  }
}

}  // namespace

extern "C" int dmm_eval_mask_scores(dmm_ctx* ctx, const int64_t* indptr, const int32_t* indices, const int64_t* row_ids,
                                    int64_t n_rows, int64_t n_cols, float* scores, int64_t ld, float fill, void* stream) {
  DMM_CHECK_ARG(ctx && indptr && indices && row_ids && scores, "dmm_eval_mask_scores: null argument");
  DMM_CHECK_ARG(n_rows >= 0 && n_cols > 0 && ld >= n_cols, "dmm_eval_mask_scores: bad shape");
  if (n_rows == 0) return DMM_OK;
  eval_mask_kernel<<<(unsigned)dmm_ceil_div(n_rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(indptr, indices, row_ids, n_rows,
                                                                                             n_cols, scores, ld, fill);
  DMM_LAUNCH_CHECK();
  return DMM_OK;
}

extern "C" int dmm_eval_metrics(dmm_ctx* ctx, const float* scores, int64_t ld, int64_t n_rows, const int32_t* top_items,
                                int64_t K, const int64_t* row_ids, const int64_t* test_ptr, const int32_t* test_items,
                                const double* inv_log2, const double* max_dcg, double* out, void* stream) {
  DMM_CHECK_ARG(ctx && scores && top_items && row_ids && test_ptr && test_items && inv_log2 && max_dcg && out,
                "dmm_eval_metrics: null argument");
  DMM_CHECK_ARG(K >= 1 && K <= 32, "dmm_eval_metrics: top-K must be 1..32 (got %lld)", (long long)K);
  DMM_CHECK_ARG(n_rows >= 0, "dmm_eval_metrics: bad n_rows");
  if (n_rows == 0) return DMM_OK;
  eval_metrics_kernel<<<(unsigned)dmm_ceil_div(n_rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      scores, ld, n_rows, top_items, (int)K, row_ids, test_ptr, test_items, inv_log2, max_dcg, out);
  DMM_LAUNCH_CHECK();
  return DMM_OK;
}
