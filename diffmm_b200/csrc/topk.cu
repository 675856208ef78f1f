// Per-user variable-k top-k -> edge list (rebuild of the modality-aware user-item graph).
//
// Replaces the Python loop of Main.py:224-230 (one torch.topk launch per user and one
// int(tensor) host sync per emitted edge).  One CTA per user row:
//   1. the fp32 score row is read ONCE from HBM (coalesced) into shared memory as order-preserving
//      uint32 keys (rows that do not fit stay in global/L2 and are re-read per pass);
//   2. MSB-first radix select (4 passes x 8 bits, warp-aggregated shared-memory histograms) finds
//      the exact k-th largest key T and how many ties at T must be taken;
//   3. an ordered block scan emits the column indices of {key > T} U {first ties at T} in ascending
//      column order straight into the CSR slot out_ptr[r] .. out_ptr[r+1] (k_r = deg(u), so the
//      output offsets are the train CSR indptr: no atomics, no host sync, deterministic).
// Tie-break: value descending, then column ascending (-0.0 == +0.0).  HBM-bound: 4*I bytes read
// and 8*k bytes written per user and modality.
#include "common.cuh"

namespace {

constexpr int TOPK_THREADS = 256;

__device__ __forceinline__ uint32_t order_key(float f) {
  uint32_t u = __float_as_uint(f);
  if (u == 0x80000000u) u = 0u;  // -0.0 ties with +0.0
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// block-wide exclusive scan of one int per thread (256 threads); returns exclusive prefix, total in *total
__device__ __forceinline__ int block_excl_scan(int v, int* warp_sums, int* total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_sums[w] = incl;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < TOPK_THREADS / 32; ++i) {
    const int s = warp_sums[i];
    if (i < w) base += s;
    tot += s;
  }
  __syncthreads();
  *total = tot;
  return base + incl - v;
}

template <bool IN_SMEM>
__global__ void __launch_bounds__(TOPK_THREADS) topk_edges_kernel(const float* __restrict__ scores, int64_t ld,
                                                                  int64_t n_rows, int n_cols,
                                                                  const int64_t* __restrict__ out_ptr, int64_t row_base,
                                                                  int32_t* __restrict__ out_users,
                                                                  int32_t* __restrict__ out_items,
                                                                  int32_t* __restrict__ status) {
  extern __shared__ uint32_t s_keys[];  // n_cols keys when IN_SMEM
  __shared__ int hist[256];
  __shared__ int warp_sums[TOPK_THREADS / 32];
  __shared__ uint32_t s_prefix;
  __shared__ int s_kk;

  const int64_t r = blockIdx.x;
  if (r >= n_rows) return;
  const int64_t o0 = out_ptr[r], o1 = out_ptr[r + 1];
  int k = (int)(o1 - o0);
  if (k <= 0) return;
  if (k > n_cols) {
    if (status && threadIdx.x == 0) atomicExch(status, 1);
    k = n_cols;
  }
  const float* row = scores + r * ld;
  const int tid = threadIdx.x;

  if (IN_SMEM) {
    // single HBM read of the row; float4 when the row start is 16 B aligned
    const bool vec = ((reinterpret_cast<uintptr_t>(row) & 15u) == 0);
    if (vec) {
      const int n4 = n_cols >> 2;
      const float4* row4 = reinterpret_cast<const float4*>(row);
      for (int i = tid; i < n4; i += TOPK_THREADS) {
        const float4 v = __ldcs(row4 + i);
        s_keys[4 * i + 0] = order_key(v.x);
        s_keys[4 * i + 1] = order_key(v.y);
        s_keys[4 * i + 2] = order_key(v.z);
        s_keys[4 * i + 3] = order_key(v.w);
      }
      for (int i = (n4 << 2) + tid; i < n_cols; i += TOPK_THREADS) s_keys[i] = order_key(__ldcs(row + i));
    } else {
      for (int i = tid; i < n_cols; i += TOPK_THREADS) s_keys[i] = order_key(__ldcs(row + i));
    }
  }
  if (tid == 0) {
    s_prefix = 0u;
    s_kk = k;
  }
  __syncthreads();

  auto key_at = [&](int i) -> uint32_t { return IN_SMEM ? s_keys[i] : order_key(__ldg(row + i)); };

  uint32_t prefix = 0u, mask = 0u;
  int kk = k;
  if (k < n_cols) {
#pragma unroll 1
    for (int shift = 24; shift >= 0; shift -= 8) {
      hist[tid] = 0;
      __syncthreads();
      for (int i0 = 0; i0 < n_cols; i0 += TOPK_THREADS) {
        const int i = i0 + tid;
        const bool in = i < n_cols;
        const uint32_t key = in ? key_at(i) : 0u;
        const bool cand = in && ((key & mask) == prefix);
        const uint32_t digit = (key >> shift) & 0xFFu;
        // warp-aggregated histogram: one shared-memory atomic per distinct digit per warp
        const uint32_t active = __ballot_sync(0xffffffffu, cand);
        if (cand) {
          const uint32_t peers = __match_any_sync(active, digit);
          if ((int)(__ffs(peers) - 1) == (tid & 31)) atomicAdd(&hist[digit], __popc(peers));
        }
      }
      __syncthreads();
      if (tid < 32) {
        // descending scan: lane l owns bins 255-8l .. 248-8l
        int loc[8], sum = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          loc[j] = hist[255 - 8 * tid - j];
          sum += loc[j];
        }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, o);
          if (tid >= o) incl += t;
        }
        const int excl = incl - sum;
        const int want = kk;  // == s_kk, kept in a register by every thread
        if (excl < want && want <= incl) {
          int cum = excl;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (cum < want && want <= cum + loc[j]) {
              s_prefix = prefix | ((uint32_t)(255 - 8 * tid - j) << shift);
              s_kk = want - cum;
            }
            cum += loc[j];
          }
        }
      }
      __syncthreads();
      prefix = s_prefix;
      kk = s_kk;
      mask |= 0xFFu << shift;
    }
  }
  // threshold key T = prefix; take every key > T and the first kk keys == T (ascending column)
  const uint32_t T = (k < n_cols) ? prefix : 0u;
  const int need_eq = (k < n_cols) ? kk : 0;
  const bool take_all = !(k < n_cols);

  // contiguous segment per thread, odd length => conflict-free strided shared-memory reads
  int seg = (n_cols + TOPK_THREADS - 1) / TOPK_THREADS;
  seg |= 1;
  const int b = min(tid * seg, n_cols), e = min(b + seg, n_cols);
  int c_gt = 0, c_eq = 0;
  for (int i = b; i < e; ++i) {
    const uint32_t key = key_at(i);
    c_gt += (take_all || key > T) ? 1 : 0;
    c_eq += (!take_all && key == T) ? 1 : 0;
  }
  int tot;
  const int eq_before = block_excl_scan(c_eq, warp_sums, &tot);
  int eq_take = need_eq - eq_before;
  eq_take = eq_take < 0 ? 0 : (eq_take > c_eq ? c_eq : eq_take);
  const int pos0 = block_excl_scan(c_gt + eq_take, warp_sums, &tot);
  int64_t w = o0 + pos0;
  const int32_t user = (int32_t)(row_base + r);
  for (int i = b; i < e; ++i) {
    const uint32_t key = key_at(i);
    bool sel = take_all || key > T;
    if (!sel && key == T && eq_take > 0) {
      sel = true;
      --eq_take;
    }
    if (sel) {
      out_items[w] = i;
      if (out_users) out_users[w] = user;
      ++w;
    }
  }
}

}  // namespace

extern "C" int dmm_topk_edges(dmm_ctx* ctx, const float* scores, int64_t ld, int64_t n_rows, int64_t n_cols,
                              const int64_t* out_ptr, int64_t row_base, int32_t* out_users, int32_t* out_items,
                              int32_t* status, void* stream) {
  DMM_CHECK_ARG(ctx && scores && out_ptr && out_items, "dmm_topk_edges: null argument");
  DMM_CHECK_ARG(n_cols > 0 && n_cols < (1LL << 31) && ld >= n_cols, "dmm_topk_edges: bad shape n_cols=%lld ld=%lld",
                (long long)n_cols, (long long)ld);
  DMM_CHECK_ARG(n_rows >= 0 && n_rows < (1LL << 31), "dmm_topk_edges: bad n_rows");
  if (n_rows == 0) return DMM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)n_cols * sizeof(uint32_t);
  const size_t cap = (size_t)ctx->max_smem_optin > 8192 ? (size_t)ctx->max_smem_optin - 4096 : 0;
  if (smem <= cap) {
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
      DMM_CUDA(cudaFuncSetAttribute(topk_edges_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap));
      configured = cap;
    }
    topk_edges_kernel<true><<<(unsigned)n_rows, TOPK_THREADS, smem, st>>>(scores, ld, n_rows, (int)n_cols, out_ptr, row_base,
                                                                        out_users, out_items, status);
  } else {
    topk_edges_kernel<false><<<(unsigned)n_rows, TOPK_THREADS, 0, st>>>(scores, ld, n_rows, (int)n_cols, out_ptr, row_base,
                                                                      out_users, out_items, status);
  }
  DMM_LAUNCH_CHECK();
  return DMM_OK;
}
