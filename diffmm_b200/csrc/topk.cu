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
constexpr int TOPK_NB_LOG2 = 11;
constexpr int TOPK_NB = 1 << TOPK_NB_LOG2;   // range-adapted buckets of the first pass
constexpr int TOPK_CAND = 1024;              // threshold-bucket keys selected exactly in shared memory

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

// MSB-first radix select (4 passes x 8 bits, warp-aggregated shared-memory histograms) of the
// `want`-th largest of `count` keys read through `key_at`.  Returns the threshold key and, in
// *need_eq, how many keys equal to it belong to the top `want`.  Uniform across the block.
template <typename KeyAt>
__device__ __forceinline__ uint32_t radix_select(KeyAt key_at, int count, int want, int* hist, uint32_t* s_prefix,
                                                 int* s_kk, int* need_eq) {
  const int tid = threadIdx.x;
  uint32_t prefix = 0u, mask = 0u;
  int kk = want;
#pragma unroll 1
  for (int shift = 24; shift >= 0; shift -= 8) {
    hist[tid] = 0;
    __syncthreads();
    for (int i0 = 0; i0 < count; i0 += TOPK_THREADS) {
      const int i = i0 + tid;
      const bool in = i < count;
      const uint32_t key = in ? key_at(i) : 0u;
      const bool cand = in && ((key & mask) == prefix);
      const uint32_t digit = (key >> shift) & 0xFFu;
      // one shared-memory atomic per distinct digit per warp
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
      if (excl < kk && kk <= incl) {
        int cum = excl;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (cum < kk && kk <= cum + loc[j]) {
            *s_prefix = prefix | ((uint32_t)(255 - 8 * tid - j) << shift);
            *s_kk = kk - cum;
          }
          cum += loc[j];
        }
      }
    }
    __syncthreads();
    prefix = *s_prefix;
    kk = *s_kk;
    mask |= 0xFFu << shift;
  }
  *need_eq = kk;
  return prefix;
}

// Selection strategy.  The leading radix digits of fp32 scores (sign + exponent) barely discriminate, so
// a plain 4-pass radix select walks the whole row four times with almost every key a candidate.
// Instead ONE histogram pass over range-adapted buckets ((key - kmin) >> sh, NB buckets, monotone in
// the key) isolates the bucket holding the k-th largest key; only that bucket's keys (a handful for
// continuous scores) go through the exact radix select in shared memory.  Rows whose threshold bucket
// is crowded (heavy ties) fall back to the exact select over the whole row.  Either way the result is
// the exact (value desc, column asc) top-k.
template <bool IN_SMEM>
__global__ void __launch_bounds__(TOPK_THREADS) topk_edges_kernel(const float* __restrict__ scores, int64_t ld,
                                                                  int64_t n_rows, int n_cols,
                                                                  const int64_t* __restrict__ out_ptr, int64_t row_base,
                                                                  int32_t* __restrict__ out_users,
                                                                  int32_t* __restrict__ out_items,
                                                                  int32_t* __restrict__ status) {
  extern __shared__ uint32_t s_keys[];  // n_cols keys when IN_SMEM
  __shared__ int bucket[TOPK_NB];
  __shared__ uint32_t cand[TOPK_CAND];
  __shared__ int hist[256];
  __shared__ int warp_sums[TOPK_THREADS / 32];
  __shared__ uint32_t s_red[2 * (TOPK_THREADS / 32)];
  __shared__ uint32_t s_prefix;
  __shared__ int s_kk, s_bin, s_above, s_ncand;

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
  const bool take_all = !(k < n_cols);

  uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;
  auto track = [&](uint32_t key) {
    kmin = min(kmin, key);
    kmax = max(kmax, key);
  };
  if (IN_SMEM) {
    // single HBM read of the row; float4 when the row start is 16 B aligned
    const bool vec = ((reinterpret_cast<uintptr_t>(row) & 15u) == 0);
    if (vec) {
      const int n4 = n_cols >> 2;
      const float4* row4 = reinterpret_cast<const float4*>(row);
      for (int i = tid; i < n4; i += TOPK_THREADS) {
        const float4 v = __ldcs(row4 + i);
        const uint4 q = make_uint4(order_key(v.x), order_key(v.y), order_key(v.z), order_key(v.w));
        reinterpret_cast<uint4*>(s_keys)[i] = q;
        track(q.x); track(q.y); track(q.z); track(q.w);
      }
      for (int i = (n4 << 2) + tid; i < n_cols; i += TOPK_THREADS) {
        const uint32_t q = order_key(__ldcs(row + i));
        s_keys[i] = q;
        track(q);
      }
    } else {
      for (int i = tid; i < n_cols; i += TOPK_THREADS) {
        const uint32_t q = order_key(__ldcs(row + i));
        s_keys[i] = q;
        track(q);
      }
    }
  } else if (!take_all) {
    for (int i = tid; i < n_cols; i += TOPK_THREADS) track(order_key(__ldg(row + i)));
  }
  auto key_at = [&](int i) -> uint32_t { return IN_SMEM ? s_keys[i] : order_key(__ldg(row + i)); };

  uint32_t T = 0u;
  int need_eq = 0;
  if (!take_all && k <= TOPK_THREADS) {
    // Small k (almost every user): the k-th largest of the per-thread maxima is a lower bound L of the k-th
    // largest key (the k largest maxima are k distinct keys >= L).  Keys >= L are compacted (a few more than
    // k for continuous scores) and selected exactly in shared memory; no per-key atomics anywhere.
    uint32_t* const tmax = reinterpret_cast<uint32_t*>(bucket);
    tmax[tid] = kmax;   // threads that own no key publish 0, the smallest key
    if (tid == 0) s_ncand = 0;
    __syncthreads();
    int dummy;
    const uint32_t L = radix_select([&](int i) -> uint32_t { return tmax[i]; }, TOPK_THREADS, k, hist, &s_prefix, &s_kk,
                                    &dummy);
    for (int i0 = 0; i0 < n_cols; i0 += TOPK_THREADS) {
      const int i = i0 + tid;
      const uint32_t key = i < n_cols ? key_at(i) : 0u;
      const bool c = i < n_cols && key >= L;
      const uint32_t bal = __ballot_sync(0xffffffffu, c);
      if (bal) {
        int base = 0;
        if ((tid & 31) == 0) base = atomicAdd(&s_ncand, __popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        const int pos = base + __popc(bal & ((1u << (tid & 31)) - 1u));
        if (c && pos < TOPK_CAND) cand[pos] = key;
      }
    }
    __syncthreads();
    const int m = s_ncand;   // >= k
    if (m <= TOPK_CAND) {
      T = radix_select([&](int i) -> uint32_t { return cand[i]; }, m, k, hist, &s_prefix, &s_kk, &need_eq);
    } else {
      T = radix_select(key_at, n_cols, k, hist, &s_prefix, &s_kk, &need_eq);
    }
  } else if (!take_all) {
#pragma unroll
    for (int j = 0; j < TOPK_NB / TOPK_THREADS; ++j) bucket[tid + j * TOPK_THREADS] = 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
      kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    }
    if ((tid & 31) == 0) {
      s_red[tid >> 5] = kmin;
      s_red[TOPK_THREADS / 32 + (tid >> 5)] = kmax;
    }
    if (tid == 0) s_ncand = 0;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < TOPK_THREADS / 32; ++i) {
      kmin = min(kmin, s_red[i]);
      kmax = max(kmax, s_red[TOPK_THREADS / 32 + i]);
    }
    // (kmax - kmin) >> sh < TOPK_NB
    const uint32_t range = kmax - kmin;
    const int bits = 32 - __clz(range);               // range == 0 -> 0
    const int sh = bits > TOPK_NB_LOG2 ? bits - TOPK_NB_LOG2 : 0;
    for (int i = tid; i < n_cols; i += TOPK_THREADS) atomicAdd(&bucket[(key_at(i) - kmin) >> sh], 1);
    __syncthreads();
    {
      // descending scan over the buckets: thread t owns buckets NB-1-PER*t .. NB-PER*(t+1)
      constexpr int PER = TOPK_NB / TOPK_THREADS;
      int loc[PER], sum = 0;
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        loc[j] = bucket[TOPK_NB - 1 - PER * tid - j];
        sum += loc[j];
      }
      int tot;
      const int excl = block_excl_scan(sum, warp_sums, &tot);
      if (excl < k && k <= excl + sum) {
        int cum = excl;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
          if (cum < k && k <= cum + loc[j]) {
            s_bin = TOPK_NB - 1 - PER * tid - j;
            s_above = cum;
          }
          cum += loc[j];
        }
      }
    }
    __syncthreads();
    const int bin = s_bin;
    const int want = k - s_above;          // >= 1: rank of the threshold inside its bucket
    const int m = bucket[bin];
    if (m <= TOPK_CAND) {
      for (int i = tid; i < n_cols; i += TOPK_THREADS) {
        const uint32_t key = key_at(i);
        if ((int)((key - kmin) >> sh) == bin) cand[atomicAdd(&s_ncand, 1)] = key;
      }
      __syncthreads();
      T = radix_select([&](int i) -> uint32_t { return cand[i]; }, m, want, hist, &s_prefix, &s_kk, &need_eq);
    } else {
      T = radix_select(key_at, n_cols, k, hist, &s_prefix, &s_kk, &need_eq);
    }
  }

  // take every key > T and the first need_eq keys == T (ascending column)
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
  if (c_gt + eq_take == 0) return;        // no barrier follows
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
  // static shared memory of the kernel (buckets, candidates, scratch) is ~14 KB
  const size_t cap = (size_t)ctx->max_smem_optin > 32768 ? (size_t)ctx->max_smem_optin - 16384 : 0;
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
