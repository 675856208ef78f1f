// CSR SpMM for the LightGCN-style propagation over the normalised user-item graph.
//
// Replaces torch.sparse.mm on an uncoalesced COO (Model.py:90,93,105,111,114,123,130; Main.py:319),
// which coalesces (sorts) the operand on every call before cuSPARSE runs.  The adjacency is
// symmetric, so the backward pass is the same call on the incoming gradient.
//
// Layout: one warp per output row.  For D = 64 a row of X is 256 B = 16 float4: the two half warps
// gather two neighbours at a time with 128-bit loads (fully coalesced 256 B segments), 4 neighbours
// per half warp in flight; rows longer than LONG_ROW are finished by the whole CTA with a fixed-order
// shared-memory reduction (deterministic, no atomics).  HBM-bound: 8*nnz + 8*(N+1) + 2*N*D*4 bytes
// per product when X is not L2 resident.
#include "common.cuh"

namespace {

constexpr int SPMM_THREADS = 256;
constexpr int SPMM_WARPS = SPMM_THREADS / 32;
constexpr int LONG_ROW = 1024;

struct Epi {
  float alpha, beta;
  const float* z;
  int64_t ld_z;
};

// ---- D == 64 fast path --------------------------------------------------------------------
__device__ __forceinline__ float4 fma4(float a, const float4& x, const float4& acc) {
  return make_float4(fmaf(a, x.x, acc.x), fmaf(a, x.y, acc.y), fmaf(a, x.z, acc.z), fmaf(a, x.w, acc.w));
}
__device__ __forceinline__ float4 add4(const float4& a, const float4& b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// accumulates A[r, beg:end) . X over neighbours j = beg + h, beg + h + stride, ... for this half warp
__device__ __forceinline__ float4 gather64(const int32_t* __restrict__ idx, const float* __restrict__ val,
                                           const float* __restrict__ x, int64_t ld_x, int64_t beg, int64_t end,
                                           int stride, int l16) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int64_t j = beg;
  for (; j + 3 * stride < end; j += 4 * stride) {
    const int32_t c0 = __ldg(idx + j), c1 = __ldg(idx + j + stride), c2 = __ldg(idx + j + 2 * stride),
                  c3 = __ldg(idx + j + 3 * stride);
    const float v0 = __ldg(val + j), v1 = __ldg(val + j + stride), v2 = __ldg(val + j + 2 * stride),
                v3 = __ldg(val + j + 3 * stride);
    const float4 x0 = __ldg(reinterpret_cast<const float4*>(x + (int64_t)c0 * ld_x) + l16);
    const float4 x1 = __ldg(reinterpret_cast<const float4*>(x + (int64_t)c1 * ld_x) + l16);
    const float4 x2 = __ldg(reinterpret_cast<const float4*>(x + (int64_t)c2 * ld_x) + l16);
    const float4 x3 = __ldg(reinterpret_cast<const float4*>(x + (int64_t)c3 * ld_x) + l16);
    acc = fma4(v0, x0, acc);
    acc = fma4(v1, x1, acc);
    acc = fma4(v2, x2, acc);
    acc = fma4(v3, x3, acc);
  }
  for (; j < end; j += stride) {
    const int32_t c = __ldg(idx + j);
    const float v = __ldg(val + j);
    acc = fma4(v, __ldg(reinterpret_cast<const float4*>(x + (int64_t)c * ld_x) + l16), acc);
  }
  return acc;
}

__global__ void __launch_bounds__(SPMM_THREADS) spmm64_kernel(const int64_t* __restrict__ ptr,
                                                              const int32_t* __restrict__ idx,
                                                              const float* __restrict__ val, int64_t row0, int64_t row1,
                                                              const float* __restrict__ x, int64_t ld_x, Epi ep,
                                                              float* __restrict__ y, int64_t ld_y) {
  __shared__ int64_t long_rows[SPMM_WARPS];
  __shared__ int n_long;
  __shared__ float4 part[SPMM_WARPS * 2][16];
  if (threadIdx.x == 0) n_long = 0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = lane >> 4, l16 = lane & 15;
  const int64_t r = row0 + (int64_t)blockIdx.x * SPMM_WARPS + warp;
  if (r < row1) {
    const int64_t b = ptr[r], e = ptr[r + 1];
    if (e - b > LONG_ROW) {
      if (lane == 0) long_rows[atomicAdd(&n_long, 1)] = r;
    } else {
      float4 acc = gather64(idx, val, x, ld_x, b + half, e, 2, l16);
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);
      acc.z += __shfl_xor_sync(0xffffffffu, acc.z, 16);
      acc.w += __shfl_xor_sync(0xffffffffu, acc.w, 16);
      if (half == 0) {
        float4 o = make_float4(ep.alpha * acc.x, ep.alpha * acc.y, ep.alpha * acc.z, ep.alpha * acc.w);
        if (ep.z) {
          const float4 zz = __ldg(reinterpret_cast<const float4*>(ep.z + r * ep.ld_z) + l16);
          o = make_float4(fmaf(ep.beta, zz.x, o.x), fmaf(ep.beta, zz.y, o.y), fmaf(ep.beta, zz.z, o.z),
                          fmaf(ep.beta, zz.w, o.w));
        }
        reinterpret_cast<float4*>(y + r * ld_y)[l16] = o;
      }
    }
  }
  __syncthreads();
  const int nl = n_long;
  // long rows: all 16 half warps stride the neighbour list; fixed-order reduction
  for (int q = 0; q < nl; ++q) {
    // sort order of long_rows is irrelevant for determinism: each row's result is independent
    const int64_t rr = long_rows[q];
    const int64_t b = ptr[rr], e = ptr[rr + 1];
    const int hw = warp * 2 + half;
    part[hw][l16] = gather64(idx, val, x, ld_x, b + hw, e, SPMM_WARPS * 2, l16);
    __syncthreads();
    if (threadIdx.x < 16) {
      float4 acc = part[0][threadIdx.x];
      for (int k = 1; k < SPMM_WARPS * 2; ++k) acc = add4(acc, part[k][threadIdx.x]);
      float4 o = make_float4(ep.alpha * acc.x, ep.alpha * acc.y, ep.alpha * acc.z, ep.alpha * acc.w);
      if (ep.z) {
        const float4 zz = __ldg(reinterpret_cast<const float4*>(ep.z + rr * ep.ld_z) + threadIdx.x);
        o = make_float4(fmaf(ep.beta, zz.x, o.x), fmaf(ep.beta, zz.y, o.y), fmaf(ep.beta, zz.z, o.z),
                        fmaf(ep.beta, zz.w, o.w));
      }
      reinterpret_cast<float4*>(y + rr * ld_y)[threadIdx.x] = o;
    }
    __syncthreads();
  }
}

// ---- generic D (multiple of 4, <= 256): one warp per row, lanes stride the float4 columns ----------
__global__ void __launch_bounds__(SPMM_THREADS) spmm_generic_kernel(const int64_t* __restrict__ ptr,
                                                                    const int32_t* __restrict__ idx,
                                                                    const float* __restrict__ val, int64_t row0,
                                                                    int64_t row1, const float* __restrict__ x,
                                                                    int64_t ld_x, int D4, Epi ep, float* __restrict__ y,
                                                                    int64_t ld_y) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = row0 + (int64_t)blockIdx.x * SPMM_WARPS + warp;
  if (r >= row1) return;
  const int64_t b = ptr[r], e = ptr[r + 1];
  float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;  // columns lane and lane + 32 (D4 <= 64)
  for (int64_t j = b; j < e; ++j) {
    const int32_t c = __ldg(idx + j);
    const float v = __ldg(val + j);
    const float4* xr = reinterpret_cast<const float4*>(x + (int64_t)c * ld_x);
    if (lane < D4) acc0 = fma4(v, __ldg(xr + lane), acc0);
    if (lane + 32 < D4) acc1 = fma4(v, __ldg(xr + lane + 32), acc1);
  }
  for (int k = 0; k < 2; ++k) {
    const int col = lane + 32 * k;
    if (col >= D4) break;
    const float4 a = k ? acc1 : acc0;
    float4 o = make_float4(ep.alpha * a.x, ep.alpha * a.y, ep.alpha * a.z, ep.alpha * a.w);
    if (ep.z) {
      const float4 zz = __ldg(reinterpret_cast<const float4*>(ep.z + r * ep.ld_z) + col);
      o = make_float4(fmaf(ep.beta, zz.x, o.x), fmaf(ep.beta, zz.y, o.y), fmaf(ep.beta, zz.z, o.z), fmaf(ep.beta, zz.w, o.w));
    }
    reinterpret_cast<float4*>(y + r * ld_y)[col] = o;
  }
}

// ---- cross-layer CL perturbation (Main.py:320-321), one warp per row --------------------------------
__global__ void __launch_bounds__(256) sign_noise_kernel(float* __restrict__ e, int64_t ld_e,
                                                         const float* __restrict__ rnd, int64_t ld_r, int64_t n_rows,
                                                         int D, float noise_degree) {
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n_rows) return;
  float ss = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float v = __ldg(rnd + r * ld_r + c);
    ss = fmaf(v, v, ss);
  }
  ss = dmm_warp_sum(ss);
  const float inv = noise_degree / fmaxf(sqrtf(ss), 1e-12f);
  for (int c = lane; c < D; c += 32) {
    const float v = e[r * ld_e + c];
    const float sg = (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f);
    e[r * ld_e + c] = v + sg * __ldg(rnd + r * ld_r + c) * inv;
  }
}

}  // namespace

extern "C" int dmm_spmm_csr(dmm_ctx* ctx, const int64_t* adj_ptr, const int32_t* adj_idx, const float* adj_val,
                            int64_t row0, int64_t row1, const float* x, int64_t ld_x, int64_t D, float alpha,
                            float beta, const float* z, int64_t ld_z, float* y, int64_t ld_y, void* stream) {
  DMM_CHECK_ARG(ctx && adj_ptr && adj_idx && adj_val && x && y, "dmm_spmm_csr: null argument");
  DMM_CHECK_ARG(D > 0 && D % 4 == 0 && D <= 256, "dmm_spmm_csr: D must be a multiple of 4 and <= 256 (got %lld)", (long long)D);
  DMM_CHECK_ARG(ld_x % 4 == 0 && ld_y % 4 == 0 && ld_x >= D && ld_y >= D && (!z || (ld_z % 4 == 0 && ld_z >= D)),
                "dmm_spmm_csr: leading dimensions must be >= D and multiples of 4");
  auto al16 = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  DMM_CHECK_ARG(al16(x) && al16(y) && al16(z), "dmm_spmm_csr: X/Y/Z must be 16-byte aligned");
  DMM_CHECK_ARG(row0 >= 0 && row1 >= row0, "dmm_spmm_csr: bad row range");
  if (row1 == row0) return DMM_OK;
  const Epi ep{alpha, z ? beta : 0.f, z, ld_z};
  const unsigned grid = (unsigned)dmm_ceil_div(row1 - row0, SPMM_WARPS);
  if (D == 64) {
    spmm64_kernel<<<grid, SPMM_THREADS, 0, (cudaStream_t)stream>>>(adj_ptr, adj_idx, adj_val, row0, row1, x, ld_x, ep, y, ld_y);
  } else {
    spmm_generic_kernel<<<grid, SPMM_THREADS, 0, (cudaStream_t)stream>>>(adj_ptr, adj_idx, adj_val, row0, row1, x, ld_x,
                                                                         (int)(D / 4), ep, y, ld_y);
  }
  DMM_LAUNCH_CHECK();
  return DMM_OK;
}

extern "C" int dmm_sign_noise_(dmm_ctx* ctx, float* e, int64_t ld_e, const float* rnd, int64_t ld_r, int64_t n_rows,
                               int64_t D, float noise_degree, void* stream) {
  DMM_CHECK_ARG(ctx && e && rnd, "dmm_sign_noise_: null argument");
  DMM_CHECK_ARG(D > 0 && ld_e >= D && ld_r >= D, "dmm_sign_noise_: bad shape");
  if (n_rows <= 0) return DMM_OK;
  sign_noise_kernel<<<(unsigned)dmm_ceil_div(n_rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(e, ld_e, rnd, ld_r, n_rows,
                                                                                              (int)D, noise_degree);
  DMM_LAUNCH_CHECK();
  return DMM_OK;
}
