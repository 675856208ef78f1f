// Fused BPR and in-batch InfoNCE losses (forward + backward) with warp-shuffle reductions.
//
// Replaces Utils/Utils.py:57-98 (gathers, F.normalize, a B x B logits GEMM, log_softmax, diag, and
// the autograd graph behind them).  The B x B matrix is never written to HBM: every warp owns one
// anchor row, keeps its normalised vector in registers, streams the other view's normalised rows
// (B x D fp32 = 256 KB at B = 1024, L2/L1 resident) and folds them into an online log-sum-exp.
// Means are reduced in a fixed order (per-row scratch + single-CTA tree) so results are bitwise
// reproducible run to run.
#include "common.cuh"

namespace {

constexpr int MAX_D = 256;  // D <= 256: up to 8 elements per lane

__global__ void __launch_bounds__(1024) mean_reduce_kernel(const float* __restrict__ v, int64_t n, float scale,
                                                           float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 1024) s += v[i];
  s = dmm_warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = red[threadIdx.x];
    t = dmm_warp_sum(t);
    if (threadIdx.x == 0) *out = t * scale;
  }
}

// ------------------------------------------------------------------------------------------ BPR
// Utils/Utils.py:92-98: loss_b = -log(1e-5 + sigmoid(u.p - u.n)); d loss_b/dx = -s(1-s)/(1e-5+s)
__global__ void __launch_bounds__(256) bpr_kernel(const float* __restrict__ ue, int64_t ld_u,
                                                  const float* __restrict__ ie, int64_t ld_i,
                                                  const int64_t* __restrict__ users, const int64_t* __restrict__ pos,
                                                  const int64_t* __restrict__ neg, int64_t B, int D, float gscale,
                                                  float* __restrict__ row_loss, float* __restrict__ g_u,
                                                  float* __restrict__ g_p, float* __restrict__ g_n) {
  const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* u = ue + users[b] * ld_u;
  const float* p = ie + pos[b] * ld_i;
  const float* n = ie + neg[b] * ld_i;
  float dp = 0.f, dn = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float uv = __ldg(u + c);
    dp = fmaf(uv, __ldg(p + c), dp);
    dn = fmaf(uv, __ldg(n + c), dn);
  }
  dp = dmm_warp_sum(dp);
  dn = dmm_warp_sum(dn);
  const float x = dp - dn;
  const float s = 1.f / (1.f + expf(-x));
  if (lane == 0) row_loss[b] = -logf(10e-6f + s);
  if (g_u) {
    const float d = -(s * (1.f - s)) / (10e-6f + s) * gscale;  // gscale = upstream grad / B
    for (int c = lane; c < D; c += 32) {
      const float uv = __ldg(u + c), pv = __ldg(p + c), nv = __ldg(n + c);
      g_u[b * D + c] = d * (pv - nv);
      g_p[b * D + c] = d * uv;
      g_n[b * D + c] = -d * uv;
    }
  }
}

// ------------------------------------------------------------------------------------------ InfoNCE
// gathers rows idx[b] of both views, L2-normalises (eps 1e-12), writes n1/n2 [B, D] and inverse norms
__global__ void __launch_bounds__(256) nce_gather_norm_kernel(const float* __restrict__ v1, int64_t ld1,
                                                              const float* __restrict__ v2, int64_t ld2,
                                                              const int64_t* __restrict__ idx, int64_t B, int D,
                                                              float* __restrict__ n1, float* __restrict__ n2,
                                                              float* __restrict__ inv1, float* __restrict__ inv2) {
  const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* a = v1 + idx[b] * ld1;
  const float* c = v2 + idx[b] * ld2;
  float s1 = 0.f, s2 = 0.f;
  for (int k = lane; k < D; k += 32) {
    const float x = __ldg(a + k), y = __ldg(c + k);
    s1 = fmaf(x, x, s1);
    s2 = fmaf(y, y, s2);
  }
  s1 = dmm_warp_sum(s1);
  s2 = dmm_warp_sum(s2);
  const float i1 = 1.f / fmaxf(sqrtf(s1), 1e-12f), i2 = 1.f / fmaxf(sqrtf(s2), 1e-12f);
  for (int k = lane; k < D; k += 32) {
    n1[b * D + k] = __ldg(a + k) * i1;
    n2[b * D + k] = __ldg(c + k) * i2;
  }
  if (lane == 0 && inv1) {
    inv1[b] = i1;
    inv2[b] = i2;
  }
}

// One warp per anchor row i of `na` against all rows j of `nb` (both normalised [B, D]).
//   MODE 0 (forward):  lse_i = logsumexp_j(a_i.b_j / T);  row_loss_i = lse_i - a_i.b_i / T
//   MODE 1 (backward wrt a): ga_i = sum_j (p_ij - [i==j]) b_j * c,   p_ij = exp(a_i.b_j / T - lse_i)
//   MODE 2 (backward wrt b): gb_i = sum_j (p_ji - [i==j]) a_j * c,   p_ji = exp(a_j.b_i / T - lse_j)
//     (called with na/nb swapped: the anchor is b_i and the stream is a_j; lse is indexed by j)
// followed, for MODE 1/2, by the normalisation backward g_x = inv * (g - n (n.g)).
template <int MODE, int EPL /* elements per lane = D / 32 */>
__global__ void __launch_bounds__(256) nce_rows_kernel(const float* __restrict__ na, const float* __restrict__ nb,
                                                       int64_t B, float inv_temp, const float* __restrict__ lse_in,
                                                       const float* __restrict__ inv_norm, float coef,
                                                       float* __restrict__ out_lse, float* __restrict__ out_row_loss,
                                                       float* __restrict__ out_grad) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= B) return;
  constexpr int D = EPL * 32;
  float a[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) a[e] = __ldg(na + i * D + lane + 32 * e);
  float m = -INFINITY, l = 0.f, diag = 0.f;
  float g[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) g[e] = 0.f;
  const float lse_i = (MODE == 1) ? lse_in[i] : 0.f;

  for (int64_t j0 = 0; j0 < B; j0 += 4) {
    float bv[4][EPL], s[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int64_t j = j0 + q < B ? j0 + q : B - 1;
#pragma unroll
      for (int e = 0; e < EPL; ++e) bv[q][e] = __ldg(nb + j * D + lane + 32 * e);
      float d = 0.f;
#pragma unroll
      for (int e = 0; e < EPL; ++e) d = fmaf(a[e], bv[q][e], d);
      s[q] = d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int q = 0; q < 4; ++q) s[q] += __shfl_xor_sync(0xffffffffu, s[q], o);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int64_t j = j0 + q;
      if (j >= B) break;
      const float sc = s[q] * inv_temp;
      if (MODE == 0) {
        if (j == i) diag = sc;
        const float mn = fmaxf(m, sc);
        l = l * expf(m - mn) + expf(sc - mn);
        m = mn;
      } else {
        const float lse = (MODE == 1) ? lse_i : __ldg(lse_in + j);
        const float w = expf(sc - lse) - (j == i ? 1.f : 0.f);
#pragma unroll
        for (int e = 0; e < EPL; ++e) g[e] = fmaf(w, bv[q][e], g[e]);
      }
    }
  }
  if (MODE == 0) {
    if (lane == 0) {
      const float lse = m + logf(l);
      out_lse[i] = lse;
      out_row_loss[i] = lse - diag;
    }
  } else {
    // normalisation backward: x = raw row, n = x * inv;  g_x = inv * (g - n (n . g)) * coef
    float dot = 0.f;
#pragma unroll
    for (int e = 0; e < EPL; ++e) dot = fmaf(a[e], g[e], dot);
    dot = dmm_warp_sum(dot);
    const float sc = inv_norm[i] * coef;
#pragma unroll
    for (int e = 0; e < EPL; ++e) out_grad[i * D + lane + 32 * e] = sc * (g[e] - a[e] * dot);
  }
}

// ---- D == 64: register-tiled version ----------------------------------------------------------------
// The warp-per-anchor kernel above spends its time in shuffle reductions (one 5-step butterfly per logit).
// Here a CTA owns NCE_BM = 16 anchor rows and walks the other view in tiles of NCE_BN = 64 rows staged in
// shared memory; thread (r, c) accumulates the 4 logits (r, c + 16 q) with float4 reads along D (no
// reductions), the softmax statistics of a row are combined over its 16 threads once per TILE, and for the
// backward the 16 x 64 probability tile goes through shared memory into a second register-tiled product
// P . B (thread (r, d) owns 4 gradient components).  Same MODE semantics as nce_rows_kernel.
constexpr int NCE_BM = 16, NCE_BN = 64, NCE_LD = 68;   // 68-float rows: 16-byte aligned, conflict-free column access

template <int MODE>
__global__ void __launch_bounds__(256) nce_tile64_kernel(const float* __restrict__ na, const float* __restrict__ nb,
                                                         int64_t B, float inv_temp, const float* __restrict__ lse_in,
                                                         const float* __restrict__ inv_norm, float coef,
                                                         float* __restrict__ out_lse, float* __restrict__ out_row_loss,
                                                         float* __restrict__ out_grad) {
  __shared__ __align__(16) float As[NCE_BM][NCE_LD];
  __shared__ __align__(16) float Bs[NCE_BN][NCE_LD];
  __shared__ float Ps[NCE_BM][NCE_BN + 1];
  __shared__ float lse_s[NCE_BN];
  const int tid = threadIdx.x;
  const int r = tid >> 4, c4 = tid & 15;
  const int64_t i0 = (int64_t)blockIdx.x * NCE_BM;
  const int64_t i = i0 + r;
  const uint32_t hmask = 0xFFFFu << (tid & 16);   // the 16 threads of a row are one half warp
  {
    // anchor tile: 16 rows x 16 float4
    const int64_t row = i0 + (tid >> 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < B) v = __ldg(reinterpret_cast<const float4*>(na + row * 64) + (tid & 15));
    *reinterpret_cast<float4*>(&As[tid >> 4][4 * (tid & 15)]) = v;
  }
  float m = -INFINITY, l = 0.f, diag = -INFINITY;
  float g[4] = {0.f, 0.f, 0.f, 0.f};
  const float lse_i = (MODE == 1 && i < B) ? lse_in[i] : 0.f;

  for (int64_t j0 = 0; j0 < B; j0 += NCE_BN) {
    __syncthreads();   // previous tile fully consumed (also publishes As on the first pass)
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int jr = (tid >> 4) + 16 * it;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j0 + jr < B) v = __ldg(reinterpret_cast<const float4*>(nb + (j0 + jr) * 64) + (tid & 15));
      *reinterpret_cast<float4*>(&Bs[jr][4 * (tid & 15)]) = v;
    }
    if (MODE == 2 && tid < NCE_BN) lse_s[tid] = (j0 + tid < B) ? lse_in[j0 + tid] : 0.f;
    __syncthreads();
    // logits (r, c4 + 16 q)
    float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k4 = 0; k4 < 16; ++k4) {
      const float4 a = *reinterpret_cast<const float4*>(&As[r][4 * k4]);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 b = *reinterpret_cast<const float4*>(&Bs[c4 + 16 * q][4 * k4]);
        s[q] = fmaf(a.x, b.x, s[q]);
        s[q] = fmaf(a.y, b.y, s[q]);
        s[q] = fmaf(a.z, b.z, s[q]);
        s[q] = fmaf(a.w, b.w, s[q]);
      }
    }
    if (MODE == 0) {
      float tmax = -INFINITY;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int64_t j = j0 + c4 + 16 * q;
        s[q] = (j < B) ? s[q] * inv_temp : -INFINITY;
        if (j == i) diag = s[q];
        tmax = fmaxf(tmax, s[q]);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(hmask, tmax, o, 16));
      const float mn = fmaxf(m, tmax);          // finite: every tile holds at least one valid column
      float part = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) part += expf(s[q] - mn);   // exp(-inf) = 0 for masked columns
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) part += __shfl_xor_sync(hmask, part, o, 16);
      l = l * expf(m - mn) + part;
      m = mn;
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int c = c4 + 16 * q;
        const int64_t j = j0 + c;
        const float lse = (MODE == 1) ? lse_i : lse_s[c];
        float w = 0.f;
        if (j < B && i < B) w = expf(s[q] * inv_temp - lse) - (j == i ? 1.f : 0.f);
        Ps[r][c] = w;
      }
      __syncthreads();
      // g(r, 4 c4 .. 4 c4 + 3) += sum_j P(r, j) B(j, .)
#pragma unroll 8
      for (int j = 0; j < NCE_BN; ++j) {
        const float w = Ps[r][j];
        const float4 b = *reinterpret_cast<const float4*>(&Bs[j][4 * c4]);
        g[0] = fmaf(w, b.x, g[0]);
        g[1] = fmaf(w, b.y, g[1]);
        g[2] = fmaf(w, b.z, g[2]);
        g[3] = fmaf(w, b.w, g[3]);
      }
    }
  }
  if (MODE == 0) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) diag = fmaxf(diag, __shfl_xor_sync(hmask, diag, o, 16));
    if (c4 == 0 && i < B) {
      const float lse = m + logf(l);
      out_lse[i] = lse;
      out_row_loss[i] = lse - diag;
    }
  } else {
    // normalisation backward: g_x = inv * (g - n (n . g)) * coef
    const float4 a = *reinterpret_cast<const float4*>(&As[r][4 * c4]);
    float dot = a.x * g[0] + a.y * g[1] + a.z * g[2] + a.w * g[3];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) dot += __shfl_xor_sync(hmask, dot, o, 16);
    if (i < B) {
      const float sc = inv_norm[i] * coef;
      *reinterpret_cast<float4*>(out_grad + i * 64 + 4 * c4) =
          make_float4(sc * (g[0] - a.x * dot), sc * (g[1] - a.y * dot), sc * (g[2] - a.z * dot), sc * (g[3] - a.w * dot));
    }
  }
}

__global__ void __launch_bounds__(256) scatter_add_rows_kernel(const float* __restrict__ src, int64_t ld_s,
                                                               const int64_t* __restrict__ idx, int64_t B, int D,
                                                               float* __restrict__ dst, int64_t ld_d) {
  const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const int64_t r = idx[b];
  for (int c = lane; c < D; c += 32) atomicAdd(dst + r * ld_d + c, __ldg(src + b * ld_s + c));
}

template <int MODE>
int launch_nce(int64_t B, int64_t D, const float* na, const float* nb, float inv_temp, const float* lse_in,
               const float* inv_norm, float coef, float* out_lse, float* out_row_loss, float* out_grad,
               cudaStream_t st) {
  if (D == 64) {
    nce_tile64_kernel<MODE><<<(unsigned)dmm_ceil_div(B, NCE_BM), 256, 0, st>>>(na, nb, B, inv_temp, lse_in, inv_norm, coef, out_lse,
                                                                          out_row_loss, out_grad);
    DMM_LAUNCH_CHECK();
    return DMM_OK;
  }
  const unsigned grid = (unsigned)dmm_ceil_div(B * 32, 256);
  switch (D / 32) {
    case 1: nce_rows_kernel<MODE, 1><<<grid, 256, 0, st>>>(na, nb, B, inv_temp, lse_in, inv_norm, coef, out_lse, out_row_loss, out_grad); break;
    case 2: nce_rows_kernel<MODE, 2><<<grid, 256, 0, st>>>(na, nb, B, inv_temp, lse_in, inv_norm, coef, out_lse, out_row_loss, out_grad); break;
    case 4: nce_rows_kernel<MODE, 4><<<grid, 256, 0, st>>>(na, nb, B, inv_temp, lse_in, inv_norm, coef, out_lse, out_row_loss, out_grad); break;
    case 8: nce_rows_kernel<MODE, 8><<<grid, 256, 0, st>>>(na, nb, B, inv_temp, lse_in, inv_norm, coef, out_lse, out_row_loss, out_grad); break;
    default:
      dmm_set_error("InfoNCE: D must be 32, 64, 128 or 256 (got %lld)", (long long)D);
      return DMM_ERR_UNSUPPORTED;
  }
  DMM_LAUNCH_CHECK();
  return DMM_OK;
}

}  // namespace

extern "C" int dmm_bpr_fwd_bwd(dmm_ctx* ctx, const float* u_emb, int64_t ld_u, const float* i_emb, int64_t ld_i,
                               const int64_t* users, const int64_t* pos, const int64_t* neg, int64_t B, int64_t D,
                               float grad_scale, float* row_loss, float* loss, float* g_u, float* g_p, float* g_n,
                               void* stream) {
  DMM_CHECK_ARG(ctx && u_emb && i_emb && users && pos && neg && row_loss && loss, "dmm_bpr_fwd_bwd: null argument");
  DMM_CHECK_ARG(B > 0 && D > 0 && D <= MAX_D, "dmm_bpr_fwd_bwd: bad shape");
  DMM_CHECK_ARG((g_u && g_p && g_n) || (!g_u && !g_p && !g_n), "dmm_bpr_fwd_bwd: gradient outputs are all-or-none");
  cudaStream_t st = (cudaStream_t)stream;
  bpr_kernel<<<(unsigned)dmm_ceil_div(B * 32, 256), 256, 0, st>>>(u_emb, ld_u, i_emb, ld_i, users, pos, neg, B, (int)D,
                                                                 grad_scale / (float)B, row_loss, g_u, g_p, g_n);
  DMM_LAUNCH_CHECK();
  mean_reduce_kernel<<<1, 1024, 0, st>>>(row_loss, B, 1.f / (float)B, loss);
  DMM_LAUNCH_CHECK();
  return DMM_OK;
}

extern "C" int dmm_infonce_fwd(dmm_ctx* ctx, const float* v1, int64_t ld1, const float* v2, int64_t ld2,
                               const int64_t* idx, int64_t B, int64_t D, float temperature, float* workspace,
                               float* row_loss, float* loss, float* lse, float* inv1, float* inv2, void* stream) {
  DMM_CHECK_ARG(ctx && v1 && v2 && idx && workspace && row_loss && loss && lse && inv1 && inv2,
                "dmm_infonce_fwd: null argument");
  DMM_CHECK_ARG(B > 0 && temperature > 0.f, "dmm_infonce_fwd: bad B or temperature");
  DMM_CHECK_ARG(ld1 >= D && ld2 >= D, "dmm_infonce_fwd: leading dimension < D");
  cudaStream_t st = (cudaStream_t)stream;
  float* n1 = workspace;
  float* n2 = workspace + B * D;
  nce_gather_norm_kernel<<<(unsigned)dmm_ceil_div(B * 32, 256), 256, 0, st>>>(v1, ld1, v2, ld2, idx, B, (int)D, n1, n2, inv1, inv2);
  DMM_LAUNCH_CHECK();
  int rc = launch_nce<0>(B, D, n1, n2, 1.f / temperature, nullptr, nullptr, 0.f, lse, row_loss, nullptr, st);
  if (rc) return rc;
  mean_reduce_kernel<<<1, 1024, 0, st>>>(row_loss, B, 1.f / (float)B, loss);
  DMM_LAUNCH_CHECK();
  return DMM_OK;
}

extern "C" int dmm_infonce_bwd(dmm_ctx* ctx, const float* v1, int64_t ld1, const float* v2, int64_t ld2,
                               const int64_t* idx, int64_t B, int64_t D, float temperature, const float* lse,
                               const float* inv1, const float* inv2, float grad_scale, float* workspace, float* g1,
                               float* g2, void* stream) {
  DMM_CHECK_ARG(ctx && v1 && v2 && idx && lse && inv1 && inv2 && workspace && g1 && g2, "dmm_infonce_bwd: null argument");
  DMM_CHECK_ARG(B > 0 && temperature > 0.f, "dmm_infonce_bwd: bad B or temperature");
  cudaStream_t st = (cudaStream_t)stream;
  float* n1 = workspace;
  float* n2 = workspace + B * D;
  // recompute the normalised gathers (cheaper than keeping 2*B*D floats alive across the whole step)
  nce_gather_norm_kernel<<<(unsigned)dmm_ceil_div(B * 32, 256), 256, 0, st>>>(v1, ld1, v2, ld2, idx, B, (int)D, n1, n2,
                                                                              nullptr, nullptr);
  DMM_LAUNCH_CHECK();
  const float coef = grad_scale / ((float)B * temperature);
  int rc = launch_nce<1>(B, D, n1, n2, 1.f / temperature, lse, inv1, coef, nullptr, nullptr, g1, st);
  if (rc) return rc;
  return launch_nce<2>(B, D, n2, n1, 1.f / temperature, lse, inv2, coef, nullptr, nullptr, g2, st);
}

extern "C" int dmm_scatter_add_rows(dmm_ctx* ctx, const float* src, int64_t ld_s, const int64_t* idx, int64_t B,
                                    int64_t D, float* dst, int64_t ld_d, void* stream) {
  DMM_CHECK_ARG(ctx && src && idx && dst, "dmm_scatter_add_rows: null argument");
  DMM_CHECK_ARG(B >= 0 && D > 0 && ld_s >= D && ld_d >= D, "dmm_scatter_add_rows: bad shape");
  if (B == 0) return DMM_OK;
  scatter_add_rows_kernel<<<(unsigned)dmm_ceil_div(B * 32, 256), 256, 0, (cudaStream_t)stream>>>(src, ld_s, idx, B, (int)D, dst, ld_d);
  DMM_LAUNCH_CHECK();
  return DMM_OK;
}
