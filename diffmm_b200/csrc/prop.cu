// The element-wise glue of Model.gcn_MM (Model.py:60-134) between its SpMM products, forward and backward, as four
// kernels instead of ~45 ATen launches per joint-training step:
//
//   * F.normalize of the projected modality features (Model.py:89-93,104-105): y = x / max(||x||_2, eps) per row,
//     backward gx = (g - y (y . g)) / max(||x||, eps) (plain g / eps for a row clamped at eps);
//   * the modality mix (Model.py:116-119,125-127):  out = sum_m w_m (y + lam z_m)  with w = softmax(modal_weight) read
//     from the device, evaluated in the reference's order (aware_m = y + lam z_m, then the weighted terms added left to
//     right), backward  gy = sum_m w_m g,  gz_m = lam w_m g,  gw_m = sum (g . aware_m)  (per-CTA partial sums, added in a
//     fixed order by the caller: deterministic, no atomics).
// All four are HBM-bound streaming kernels over [rows, D] fp32 matrices (a few MB here: launch-latency bound in practice).
#include "common.cuh"

namespace {

constexpr int PROP_MAX_M = 4;      // modalities per mix (the reference has 2 or 3)

// one warp per row; D <= 1024 (any D, lanes stride the columns)
__global__ void __launch_bounds__(256) rownorm_fwd_kernel(const float* __restrict__ x, int64_t ld_x, int64_t n_rows, int D,
                                                          float eps, float* __restrict__ y, int64_t ld_y,
                                                          float* __restrict__ inv) {
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n_rows) return;
  const float* xr = x + r * ld_x;
  float ss = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float v = xr[c];
    ss = fmaf(v, v, ss);
  }
  ss = dmm_warp_sum(ss);
  const float nrm = sqrtf(ss);
  const bool clamped = !(nrm > eps);
  const float den = clamped ? eps : nrm;
  float* yr = y + r * ld_y;
  for (int c = lane; c < D; c += 32) yr[c] = xr[c] / den;       // the division of F.normalize, not a reciprocal multiply
  if (lane == 0) inv[r] = clamped ? -1.f / den : 1.f / den;      // sign = "clamped at eps" flag for the backward
}

__global__ void __launch_bounds__(256) rownorm_bwd_kernel(const float* __restrict__ y, int64_t ld_y, const float* __restrict__ inv,
                                                          const float* __restrict__ g, int64_t ld_g, int64_t n_rows, int D,
                                                          float* __restrict__ gx, int64_t ld_gx) {
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n_rows) return;
  const float* yr = y + r * ld_y;
  const float* gr = g + r * ld_g;
  const float iv = inv[r];
  float dot = 0.f;
  if (iv > 0.f) {
    for (int c = lane; c < D; c += 32) dot = fmaf(yr[c], gr[c], dot);
    dot = dmm_warp_sum(dot);
  }
  const float s = fabsf(iv);
  float* o = gx + r * ld_gx;
  for (int c = lane; c < D; c += 32) o[c] = (gr[c] - yr[c] * dot) * s;
}

struct MixPtrs {
  const float* z[PROP_MAX_M];
  float* gz[PROP_MAX_M];
};

// out = sum_m w_m (y + lam z_m); flat over n4 float4 elements (rows are dense: ld == D, D % 4 == 0)
__global__ void __launch_bounds__(256) modal_mix_fwd_kernel(const float4* __restrict__ y, MixPtrs p, const float* __restrict__ w,
                                                            int M, float lam, int64_t n4, float4* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 yv = y[i];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int m = 0; m < PROP_MAX_M; ++m) {
    if (m >= M) break;
    const float4 zv = reinterpret_cast<const float4*>(p.z[m])[i];
    const float wm = w[m];
    // aware = y + lam * z (separate multiply and add, like the eager expression), then w_m * aware
    const float4 aw = make_float4(__fadd_rn(yv.x, __fmul_rn(lam, zv.x)), __fadd_rn(yv.y, __fmul_rn(lam, zv.y)),
                                  __fadd_rn(yv.z, __fmul_rn(lam, zv.z)), __fadd_rn(yv.w, __fmul_rn(lam, zv.w)));
    const float4 t = make_float4(__fmul_rn(wm, aw.x), __fmul_rn(wm, aw.y), __fmul_rn(wm, aw.z), __fmul_rn(wm, aw.w));
    acc = m == 0 ? t : make_float4(__fadd_rn(acc.x, t.x), __fadd_rn(acc.y, t.y), __fadd_rn(acc.z, t.z), __fadd_rn(acc.w, t.w));
  }
  out[i] = acc;
}

// gy = (sum_m w_m) g evaluated as sum_m (w_m g), gz_m = lam (w_m g), partial[blockIdx.x, m] = sum over the CTA of g . aware_m
__global__ void __launch_bounds__(256) modal_mix_bwd_kernel(const float4* __restrict__ g, const float4* __restrict__ y, MixPtrs p,
                                                            const float* __restrict__ w, int M, float lam, int64_t n4,
                                                            float4* __restrict__ gy, float* __restrict__ partial) {
  __shared__ float red[PROP_MAX_M][8];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool in = i < n4;
  const float4 gv = in ? g[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 yv = in ? y[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 gacc = make_float4(0.f, 0.f, 0.f, 0.f);
  float dots[PROP_MAX_M];
#pragma unroll
  for (int m = 0; m < PROP_MAX_M; ++m) {
    dots[m] = 0.f;
    if (m < M) {
      const float wm = w[m];
      const float4 wg = make_float4(wm * gv.x, wm * gv.y, wm * gv.z, wm * gv.w);
      gacc = make_float4(gacc.x + wg.x, gacc.y + wg.y, gacc.z + wg.z, gacc.w + wg.w);
      if (in) {
        const float4 zv = reinterpret_cast<const float4*>(p.z[m])[i];
        if (p.gz[m]) reinterpret_cast<float4*>(p.gz[m])[i] = make_float4(lam * wg.x, lam * wg.y, lam * wg.z, lam * wg.w);
        dots[m] = gv.x * (yv.x + lam * zv.x) + gv.y * (yv.y + lam * zv.y) + gv.z * (yv.z + lam * zv.z) + gv.w * (yv.w + lam * zv.w);
      }
      dots[m] = dmm_warp_sum(dots[m]);
    }
  }
  if (in && gy) gy[i] = gacc;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) {
#pragma unroll
    for (int m = 0; m < PROP_MAX_M; ++m) red[m][wid] = dots[m];
  }
  __syncthreads();
  if (threadIdx.x < M) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[threadIdx.x][k];
    partial[(int64_t)blockIdx.x * M + threadIdx.x] = s;
  }
}

}  // namespace

extern "C" int dmm_rownorm_fwd(dmm_ctx* ctx, const float* x, int64_t ld_x, int64_t n_rows, int64_t D, float eps, float* y,
                               int64_t ld_y, float* inv, void* stream) {
  DMM_CHECK_ARG(ctx && x && y && inv, "dmm_rownorm_fwd: null argument");
  DMM_CHECK_ARG(n_rows >= 0 && D > 0 && D < (1 << 20) && ld_x >= D && ld_y >= D && eps > 0.f, "dmm_rownorm_fwd: bad shape");
  if (n_rows == 0) return DMM_OK;
  rownorm_fwd_kernel<<<(unsigned)dmm_ceil_div(n_rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(x, ld_x, n_rows, (int)D, eps, y,
                                                                                               ld_y, inv);
  DMM_LAUNCH_CHECK();
  return DMM_OK;
}

extern "C" int dmm_rownorm_bwd(dmm_ctx* ctx, const float* y, int64_t ld_y, const float* inv, const float* g, int64_t ld_g,
                               int64_t n_rows, int64_t D, float* gx, int64_t ld_gx, void* stream) {
  DMM_CHECK_ARG(ctx && y && inv && g && gx, "dmm_rownorm_bwd: null argument");
  DMM_CHECK_ARG(n_rows >= 0 && D > 0 && D < (1 << 20) && ld_y >= D && ld_g >= D && ld_gx >= D, "dmm_rownorm_bwd: bad shape");
  if (n_rows == 0) return DMM_OK;
  rownorm_bwd_kernel<<<(unsigned)dmm_ceil_div(n_rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(y, ld_y, inv, g, ld_g, n_rows,
                                                                                               (int)D, gx, ld_gx);
  DMM_LAUNCH_CHECK();
  return DMM_OK;
}

extern "C" int dmm_modal_mix_fwd(dmm_ctx* ctx, const float* y, const float* const* z, const float* w, int32_t n_modal, float lam,
                                 int64_t n_elems, float* out, void* stream) {
  DMM_CHECK_ARG(ctx && y && z && w && out, "dmm_modal_mix_fwd: null argument");
  DMM_CHECK_ARG(n_modal >= 1 && n_modal <= PROP_MAX_M, "dmm_modal_mix_fwd: 1 .. %d modalities (got %d)", PROP_MAX_M, n_modal);
  DMM_CHECK_ARG(n_elems >= 0 && n_elems % 4 == 0, "dmm_modal_mix_fwd: element count must be a multiple of 4");
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  MixPtrs p{};
  bool aligned = al16(y) && al16(out);
  for (int m = 0; m < n_modal; ++m) {
    DMM_CHECK_ARG(z[m] != nullptr, "dmm_modal_mix_fwd: null z[%d]", m);
    p.z[m] = z[m];
    aligned = aligned && al16(z[m]);
  }
  DMM_CHECK_ARG(aligned, "dmm_modal_mix_fwd: buffers must be 16-byte aligned");
  if (n_elems == 0) return DMM_OK;
  const int64_t n4 = n_elems / 4;
  modal_mix_fwd_kernel<<<(unsigned)dmm_ceil_div(n4, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(y), p, w, n_modal, lam, n4, reinterpret_cast<float4*>(out));
  DMM_LAUNCH_CHECK();
  return DMM_OK;
}

extern "C" int64_t dmm_modal_mix_partial_rows(int64_t n_elems) { return dmm_ceil_div(n_elems / 4, 256); }

extern "C" int dmm_modal_mix_bwd(dmm_ctx* ctx, const float* g, const float* y, const float* const* z, const float* w,
                                 int32_t n_modal, float lam, int64_t n_elems, float* gy, float* const* gz, float* partial,
                                 void* stream) {
  DMM_CHECK_ARG(ctx && g && y && z && w && gz && partial, "dmm_modal_mix_bwd: null argument");
  DMM_CHECK_ARG(n_modal >= 1 && n_modal <= PROP_MAX_M, "dmm_modal_mix_bwd: 1 .. %d modalities (got %d)", PROP_MAX_M, n_modal);
  DMM_CHECK_ARG(n_elems >= 0 && n_elems % 4 == 0, "dmm_modal_mix_bwd: element count must be a multiple of 4");
  auto al16 = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  MixPtrs p{};
  bool aligned = al16(g) && al16(y) && al16(gy);
  for (int m = 0; m < n_modal; ++m) {
    DMM_CHECK_ARG(z[m] != nullptr, "dmm_modal_mix_bwd: null z[%d]", m);
    p.z[m] = z[m];
    p.gz[m] = gz[m];
    aligned = aligned && al16(z[m]) && al16(gz[m]);
  }
  DMM_CHECK_ARG(aligned, "dmm_modal_mix_bwd: buffers must be 16-byte aligned");
  if (n_elems == 0) return DMM_OK;
  const int64_t n4 = n_elems / 4;
  modal_mix_bwd_kernel<<<(unsigned)dmm_ceil_div(n4, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(g), reinterpret_cast<const float4*>(y), p, w, n_modal, lam, n4,
      reinterpret_cast<float4*>(gy), partial);
  DMM_LAUNCH_CHECK();
  return DMM_OK;
}
