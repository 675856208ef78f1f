// One-launch Adam step over all tensors of an optimiser (Main.py:92-110,189-192,375-377).
//
// torch's capturable foreach Adam walks the parameters fourteen times per step (lerp, mul, addcmul, sqrt into a new tensor,
// div, add, div, addcdiv, each a read-modify-write of full-size tensors) plus a dozen launches on the one-element step
// tensors: at the Denoise shapes (2 x 7.2 M weights per modality) that is ~1.5 GB of HBM traffic per modality and batch.
// This kernel reads p, g, m, v once and writes p, m, v once -- 28 bytes per weight -- and evaluates EXACTLY the operation
// sequence of torch.optim.adam._multi_tensor_adam (capturable branch, weight_decay = 0, amsgrad = False, maximize = False)
// in fp32, operation for operation (same roundings, same fused multiply-adds the ATen kernels contract to), so the update is
// bit-identical to the stock optimiser's (tests/test_optim_gpu.py: 40 steps, eager and under CUDA-graph replay).
//   m   = m + w1 (g - m)                         w1 = float(1 - beta1)            (_foreach_lerp_)
//   v   = v * beta2;  v = v + w2 (g g)           w2 = float(1 - beta2)            (_foreach_mul_, _foreach_addcmul_)
//   ss  = 1 / ((beta1^t - 1) / lr)               = -lr / (1 - beta1^t)            (_foreach_pow, sub_, div_, reciprocal_)
//   c2  = sqrt(-(beta2^t - 1))                                                    (_foreach_pow, sub_, neg_, sqrt_)
//   den = ((sqrt(v) / c2) + eps) / ss                                             (_foreach_sqrt, div_, add_, div_)
//   p   = p + m / den                                                             (_foreach_addcdiv_)
// t (fp32, already incremented by the caller) and lr (fp32) are DEVICE scalars: capturable in a CUDA graph.
#include "common.cuh"

namespace {

constexpr int ADAM_MAX_TENSORS = 24;
constexpr int ADAM_CHUNK = 256 * 4 * 4;      // elements per CTA: 256 threads x 4 float4

struct AdamArgs {
  float* p[ADAM_MAX_TENSORS];
  const float* g[ADAM_MAX_TENSORS];
  float* m[ADAM_MAX_TENSORS];
  float* v[ADAM_MAX_TENSORS];
  int64_t n[ADAM_MAX_TENSORS];
  int block0[ADAM_MAX_TENSORS + 1];          // first CTA of every tensor
  int n_tensors;
};

__device__ __forceinline__ void adam_elem(float& p, const float g, float& m, float& v, const float w1, const float beta2,
                                          const float w2, const float c2, const float eps, const float ss) {
  m = fmaf(w1, __fsub_rn(g, m), m);                 // lerp, |weight| < 0.5: self + weight * (end - self)
  v = __fmul_rn(v, beta2);
  v = fmaf(w2, __fmul_rn(g, g), v);                 // foreach addcmul: self + value * (t1 * t2)
  float den = __fsqrt_rn(v);
  den = __fdiv_rn(den, c2);
  den = __fadd_rn(den, eps);
  den = __fdiv_rn(den, ss);
  p = __fadd_rn(p, __fdiv_rn(m, den));              // addcdiv, value = 1
}

// torch's NON-capturable foreach sequence (eager trainer: python-float lr, step counters on the host): the bias corrections
// are python doubles computed on the host and enter the kernels as fp32 scalars,
//   den = (sqrt(v) / float(c2)) + eps;   p = p + float(ss) * (m / den)          (_foreach_div_, add_, addcdiv_(.., step_size))
__device__ __forceinline__ void adam_elem_host(float& p, const float g, float& m, float& v, const float w1, const float beta2,
                                               const float w2, const float c2, const float eps, const float ss) {
  m = fmaf(w1, __fsub_rn(g, m), m);
  v = __fmul_rn(v, beta2);
  v = fmaf(w2, __fmul_rn(g, g), v);
  float den = __fsqrt_rn(v);
  den = __fdiv_rn(den, c2);
  den = __fadd_rn(den, eps);
  p = fmaf(ss, __fdiv_rn(m, den), p);
}

template <bool HOST>
__global__ void __launch_bounds__(256) adam_step_kernel(const AdamArgs a, const float* __restrict__ step, const float* __restrict__ lr,
                                                        const float beta1, const float beta2, const float w1, const float w2,
                                                        const float eps, const float host_ss, const float host_c2) {
  // which tensor: the block-start table is tiny and block-uniform
  int t = 0;
#pragma unroll 1
  while (t + 1 < a.n_tensors && (int)blockIdx.x >= a.block0[t + 1]) ++t;
  float ss, c2;
  if (HOST) {
    ss = host_ss, c2 = host_c2;                       // float(-lr / (1 - beta1^t)), float(sqrt(1 - beta2^t)) from python doubles
  } else {
    const float st = __ldg(step), l = __ldg(lr);
    // the scalar chain of the capturable branch, in its order of operations (all fp32 like the one-element tensors)
    float bc1 = __fsub_rn(powf(beta1, st), 1.0f);
    float bc2 = __fsub_rn(powf(beta2, st), 1.0f);
    bc2 = -bc2;
    bc1 = __fdiv_rn(bc1, l);
    ss = __fdiv_rn(1.0f, bc1);                        // reciprocal_
    c2 = __fsqrt_rn(bc2);
  }
  float* __restrict__ p = a.p[t];
  const float* __restrict__ g = a.g[t];
  float* __restrict__ m = a.m[t];
  float* __restrict__ v = a.v[t];
  const int64_t n = a.n[t];
  const int64_t base = (int64_t)((int)blockIdx.x - a.block0[t]) * ADAM_CHUNK;
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15u) == 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t i = base + ((int64_t)k * 256 + threadIdx.x) * 4;
    if (i >= n) break;
    if (vec && i + 4 <= n) {
      float4 pv = *reinterpret_cast<float4*>(p + i), mv = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
      const float4 gv = __ldcs(reinterpret_cast<const float4*>(g + i));
      if (HOST) {
        adam_elem_host(pv.x, gv.x, mv.x, vv.x, w1, beta2, w2, c2, eps, ss);
        adam_elem_host(pv.y, gv.y, mv.y, vv.y, w1, beta2, w2, c2, eps, ss);
        adam_elem_host(pv.z, gv.z, mv.z, vv.z, w1, beta2, w2, c2, eps, ss);
        adam_elem_host(pv.w, gv.w, mv.w, vv.w, w1, beta2, w2, c2, eps, ss);
      } else {
        adam_elem(pv.x, gv.x, mv.x, vv.x, w1, beta2, w2, c2, eps, ss);
        adam_elem(pv.y, gv.y, mv.y, vv.y, w1, beta2, w2, c2, eps, ss);
        adam_elem(pv.z, gv.z, mv.z, vv.z, w1, beta2, w2, c2, eps, ss);
        adam_elem(pv.w, gv.w, mv.w, vv.w, w1, beta2, w2, c2, eps, ss);
      }
      *reinterpret_cast<float4*>(p + i) = pv;
      *reinterpret_cast<float4*>(m + i) = mv;
      *reinterpret_cast<float4*>(v + i) = vv;
    } else {
      for (int64_t j = i; j < n && j < i + 4; ++j) {
        float pv = p[j], mv = m[j], vv = v[j];
        if (HOST) adam_elem_host(pv, g[j], mv, vv, w1, beta2, w2, c2, eps, ss);
        else adam_elem(pv, g[j], mv, vv, w1, beta2, w2, c2, eps, ss);
        p[j] = pv, m[j] = mv, v[j] = vv;
      }
    }
  }
}

}  // namespace

static int adam_launch(dmm_ctx* ctx, const char* who, int32_t n_tensors, float* const* params, const float* const* grads,
                       float* const* exp_avg, float* const* exp_avg_sq, const int64_t* numel, const float* step, const float* lr,
                       bool host, double host_ss, double host_c2, double beta1, double beta2, double eps, void* stream) {
  DMM_CHECK_ARG(ctx && params && grads && exp_avg && exp_avg_sq && numel && (host || (step && lr)), "%s: null argument", who);
  DMM_CHECK_ARG(n_tensors >= 0, "%s: bad tensor count", who);
  cudaStream_t st = (cudaStream_t)stream;
  const float w1 = (float)(1.0 - beta1), w2 = (float)(1.0 - beta2);
  DMM_CHECK_ARG(w1 > 0.f && w1 < 0.5f, "%s: beta1 must lie in (0.5, 1) (the lerp branch of |weight| < 0.5)", who);
  int t = 0;
  while (t < n_tensors) {
    AdamArgs a{};
    int nt = 0;
    int64_t blocks = 0;
    for (; t < n_tensors && nt < ADAM_MAX_TENSORS; ++t) {
      DMM_CHECK_ARG(params[t] && grads[t] && exp_avg[t] && exp_avg_sq[t] && numel[t] >= 0, "%s: null tensor %d", who, t);
      if (numel[t] == 0) continue;
      a.p[nt] = params[t], a.g[nt] = grads[t], a.m[nt] = exp_avg[t], a.v[nt] = exp_avg_sq[t], a.n[nt] = numel[t];
      a.block0[nt] = (int)blocks;
      blocks += dmm_ceil_div(numel[t], ADAM_CHUNK);
      DMM_CHECK_ARG(blocks < (1LL << 31), "%s: too many elements", who);
      ++nt;
    }
    if (nt == 0) continue;
    a.block0[nt] = (int)blocks;
    a.n_tensors = nt;
    if (host)
      adam_step_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(a, nullptr, nullptr, (float)beta1, (float)beta2, w1, w2, (float)eps,
                                                              (float)host_ss, (float)host_c2);
    else
      adam_step_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(a, step, lr, (float)beta1, (float)beta2, w1, w2, (float)eps, 0.f, 0.f);
    DMM_LAUNCH_CHECK();
  }
  return DMM_OK;
}

extern "C" int dmm_adam_step(dmm_ctx* ctx, int32_t n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                             float* const* exp_avg_sq, const int64_t* numel, const float* step, const float* lr, double beta1,
                             double beta2, double eps, void* stream) {
  return adam_launch(ctx, "dmm_adam_step", n_tensors, params, grads, exp_avg, exp_avg_sq, numel, step, lr, false, 0.0, 0.0, beta1,
                     beta2, eps, stream);
}

extern "C" int dmm_adam_step_host(dmm_ctx* ctx, int32_t n_tensors, float* const* params, const float* const* grads,
                                  float* const* exp_avg, float* const* exp_avg_sq, const int64_t* numel, double step_size,
                                  double bias_correction2_sqrt, double beta1, double beta2, double eps, void* stream) {
  return adam_launch(ctx, "dmm_adam_step_host", n_tensors, params, grads, exp_avg, exp_avg_sq, numel, nullptr, nullptr, true, step_size,
                     bias_correction2_sqrt, beta1, beta2, eps, stream);
}
