// Context, error reporting and version of the C ABI.
#include "common.cuh"

#include <string.h>

namespace {
thread_local char g_err[512] = "";
}

void dmm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int dmm_version(void) { return 100; }  // 0.1.0

extern "C" const char* dmm_last_error(void) { return g_err; }

extern "C" int dmm_init(int device, dmm_ctx** out) {
  DMM_CHECK_ARG(out, "dmm_init: null out pointer");
  *out = nullptr;
  int n = 0;
  DMM_CUDA(cudaGetDeviceCount(&n));
  DMM_CHECK_ARG(device >= 0 && device < n, "dmm_init: device %d out of range (%d visible)", device, n);
  DMM_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  DMM_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    dmm_set_error("dmm_init: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major,
                  prop.minor);
    return DMM_ERR_UNSUPPORTED;
  }
  dmm_ctx* c = new dmm_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  c->encode_tiled = nullptr;
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
    delete c;
    dmm_set_error("dmm_init: cuTensorMapEncodeTiled not available from the driver (%s)", cudaGetErrorString(e));
    return DMM_ERR_CUDA;
  }
  c->encode_tiled = fn;
  *out = c;
  return DMM_OK;
}

extern "C" void dmm_destroy(dmm_ctx* ctx) { delete ctx; }

extern "C" int dmm_num_sms(const dmm_ctx* ctx) { return ctx ? ctx->num_sms : 0; }
