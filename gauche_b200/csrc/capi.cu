// capi.cu — error plumbing and device queries of the C ABI (include/gauche_b200.h).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace gk {

static thread_local char g_err[512] = "";

char* err_buf() { return g_err; }

int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_err(cudaError_t e, const char* where) {
  if (e == cudaSuccess) return 0;
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorString(e), cudaGetErrorName(e));
  return (int)e;
}

int sm_count() {
  // per-device cache; one process per GPU is the deployment model but stay correct otherwise
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cache[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)
      v = 148;
    cache[dev] = v;
  }
  return cache[dev];
}

}  // namespace gk

extern "C" const char* gk_last_error(void) { return gk::err_buf(); }

extern "C" int gk_query(int what, int64_t* out) {
  using namespace gk;
  GK_CHECK_ARG(out != nullptr, GK_EINVAL, "gk_query: null out");
  switch (what) {
    case GK_Q_ABI_VERSION: *out = GK_ABI_VERSION; return 0;
    case GK_Q_SM_COUNT: *out = sm_count(); return 0;
    case GK_Q_CC: {
      int dev = 0, major = 0, minor = 0;
      GK_CUDA(cudaGetDevice(&dev));
      GK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
      GK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
      *out = major * 10 + minor;
      return 0;
    }
    case GK_Q_UMMA_TILE_M: *out = 128; return 0;
    case GK_Q_UMMA_TILE_N: *out = 256; return 0;
    case GK_Q_K_ALIGN: *out = 128; return 0;
    default: return set_err(GK_EINVAL, "gk_query: unknown selector %d", what);
  }
}
