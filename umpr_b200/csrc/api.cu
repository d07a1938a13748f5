#include <stdlib.h>
// Error reporting and version for the C-ABI (include/umpr_b200.h).
#include <stdarg.h>
#include <stdio.h>
#include "common.cuh"
#include "../../include/umpr_b200.h"

namespace umpr {
int dbg_flags() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("UMPR_DBG"); v = e ? atoi(e) : 0; }
  return v;
}

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int fail_arg(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return UMPR_ERR_ARG;
}
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}
}  // namespace umpr

extern "C" int umpr_version(void) { return UMPR_B200_VERSION; }
extern "C" const char* umpr_last_error(void) { return umpr::g_err; }
extern "C" int umpr_sm_count(int device, int* out) {
  int n = 0;
  cudaError_t e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
  if (e != cudaSuccess) { umpr::set_error("sm_count: %s", cudaGetErrorString(e)); return (int)e; }
  *out = n;
  return 0;
}
