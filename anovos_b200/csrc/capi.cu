// Library-level entry points and error plumbing of libanovos_b200.so.
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

namespace anv {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return ANV_ERR_CUDA;
}
}  // namespace anv

extern "C" int anv_version(void) { return ANV_VERSION; }
#ifndef ANV_SOURCE_HASH
#define ANV_SOURCE_HASH "unknown"
#endif
// sha1 over the .cu / .cuh sources + include/anovos_b200.h this binary was built from (anovos_b200/build.py): the Python
// binding refuses a stale .so whose struct layouts or workspace sizing may have drifted from the sources next to it.
extern "C" const char* anv_source_hash(void) { return ANV_SOURCE_HASH; }
extern "C" const char* anv_last_error(void) { return anv::g_err; }

extern "C" int anv_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem) {
  int dev = 0;
  ANV_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  ANV_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (total_mem) *total_mem = p.totalGlobalMem;
  return ANV_OK;
}
