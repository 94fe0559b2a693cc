// K1+K2 instantiation: moments and private-counter histogram in one read.
// Two kernels: the cp.async-staged loop (default) and the register-staged loop (ANV_FUSED_STAGED=0), same results bit for bit.
#include <stdlib.h>
#include "scan_impl.cuh"
namespace anv {
int launch_fused(ScanParams& P, size_t smem, cudaStream_t st) {
  const char* e = getenv("ANV_FUSED_STAGED");   // read per call (a test flips it between two calls of one process)
  const int staged = e ? atoi(e) : ANV_FUSED_STAGED_DEFAULT;
  // the ring needs ST_D * ST_CH * 4 KB behind the counters: fall back to register staging when it does not fit one CTA
  if (staged && smem + STAGE_BYTES + 16 <= 200 * 1024) return launch_scan<true, 0, false, true>(P, smem, st);
  return launch_scan<true, 0, false, false>(P, smem, st);
}
}  // namespace anv
