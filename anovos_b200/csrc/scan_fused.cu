// K1+K2 instantiation: moments and private-counter histogram in one read.
// Two kernels, same results bit for bit: the register-staged loop (default) and the cp.async-staged loop (ANV_FUSED_STAGED=1:
// null-free columns go through a thread-private shared-memory ring).  Measured on B200 (profiles/r2b_fused_ab.md): the ring
// removes the long-scoreboard stalls (4.8 -> 0.9 per issue) and wins 2-9 % on null-free frames, but costs 8 % more
// instructions; at the north-star configuration (3 of 4 numeric columns carry a bitmap, the SM clock under the power cap) the
// register-staged loop is 1 % ahead, so it stays the default.
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
