// K2 instantiations: histogram with private counters / CTA atomics / global atomics.
#include "scan_impl.cuh"
namespace anv {
int launch_hist(ScanParams& P, int path, size_t smem, cudaStream_t st) {
  if (path == 0) return launch_scan<false, 0, false>(P, smem, st);
  if (path == 1) return launch_scan<false, 1, false>(P, smem, st);
  return launch_scan<false, 2, false>(P, smem, st);
}
}  // namespace anv
