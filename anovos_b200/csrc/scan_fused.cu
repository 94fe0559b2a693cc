// K1+K2 instantiation: moments and private-counter histogram in one read.
#include "scan_impl.cuh"
namespace anv {
int launch_fused(ScanParams& P, size_t smem, cudaStream_t st) { return launch_scan<true, 0, false>(P, smem, st); }
}  // namespace anv
