// bin-id materialisation instantiation (attribute_binning's returned frame).
#include "scan_impl.cuh"
namespace anv {
int launch_assign(ScanParams& P, size_t smem, cudaStream_t st) { return launch_scan<false, -1, true>(P, smem, st); }
}  // namespace anv
