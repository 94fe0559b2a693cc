// K1 instantiation: moments only.
#include "scan_impl.cuh"
namespace anv {
int launch_mom(ScanParams& P, cudaStream_t st) { return launch_scan<true, -1, false>(P, 0, st); }
}  // namespace anv
