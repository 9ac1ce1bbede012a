// Shared-memory staged height-scan kernel (variant 1).  Placeholder until the staged path lands:
// it must fail loudly rather than silently run another kernel.
#include "scan_common.cuh"

namespace rover {
int launch_height_scan_staged(const float*, const float*, int, const float*, int, const ScanGridDev&, float, float,
                              float*, int, float*, cudaStream_t) {
    return fail("rover_height_scan: variant 1 (staged) is not built yet");
}
}  // namespace rover
