// Version / error reporting of the rover_b200 C-ABI library.
#include "common.cuh"

namespace rover {
char* error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}
}  // namespace rover

extern "C" int rover_abi_version(void) { return ROVER_B200_ABI_VERSION; }
extern "C" const char* rover_last_error(void) { return rover::error_buffer(); }
