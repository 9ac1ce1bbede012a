// Shared helpers of the rover_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/rover_b200.h"

namespace rover {

char* error_buffer();  // thread-local, defined in abi.cu

inline int fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(error_buffer(), 512, fmt, ap);
    va_end(ap);
    return 1;
}

#define ROVER_CHECK(cond, ...)                        \
    do {                                              \
        if (!(cond)) return ::rover::fail(__VA_ARGS__); \
    } while (0)

#define ROVER_CUDA(call)                                                                         \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess) return ::rover::fail("%s: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail("%s launch failed: %s", what, cudaGetErrorString(e));
    return 0;
}

// ---- programmatic dependent launch (PDL).  A kernel launched with launch_overlapped() may begin -- CTAs resident,
// prologue running -- while the kernel in front of it on the stream is still finishing; it must execute
// grid_dependency_wait() before it touches anything that kernel writes (the wait returns when the predecessor has
// completed and its memory is visible; immediately if there is none).  The predecessor may call
// grid_dependency_trigger() to let the dependents start before it exits (otherwise they start when it does).
// What a kernel may read BEFORE the wait is launch-invariant data only (ray pattern, lattice lines, tensor maps, packed
// network weights): data that was complete before the kernel in front of it on the stream was launched.  The Python layer
// guarantees that by synchronising the stream after it builds such data (ScanGridHandle, policy._pack); a C-ABI caller
// that rebuilds tables or weights must not let the kernel that writes them be the one directly in front of a step kernel.
// Works under stream capture (a programmatic edge in the graph).  ROVER_PDL=0 in the environment turns it off.
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dependency_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("ROVER_PDL");
        return !(e != nullptr && e[0] == '0');
    }();
    return on;
}

template <class... Params, class... Args>
inline cudaError_t launch_overlapped(void (*kernel)(Params...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<Params>(args)...);
}

}  // namespace rover
