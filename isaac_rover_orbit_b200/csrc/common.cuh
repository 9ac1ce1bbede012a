// Shared helpers of the rover_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
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

}  // namespace rover
