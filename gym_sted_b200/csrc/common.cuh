// Shared host-side plumbing of the C ABI (error text, CUDA call checking).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>

#include "../../include/sted_b200.h"

extern thread_local char sted_err_buf[512];

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(sted_err_buf, sizeof(sted_err_buf), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess)                                                              \
            return fail(_e == cudaErrorNoDevice || _e == cudaErrorInsufficientDriver        \
                            ? STED_ENODEVICE : STED_ECUDA,                                  \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,   \
                        __LINE__);                                                          \
    } while (0)

inline int check_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(STED_ENODEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    return STED_OK;
}
