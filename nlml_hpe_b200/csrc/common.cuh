// Shared host-side plumbing for the C ABI: error reporting, device guard, pinned staging.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/nlml_hpe_b200.h"

namespace nlml {

inline char* last_error_buf() {
    static thread_local char buf[512] = "";
    return buf;
}
inline int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

#define NLML_CUDA(expr)                                                                          \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess)                                                                   \
            return nlml::set_error((int)_e, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                   __FILE__, __LINE__);                                          \
    } while (0)

// Select `device` for the scope, restore the previous one on exit.
struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) return;
        ok = cudaSetDevice(device) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

inline int check_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return set_error(NLML_E_NO_DEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= n) return set_error(NLML_E_INVALID, "device %d out of range [0,%d)", device, n);
    cudaDeviceProp prop;
    NLML_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return set_error(NLML_E_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                         prop.major, prop.minor);
    return 0;
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace nlml
