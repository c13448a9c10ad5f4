// Host-side helpers shared by the C-ABI translation units: thread-local error string, CUDA error mapping.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include "../../include/srb200.h"

namespace srb {

char* error_buffer();  // thread-local, 512 bytes (defined in api.cu)

inline int32_t fail(int32_t code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(error_buffer(), 512, fmt, ap);
    va_end(ap);
    return code;
}

#define SR_CUDA_OK(expr)                                                                         \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess)                                                                  \
            return ::srb::fail(SR_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                               __FILE__, __LINE__);                                              \
    } while (0)

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// One-time set-up PER DEVICE (function attributes belong to a device's context, SM counts to a device): `state` is a
// static object at the call site, `f` runs once for each device the calling code is used on.
constexpr int kMaxDevices = 64;
struct PerDeviceOnce {
    std::mutex m;
    bool done[kMaxDevices] = {};
};
template <class F>
inline void once_per_device(PerDeviceOnce& state, F f) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) {
        f();
        return;
    }
    std::lock_guard<std::mutex> lock(state.m);
    if (!state.done[dev]) {
        f();
        state.done[dev] = true;
    }
}

inline int current_device_sms() {
    static int sms[kMaxDevices] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 148;
    if (!sms[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) sms[dev] = n;
    }
    return sms[dev] ? sms[dev] : 148;
}

}  // namespace srb
