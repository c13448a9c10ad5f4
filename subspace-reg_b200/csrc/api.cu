// Library-wide C-ABI entry points: error string, version, device check.
#include "common.h"

namespace srb {
char* error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}
}  // namespace srb

extern "C" const char* sr_last_error(void) { return srb::error_buffer(); }

extern "C" int32_t sr_version(void) { return 100; /* 0.1.0 */ }

extern "C" int32_t sr_check_device(int32_t dev) {
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) return srb::fail(SR_E_CUDA, "cudaGetDeviceProperties(%d): %s", dev, cudaGetErrorString(e));
    if (prop.major != 10)
        return srb::fail(SR_E_DEVICE, "device %d is sm_%d%d; srb200 is built for sm_100a only (no fallback path)", dev,
                         prop.major, prop.minor);
    return SR_OK;
}
