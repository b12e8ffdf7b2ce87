// Library-level entry points of the C ABI (error string, version, device check).
#include "common.cuh"

namespace pdse {
char* error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}
}  // namespace pdse

extern "C" const char* pdse_last_error(void) { return pdse::error_buffer(); }

extern "C" int pdse_abi_version(void) { return 1; }

// 0 when the current device is sm_100 (B200); negative otherwise.
extern "C" int pdse_check_device(void) {
    using namespace pdse;
    int dev = 0, major = 0;
    PDSE_CUDA(cudaGetDevice(&dev));
    PDSE_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));   // cheap: called on every op
    if (major != 10) return set_error("pdse: this library is built for sm_100a (B200) only");
    return PDSE_OK;
}

extern "C" int pdse_sm_count(void) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    return n;
}
