// Library-level entry points of the C ABI (error string, version, device check).
#include "common.cuh"
#include "opfmt.h"

namespace pdse {
char* error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}
int sm_count() {
    static std::atomic<int> cache[MAX_DEVICES];
    const int dev = current_device();
    int n = cache[dev].load(std::memory_order_relaxed);
    if (n <= 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cache[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}
}  // namespace pdse

extern "C" const char* pdse_last_error(void) { return pdse::error_buffer(); }

extern "C" int pdse_abi_version(void) { return 3; }
extern "C" int pdse_operand_format(void) { return PDSE_OP_FP16; }

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

// Kernel-side errors.  A persistent kernel cannot return a status, so it records the first failure in a caller-owned
// STICKY device block int32[8] = {code, detail0, detail1, count, timeout_us, ...} that no entry point ever clears: the caller copies
// it to the host at its own synchronisation point and passes the copy here.  0 when clean; otherwise negative with the
// message available from pdse_last_error().
extern "C" int pdse_status_check(const int* status_host) {
    using namespace pdse;
    if (!status_host || status_host[0] == 0) return PDSE_OK;
    if (status_host[0] == PDSE_STATUS_TCM_TIMEOUT)
        snprintf(error_buffer(), 512,
                 "pdse_tcm_flow: dependency wait timed out (launch %d, tile %d; %d waits failed): the output of this and "
                 "every later call sharing the status word is invalid",
                 status_host[1], status_host[2], status_host[3]);
    else
        snprintf(error_buffer(), 512, "pdse: kernel-side error code %d (%d, %d)", status_host[0], status_host[1], status_host[2]);
    return PDSE_EKERNEL;
}
