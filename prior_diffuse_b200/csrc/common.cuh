// Shared host-side plumbing for the C-ABI: thread-local error string, launch checks.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdlib>
#include <cstdint>
#include <cstdio>

namespace pdse {

enum : int { PDSE_OK = 0, PDSE_EINVAL = -1, PDSE_ECUDA = -2, PDSE_EKERNEL = -3 };
enum : int { PDSE_STATUS_TCM_TIMEOUT = 1 };   // codes a kernel writes into a caller-owned sticky status word

char* error_buffer();  // thread-local, 512 bytes (defined in api.cu)

inline int set_error(const char* msg) {
    snprintf(error_buffer(), 512, "%s", msg);
    return PDSE_EINVAL;
}
inline int set_cuda_error(const char* what, cudaError_t e) {
    snprintf(error_buffer(), 512, "%s: %s", what, cudaGetErrorString(e));
    return PDSE_ECUDA;
}
// Launch errors only (no sync: every entry point must be legal under stream capture).
inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(what, e);
    return PDSE_OK;
}

#define PDSE_CUDA(expr)                                              \
    do {                                                             \
        cudaError_t _e = (expr);                                     \
        if (_e != cudaSuccess) return ::pdse::set_cuda_error(#expr, _e); \
    } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Per-device state: kernel attributes (maximum dynamic shared memory, non-portable cluster size) and the SM count belong
// to a DEVICE, and entry points may be called from several host threads (one per stream / GPU), so every cache below is
// an array of atomics indexed by the current device.
constexpr int MAX_DEVICES = 64;
inline int current_device() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= MAX_DEVICES) d = 0;
    return d;
}
int sm_count();   // multiprocessors of the current device (api.cu)

struct SmemCache {
    std::atomic<int> hw[MAX_DEVICES];   // static storage: zero-initialised
};
// Opt in to > 48 KB dynamic shared memory.  The attribute is only (re)set when a larger size than ever before is
// requested on this device, so steady-state calls (and CUDA-graph capture after one warm-up) issue no non-stream API
// call.  Two threads racing here both set the attribute to a sufficient size: benign.
template <typename K>
inline int ensure_smem(K kernel, size_t bytes, SmemCache* cache) {
    if (bytes > 227 * 1024) return set_error("shared memory request exceeds 227 KB");
    std::atomic<int>& hw = cache->hw[current_device()];
    if ((int)bytes <= hw.load(std::memory_order_acquire)) return PDSE_OK;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute", e);
    int seen = hw.load(std::memory_order_relaxed);
    while (seen < (int)bytes && !hw.compare_exchange_weak(seen, (int)bytes, std::memory_order_release)) {
    }
    return PDSE_OK;
}

// Programmatic dependent launch.  A kernel launched through launch_pdl may start (block scheduling, barrier / tensor-memory
// set-up, weight loads -- everything that does not touch what earlier kernels of the stream wrote) while its predecessor
// is still draining; it must execute pdl_wait() on EVERY thread before its first access to such data and before its
// first global write.  pdl_wait returns once the predecessor grid has completed and its writes are visible (a no-op for
// an ordinary launch), so completion stays transitive along the stream.  pdl_trigger() lets the successor's blocks be
// scheduled as soon as this kernel's blocks have all started and room frees up (i.e. under this kernel's tail).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KA, typename... A>
inline cudaError_t launch_pdl(void (*kernel)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A&&... args) {
    static const bool off = getenv("PDSE_PDL_OFF") != nullptr;    // A/B switch: ordinary stream-ordered launches
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = off ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KA>(args)...);
}

}  // namespace pdse
