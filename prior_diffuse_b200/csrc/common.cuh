// Shared host-side plumbing for the C-ABI: thread-local error string, launch checks.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

namespace pdse {

enum : int { PDSE_OK = 0, PDSE_EINVAL = -1, PDSE_ECUDA = -2 };

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

// Opt in to > 48 KB dynamic shared memory.  The attribute is only (re)set when a larger size than ever
// before is requested, so steady-state calls (and CUDA-graph capture after one warm-up) issue no
// non-stream API call.
template <typename K>
inline int ensure_smem(K kernel, size_t bytes, int* high_water) {
    if (bytes > 227 * 1024) return set_error("shared memory request exceeds 227 KB");
    if ((int)bytes <= *high_water) return PDSE_OK;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute", e);
    *high_water = (int)bytes;
    return PDSE_OK;
}

}  // namespace pdse
