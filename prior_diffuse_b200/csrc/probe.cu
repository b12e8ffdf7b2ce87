// pdse_probe_gemm: smallest possible exercise of the tcgen05 path used everywhere
// else (bulk copy -> CP8 planes in smem -> descriptor with a row shift -> UMMA ->
// TMEM -> registers).  tests/test_gpu_probe.py checks it against a float matmul for
// several row shifts; if this is wrong nothing GEMM-shaped in the library is right.
#include "common.cuh"
#include "umma.cuh"

namespace pdse {

// A: [K/8][a_rows][8] bf16, B: [K/8][N][8] bf16, D: [128][N] fp32 (row-major)
// D[m][n] = sum_k A[m + row_shift][k] * B[n][k]
__global__ void __launch_bounds__(128, 1)
probe_gemm_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ Bm, float* __restrict__ D,
                  int a_rows, int N, int K, int row_shift, int swap_lbo_sbo) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar_ld, bar_mma;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x;
    const int kc_n = K / 8;
    uint8_t* sA = smem;
    uint8_t* sB = smem + (size_t)kc_n * a_rows * 16;
    const uint32_t a_bytes = kc_n * a_rows * 16, b_bytes = kc_n * N * 16;

    if (tid == 0) {
        mbar_init(&bar_ld, 1);
        mbar_init(&bar_mma, 1);
        fence_mbar_init();
    }
    __syncwarp();
    if (tid < 32) tmem_alloc(&tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (tid == 0) {
        mbar_arrive_expect_tx(&bar_ld, a_bytes + b_bytes);
        bulk_g2s(sA, A, a_bytes, &bar_ld);
        bulk_g2s(sB, Bm, b_bytes, &bar_ld);
    }
    mbar_wait(&bar_ld, 0);

    uint32_t parity = 0;
    phase_begin();
    if (tid == 0) {
        const uint32_t idesc = make_idesc_bf16(128, N);
        for (int ks = 0; ks < K / 16; ++ks) {
            uint32_t a_addr = smem_u32(sA) + (2 * ks) * a_rows * 16 + row_shift * 16;
            uint32_t b_addr = smem_u32(sB) + (2 * ks) * N * 16;
            uint64_t ad, bd;
            if (!swap_lbo_sbo) {
                ad = make_smem_desc(a_addr, a_rows * 16, 128);
                bd = make_smem_desc(b_addr, N * 16, 128);
            } else {
                ad = make_smem_desc(a_addr, 128, a_rows * 16);
                bd = make_smem_desc(b_addr, 128, N * 16);
            }
            umma_bf16(tmem, ad, bd, idesc, ks > 0);
        }
    }
    phase_end(&bar_mma, parity);

    const int warp = tid >> 5;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tmem_ld16(trow + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) D[(size_t)tid * N + c0 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem, 256);
}

}  // namespace pdse

extern "C" int pdse_probe_gemm(const void* A, const void* B, float* D, int a_rows, int N, int K, int row_shift,
                               int swap_lbo_sbo, void* stream) {
    using namespace pdse;
    if (K % 16 || N % 16 || N > 256 || N < 16 || a_rows < 128 + row_shift) return set_error("probe_gemm: bad shape");
    size_t smem = (size_t)(K / 8) * (a_rows + N) * 16;
    if (smem > 200 * 1024) return set_error("probe_gemm: too large");
    static int hw = 0;
    if (int e = ensure_smem(probe_gemm_kernel, smem, &hw)) return e;
    probe_gemm_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)A, (const __nv_bfloat16*)B, D,
                                                              a_rows, N, K, row_shift, swap_lbo_sbo);
    return check_launch("probe_gemm");
}
