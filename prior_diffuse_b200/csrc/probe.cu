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
    if (swap_lbo_sbo == 2) {   // A operand from tensor memory: row tid's K elements as bf16 pairs in columns [128, 128 + K/2)
        const uint32_t trow_a = tmem + ((uint32_t)((tid >> 5) * 32) << 16) + 128;
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint4 u0 = *reinterpret_cast<const uint4*>(sA + (size_t)(2 * ks) * a_rows * 16 + (tid + row_shift) * 16);
            const uint4 u1 = *reinterpret_cast<const uint4*>(sA + (size_t)(2 * ks + 1) * a_rows * 16 + (tid + row_shift) * 16);
            const uint32_t w[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
            tmem_st8(trow_a + ks * 8, reinterpret_cast<const float*>(w));
        }
        tmem_st_wait();
    }
    phase_begin();
    if (tid == 0 && swap_lbo_sbo == 2) {
        const uint32_t idesc = make_idesc_bf16(128, N);
        for (int ks = 0; ks < K / 16; ++ks)
            umma_bf16_ta(tmem, tmem + 128 + ks * 8, make_smem_desc(smem_u32(sB) + (2 * ks) * N * 16, N * 16, 128), idesc, ks > 0);
    } else if (tid == 0) {
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
    if (K % 16 || N % 16 || N > 256 || N < 16 || a_rows < 128 + row_shift || (swap_lbo_sbo == 2 && (N > 128 || K > 256))) return set_error("probe_gemm: bad shape");
    size_t smem = (size_t)(K / 8) * (a_rows + N) * 16;
    if (smem > 200 * 1024) return set_error("probe_gemm: too large");
    static SmemCache hw;
    if (int e = ensure_smem(probe_gemm_kernel, smem, &hw)) return e;
    probe_gemm_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)A, (const __nv_bfloat16*)B, D,
                                                              a_rows, N, K, row_shift, swap_lbo_sbo);
    return check_launch("probe_gemm");
}

// ---------------------------------------------------------------------------------------------------------
// pdse_probe_tmem: measurement hook (tests/gpu_probe_tmem.py).  How fast can four warps drain a 128 x 256 fp32
// accumulator with tcgen05.ld, alone and while one lane keeps the tensor core busy with back-to-back
// 128 x N x 16 MMAs (four committed batches of 16 in flight) into the other half of TMEM?  And how fast does one lane pull bulk copies into shared memory?
// out[blockIdx.x * 4 + {0,1,2,3}] = cycles per drain | MMAs issued | cycles of the MMA lane | bulk bytes per cycle x 1000
namespace pdse {
__global__ void __launch_bounds__(192, 1)
probe_tmem_kernel(long long* __restrict__ out, const uint8_t* __restrict__ src, int mode, int iters, int mma_n, int ld_cols,
                  int copy_bytes, long cta_stride, int nblk) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar_mma[4], bar_cp[4];
    __shared__ uint32_t tmem_slot;
    __shared__ volatile int stop;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(&bar_mma[i], 1);
        for (int i = 0; i < 4; ++i) mbar_init(&bar_cp[i], 1);
        fence_mbar_init();
        stop = 0;
    }
    for (int i = tid; i < 65536 / 16; i += 192) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    __syncwarp();
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp == 0) {
        if (lane == 0 && (mode & 1)) {       // MMA stream: batches of 8 MMAs, at most two batches in flight
            const uint32_t idesc = make_idesc_bf16(128, mma_n);
            // mode bits [8,16): A window start in 16-byte rows (shifted-window convolution taps); [16,32): A plane stride in rows
            const uint32_t a_shift = (mode >> 8) & 0xff, a_lbo = (mode >> 16) ? (uint32_t)(mode >> 16) * 16 : 2048;
            const uint64_t ad = make_smem_desc(smem_u32(smem) + a_shift * 16, a_lbo, 128), bd = make_smem_desc(smem_u32(smem) + 32768, mma_n * 16, 128);
            long long n = 0;
            const long long t0 = clock64();
            int batch = 0;
            while (!stop) {
                if (batch >= 4) mbar_wait(&bar_mma[batch & 3], ((batch >> 2) - 1) & 1);   // four batches of 16 in flight
                if (mode & 8) {                  // four independent accumulators in turn (is the ~60-cycle floor a D dependency?)
#pragma unroll
                    for (int i = 0; i < 16; ++i) umma_bf16(tmem + (i & 3) * 64, ad, bd, idesc, 1);
                } else if (mode & 16) {          // A operand from tensor memory (columns [256, 264))
#pragma unroll
                    for (int i = 0; i < 16; ++i) umma_bf16_ta(tmem, tmem + 256, bd, idesc, 1);
                } else if (mode & 32) {          // both
#pragma unroll
                    for (int i = 0; i < 16; ++i) umma_bf16_ta(tmem + (i & 3) * 64, tmem + 256, bd, idesc, 1);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) umma_bf16(tmem, ad, bd, idesc, 1);
                }
                umma_commit(&bar_mma[batch & 3]);
                ++batch;
                n += 16;
            }
            const long long t1 = clock64();
            out[blockIdx.x * 4 + 1] = n;
            out[blockIdx.x * 4 + 2] = t1 - t0;
        }
        __syncwarp();
    } else if (warp == 5 && (mode & 64)) {
        if (lane == 0) {                     // a second MMA issuer on a warp of its own (own accumulator, own barriers)
            const uint32_t idesc = make_idesc_bf16(128, mma_n);
            const uint64_t ad = make_smem_desc(smem_u32(smem), 2048, 128), bd = make_smem_desc(smem_u32(smem) + 32768, mma_n * 16, 128);
            long long n = 0;
            const long long t0 = clock64();
            int batch = 0;
            while (!stop) {
                if (batch >= 4) mbar_wait(&bar_cp[batch & 3], ((batch >> 2) - 1) & 1);
#pragma unroll
                for (int i = 0; i < 16; ++i) umma_bf16(tmem + 128, ad, bd, idesc, 1);
                umma_commit(&bar_cp[batch & 3]);
                ++batch;
                n += 16;
            }
            const long long t1 = clock64();
            out[blockIdx.x * 4 + 3] = (t1 - t0) * 1000 / n;
        }
        __syncwarp();
    } else if (warp == 5) {
        if (lane == 0 && (mode & 4)) {       // bulk-copy stream: 4 copies in flight into the upper half of smem
            const long long t0 = clock64();
            long long bytes = 0;
            for (int i = 0; i < iters * 8; ++i) {
                const int s = i & 3;
                if (i >= 4) mbar_wait(&bar_cp[s], ((i >> 2) - 1) & 1);
                mbar_arrive_expect_tx(&bar_cp[s], copy_bytes);
                bulk_g2s(smem + 65536 + s * 32768, src + (size_t)blockIdx.x * cta_stride + (size_t)(i % nblk) * 32768, copy_bytes, &bar_cp[s]);
                bytes += copy_bytes;
            }
            for (int s = 0; s < 4; ++s) mbar_wait(&bar_cp[s], ((iters * 8 - 4 + ((s - iters * 8 % 4 + 4) % 4)) >> 2) & 1);
            const long long t1 = clock64();
            out[blockIdx.x * 4 + 3] = bytes * 1000 / (t1 - t0);
        }
        __syncwarp();
    } else {
        const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256;
        float acc = 0.f;
        __syncwarp();
        const long long t0 = clock64();
        if (mode & 2) {
            for (int it = 0; it < iters; ++it) {
                if (ld_cols == 32) {
                    for (int c0 = 0; c0 < 256; c0 += 32) {
                        float v[32];
                        tmem_ld32(trow + c0, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc += v[j];
                    }
                } else {
                    for (int c0 = 0; c0 < 256; c0 += 16) {
                        float v[16];
                        tmem_ld16(trow + c0, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) acc += v[j];
                    }
                }
            }
        } else {
            for (int it = 0; it < iters * 50; ++it) acc += __sinf(acc);
        }
        const long long t1 = clock64();
        if (tid == 32) out[blockIdx.x * 4 + 0] = (t1 - t0) / iters;
        if (acc == 123.456f) out[0] = 0;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (tid == 32) stop = 1;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}
}  // namespace pdse

extern "C" int pdse_probe_tmem(long long* out, const void* src, int mode, int iters, int mma_n, int ld_cols, int copy_bytes,
                               int ctas, long cta_stride, int nblk, void* stream) {
    using namespace pdse;
    if (mma_n % 16 || mma_n > 256 || copy_bytes > 32768 || copy_bytes % 16) return set_error("probe_tmem: bad arguments");
    static SmemCache hw;
    const size_t smem = 65536 + 4 * 32768;
    if (int e = ensure_smem(probe_tmem_kernel, smem, &hw)) return e;
    probe_tmem_kernel<<<ctas, 192, smem, (cudaStream_t)stream>>>(out, (const uint8_t*)src, mode, iters, mma_n, ld_cols, copy_bytes,
                                                                 cta_stride, nblk > 0 ? nblk : 1);
    return check_launch("probe_tmem");
}
