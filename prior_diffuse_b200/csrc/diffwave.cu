// diff2.DiffWave (model/diff2.py:12-158), the time-domain gated-tanh residual stack north_star describes -- SURVEY 8(f) item 4.
//
//   x = relu(W_in audio + b)               cond = relu(W_in audio_init + b)                      (:29-31, :38-40)
//   per layer i, dilation d = 2^(i % cycle):
//     z = dilated_conv_i(x + W_d,i e + b_d,i) + conditioner_projection_i(cond)   [2C]            (:132-136)
//     g = sigmoid(z[:C]) * tanh(z[C:])                                                            (:138-139)
//     (residual | skip) = output_projection_i(g) ;  x = (x + residual) / sqrt(2) ;  S += skip     (:154-158, :46-48)
//   out = W_out relu(W_skip S / sqrt(layers) + b) + b                                             (:50-55)
//
// One launch per layer.  Rows = time samples (3 s = 48 000 rows per utterance, C = 64 channels): the two k = 3 dilated
// convolutions of a layer are ONE implicit GEMM with K = 2 x 3 x 64 = 384 and N = 128 per 128-row tile -- six 128-row
// windows (x+e at t-d, t, t+d; cond at t-d, t, t+d) streamed through a shared-memory ring by bulk copies, the weights of
// the layer (112 KB) resident in shared memory, accumulators in tensor memory; the gate, the 64 -> 128 output projection
// (a second tcgen05 GEMM on the staged g), the residual / skip update and the next layer's bf16 operand (x + e_next) are
// fused behind it.  Zero padding of the convolutions = zero guard rows around every utterance's operand planes.
// The residual stream and the skip sum stay fp32 in HBM (planes of float4: coalesced row-per-thread access); a layer moves
// ~1.4 KB per row for 115 kFLOP, so the stack is HBM-bound (ridge of the B200: ~214 flop/B).
#include "common.cuh"
#include "umma.cuh"

namespace pdse {

constexpr int DW_C = 64;            // residual channels (the DiffWave base configuration; the kernels are built for it)
constexpr int DW_GUARD = 640;       // zero rows in front of / behind every utterance's operand plane: >= max dilation (512) + a tile
constexpr int DW_WIN = 8 * 128 * 16;          // one window: 8 chunk planes x 128 rows x 16 B
constexpr int DW_RING = 4;                    // windows in flight
constexpr int DW_WB = (48 + 8) * 128 * 16;    // W_cat [48 planes][128][8] | W_o [8][128][8]  (bf16)
constexpr int DW_BIAS = 2 * (2 * 128 * 16);   // bias blocks of the conv pair and of the output projection
constexpr int DW_GROUPS = 2;                  // row-thread groups taking alternate tiles (256 TMEM columns and 16 KB staging each)
constexpr int DW_SMEM = DW_WB + DW_BIAS + 4096 + DW_GROUPS * 16384 + DW_RING * DW_WIN;
constexpr int DW_THR = 128 * DW_GROUPS + 64;  // 4 warps per group: one thread per row; last two warps: loader lanes (4 planes each)

// ---------------------------------------------------------------------------------------------- diffusion embedding
// DiffusionEmbedding (:71-95: table lookup / lerp, two SiLU linears) + every layer's diffusion_projection (:132):
// t [B] -> dtab [B][n_rows], n_rows = layers * 64
__device__ __forceinline__ float dw_warp_sum(float v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__global__ void __launch_bounds__(512)
dw_embed_kernel(const float* __restrict__ t, const float* __restrict__ table, const float* __restrict__ p1w,
                const float* __restrict__ p1b, const float* __restrict__ p2w, const float* __restrict__ p2b,
                const float* __restrict__ rows, const float* __restrict__ rbias, int n_rows, float* __restrict__ out) {
    __shared__ float e[128], h1[512], h2[512];
    const int n = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float tv = t[n];
    int lo = (int)floorf(tv), hi = (int)ceilf(tv);
    lo = min(max(lo, 0), 49);
    hi = min(max(hi, 0), 49);
    if (tid < 128) {
        const float a = table[lo * 128 + tid], b = table[hi * 128 + tid];
        e[tid] = a + (b - a) * (tv - (float)lo);
    }
    __syncthreads();
    for (int r = warp; r < 512; r += 16) {
        float acc = 0.f;
        for (int k = lane; k < 128; k += 32) acc = fmaf(p1w[r * 128 + k], e[k], acc);
        acc = dw_warp_sum(acc);
        if (lane == 0) {
            const float v = acc + p1b[r];
            h1[r] = v / (1.f + expf(-v));
        }
    }
    __syncthreads();
    for (int r = warp; r < 512; r += 16) {
        float acc = 0.f;
        for (int k = lane; k < 512; k += 32) acc = fmaf(p2w[r * 512 + k], h1[k], acc);
        acc = dw_warp_sum(acc);
        if (lane == 0) {
            const float v = acc + p2b[r];
            h2[r] = v / (1.f + expf(-v));
        }
    }
    __syncthreads();
    for (int r = warp; r < n_rows; r += 16) {
        float acc = 0.f;
        for (int k = lane; k < 512; k += 32) acc = fmaf(rows[(size_t)r * 512 + k], h2[k], acc);
        acc = dw_warp_sum(acc);
        if (lane == 0) out[(size_t)n * n_rows + r] = acc + rbias[r];
    }
}

// ---------------------------------------------------------------------------------------------- input projection
// x = relu(w audio + b) -> fp32 planes [B][16][L][4] and the first layer's operand y = bf16(x + e_0) [B][8][Lg][8];
// cond = relu(w audio_init + b) -> bf16 [B][8][Lg][8]
__global__ void __launch_bounds__(256)
dw_pre_kernel(const float* __restrict__ audio, const float* __restrict__ init, const float* __restrict__ win,
              const float* __restrict__ dtab, int dstride, float* __restrict__ x, __nv_bfloat16* __restrict__ y,
              __nv_bfloat16* __restrict__ cond, int L, int Lg) {
    __shared__ float sw[128], sd[64];
    const int b = blockIdx.y;
    if (threadIdx.x < 128) sw[threadIdx.x] = win[threadIdx.x];
    if (threadIdx.x < 64) sd[threadIdx.x] = dtab[(size_t)b * dstride + threadIdx.x];
    __syncthreads();
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= L) return;
    const float a = audio[(size_t)b * L + t], ai = init[(size_t)b * L + t];
#pragma unroll
    for (int kc = 0; kc < 8; ++kc) {
        float xv[8], yv[8], cv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = kc * 8 + j;
            xv[j] = fmaxf(fmaf(sw[c], a, sw[64 + c]), 0.f);
            cv[j] = fmaxf(fmaf(sw[c], ai, sw[64 + c]), 0.f);
            yv[j] = xv[j] + sd[c];
        }
        float4* xp = reinterpret_cast<float4*>(x + (((size_t)b * 16 + 2 * kc) * L + t) * 4);
        xp[0] = make_float4(xv[0], xv[1], xv[2], xv[3]);
        xp[(size_t)L] = make_float4(xv[4], xv[5], xv[6], xv[7]);
        const size_t o = (((size_t)b * 8 + kc) * Lg + DW_GUARD + t) * 8;
        *reinterpret_cast<uint4*>(y + o) = pack8(yv);
        *reinterpret_cast<uint4*>(cond + o) = pack8(cv);
    }
}

// ---------------------------------------------------------------------------------------------- one residual layer
struct DwLayerArgs {
    const __nv_bfloat16* y_in;    // [B][8][Lg][8]  bf16(x + e_i), zero guards
    __nv_bfloat16* y_out;         // the next layer's operand (another buffer: neighbouring tiles still read y_in)
    const __nv_bfloat16* cond;    // [B][8][Lg][8]
    float* x;                     // [B][16][L][4]  residual stream, in place
    float* skip;                  // [B][16][L][4]  skip sum
    const __nv_bfloat16* wb;      // W_cat | W_o | bias block (conv pair) | bias block (output projection)
    const float* dnext;           // e_{i+1}: [B][dstride] (NULL for the last layer)
    int dstride, B, L, Lg, dil, first, last;
};

__global__ void __launch_bounds__(DW_THR, 1) dw_layer_kernel(DwLayerArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar_w, bar_mma[DW_GROUPS], full[DW_RING], empty[DW_RING];
    __shared__ uint32_t tmem_slot;
    __shared__ volatile int issue_turn;            // tiles whose conv MMAs have been issued (keeps the ring's consumers in tile order)
    __shared__ float sD[DW_GROUPS][64];            // e_{i+1} of the tile's utterance
    uint8_t* sW = smem;                            // W_cat (48 planes) | W_o (8 planes)
    uint8_t* sBias = sW + DW_WB;                   // conv bias block | output bias block
    uint8_t* sOnes = sBias + DW_BIAS;
    uint8_t* sG = sOnes + 4096;                    // g staging [groups][8][128][16 B]
    uint8_t* sRing = sG + DW_GROUPS * 16384;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(&bar_w, 1);
        for (int g = 0; g < DW_GROUPS; ++g) mbar_init(&bar_mma[g], 1);
        for (int s = 0; s < DW_RING; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        issue_turn = 0;
        fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) tmem_alloc(&tmem_slot, 256 * DW_GROUPS);
    init_ones_plane(sOnes, tid, DW_THR);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const int tiles_t = (a.L + 127) / 128, total = a.B * tiles_t;

    if (warp >= 4 * DW_GROUPS) {
        // ------------------------------------------------------------------ loader lanes: windows through the ring
        if (lane == 0) {
            const int half = warp - 4 * DW_GROUPS;             // planes [4 half, 4 half + 4) of every window
            if (half == 0) {
                mbar_arrive_expect_tx(&bar_w, DW_WB + DW_BIAS);
                bulk_g2s(sW, a.wb, DW_WB + DW_BIAS, &bar_w);
            }
            uint32_t cnt = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
                const int b = tile / tiles_t, t0 = (tile % tiles_t) * 128;
                for (int w = 0; w < 6; ++w, ++cnt) {
                    const uint32_t s = cnt % DW_RING;
                    if (cnt >= DW_RING) mbar_wait(&empty[s], ((cnt / DW_RING) - 1) & 1);
                    if (half == 0) mbar_arrive_expect_tx(&full[s], DW_WIN);
                    const __nv_bfloat16* src = (w < 3 ? a.y_in : a.cond) +
                                               (((size_t)b * 8) * a.Lg + DW_GUARD + t0 + ((w % 3) - 1) * a.dil) * 8;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int kc = half * 4 + k;
                        bulk_g2s(sRing + s * DW_WIN + kc * 2048, src + (size_t)kc * a.Lg * 8, 2048, &full[s]);
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ row threads (thread = accumulator row);
        // the groups take alternate tiles, so one group's memory-bound update runs under the other's MMAs and gates
        const int g = warp >> 2, r = tid & 127;
        const uint32_t tacc = tmem_slot + 256 * g;             // [0,128): gate | filter ; [128,256): residual | skip
        const uint32_t trow = tacc + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t b_conv = smem_u32(sBias), b_out = b_conv + 4096, ones = smem_u32(sOnes);
        uint8_t* sA2 = sG + g * 16384;
        const int bar_id = 1 + g;
        mbar_wait(&bar_w, 0);
        uint32_t par = 0;
        for (int i = g;; i += DW_GROUPS) {
            const int tile = blockIdx.x + i * gridDim.x;
            if (tile >= total) break;
            const int b = tile / tiles_t, t0 = (tile % tiles_t) * 128, t = t0 + r;
            const bool valid = t < a.L;
            const size_t xo = (((size_t)b * 16) * a.L + t) * 4;       // + plane * 4 L
            if (r < 64 && a.dnext) sD[g][r] = __ldg(a.dnext + (size_t)b * a.dstride + r);
            // the tile's residual rows: nothing in this tile depends on them until the very end -- in flight under the MMAs
            float4 xv[16];
            if (valid) {
#pragma unroll
                for (int p = 0; p < 16; ++p) xv[p] = __ldcs(reinterpret_cast<const float4*>(a.x + xo + (size_t)p * 4 * a.L));
            }
            if (r == 0) {
                const uint32_t idesc = make_idesc_op(128, 128);
                while (issue_turn != i) {}
                umma_bias(tacc, ones, b_conv, 128, 0);
                uint32_t cnt = 6u * (uint32_t)i;
                for (int w = 0; w < 6; ++w, ++cnt) {
                    const uint32_t s = cnt % DW_RING;
                    mbar_wait(&full[s], (cnt / DW_RING) & 1);
                    tc_fence_after();
                    const uint64_t aD = make_smem_desc(smem_u32(sRing) + s * DW_WIN, 2048, 128);
                    const uint64_t bD = make_smem_desc(smem_u32(sW) + w * 8 * 2048, 2048, 128);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) umma_bf16(tacc, dadd(aD, 2 * ks * 2048), dadd(bD, 2 * ks * 2048), idesc, 1);
                    umma_commit(&empty[s]);
                }
                umma_commit(&bar_mma[g]);
                issue_turn = i + 1;
            }
            phase_wait(&bar_mma[g], par);
            // g = sigmoid(gate) * tanh(filter)  (:138-139; sigmoid(z) = 0.5 tanh(z / 2) + 0.5) -> bf16 A operand
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 16) {
                float gt[16], fl[16];
                tmem_ld16(trow + c0, gt);
                tmem_ld16(trow + 64 + c0, fl);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) gt[j] = fmaf(0.5f, tanh_fast(0.5f * gt[j]), 0.5f) * tanh_fast(fl[j]);
                *reinterpret_cast<uint4*>(sA2 + (c0 / 8) * 2048 + r * 16) = pack8(gt);
                *reinterpret_cast<uint4*>(sA2 + (c0 / 8 + 1) * 2048 + r * 16) = pack8(gt + 8);
            }
            // (the group's 128 row threads only: nobody else joins these barriers)
            fence_proxy_async_smem();
            tc_fence_before();
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
            tc_fence_after();
            if (r == 0) {
                const uint32_t idesc = make_idesc_op(128, 128);
                const uint64_t aD = make_smem_desc(smem_u32(sA2), 2048, 128), bD = make_smem_desc(smem_u32(sW) + 48 * 2048, 2048, 128);
                umma_bias(tacc + 128, ones, b_out, 128, 0);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma_bf16(tacc + 128, dadd(aD, 2 * ks * 2048), dadd(bD, 2 * ks * 2048), idesc, 1);
                umma_commit(&bar_mma[g]);
            }
            // the first half of the skip rows, in flight under the output projection
            float4 sv[8];
            const bool ldskip = valid && !a.first;
#pragma unroll
            for (int q = 0; q < 8; ++q)
                sv[q] = ldskip ? __ldcs(reinterpret_cast<const float4*>(a.skip + xo + (size_t)q * 4 * a.L)) : make_float4(0.f, 0.f, 0.f, 0.f);
            phase_wait(&bar_mma[g], par);
            // x = (x + residual) / sqrt(2) ; skip += skip_i ; y_next = bf16(x + e_{i+1})
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c0 = 32 * h;
                float rs[32], sk[32];
                tmem_ld32(trow + 128 + c0, rs);
                tmem_ld32(trow + 192 + c0, sk);
                tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int p = 8 * h + q;
                        float4 v = xv[p];
                        v.x = (v.x + rs[4 * q + 0]) * 0.70710678118654752f;
                        v.y = (v.y + rs[4 * q + 1]) * 0.70710678118654752f;
                        v.z = (v.z + rs[4 * q + 2]) * 0.70710678118654752f;
                        v.w = (v.w + rs[4 * q + 3]) * 0.70710678118654752f;
                        __stcs(reinterpret_cast<float4*>(a.x + xo + (size_t)p * 4 * a.L), v);
                        float4 u = sv[q];
                        u.x += sk[4 * q + 0], u.y += sk[4 * q + 1], u.z += sk[4 * q + 2], u.w += sk[4 * q + 3];
                        __stcs(reinterpret_cast<float4*>(a.skip + xo + (size_t)p * 4 * a.L), u);
                        rs[4 * q + 0] = v.x + sD[g][c0 + 4 * q + 0];
                        rs[4 * q + 1] = v.y + sD[g][c0 + 4 * q + 1];
                        rs[4 * q + 2] = v.z + sD[g][c0 + 4 * q + 2];
                        rs[4 * q + 3] = v.w + sD[g][c0 + 4 * q + 3];
                    }
                    if (h == 0 && !a.first) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) sv[q] = __ldcs(reinterpret_cast<const float4*>(a.skip + xo + (size_t)(8 + q) * 4 * a.L));
                    }
                    if (!a.last) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            *reinterpret_cast<uint4*>(a.y_out + (((size_t)b * 8 + c0 / 8 + k) * a.Lg + DW_GUARD + t) * 8) = pack8(rs + 8 * k);
                    }
                }
            }
            // the group's next tile overwrites its accumulators, sD and the g staging: every row thread is done with them
            tc_fence_before();
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
            tc_fence_after();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_slot, 256 * DW_GROUPS);
}

// ---------------------------------------------------------------------------------------------- skip / output projections
// out = w_out . relu(W_skip (S * scale) + b_skip) + b_out      (:50-55)
__global__ void __launch_bounds__(128)
dw_post_kernel(const float* __restrict__ skip, const __nv_bfloat16* __restrict__ wb, const float* __restrict__ wout, float scale,
               float* __restrict__ out, int L) {
    __shared__ __align__(128) uint8_t sA[16384];     // [8][128][16 B]
    __shared__ __align__(128) uint8_t sWk[8192 + 2048];
    __shared__ __align__(128) uint8_t sOnes[4096];
    __shared__ float so[68];
    __shared__ uint64_t bar_w, bar_mma;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, b = blockIdx.y, t = blockIdx.x * 128 + tid;
    if (tid == 0) {
        mbar_init(&bar_w, 1);
        mbar_init(&bar_mma, 1);
        fence_mbar_init();
    }
    __syncwarp();
    if (tid < 32) tmem_alloc(&tmem_slot, 64);
    init_ones_plane(sOnes, tid, 128);
    if (tid < 65) so[tid] = wout[tid];
    if (tid == 0) {
        mbar_arrive_expect_tx(&bar_w, 8192 + 2048);
        bulk_g2s(sWk, wb, 8192 + 2048, &bar_w);
    }
#pragma unroll
    for (int kc = 0; kc < 8; ++kc) {
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (t < L) {
            const float4* sp = reinterpret_cast<const float4*>(skip + (((size_t)b * 16 + 2 * kc) * L + t) * 4);
            const float4 u0 = sp[0], u1 = sp[(size_t)L];
            v[0] = u0.x * scale, v[1] = u0.y * scale, v[2] = u0.z * scale, v[3] = u0.w * scale;
            v[4] = u1.x * scale, v[5] = u1.y * scale, v[6] = u1.z * scale, v[7] = u1.w * scale;
        }
        *reinterpret_cast<uint4*>(sA + kc * 2048 + tid * 16) = pack8(v);
    }
    mbar_wait(&bar_w, 0);
    phase_begin();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_op(128, 64);
        const uint64_t aD = make_smem_desc(smem_u32(sA), 2048, 128), bD = make_smem_desc(smem_u32(sWk), 1024, 128);
        umma_bias(tmem, smem_u32(sOnes), smem_u32(sWk) + 8192, 64, 0);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_bf16(tmem, dadd(aD, 2 * ks * 2048), dadd(bD, 2 * ks * 1024), idesc, 1);
    }
    uint32_t par = 0;
    phase_end(&bar_mma, par);
    const uint32_t trow = tmem + ((uint32_t)((tid >> 5) * 32) << 16);
    float acc = so[64];
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
        float v[32];
        tmem_ld32(trow + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc = fmaf(fmaxf(v[j], 0.f), so[c0 + j], acc);
    }
    if (t < L) out[(size_t)b * L + t] = acc;
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem, 64);
}

}  // namespace pdse

// ============================================================================ C ABI
using namespace pdse;

extern "C" int pdse_dw_guard_rows(void) { return DW_GUARD; }

extern "C" int pdse_dw_embed(const float* t, int B, const float* table, const float* p1w, const float* p1b, const float* p2w,
                             const float* p2b, const float* rows, const float* rbias, int n_rows, float* dtab, void* stream) {
    if (B <= 0 || n_rows <= 0) return set_error("pdse_dw_embed: empty input");
    dw_embed_kernel<<<B, 512, 0, (cudaStream_t)stream>>>(t, table, p1w, p1b, p2w, p2b, rows, rbias, n_rows, dtab);
    return check_launch("pdse_dw_embed");
}

extern "C" int pdse_dw_pre_fwd(const float* audio, const float* init, const float* win, const float* dtab, int dstride, float* x,
                               void* y, void* cond, int B, int L, void* stream) {
    if (B <= 0 || L <= 0) return set_error("pdse_dw_pre_fwd: empty input");
    dim3 grid(ceil_div(L, 256), B);
    dw_pre_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(audio, init, win, dtab, dstride, x, (__nv_bfloat16*)y, (__nv_bfloat16*)cond, L,
                                                          L + 2 * DW_GUARD);
    return check_launch("pdse_dw_pre_fwd");
}

extern "C" int pdse_dw_layer_fwd(const void* y_in, void* y_out, const void* cond, float* x, float* skip, const void* wb,
                                 const float* dnext, int dstride, int B, int L, int dilation, int first, int last, void* stream) {
    if (B <= 0 || L <= 0) return set_error("pdse_dw_layer_fwd: empty input");
    if (dilation < 1 || dilation + 128 > DW_GUARD) return set_error("pdse_dw_layer_fwd: dilation must be in [1, 512]");
    if (!last && (!dnext || !y_out)) return set_error("pdse_dw_layer_fwd: the next layer's embedding and operand buffer are required");
    DwLayerArgs a;
    a.y_in = (const __nv_bfloat16*)y_in;
    a.y_out = (__nv_bfloat16*)y_out;
    a.cond = (const __nv_bfloat16*)cond;
    a.x = x;
    a.skip = skip;
    a.wb = (const __nv_bfloat16*)wb;
    a.dnext = last ? nullptr : dnext;
    a.dstride = dstride;
    a.B = B;
    a.L = L;
    a.Lg = L + 2 * DW_GUARD;
    a.dil = dilation;
    a.first = first;
    a.last = last;
    static SmemCache hw;
    if (int e = ensure_smem(dw_layer_kernel, (size_t)DW_SMEM, &hw)) return e;
    const int tiles = B * ceil_div(L, 128);
    dw_layer_kernel<<<min(tiles, sm_count()), DW_THR, DW_SMEM, (cudaStream_t)stream>>>(a);
    return check_launch("pdse_dw_layer_fwd");
}

extern "C" int pdse_dw_post_fwd(const float* skip, const void* wb, const float* wout, float scale, float* out, int B, int L,
                                void* stream) {
    if (B <= 0 || L <= 0) return set_error("pdse_dw_post_fwd: empty input");
    dim3 grid(ceil_div(L, 128), B);
    dw_post_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(skip, (const __nv_bfloat16*)wb, wout, scale, out, L);
    return check_launch("pdse_dw_post_fwd");
}
