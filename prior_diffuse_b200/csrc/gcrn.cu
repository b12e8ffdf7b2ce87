// GCRN prior (model/gcrn.py:87-166) on sm_100a.
//
//   conv1                      : explicit 6-wide im2col (Cin = 2) + one UMMA, GLU/BN/ELU epilogue
//   conv2..5, conv5_t..conv2_t : ONE streaming implicit-GEMM kernel (stream_kernel): the tile's input planes
//                                stay resident in shared memory (taps = shifted windows), the weights stream
//                                through a 4-stage UBLKCP ring, accumulators are double-buffered in TMEM so
//                                the GLU + BN + ELU epilogue of one pass overlaps the MMAs of the next
//   LSTM input projections     : the same kernel in LIN mode (M = T*B rows, N = 2048, K = 512)
//   LSTM recurrence            : lstm_rec_kernel, 16 CTAs per group each holding a 128x512 slice of W_hh in
//                                shared memory for the whole sequence; h_t exchanged through L2
//   LayerNorm + group shuffles : ln_kernel (gcrn.py:29-35)
//   conv1_t + BN + ELU + fc    : out_kernel (CUDA cores; 1 output channel)
//
// Activation layouts (bf16 CP8, see umma.cuh):
//   SO  "split-outer"      [B][C/8][2][T*Q][8]   row(t,f) = t*Q + (f>>1) in parity plane f&1   (strided conv input)
//   UG  "unsplit guarded"  [B][C/8][T*P+1][8]    row(t,f) = t*P + 1 + f, P = F+1, row t*P = 0  (transposed conv input)
#include "common.cuh"
#include "umma.cuh"
#include <cstdlib>

namespace pdse {

__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : __expf(x) - 1.f; }

// ============================================================================ conv1
struct GConv1Args {
    const float* y;          // [B][2][T][161]
    __nv_bfloat16* out_so;   // SO F=80 (Q=40), 16 ch
    __nv_bfloat16* out_ug;   // UG F=80 (P=81), 16 ch, ELU applied a second time (gcrn.py:153)
    const __nv_bfloat16* wb; // [2][32][8]
    const float* ep;         // bv[16] bg[16] scale[16] shift[16]
    int B, T;
};

__global__ void __launch_bounds__(128) gconv1_kernel(GConv1Args a) {
    __shared__ __align__(128) uint8_t sA[2 * 2048];
    __shared__ __align__(128) uint8_t sW[2 * 32 * 16];
    __shared__ float sy[3 * 2 * 164];
    __shared__ uint64_t bar_mma;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&bar_mma, 1);
        fence_mbar_init();
    }
    __syncwarp();
    if (tid < 32) tmem_alloc(&tmem_slot, 32);
    for (int i = tid; i < 64; i += 128) reinterpret_cast<uint4*>(sW)[i] = reinterpret_cast<const uint4*>(a.wb)[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t trow = tmem + ((uint32_t)((tid >> 5) * 32) << 16);
    uint32_t par = 0;
    const int rows_b = a.T * 80, tiles_b = (rows_b + 127) / 128;
    for (int tile = blockIdx.x; tile < a.B * tiles_b; tile += gridDim.x) {
        const int b = tile / tiles_b, m0 = (tile % tiles_b) * 128, tA = m0 / 80;
        __syncthreads();
        for (int i = tid; i < 3 * 2 * 161; i += 128) {
            const int rr = i / 322, rem = i % 322, c = rem / 161, f = rem % 161, t = tA + rr;
            sy[(rr * 2 + c) * 164 + f] = t < a.T ? a.y[(((size_t)b * 2 + c) * a.T + t) * 161 + f] : 0.f;
        }
        __syncthreads();
        const int m = m0 + tid, t = m / 80, j = m % 80;
        const bool valid = m < rows_b;
        float v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = 0.f;
        if (valid) {
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int df = 0; df < 3; ++df) v[c * 3 + df] = sy[((t - tA) * 2 + c) * 164 + 2 * j + df];
        }
        *reinterpret_cast<uint4*>(sA + tid * 16) = pack8(v);
        *reinterpret_cast<uint4*>(sA + 2048 + tid * 16) = pack8(v + 8);
        phase_begin();
        if (tid == 0)
            umma_bf16(tmem, make_smem_desc(smem_u32(sA), 2048, 128), make_smem_desc(smem_u32(sW), 512, 128),
                      make_idesc_op(128, 32), 0);
        phase_end(&bar_mma, par);
        float d[32];
        tmem_ld32(trow, d);
        tmem_ld_wait();
        if (valid) {
            float e[16], e2[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const float g = (d[c] + __ldg(a.ep + c)) * fast_sigmoid(d[16 + c] + __ldg(a.ep + 16 + c));
                e[c] = elu1(fmaf(g, __ldg(a.ep + 32 + c), __ldg(a.ep + 48 + c)));
                e2[c] = elu1(e[c]);
            }
            const size_t so_rows = (size_t)a.T * 40, ug_rows = (size_t)a.T * 81 + 1;
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                *reinterpret_cast<uint4*>(a.out_so + ((((size_t)b * 2 + cc) * 2 + (j & 1)) * so_rows + (size_t)t * 40 + (j >> 1)) * 8) =
                    pack8(e + 8 * cc);
                *reinterpret_cast<uint4*>(a.out_ug + (((size_t)b * 2 + cc) * ug_rows + (size_t)t * 81 + 1 + j) * 8) = pack8(e2 + 8 * cc);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem, 32);
}

// ============================================================================ streaming implicit GEMM
enum : int { MODE_ENC = 0, MODE_DEC = 1, MODE_LIN = 2 };
// warp 0 = producer lane, warp 1 = MMA issuer lane, then 4 * NH epilogue warps

struct StreamArgs {
    const __nv_bfloat16* src[2];   // A sources (concatenated along K)
    int nc[2];                      // chunk planes per source
    int npar;                       // parity sub-planes per chunk plane (2 = SO input, 1 otherwise)
    long plane_rows;                // rows of one (chunk, parity) plane in global memory
    int B, T, P, nt;                // tile = nt time rows of pitch P (LIN: T = total rows, P = 1, nt = 128, B = 1)
    int R;                          // shared-memory rows per plane (tile rows + 1)
    int n_out_par;                  // 2 for transposed convs (even / odd outputs), else 1
    int ntap[2], tap_par[2][3], tap_shift[2][3];
    const __nv_bfloat16* w[2];      // weight stream per output parity
    int ntile, n_ntiles, kb;        // n-tile width (value|gate), number of n-tiles, k-steps per streamed block
    const float* ep;                // per n-tile: bv | bg | scale | shift (ct each); LIN: bias[N]
    int ep_floats;                  // staged into shared memory once per CTA (the epilogues read nothing else)
    int stages;                     // depth of the weight ring (4 or 8)
    int abufs;                      // A tile buffers (2 = the next tile's planes are prefetched)
    int n_split;                    // work unit = (tile, 1/n_split of the n-tiles): evens out the last wave (LIN)
    int mode, elu2;
    int Fo[2];                      // valid outputs per virtual row for each output parity
    int C_out;                      // output channels (all n-tiles)
    __nv_bfloat16* out_so;          // ENC: next conv's input, SO with F = Fo[0]
    __nv_bfloat16* out_ug;          // ENC: decoder skip (UG, F = Fo[0]); DEC: next layer's input (UG, F = F_out)
    int ug_P;
    __nv_bfloat16* out_xl[2];       // conv5: LSTM layer-1 A operand per group, [64][T*B][8], row = t*B + b
    float* out_f32;                 // LIN: [T][N_total][Bp]   (row = t*Bl + b)
    int Bl, Bp, N_total;
};

// NH: epilogue threads per accumulator row -- 1: 192-thread CTAs, up to four per SM
// for the small layers; 2: 320-thread CTAs for the layers that run one CTA per SM (ncu source page, round 2: their four
// epilogue warps were busy 68 % of the time and the MMA issuer lane waited on them for 59 % of its own).
// Measured and NOT the limit of the large-K layers: the ring depth (4, 5, 8, 12, 16, 20 stages: same time) and a weight
// stream shared by 2 or 4 CTAs of a cluster through multicast copies (slower: the slots then wait for the slowest consumer).
constexpr int MAX_ST = 16;
template <int NH>
__global__ void __launch_bounds__(64 + 128 * NH, NH == 1 ? 4 : 1) stream_kernel(StreamArgs a) {
    constexpr int SNTHR = 64 + 128 * NH;
    const uint32_t ST = a.stages;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar_a[2], bar_adone[2], bar_full[MAX_ST], bar_empty[MAX_ST], bar_acc[2], bar_free[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NC = a.nc[0] + a.nc[1];
    const uint32_t RB = a.R * 16;
    const uint32_t a_bytes = ((uint32_t)NC * a.npar * RB + 127) & ~127u;
    const uint32_t blk_bytes = (uint32_t)a.ntile * 2 * a.kb * 16;
    const uint32_t abufs = a.abufs;
    uint8_t* sA = smem;
    uint8_t* sB = smem + abufs * a_bytes;
    float* sEp = reinterpret_cast<float*>(sB + ST * blk_bytes);
    for (int i = tid; i < a.ep_floats; i += SNTHR) sEp[i] = __ldg(a.ep + i);
    const uint32_t acc_cols = a.ntile < 32 ? 32 : a.ntile;     // per accumulator buffer
    const uint32_t tmem_cols = acc_cols * 2 <= 32 ? 32 : acc_cols * 2 <= 64 ? 64 : acc_cols * 2 <= 128 ? 128
                               : acc_cols * 2 <= 256 ? 256 : 512;
    if (tid == 0) {
        for (uint32_t s = 0; s < ST; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_a[i], 1);
            mbar_init(&bar_adone[i], 1);
            mbar_init(&bar_acc[i], 1);
            mbar_init(&bar_free[i], 128 * NH);
        }
        fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) tmem_alloc(&tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    const int tiles_t = (a.T + a.nt - 1) / a.nt;
    const int nsp = a.n_split, npt = a.n_ntiles / nsp;             // n-tiles per work unit
    const int total_tiles = a.B * tiles_t * nsp;                   // work units (unit u: tile u / nsp, n-tile group u % nsp)
    const int kpt = NC / (2 * a.kb);                 // streamed blocks per tap
    const size_t blk_elems = (size_t)a.ntile * 2 * a.kb * 8;

    if (warp == 0) {
        // ------------------------------------------------------------ producer lane: runs ahead of the MMAs across
        // n-tile, parity and tile boundaries (a bulk copy takes 1.2-1.7k cycles whatever its size, so the ring must
        // never drain); the next tile's A planes go into the other A buffer while this tile's MMAs run
        if (lane == 0) {
            uint32_t cnt = 0, tile_it = 0, pst = 0, prd = 0;     // (stage, round) of block cnt without a runtime division
            auto load_A = [&](int unit, uint32_t it) {
                const int tile = unit / nsp;
                const int b = tile / tiles_t, t0 = (tile % tiles_t) * a.nt;
                const long row0 = (long)t0 * a.P;
                long nrows = (long)min(a.nt, a.T - t0) * a.P + 1;
                if (row0 + nrows > a.plane_rows) nrows = a.plane_rows - row0;
                const uint32_t ab = abufs == 2 ? (it & 1) : 0, ause = abufs == 2 ? (it >> 1) : it;
                if (ause > 0) mbar_wait(&bar_adone[ab], (ause - 1) & 1);      // MMAs of the tile that used this buffer are done
                mbar_arrive_expect_tx(&bar_a[ab], (uint32_t)NC * a.npar * (uint32_t)nrows * 16);
                uint8_t* dstA = sA + ab * a_bytes;
                for (int s = 0, kc0 = 0; s < 2; ++s) {
                    for (int kc = 0; kc < a.nc[s]; ++kc)
                        for (int p = 0; p < a.npar; ++p)
                            bulk_g2s(dstA + ((kc0 + kc) * a.npar + p) * RB,
                                     a.src[s] + ((((size_t)b * a.nc[s] + kc) * a.npar + p) * a.plane_rows + row0) * 8,
                                     (uint32_t)nrows * 16, &bar_a[ab]);
                    kc0 += a.nc[s];
                }
            };
            if ((int)blockIdx.x < total_tiles) load_A(blockIdx.x, 0);
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_it) {
                if (abufs == 2) {
                    if (tile + (int)gridDim.x < total_tiles) load_A(tile + gridDim.x, tile_it + 1);
                } else if (tile_it > 0) {
                    load_A(tile, tile_it);
                }
                for (int op = 0; op < a.n_out_par; ++op) {
                    const int nblk = a.ntap[op] * kpt;
                    const __nv_bfloat16* wsrc = a.w[op] + (size_t)(tile % nsp) * npt * nblk * blk_elems;   // `tile` counts work units here
                    for (int i = 0; i < npt * nblk; ++i, ++cnt) {             // n-tiles are contiguous in the stream
                        const uint32_t s = pst;
                        if (prd > 0) mbar_wait(&bar_empty[s], (prd - 1) & 1);
                        if (++pst == ST) pst = 0, ++prd;
                        mbar_arrive_expect_tx(&bar_full[s], blk_bytes);
                        bulk_g2s(sB + s * blk_bytes, wsrc + (size_t)i * blk_elems, blk_bytes, &bar_full[s]);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer lane
        if (lane == 0) {
            uint32_t cnt = 0, pass = 0, tile_it = 0, mst = 0, mph = 0;
            const uint32_t idesc = make_idesc_op(128, a.ntile);
            const uint32_t a_lbo = a.npar * RB, kstep_a = 2 * a.npar * RB, kstep_b = 2 * a.ntile * 16;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_it) {
                const uint32_t ab = abufs == 2 ? (tile_it & 1) : 0, ause = abufs == 2 ? (tile_it >> 1) : tile_it;
                mbar_wait(&bar_a[ab], ause & 1);
                const uint64_t adesc0 = make_smem_desc(smem_u32(sA) + ab * a_bytes, a_lbo, 128);
                for (int op = 0; op < a.n_out_par; ++op) {
                    for (int nti = 0; nti < npt; ++nti, ++pass) {
                        const uint32_t buf = pass & 1;
                        if (pass >= 2) mbar_wait(&bar_free[buf], ((pass >> 1) - 1) & 1);   // epilogue drained this buffer
                        tc_fence_after();
                        const uint32_t d_tmem = tmem + buf * acc_cols;
                        uint32_t first = 0;
                        for (int tap = 0; tap < a.ntap[op]; ++tap) {
                            const uint64_t adesc_t = dadd(adesc0, a.tap_par[op][tap] * RB + a.tap_shift[op][tap] * 16);
                            for (int kblk = 0; kblk < kpt; ++kblk, ++cnt) {
                                const uint32_t s = mst;
                                mbar_wait(&bar_full[s], mph);
                                if (++mst == ST) mst = 0, mph ^= 1u;
                                tc_fence_after();
                                const uint64_t bdesc = make_smem_desc(smem_u32(sB) + s * blk_bytes, a.ntile * 16, 128);
                                const uint64_t adesc = dadd(adesc_t, (uint32_t)kblk * a.kb * kstep_a);
                                for (int ks = 0; ks < a.kb; ++ks) {
                                    umma_bf16(d_tmem, dadd(adesc, ks * kstep_a), dadd(bdesc, ks * kstep_b), idesc, first);
                                    first = 1;
                                }
                                umma_commit(&bar_empty[s]);
                            }
                        }
                        umma_commit(&bar_acc[buf]);
                    }
                }
                umma_commit(&bar_adone[ab]);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------ epilogue warps
        const int q = warp & 3;                       // TMEM lane quarter this warp may read
        const int hc = (warp - 2) >> 2;               // which 1 / NH of the channels of a row this thread handles
        const int row = q * 32 + lane;
        const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
        const int ct = a.ntile / 2;
        uint32_t pass = 0;
        for (int unit = blockIdx.x; unit < total_tiles; unit += gridDim.x) {
            const int tile = unit / nsp, n0 = (unit % nsp) * npt;
            const int b = tile / tiles_t, t0 = (tile % tiles_t) * a.nt;
            const int ntv = min(a.nt, a.T - t0);
            const int tl = row / a.P, j = row - tl * a.P, t = t0 + tl;
            for (int op = 0; op < a.n_out_par; ++op)
                for (int nti = n0; nti < n0 + npt; ++nti, ++pass) {
                    const uint32_t buf = pass & 1;
                    mbar_wait(&bar_acc[buf], (pass >> 1) & 1);
                    __syncwarp();
                    tc_fence_after();
                    const uint32_t tcol = tlane + buf * acc_cols;
                    if (a.mode == MODE_LIN) {
                        const bool valid = t < a.T;
                        const int tt = t / a.Bl, bb = t - tt * a.Bl;
                        float* dst = a.out_f32 + ((size_t)tt * a.N_total + (size_t)nti * a.ntile) * a.Bp + bb;
                        const float* bias = sEp + nti * a.ntile;
                        for (int c0 = hc * (a.ntile / NH); c0 < (hc + 1) * (a.ntile / NH); c0 += 32) {   // few, wide TMEM loads
                            float v[32];
                            tmem_ld32(tcol + c0, v);
                            tmem_ld_wait();
                            if (valid) {
#pragma unroll
                                for (int i = 0; i < 32; ++i) dst[(size_t)(c0 + i) * a.Bp] = v[i] + bias[c0 + i];
                            }
                        }
                    } else {
                        const int fo = a.mode == MODE_DEC ? 2 * j + op : j;
                        const bool valid = tl < ntv && j < a.Fo[op];
                        const float* ep = sEp + (size_t)nti * 4 * ct;
                        for (int c0 = hc * (ct / NH); c0 < (hc + 1) * (ct / NH); c0 += 8) {
                            float v[8], g[8], e2[8];
                            tmem_ld8(tcol + c0, v);
                            tmem_ld8(tcol + ct + c0, g);
                            tmem_ld_wait();
                            if (valid) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const float y = (v[i] + ep[c0 + i]) * fast_sigmoid(g[i] + ep[ct + c0 + i]);
                                    v[i] = elu1(fmaf(y, ep[2 * ct + c0 + i], ep[3 * ct + c0 + i]));
                                    e2[i] = a.elu2 ? elu1(v[i]) : v[i];
                                }
                                const int cc = (nti * ct + c0) >> 3, ccn = a.C_out >> 3;
                                if (a.out_so) {
                                    const int Qn = (a.Fo[0] + 1) >> 1;
                                    *reinterpret_cast<uint4*>(a.out_so + ((((size_t)b * ccn + cc) * 2 + (fo & 1)) * ((size_t)a.T * Qn) +
                                                                          (size_t)t * Qn + (fo >> 1)) * 8) = pack8(v);
                                }
                                if (a.out_ug)
                                    *reinterpret_cast<uint4*>(a.out_ug + (((size_t)b * ccn + cc) * ((size_t)a.T * a.ug_P + 1) +
                                                                          (size_t)t * a.ug_P + 1 + fo) * 8) = pack8(e2);
                                if (a.out_xl[0]) {   // conv5 -> LSTM layer-1 operand: group = c / 128, kk = f*128 + (c % 128)
                                    const int grp = cc >> 4, kc = fo * 16 + (cc & 15);
                                    *reinterpret_cast<uint4*>(a.out_xl[grp] + ((size_t)kc * ((size_t)a.T * a.B) + (size_t)t * a.B + b) * 8) =
                                        pack8(v);
                                }
                            }
                        }
                    }
                    tc_fence_before();
                    mbar_arrive(&bar_free[buf]);
                }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

// ============================================================================ LSTM input projection, A operand in tensor memory
// pre[T*B][2048] = x[T*B][512] . W_ih^T + b  (gcrn.py:12-15, the input half of nn.LSTM) as a 128-row-tile GEMM.
// stream_kernel keeps the 128 x 512 A tile in shared memory (128 KB), which leaves a 4-stage weight ring: 48 KB in
// flight per SM = one 16 KB block per ~500 cycles against 256 cycles of MMAs per block (ncu source page: issuer and
// epilogue lanes wait on the ring; 0.114 ms per launch, tensor pipe 21 %).  This layer has no taps, so its A tile can
// live in TENSOR MEMORY instead (tcgen05.mma with [a_tmem]: row m = lane m, 32-bit column c = the K elements 2c, 2c+1;
// 256 columns for K = 512), loaded straight from global memory into registers and stored with tcgen05.st.  Shared
// memory then holds nothing but a 16-stage ring of 8 KB weight blocks (N = 128 n-tiles: the other 256 columns are the two
// accumulators), i.e. 120 KB in flight per SM.
struct LinArgs {
    const __nv_bfloat16* x;       // A: CP8 planes [64][rows][8]
    const __nv_bfloat16* w;       // weight stream [16 n-tiles][64 planes][128][8] (blocks of LIN_PL planes are contiguous)
    const float* bias;            // [2048]
    float* out;                   // [T][2048][Bp], row = t*Bl + b
    int rows, Bl, Bp;
    int rounds, left_split;       // schedule: `rounds` whole tiles per CTA, then the left-over tiles cut into `left_split` n-tile groups
    int debug;                    // measurement hook (PDSE_LIN_DEBUG): 1 = no output stores, 2 = no MMAs, 4 = no weight copies
};
// Work units.  Loading the A tile (global -> registers -> tcgen05.st) cannot overlap the previous unit's MMAs (they read the
// same tensor-memory columns), so it is paid once per UNIT: 8 k cycles against 2.4 k per n-tile.  Cutting every tile into
// n-tile groups to even out the last wave (as stream_kernel does) made that the largest item of the kernel (ablation: with
// stores, MMAs and weight copies all switched off 2/3 of the time remained).  Instead CTA c first does whole tiles
// c, c + G, ... (`rounds` of them), and only the tiles left over after the last full round are cut, into the largest
// power-of-two number of n-tile groups that still gives every unit its own CTA.
struct LinUnit {
    int tile, n0, npt;
};
__device__ __forceinline__ int lin_units(const LinArgs& a, int cta, int G) {
    const int tiles = (a.rows + 127) / 128, left = tiles - a.rounds * G;
    return a.rounds + (cta < left * a.left_split ? 1 : 0);
}
__device__ __forceinline__ LinUnit lin_unit(const LinArgs& a, int cta, int G, int it) {
    if (it < a.rounds) return {it * G + cta, 0, 16};
    const int npt = 16 / a.left_split;
    return {a.rounds * G + cta / a.left_split, (cta % a.left_split) * npt, npt};
}
constexpr int LIN_PL = 16;                 // chunk planes per streamed block = LIN_PL / 2 K = 16 steps: every block costs the issuer lane a
                                           // try_wait + fence + commit (~200 cycles) whatever its size, so blocks are large (8 MMAs each)
constexpr int LIN_ST = 4;                  // ring stages (power of two)
constexpr int LIN_LOG = 2;
constexpr int LIN_BPT = 64 / LIN_PL;       // blocks per n-tile (K = 512 = 64 planes)
constexpr int LIN_BLK = LIN_PL * 128 * 16; // bytes of one block of one 128-wide n-tile
constexpr int LIN_THR = 320;               // warp 0: producer lane, warp 1: MMA issuer lane, warps 2..9: A loaders + epilogue
                                           // (two threads per row: the epilogue's instruction stream, not the MMAs, set the pace)
constexpr int LIN_SMEM = LIN_ST * LIN_BLK + 2048 * 4;

template <int BP>     // batch pitch of the output (compile time: the 128 stores per row and n-tile use immediate offsets)
__global__ void __launch_bounds__(LIN_THR, 1) lin_ta_kernel(LinArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar_full[LIN_ST], bar_empty[LIN_ST], bar_aready, bar_adone, bar_acc[2], bar_free[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* sB = smem;
    float* sBias = reinterpret_cast<float*>(smem + LIN_ST * LIN_BLK);
    for (int i = tid; i < 2048; i += LIN_THR) sBias[i] = __ldg(a.bias + i);
    if (tid == 0) {
        for (int s = 0; s < LIN_ST; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], 1);
        }
        mbar_init(&bar_aready, 256);
        mbar_init(&bar_adone, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_acc[i], 1);
            mbar_init(&bar_free[i], 256);
        }
        fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;            // columns [0,256): A; [256,384) and [384,512): accumulators
    const int n_units = lin_units(a, blockIdx.x, gridDim.x);

    if (warp == 0) {
        // ------------------------------------------------------------ producer lane: never drains across units
        if (lane == 0) {
            uint32_t cnt = 0;
            for (int it = 0; it < n_units; ++it) {
                const LinUnit u = lin_unit(a, blockIdx.x, gridDim.x, it);
                const __nv_bfloat16* wsrc = a.w + (size_t)u.n0 * LIN_BPT * (LIN_BLK / 2);
                for (int i = 0; i < u.npt * LIN_BPT; ++i, ++cnt) {
                    const uint32_t s = cnt & (LIN_ST - 1);
                    if (cnt >= LIN_ST) mbar_wait(&bar_empty[s], ((cnt >> LIN_LOG) - 1) & 1);
                    mbar_arrive_expect_tx(&bar_full[s], LIN_BLK);
                    if (a.debug & 4) {
                        asm volatile("mbarrier.complete_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar_full[s])), "r"(LIN_BLK) : "memory");
                    } else {
                        bulk_g2s(sB + s * LIN_BLK, wsrc + (size_t)i * (LIN_BLK / 2), LIN_BLK, &bar_full[s]);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer lane
        if (lane == 0) {
            uint32_t cnt = 0, pass = 0;
            const uint32_t idesc = make_idesc_op(128, 128);
            for (int it = 0; it < n_units; ++it) {
                const int npt = lin_unit(a, blockIdx.x, gridDim.x, it).npt;
                mbar_wait(&bar_aready, it & 1);              // this unit's A tile is in tensor memory
                tc_fence_after();
                for (int nti = 0; nti < npt; ++nti, ++pass) {
                    const uint32_t buf = pass & 1;
                    if (pass >= 2) mbar_wait(&bar_free[buf], ((pass >> 1) - 1) & 1);   // epilogue drained this accumulator
                    tc_fence_after();
                    const uint32_t d_tmem = tmem + 256 + buf * 128;
                    for (int kblk = 0; kblk < LIN_BPT; ++kblk, ++cnt) {
                        const uint32_t s = cnt & (LIN_ST - 1);
                        mbar_wait(&bar_full[s], (cnt >> LIN_LOG) & 1);
                        tc_fence_after();
                        const uint64_t bdesc = make_smem_desc(smem_u32(sB) + s * LIN_BLK, 128 * 16, 128);
                        if (!(a.debug & 2)) {
#pragma unroll
                            for (int ks = 0; ks < LIN_PL / 2; ++ks)
                                umma_bf16_ta(d_tmem, tmem + (kblk * (LIN_PL / 2) + ks) * 8, dadd(bdesc, ks * 2 * 128 * 16), idesc, (kblk | ks) > 0);
                        }
                        umma_commit(&bar_empty[s]);
                    }
                    umma_commit(&bar_acc[buf]);
                }
                umma_commit(&bar_adone);                      // every MMA that reads this A tile has completed
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------ A loaders + epilogue warps
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int hc = (warp - 2) >> 2;               // which half of the planes (A load) / columns (epilogue) this thread takes
        const int row = q * 32 + lane;
        const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t pass = 0;
        for (int it = 0; it < n_units; ++it) {
            const LinUnit u = lin_unit(a, blockIdx.x, gridDim.x, it);
            const int n0 = u.n0, npt = u.npt;
            const long grow = (long)u.tile * 128 + row;
            const bool valid = grow < a.rows;
            if (it > 0) {
                mbar_wait(&bar_adone, (it - 1) & 1);
                __syncwarp();
                tc_fence_after();
            }
            // A row: 64 chunk planes of 16 bytes -> 256 columns of bf16 pairs; this thread: planes [32 hc, 32 hc + 32),
            // eight planes (32 columns) in flight at a time
            const uint4* src = reinterpret_cast<const uint4*>(a.x) + (valid ? grow : 0) + (size_t)hc * 32 * a.rows;
#pragma unroll 1
            for (int k8 = 0; k8 < 4; ++k8) {
                uint4 u[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) u[j] = valid ? __ldg(src + (size_t)(k8 * 8 + j) * a.rows) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t w8[8] = {u[2 * j].x, u[2 * j].y, u[2 * j].z, u[2 * j].w, u[2 * j + 1].x, u[2 * j + 1].y, u[2 * j + 1].z, u[2 * j + 1].w};
                    tmem_st8(tlane + (hc * 16 + k8 * 4 + j) * 8, reinterpret_cast<const float*>(w8));
                }
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&bar_aready);
            const int tt = (int)(grow / a.Bl), bb = (int)(grow - (long)tt * a.Bl);
            for (int nti = n0; nti < n0 + npt; ++nti, ++pass) {
                const uint32_t buf = pass & 1;
                mbar_wait(&bar_acc[buf], (pass >> 1) & 1);
                __syncwarp();
                tc_fence_after();
                float* dst = a.out + ((size_t)tt * 2048 + (size_t)nti * 128 + hc * 64) * BP + bb;
                const float4* bias4 = reinterpret_cast<const float4*>(sBias + nti * 128 + hc * 64);
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 32) {
                    float v[32];
                    tmem_ld32(tlane + 256 + buf * 128 + hc * 64 + c0, v);
                    tmem_ld_wait();
                    if (valid && !(a.debug & 1)) {
#pragma unroll
                        for (int i4 = 0; i4 < 8; ++i4) {
                            const float4 b4 = bias4[c0 / 4 + i4];
                            dst[(c0 + 4 * i4 + 0) * BP] = v[4 * i4 + 0] + b4.x;
                            dst[(c0 + 4 * i4 + 1) * BP] = v[4 * i4 + 1] + b4.y;
                            dst[(c0 + 4 * i4 + 2) * BP] = v[4 * i4 + 2] + b4.z;
                            dst[(c0 + 4 * i4 + 3) * BP] = v[4 * i4 + 3] + b4.w;
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(&bar_free[buf]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ============================================================================ LSTM recurrence
// One cooperative launch runs a whole layer: grid = (16, G groups).  CTA (c, g) keeps rows
// {gate*512 + 32c + u} of W_hh[g] (128 x 512 bf16 = 128 KB) in shared memory and, per step, computes
// D[128 gate rows][Bp] = W_slice . h_{t-1}^T on the tensor core, adds the precomputed input projection,
// applies the gates and publishes its 32 units of h_t (bf16, CP8 [64][Bp][8]) for the other CTAs.
struct LstmArgs {
    const __nv_bfloat16* whh[2];   // per group: [16][64][128][8]
    const float* pre[2];           // per group: [T][2048][Bp]  (row order n = cta*128 + lane)
    float* hout[2];                // per group: [T*B][512] fp32, row = t*B + b
    __nv_bfloat16* hbuf;           // [G][2][64][Bp][8]  ping-pong h operand
    unsigned int* sync;            // [G] arrival counters (zeroed before launch)
    int B, Bp, T, stage_bytes;
    long long* prof;               // debug: per-phase cycle counters of CTA (0,0) (NULL in production)
};

constexpr int LSTM_THR = 256;   // two threads per gate row (each owns half of the batch columns)

__device__ __forceinline__ float fast_tanh(float x) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x));
    return t;
}

// CLUSTER = true: the 16 CTAs of a group form one thread-block cluster and synchronise each step with the hardware
// cluster barrier (arrive.release / wait.acquire); CLUSTER = false: cooperative launch + release/acquire counter in L2.
template <int BP, bool CLUSTER>
__global__ void __launch_bounds__(LSTM_THR, 1) lstm_rec_kernel(LstmArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar_ld, bar_mma;
    __shared__ uint32_t tmem_slot;
    constexpr int HB = BP / 2;                 // batch columns per thread
    constexpr int NV = HB / 4;                 // float4 per thread per step
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = tid & 127, half = tid >> 7;
    const int c = blockIdx.x, g = blockIdx.y;
    uint8_t* sW = smem;                       // [64][128][16B]
    uint8_t* sH = smem + 131072;              // [64][BP][16B]; reused as gate staging fp32 [4][32][BP+1]
    float* sG = reinterpret_cast<float*>(sH);
    float* sC = reinterpret_cast<float*>(smem + 131072 + a.stage_bytes);   // cell state [BP][32 units]
    constexpr uint32_t tmem_cols = BP <= 32 ? 32 : 64;
    if (tid == 0) {
        mbar_init(&bar_ld, 1);
        mbar_init(&bar_mma, 1);
        fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) tmem_alloc(&tmem_slot, tmem_cols);
    for (int i = tid; i < 32 * BP; i += LSTM_THR) sC[i] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    if (tid == 0) {
        mbar_arrive_expect_tx(&bar_ld, 131072);
        bulk_g2s(sW, a.whh[g] + (size_t)c * 65536, 131072, &bar_ld);
    }
    const float* pre = a.pre[g] + ((size_t)c * 128 + row) * BP + half * HB;   // + t * 2048 * BP
    __nv_bfloat16* hb = a.hbuf + (size_t)g * 2 * 64 * BP * 8;
    unsigned int* cnt = a.sync + g;
    const uint32_t idesc = make_idesc_op(128, BP);
    const int gate = (warp & 3);              // TMEM lane quarter = gate type (i, f, g, o)

    float4 pcur[NV], pnext[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) pcur[i] = __ldg(reinterpret_cast<const float4*>(pre) + i);
    mbar_wait(&bar_ld, 0);
    uint32_t par = 0;

    long long pc[6] = {0, 0, 0, 0, 0, 0}, tk = clock64();
    const bool profiling = a.prof != nullptr && tid == 0 && c == 0 && g == 0;
#define PDSE_TICK(i) if (profiling) { const long long n_ = clock64(); pc[i] += n_ - tk; tk = n_; }
    for (int t = 0; t < a.T; ++t) {
        // prefetch the next step's input projection while this step runs
        if (t + 1 < a.T) {
            const float4* pn = reinterpret_cast<const float4*>(pre + (size_t)(t + 1) * 2048 * BP);
#pragma unroll
            for (int i = 0; i < NV; ++i) pnext[i] = __ldg(pn + i);
        }
        if (t > 0) {
            // wait until all 16 CTAs of this group have published h_{t-1}
            if constexpr (CLUSTER) {
                asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
            } else {
                if (tid == 0) {
                    const unsigned int target = 16u * (unsigned)t;
                    unsigned int v;
                    do {
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(cnt) : "memory");
                    } while (v < target);
                }
                __syncthreads();
            }
            PDSE_TICK(0)   // barrier wait
            const uint4* src = reinterpret_cast<const uint4*>(hb + (size_t)((t - 1) & 1) * 64 * BP * 8);
            constexpr int NL = 64 * BP / LSTM_THR;
            uint4 tmp[NL];
#pragma unroll
            for (int i = 0; i < NL; ++i) tmp[i] = __ldcg(src + tid + i * LSTM_THR);
#pragma unroll
            for (int i = 0; i < NL; ++i) reinterpret_cast<uint4*>(sH)[tid + i * LSTM_THR] = tmp[i];
            PDSE_TICK(1)   // h load issue + smem store
            phase_begin();
            if (tid == 0) {
                // descriptors differ only in the start-address field: one 64-bit add per operand per MMA
                const uint64_t ad = make_smem_desc(smem_u32(sW), 2048, 128), bd = make_smem_desc(smem_u32(sH), BP * 16, 128);
#pragma unroll
                for (int ks = 0; ks < 32; ++ks)
                    umma_bf16(tmem, ad + (uint64_t)(ks * ((2 * 2048) >> 4)), bd + (uint64_t)(ks * ((2 * BP * 16) >> 4)), idesc, ks > 0);
            }
            phase_end(&bar_mma, par);
            PDSE_TICK(2)   // sync + MMA
        }
        // gate pre-activations of row `row`, batch half `half` -> activation -> staging [gate][unit][BP+1]
#pragma unroll
        for (int b0 = 0; b0 < HB; b0 += 16) {
            float v[16];
            if (t > 0) {
                tmem_ld16(trow + half * HB + b0, v);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = 0.f;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 pv = pcur[b0 / 4 + i];
                v[4 * i + 0] += pv.x;
                v[4 * i + 1] += pv.y;
                v[4 * i + 2] += pv.z;
                v[4 * i + 3] += pv.w;
            }
            float* dst = sG + (gate * 32 + (row & 31)) * (BP + 1) + half * HB + b0;
#pragma unroll
            for (int i = 0; i < 16; ++i) dst[i] = gate == 2 ? fast_tanh(v[i]) : fast_sigmoid(v[i]);
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) pcur[i] = pnext[i];
        tc_fence_before();
        __syncthreads();
        PDSE_TICK(3)   // gate epilogue + sync
        // cell update: thread (unit = lane, bq = warp) owns batch entries b = bq + 8*i
        __nv_bfloat16* hdst = hb + (size_t)(t & 1) * 64 * BP * 8;
        float* hout = a.hout[g];
        const int unit = c * 32 + lane;
        float hv[BP / 8];
        op_t* sHb = reinterpret_cast<op_t*>(sC + 32 * BP);   // [4 planes][BP][8] staging of this CTA's h slice
#pragma unroll
        for (int i = 0; i < BP / 8; ++i) {
            const int b = warp + 8 * i;
            const float gi = sG[(0 * 32 + lane) * (BP + 1) + b], gf = sG[(1 * 32 + lane) * (BP + 1) + b];
            const float gg = sG[(2 * 32 + lane) * (BP + 1) + b], go = sG[(3 * 32 + lane) * (BP + 1) + b];
            const float cn = gf * sC[b * 32 + lane] + gi * gg;   // [b][unit]: conflict-free across the warp
            sC[b * 32 + lane] = cn;
            hv[i] = go * fast_tanh(cn);
            sHb[((lane >> 3) * BP + b) * 8 + (lane & 7)] = to_op(hv[i]);
        }
        __syncthreads();
        // the slice (units 32c..32c+31 = chunk planes 4c..4c+3) is contiguous in the exchange buffer: 16-byte stores
        {
            uint4* dst = reinterpret_cast<uint4*>(hdst + (size_t)c * 4 * BP * 8);
            const uint4* srcs = reinterpret_cast<const uint4*>(sHb);
            for (int i = tid; i < 4 * BP; i += LSTM_THR) dst[i] = srcs[i];
        }
        PDSE_TICK(4)   // cell update + h stores
        // publish h_t to the other CTAs of the group, then (off the critical path) write the fp32 copy for LayerNorm
        if constexpr (CLUSTER) {
            __syncwarp();
            asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        } else {
            __syncthreads();
            if (tid == 0) {
                __threadfence();
                atomicAdd(cnt, 1u);
            }
        }
#pragma unroll
        for (int i = 0; i < BP / 8; ++i) {
            const int b = warp + 8 * i;
            if (b < a.B) hout[((size_t)t * a.B + b) * 512 + unit] = hv[i];
        }
    }
    if constexpr (CLUSTER) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (profiling) {
        PDSE_TICK(5)
        for (int i = 0; i < 6; ++i) a.prof[i] = pc[i];
    }
#undef PDSE_TICK
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

// ---------------------------------------------------------------------------- DSMEM variant
// The 16 CTAs of a (group, 32-sequence batch half) form one thread-block cluster and exchange h_t WITHOUT global
// memory: every CTA stages its 32 units x 32 sequences (2 KB, already in the MMA's B-operand layout) in shared
// memory and pushes it into all 16 peers with cp.async.bulk shared::cta -> shared::cluster; the copies complete_tx on
// the destination's mbarrier, so "data arrived" and "barrier" are the same event and the receiver issues its MMAs
// straight from the landed planes.  h buffers and the staging slice are double-buffered (a peer can only be one
// step ahead because it needs this CTA's h_{t-1} to produce h_t).
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t cta_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void bulk_s2cluster(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
                 "r"(src_cta), "r"(bytes), "r"(bar_cluster)
                 : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// BP = sequences per cluster = N of the recurrence MMA.  The MMA issue floor does not depend on N (~50-62 cycles per
// K = 16 instruction for N <= 64) while the gate / cell epilogue and the h push scale with it, so 16
// sequences per cluster on 8 clusters (128 CTAs) beat 32 on 4 (64 CTAs of 148): tests/gpu_lstm_prof.py.
template <int BP> struct LdCfg {
    static constexpr int SLICE = 4 * BP * 16;          // bytes one CTA contributes to h_t (4 chunk planes)
    static constexpr int HBUF = 64 * BP * 16;          // one full h operand
    static constexpr int SMEM = 131072 + 2 * HBUF + 128 * (BP + 1) * 4 + 32 * BP * 4 + 2 * SLICE;
};

template <int BP>
__global__ void __launch_bounds__(LSTM_THR, 1) lstm_dsmem_kernel(LstmArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar_ld, bar_mma, h_bar[2];
    __shared__ uint32_t tmem_slot;
    constexpr int LD_SLICE = LdCfg<BP>::SLICE, LD_HBUF = LdCfg<BP>::HBUF;
    constexpr int HC = BP / 2;                                 // batch columns per thread (two threads per gate row)
    constexpr int NV = HC / 4;                                 // float4 of the input projection per thread and step
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = tid & 127, halfc = tid >> 7;
    const int c = blockIdx.x;                                  // rank in the cluster = slice of hidden units
    const int parts = gridDim.y / 2, g = blockIdx.y / parts, bh = blockIdx.y - g * parts;
    uint8_t* sW = smem;                                        // [64][128][16B]
    uint8_t* sH = sW + 131072;                                 // 2 x [64][BP][16B]
    float* sG = reinterpret_cast<float*>(sH + 2 * LD_HBUF);    // gate staging [4][32][BP+1]
    float* sC = sG + 128 * (BP + 1);                           // cell state [BP][32]
    uint8_t* sOut = reinterpret_cast<uint8_t*>(sC + 32 * BP);  // 2 x [4][BP][16B]
    if (tid == 0) {
        mbar_init(&bar_ld, 1);
        mbar_init(&bar_mma, 1);
        mbar_init(&h_bar[0], 1);
        mbar_init(&h_bar[1], 1);
        fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) tmem_alloc(&tmem_slot, 32);
    for (int i = tid; i < 32 * BP; i += LSTM_THR) sC[i] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    if (tid == 0) {
        mbar_arrive_expect_tx(&bar_ld, 131072);
        bulk_g2s(sW, a.whh[g] + (size_t)c * 65536, 131072, &bar_ld);
    }
    const float* pre = a.pre[g] + ((size_t)c * 128 + row) * a.Bp + bh * BP + halfc * HC;   // + t * 2048 * Bp
    const uint32_t idesc = make_idesc_op(128, BP);
    const int gate = warp & 3;
    float4 pcur[NV], pnext[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) pcur[i] = __ldg(reinterpret_cast<const float4*>(pre) + i);
    mbar_wait(&bar_ld, 0);
    cluster_sync_all();                    // every peer's mbarriers are initialised before anyone pushes
    uint32_t par = 0;
    float* hout = a.hout[g];
    const int unit = c * 32 + lane;
    const bool profiling = a.prof != nullptr && tid == 0 && c == 0 && blockIdx.y == 0;
    long long pc[6] = {0, 0, 0, 0, 0, 0}, tk = profiling ? clock64() : 0;
#define PDSE_TICK(i) if (profiling) { const long long n_ = clock64(); pc[i] += n_ - tk; tk = n_; }

    for (int t = 0; t < a.T; ++t) {
        if (t + 1 < a.T) {
            const float4* pn = reinterpret_cast<const float4*>(pre + (size_t)(t + 1) * 2048 * a.Bp);
#pragma unroll
            for (int i = 0; i < NV; ++i) pnext[i] = __ldg(pn + i);
        }
        if (t > 0) {
            const int bin = (t - 1) & 1;
            mbar_wait(&h_bar[bin], ((t - 1) >> 1) & 1);        // all 16 slices of h_{t-1} have landed in sH[bin]
            PDSE_TICK(0)   // wait for the peers' pushes
            tc_fence_before();
            __syncthreads();                                   // (also: everyone is done with the previous TMEM reads)
            tc_fence_after();
            PDSE_TICK(1)   // CTA barrier
            if (tid == 0) {
                const uint64_t ad = make_smem_desc(smem_u32(sW), 2048, 128);
                const uint64_t bd = make_smem_desc(smem_u32(sH) + bin * LD_HBUF, BP * 16, 128);
#pragma unroll
                for (int ks = 0; ks < 32; ++ks)
                    umma_bf16(tmem, ad + (uint64_t)(ks * ((2 * 2048) >> 4)), bd + (uint64_t)(ks * ((2 * BP * 16) >> 4)), idesc, ks > 0);
            }
            phase_end(&bar_mma, par);
            PDSE_TICK(2)   // MMA issue + completion
        }
        {   // gate pre-activations of row `row`, HC batch columns -> activation -> staging [gate][unit][BP+1]
            float v[HC];
            if (t > 0) {
                if constexpr (HC == 16) tmem_ld16(trow + halfc * HC, v);
                else tmem_ld8(trow + halfc * HC, v);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int i = 0; i < HC; ++i) v[i] = 0.f;
            }
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                v[4 * i + 0] += pcur[i].x;
                v[4 * i + 1] += pcur[i].y;
                v[4 * i + 2] += pcur[i].z;
                v[4 * i + 3] += pcur[i].w;
            }
            float* dst = sG + (gate * 32 + (row & 31)) * (BP + 1) + halfc * HC;
#pragma unroll
            for (int i = 0; i < HC; ++i) dst[i] = gate == 2 ? fast_tanh(v[i]) : fast_sigmoid(v[i]);
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) pcur[i] = pnext[i];
        tc_fence_before();
        __syncthreads();
        PDSE_TICK(3)   // gate epilogue + barrier
        // cell update: thread (unit = lane, bq = warp) owns batch entries b = bq + 8*i
        op_t* so = reinterpret_cast<op_t*>(sOut + (t & 1) * LD_SLICE);
        float hv[BP / 8];
#pragma unroll
        for (int i = 0; i < BP / 8; ++i) {
            const int b = warp + 8 * i;
            const float gi = sG[(0 * 32 + lane) * (BP + 1) + b], gf = sG[(1 * 32 + lane) * (BP + 1) + b];
            const float gg = sG[(2 * 32 + lane) * (BP + 1) + b], go = sG[(3 * 32 + lane) * (BP + 1) + b];
            const float cn = gf * sC[b * 32 + lane] + gi * gg;
            sC[b * 32 + lane] = cn;
            hv[i] = go * fast_tanh(cn);
            so[((lane >> 3) * BP + b) * 8 + (lane & 7)] = to_op(hv[i]);
        }
        if (t + 1 < a.T) {
            fence_proxy_async_smem();          // the staged slice is read by the async proxy (bulk copy)
            __syncthreads();
            const int bout = t & 1;
            if (tid == 0) mbar_arrive_expect_tx(&h_bar[bout], 16 * LD_SLICE);   // my own inbox for h_t
            // push my slice into all 16 peers (including myself).  A bulk copy is a uniform-datapath instruction: lanes
            // of one warp that issue it with different operands are serialised (~65 cycles each, 1.07 k cycles when 16
            // lanes of warp 0 did all the pushes), so every warp pushes to two peers
            if (lane < 2) {
                const int peer = warp * 2 + lane;
                bulk_s2cluster(mapa_cluster(smem_u32(sH) + bout * LD_HBUF + c * LD_SLICE, peer), smem_u32(so), LD_SLICE,
                               mapa_cluster(smem_u32(&h_bar[bout]), peer));
            }
        }
        PDSE_TICK(4)   // cell update + staging + push
#pragma unroll
        for (int i = 0; i < BP / 8; ++i) {
            const int bg = bh * BP + warp + 8 * i;
            if (bg < a.B) hout[((size_t)t * a.B + bg) * 512 + unit] = hv[i];
        }
        PDSE_TICK(5)   // h stores
    }
    if (profiling)
        for (int i = 0; i < 6; ++i) a.prof[i] = pc[i];
#undef PDSE_TICK
    cluster_sync_all();                        // no CTA leaves while a peer may still push into it
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 32);
}

// ---------------------------------------------------------------------------- DSMEM variant, two interleaved streams
// The tcgen05 issue floor makes one recurrence step cost the same 32 x ~50 cycles of tensor-pipe time for 16 as for 32
// sequences, and a B200 holds only 7 co-resident 16-CTA clusters (pdse_debug_lstm_clusters), so "16 sequences per
// cluster on 8 clusters" does not fit.  Instead every cluster runs TWO independent 16-sequence recurrences (streams)
// through the same resident W_hh slice: a dedicated issuer lane (warp 8) queues the 32 MMAs of stream s, step t as soon
// as that stream's h_{t-1} has landed, so the MMAs of one stream run underneath the gate / cell epilogue, the DSMEM push
// and the landing latency of the other.  Per stream everything is as in lstm_dsmem_kernel: h_t slices pushed into all 16
// peers' next-step operand by cp.async.bulk shared::cta -> shared::cluster with complete_tx on the receiver's mbarrier.
constexpr int L2_BP = 16;                         // sequences per stream
constexpr int L2_NI = 2;                          // MMA issuer lanes (each a K slice of every step's GEMM, an accumulator of its own)
constexpr int L2_THR = 256 + 32 * L2_NI;           // warps 0..7: epilogue (two threads per gate row); then the issuer warps
constexpr int L2_SLICE = 4 * L2_BP * 16;          // bytes one CTA contributes to one stream's h_t
constexpr int L2_HBUF = 64 * L2_BP * 16;          // one h operand of one stream
constexpr int L2_SMEM = 131072 + 4 * L2_HBUF + 128 * (L2_BP + 1) * 4 + 2 * 32 * L2_BP * 4 + 4 * L2_SLICE;

__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__global__ void __launch_bounds__(L2_THR, 1) lstm_dsmem2_kernel(LstmArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar_ld, bar_mma[2], h_bar[2][2];
    __shared__ uint32_t tmem_slot;
    constexpr int BP = L2_BP;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c = blockIdx.x;                                  // rank in the cluster = slice of hidden units
    const int parts = gridDim.y / 2, g = blockIdx.y / parts, bh = blockIdx.y - g * parts;
    const int ns = a.B - bh * 2 * BP > BP ? 2 : 1;             // streams with at least one live sequence
    uint8_t* sW = smem;                                        // [64][128][16B]
    uint8_t* sH = sW + 131072;                                 // [stream][bin] x [64][BP][16B]
    float* sG = reinterpret_cast<float*>(sH + 4 * L2_HBUF);    // gate staging [4][32][BP+1] (the streams' epilogues alternate)
    float* sC = sG + 128 * (BP + 1);                           // cell state [stream][BP][32]
    uint8_t* sOut = reinterpret_cast<uint8_t*>(sC + 2 * 32 * BP);   // [stream][bin] x [4][BP][16B]
    if (tid == 0) {
        mbar_init(&bar_ld, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bar_mma[s], L2_NI);         // one commit per issuer lane
            mbar_init(&h_bar[s][0], 1);
            mbar_init(&h_bar[s][1], 1);
        }
        fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) tmem_alloc(&tmem_slot, 2 * 16 * L2_NI);
    for (int i = tid; i < 2 * 32 * BP; i += L2_THR) sC[i] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        mbar_arrive_expect_tx(&bar_ld, 131072);
        bulk_g2s(sW, a.whh[g] + (size_t)c * 65536, 131072, &bar_ld);
    }
    mbar_wait(&bar_ld, 0);
    cluster_sync_all();                    // every peer's mbarriers are initialised before anyone pushes

    if (warp >= 8) {
        // ------------------------------------------------------------------ MMA issuer lanes.  One thread gets a 128 x N x 16
        // tcgen05.mma out every ~55-62 cycles whatever N; two issuing warps together reach ~43 (tests/gpu_probe_mma_dep.py).
        // All lanes work on the SAME stream: lane h issues the K slice [512 h / L2_NI, 512 (h + 1) / L2_NI) into an accumulator
        // of its own (the epilogue adds them up), so the two streams keep alternating
        const int kh = warp - 8;
        if (lane == 0) {
            const uint32_t idesc = make_idesc_op(128, BP);
            const uint64_t ad = make_smem_desc(smem_u32(sW), 2048, 128);
            for (int t = 1; t < a.T; ++t) {
                const int bin = (t - 1) & 1;
                for (int s = 0; s < ns; ++s) {
                    // all 16 slices of this stream's h_{t-1} have landed (that includes this CTA's own push, which its
                    // epilogue threads issue after they have read the previous accumulators: no separate "free" signal)
                    mbar_wait(&h_bar[s][bin], ((t - 1) >> 1) & 1);
                    tc_fence_after();
                    const uint64_t bd = make_smem_desc(smem_u32(sH) + (s * 2 + bin) * L2_HBUF, BP * 16, 128);
#pragma unroll
                    for (int k = 0; k < 32 / L2_NI; ++k) {
                        const int ks = kh * (32 / L2_NI) + k;
                        umma_bf16(tmem + (s * L2_NI + kh) * 16, ad + (uint64_t)(ks * ((2 * 2048) >> 4)), bd + (uint64_t)(ks * ((2 * BP * 16) >> 4)), idesc, k > 0);
                    }
                    umma_commit(&bar_mma[s]);
                }
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue threads
        const int row = tid & 127, halfc = tid >> 7;            // two threads per gate row, 8 batch columns each
        const int gate = warp & 3;
        const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const float* pre = a.pre[g] + ((size_t)c * 128 + row) * a.Bp + bh * 2 * BP + halfc * 8;   // + s*BP + t * 2048 * Bp
        float* hout = a.hout[g];
        const int unit = c * 32 + lane;
        float4 pcur[2][2], pnext[2][2];
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int i = 0; i < 2; ++i) pcur[s][i] = __ldg(reinterpret_cast<const float4*>(pre + s * BP) + i);
        uint32_t par[2] = {0u, 0u};
        const bool profiling = a.prof != nullptr && tid == 0 && c == 0 && blockIdx.y == 0;
        long long pc[6] = {0, 0, 0, 0, 0, 0}, tk = profiling ? clock64() : 0;
#define PDSE_TICK(i) if (profiling) { const long long n_ = clock64(); pc[i] += n_ - tk; tk = n_; }
        for (int t = 0; t < a.T; ++t) {
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (s >= ns) break;
                if (t + 1 < a.T) {
                    const float4* pn = reinterpret_cast<const float4*>(pre + s * BP + (size_t)(t + 1) * 2048 * a.Bp);
#pragma unroll
                    for (int i = 0; i < 2; ++i) pnext[s][i] = __ldg(pn + i);
                }
                float v[8];
                if (t > 0) {
                    mbar_wait(&bar_mma[s], par[s]);
                    par[s] ^= 1u;
                    __syncwarp();
                    tc_fence_after();
                    PDSE_TICK(0)   // wait for this stream's MMAs
                    float vp[L2_NI - 1][8];
                    tmem_ld8(trow + s * L2_NI * 16 + halfc * 8, v);
#pragma unroll
                    for (int h = 1; h < L2_NI; ++h) tmem_ld8(trow + (s * L2_NI + h) * 16 + halfc * 8, vp[h - 1]);
                    tmem_ld_wait();
#pragma unroll
                    for (int h = 1; h < L2_NI; ++h)
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] += vp[h - 1][i];
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = 0.f;
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    v[4 * i + 0] += pcur[s][i].x;
                    v[4 * i + 1] += pcur[s][i].y;
                    v[4 * i + 2] += pcur[s][i].z;
                    v[4 * i + 3] += pcur[s][i].w;
                    pcur[s][i] = pnext[s][i];
                }
                float* dst = sG + (gate * 32 + (row & 31)) * (BP + 1) + halfc * 8;
#pragma unroll
                for (int i = 0; i < 8; ++i) dst[i] = gate == 2 ? fast_tanh(v[i]) : fast_sigmoid(v[i]);
                tc_fence_before();
                epi_sync();
                PDSE_TICK(1)   // gates + barrier
                // cell update: thread (unit = lane, bq = warp) owns batch entries b = bq + 8*i
                op_t* so = reinterpret_cast<op_t*>(sOut + (s * 2 + (t & 1)) * L2_SLICE);
                float* sCs = sC + s * 32 * BP;
                float hv[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int b = warp + 8 * i;
                    const float gi = sG[(0 * 32 + lane) * (BP + 1) + b], gf = sG[(1 * 32 + lane) * (BP + 1) + b];
                    const float gg = sG[(2 * 32 + lane) * (BP + 1) + b], go = sG[(3 * 32 + lane) * (BP + 1) + b];
                    const float cn = gf * sCs[b * 32 + lane] + gi * gg;
                    sCs[b * 32 + lane] = cn;
                    hv[i] = go * fast_tanh(cn);
                    so[((lane >> 3) * BP + b) * 8 + (lane & 7)] = to_op(hv[i]);
                }
                fence_proxy_async_smem();          // the staged slice is read by the async proxy (bulk copy)
                epi_sync();                        // (also: the gate staging is free for the other stream)
                if (t + 1 < a.T) {
                    const int bout = t & 1;
                    if (tid == 0) mbar_arrive_expect_tx(&h_bar[s][bout], 16 * L2_SLICE);   // my own inbox for h_t
                    // every warp pushes to two peers (bulk copies issued by lanes of one warp serialise)
                    if (lane < 2) {
                        const int peer = warp * 2 + lane;
                        bulk_s2cluster(mapa_cluster(smem_u32(sH) + (s * 2 + bout) * L2_HBUF + c * L2_SLICE, peer), smem_u32(so),
                                       L2_SLICE, mapa_cluster(smem_u32(&h_bar[s][bout]), peer));
                    }
                }
                PDSE_TICK(2)   // cell update + staging + push
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int bg = (bh * 2 + s) * BP + warp + 8 * i;
                    if (bg < a.B) hout[((size_t)t * a.B + bg) * 512 + unit] = hv[i];
                }
                PDSE_TICK(3)   // h stores
            }
        }
        if (profiling)
            for (int i = 0; i < 6; ++i) a.prof[i] = pc[i];
#undef PDSE_TICK
    }
    cluster_sync_all();                        // no CTA leaves while a peer may still push into it
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 2 * 16 * L2_NI);
}

// ============================================================================ LayerNorm + shuffles
// mode 1 (gcrn.py:29-31): features' = 2*j + g  -> LN1 -> layer-2 operand XL2[g'][64][rows][8]
// mode 2 (gcrn.py:33-38): features'' = 512*g' + j -> LN2 -> UG planes (256 ch, F = 4): feature = c*4 + f
struct LnArgs {
    const float* h[2];      // [rows][512], row = t*B + b
    const float* w;         // [1024]
    const float* bia;       // [1024]
    __nv_bfloat16* xl[2];   // mode 1
    __nv_bfloat16* ug;      // mode 2: [B][32][T*5+1][8]
    int rows, B, T, mode;
};

constexpr int LN_PITCH = 1028;     // row pitch of the staging tile: 4 banks of skew per row for the row-crossing store phase

__global__ void __launch_bounds__(256) ln_kernel(LnArgs a) {
    __shared__ __align__(16) float sv[8][LN_PITCH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * 8, row = row0 + warp;
    float* v = sv[warp];
    if (row < a.rows) {
        // one warp per row; both groups' h rows as float4 (512 B per load instruction)
        float sum = 0.f;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const float4* src = reinterpret_cast<const float4*>(a.h[g] + (size_t)row * 512);
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int j4 = it * 32 + lane;
                const float4 x = __ldg(src + j4);
                sum += (x.x + x.y) + (x.z + x.w);
                if (a.mode == 1) {            // feature 2 j + g
                    v[2 * (4 * j4 + 0) + g] = x.x;
                    v[2 * (4 * j4 + 1) + g] = x.y;
                    v[2 * (4 * j4 + 2) + g] = x.z;
                    v[2 * (4 * j4 + 3) + g] = x.w;
                } else {                      // feature 512 g + j
                    *reinterpret_cast<float4*>(v + 512 * g + 4 * j4) = x;
                }
            }
        }
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mean = sum * (1.f / 1024.f);
        __syncwarp();
        float var = 0.f;
        for (int i = lane; i < 1024; i += 32) {
            const float d = v[i] - mean;
            var = fmaf(d, d, var);
        }
        for (int o = 16; o; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
        const float rstd = rsqrtf(var * (1.f / 1024.f) + 1e-5f);
        for (int i = lane; i < 1024; i += 32) v[i] = fmaf((v[i] - mean) * rstd, __ldg(a.w + i), __ldg(a.bia + i));
    }
    if (a.mode == 1) {
        // layer-2 operand XL2[g'][kc][row][8]: for one chunk plane the CTA's 8 rows are 128 contiguous bytes, so the store
        // phase runs across rows (thread = (chunk, row)) instead of one row per warp (32 scattered 16-byte stores)
        __syncthreads();
        const int r = threadIdx.x & 7;
        if (row0 + r < a.rows) {
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int ch = (threadIdx.x >> 3) + 32 * it, g2 = ch >> 6, kc = ch & 63;
                *reinterpret_cast<uint4*>(a.xl[g2] + ((size_t)kc * a.rows + row0 + r) * 8) = pack8(sv[r] + ch * 8);
            }
        }
    } else if (row < a.rows) {
        __syncwarp();
        const int t = row / a.B, b = row - t * a.B;
        for (int ch = lane; ch < 128; ch += 32) {          // (cc, f): 8 channels c = 8cc..8cc+7 at frequency f
            const int cc = ch >> 2, f = ch & 3;
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = v[(cc * 8 + i) * 4 + f];
            *reinterpret_cast<uint4*>(a.ug + (((size_t)b * 32 + cc) * ((size_t)a.T * 5 + 1) + (size_t)t * 5 + 1 + f) * 8) = pack8(o);
        }
    }
}

// ============================================================================ conv1_t + BN + ELU + fc
// gcrn.py:154,160,162-163: d1 = ELU(bn(GLU convT(32 -> 1, k3, s2))) ; out = fc(d1) ; X_init = out / 11
struct GOutArgs {
    const __nv_bfloat16* d2[2];   // per branch: UG F=80 (P=81), 16 ch
    const __nv_bfloat16* e1;      // UG F=80, 16 ch (ELU'd skip)
    const float* wf[2];           // per branch: wv[32][3] wg[32][3] misc[4] fcw[161][161] fcb[164]
    float* xinit;                 // [B][2][T][161]
    int B, T;
};
constexpr int OUT_FR = 32;        // frames per CTA = two M = 16 tiles of the fc MMA (each weight fragment is used twice)
constexpr int OUT_ST = 16;        // frames staged at a time for the transposed conv
constexpr int OUT_DP = 40;        // pitch of the [bin][frame] tile: conflict-free A fragments
constexpr int OUT_CH = OUT_ST * 81 + 1;   // staged 16-byte units per channel chunk: OUT_ST frames of 81 positions + the closing guard
constexpr int OUT_SMEM = 4 * OUT_CH * 16 + 168 * OUT_DP * 4 + 196 * 4;

__global__ void __launch_bounds__(192) gout_kernel(GOutArgs a) {
    extern __shared__ __align__(16) uint8_t osm[];
    __shared__ uint64_t bar_in;
    uint4* sin_ = reinterpret_cast<uint4*>(osm);                                   // [4 chunks][OUT_ST frames x 81 positions + 1]
    float* sd1 = reinterpret_cast<float*>(osm + 4 * OUT_CH * 16);                  // [bin 0..167][OUT_DP]
    float* scw = sd1 + 168 * OUT_DP;                                               // conv weights + BN (196 floats)
    const int tid = threadIdx.x, br = blockIdx.z, b = blockIdx.y, t0 = blockIdx.x * OUT_FR;
    const float* wf = a.wf[br];
    const size_t rows = (size_t)a.T * 81 + 1;
    if (tid == 0) {
        mbar_init(&bar_in, 1);
        fence_mbar_init();
    }
    for (int i = tid; i < 196; i += 192) scw[i] = __ldg(wf + i);
    for (int i = tid; i < 7 * OUT_DP; i += 192) sd1[161 * OUT_DP + i] = 0.f;      // K padding rows 161..167
    for (int pass = 0; pass < OUT_FR / OUT_ST; ++pass) {
        const int tp = t0 + pass * OUT_ST;
        if (tp >= a.T) break;
        __syncthreads();      // the previous pass no longer reads the staging buffer (and scw / the barrier are set up)
        // stage [d2 | e1] for OUT_ST frames: in the unsplit guarded layout the frames of one channel chunk are one contiguous
        // run (81 positions each: leading guard + 80 values; the next frame's guard closes the last one) -> one bulk copy per
        // chunk.  (Per-thread 16-byte loads left only two loads in flight per thread: half of the kernel time.)
        if ((tid & 31) == 0 && (tid >> 5) < 4) {
            const int cc = tid >> 5;
            const uint32_t bytes = (uint32_t)(min(OUT_ST, a.T - tp) * 81 + 1) * 16;
            if (cc == 0) mbar_arrive_expect_tx(&bar_in, 4 * bytes);
            const __nv_bfloat16* src = cc < 2 ? a.d2[br] + (((size_t)b * 2 + cc) * rows + (size_t)tp * 81) * 8
                                              : a.e1 + (((size_t)b * 2 + (cc - 2)) * rows + (size_t)tp * 81) * 8;
            bulk_g2s(sin_ + cc * OUT_CH, src, bytes, &bar_in);
        }
        mbar_wait(&bar_in, pass & 1);
        const float bv = scw[192], bg = scw[193], bs = scw[194], bsh = scw[195];
        // Transposed conv (32 -> 1 value + 1 gate, k3, s2) as a GEMM on mma.sync m16n8k16 bf16: rows = input bin j,
        // K = (tap h[j] | tap h[j-1]) x 32 channels = 64, N = (v_even, g_even, v_odd, g_odd, 0...): output bin 2j reads h[j]
        // through tap 0 and h[j-1] through tap 2, bin 2j+1 reads h[j] through tap 1.  The A fragments are 32-bit words of
        // the staged 16-byte channel units; the fp32 taps are split into bf16 hi + lo (two MMAs), so the products are exact.
        {
            const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
            uint32_t bh[4][2], bl[4][2];          // B fragments of the 4 k-steps: rows k = 2t, 2t+1 (and +8), column n = g
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float w2[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int c = (s4 & 1) * 16 + h * 8 + 2 * t + e, tapB = s4 >> 1;
                        // n = g: 0 v_even 1 g_even 2 v_odd 3 g_odd; tap A (h[j]) feeds taps 0 (even) / 1 (odd), tap B (h[j-1]) tap 2 (even only)
                        float w = 0.f;
                        if (g < 4 && !(tapB && g >= 2)) w = scw[(g & 1) * 96 + c * 3 + (tapB ? 2 : (g >> 1))];
                        w2[e] = w;
                    }
                    const op_t h0 = to_op(w2[0]), h1 = to_op(w2[1]);
                    const op_t l0 = to_op(w2[0] - op_to_float(h0)), l1 = to_op(w2[1] - op_to_float(h1));
                    bh[s4][h] = (uint32_t)op_bits(h0) | ((uint32_t)op_bits(h1) << 16);
                    bl[s4][h] = (uint32_t)op_bits(l0) | ((uint32_t)op_bits(l1) << 16);
                }
            const uint32_t* sw = reinterpret_cast<const uint32_t*>(sin_);
            for (int item = warp; item < OUT_ST * 6; item += 6) {
                const int fr = item / 6, m0 = (item - fr * 6) * 16;
                float d[4] = {t < 2 ? bv : 0.f, t < 2 ? bg : 0.f, t < 2 ? bv : 0.f, t < 2 ? bg : 0.f};
#pragma unroll
                for (int s4 = 0; s4 < 4; ++s4) {
                    const int cc = (s4 & 1) * 2, pos = m0 + g + 1 - (s4 >> 1);
                    const uint32_t* r0 = sw + ((size_t)(cc * OUT_CH + fr * 81 + pos) * 4 + t);
                    const uint32_t a0 = r0[0], a1 = r0[8 * 4], a2 = r0[OUT_CH * 4], a3 = r0[OUT_CH * 4 + 8 * 4];
                    mma_bf16_16816(d, a0, a1, a2, a3, bh[s4][0], bh[s4][1]);
                    mma_bf16_16816(d, a0, a1, a2, a3, bl[s4][0], bl[s4][1]);
                }
                if (t < 2) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int fo = 2 * (m0 + g + 8 * h) + t;
                        if (fo < 161) {
                            const float y = d[2 * h] / (1.f + __expf(-d[2 * h + 1]));
                            sd1[fo * OUT_DP + pass * OUT_ST + fr] = elu1(fmaf(y, bs, bsh));   // transposed: [bin][frame]
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    // fc (gcrn.py:162-163) as [32 frames x 161] x [161 x 161] on mma.sync m16n8k8 with 3xTF32 operand splitting
    // (a_hi w_hi + a_lo w_hi + a_hi w_lo): fp32-level accuracy for the last linear layer of the prior.  Two M = 16 tiles
    // per weight fragment: the weights come from L2 and their loads bound this phase
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const float* fcw = wf + 196;
    const float* fcb = fcw + 161 * 161;
    for (int nt = warp; nt < 21; nt += 6) {
        const int n0 = nt * 8, n = n0 + g;
        const bool nv = n < 161;
        float d[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll 3
        for (int ks = 0; ks < 21; ++ks) {
            const int k0 = ks * 8;
            const float w0 = (nv && k0 + t < 161) ? __ldg(fcw + (k0 + t) * 161 + n) : 0.f;
            const float w1 = (nv && k0 + t + 4 < 161) ? __ldg(fcw + (k0 + t + 4) * 161 + n) : 0.f;
            const float w0h = to_tf32(w0), w1h = to_tf32(w1);
            const float w0l = to_tf32(w0 - w0h), w1l = to_tf32(w1 - w1h);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                float ah[4], al[4];
                const float* r0 = sd1 + (k0 + t) * OUT_DP + mt * 16 + g;
                const float* r1 = sd1 + (k0 + t + 4) * OUT_DP + mt * 16 + g;
                const float a0 = r0[0], a1 = r0[8], a2 = r1[0], a3 = r1[8];
                ah[0] = to_tf32(a0); ah[1] = to_tf32(a1); ah[2] = to_tf32(a2); ah[3] = to_tf32(a3);
                al[0] = to_tf32(a0 - ah[0]); al[1] = to_tf32(a1 - ah[1]); al[2] = to_tf32(a2 - ah[2]); al[3] = to_tf32(a3 - ah[3]);
                mma_tf32(d[mt], ah, w0h, w1h);
                mma_tf32(d[mt], al, w0h, w1h);
                mma_tf32(d[mt], ah, w0l, w1l);
            }
        }
        const int fo = n0 + 2 * t;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int tt = t0 + mt * 16 + g + 8 * h;
                if (tt < a.T) {
                    float* o = a.xinit + (((size_t)b * 2 + br) * a.T + tt) * 161;
                    if (fo < 161) o[fo] = d[mt][2 * h] + __ldg(fcb + fo);
                    if (fo + 1 < 161) o[fo + 1] = d[mt][2 * h + 1] + __ldg(fcb + fo + 1);
                }
            }
    }
}

}  // namespace pdse

// ============================================================================ C ABI
using namespace pdse;

// gcrn.py:137 conv1 + bn1 + ELU.  y [B][2][T][161] fp32 -> SO(F=80) and UG(F=80, ELU'd twice) 16-channel maps
extern "C" int pdse_gcrn_conv1_fwd(const float* y, void* out_so, void* out_ug, const void* wb, const float* ep, int B,
                                   int T, void* stream) {
    if (B <= 0 || T <= 0) return set_error("pdse_gcrn_conv1_fwd: empty input");
    GConv1Args a{y, (__nv_bfloat16*)out_so, (__nv_bfloat16*)out_ug, (const __nv_bfloat16*)wb, ep, B, T};
    const int tiles = B * ((T * 80 + 127) / 128);
    gconv1_kernel<<<min(tiles, sm_count() * 8), 128, 0, (cudaStream_t)stream>>>(a);
    return check_launch("pdse_gcrn_conv1_fwd");
}

static int launch_stream(StreamArgs& a, cudaStream_t st) {
    const int NC = a.nc[0] + a.nc[1];
    if (NC % (2 * a.kb)) return set_error("stream_kernel: channel chunks not divisible by the k-block");
    if (a.nt * a.P > 128) return set_error("stream_kernel: tile exceeds 128 rows");
    if (a.ntile % 16 || a.ntile > 256) return set_error("stream_kernel: bad n-tile");
    a.R = a.nt * a.P + 1;
    const size_t a_bytes = ((size_t)NC * a.npar * a.R * 16 + 127) & ~(size_t)127;
    a.ep_floats = a.mode == MODE_LIN ? a.N_total : a.n_ntiles * 2 * a.ntile;
    const size_t extra = (size_t)a.ep_floats * 4 + 1024, cap = 227 * 1024;
    const int acc = a.ntile < 32 ? 32 : a.ntile;
    const int want_sm = max(1, min(4, 512 / (2 * acc)));             // CTAs per SM allowed by TMEM
    // Block size (k-steps per streamed block), ring depth, A buffers and CTAs per SM.  Measured (round 2): every streamed
    // block costs its CTA's MMA issuer lane a fixed ~400 cycles of handshake (try_wait, fence, tcgen05.commit) during which
    // the tensor pipe runs dry unless another CTA of the SM has MMAs queued -- the ring DEPTH made no difference (4 .. 20
    // stages), the number of blocks did (lin_ta: 4 -> 16 planes per block, 0.093 -> 0.054 ms).  Within an (n-tile, tap) the
    // packed planes are contiguous, so the block size is a launch-time choice.  Model: pipe share of a block =
    // mma / (mma + 400 / ctas), times a penalty when fewer than 48 KB of weights are in flight per SM.
    const int packed_kb = a.kb;
    const int mma_cycles = a.ntile > 128 ? 128 : a.ntile > 64 ? 75 : 62;
    const char* force_kb = getenv("PDSE_STREAM_KB");        // A/B switch
    double best = -1.0;
    int best_cta = 1;
    a.stages = 0;
    for (int kb = 4; kb >= 1; kb /= 2) {        // (8 k-steps per block measured slower than 4: the ring gets too shallow)
        if (NC % (2 * kb)) continue;
        if (force_kb && atoi(force_kb) != kb && !(atoi(force_kb) > kb && best < 0)) continue;
        const size_t blk = (size_t)a.ntile * 2 * kb * 16;
        for (int ab = 2; ab >= 1; --ab)
            for (int cta = want_sm; cta >= 1; --cta) {
                const size_t budget = cap / cta;
                if ((size_t)ab * a_bytes + 3 * blk + extra > budget) continue;
                const int st = (int)min((size_t)MAX_ST, (budget - (size_t)ab * a_bytes - extra) / blk);
                const double mma = (double)kb * mma_cycles, inflight = (double)cta * (st - 1) * blk;
                const double score = mma / (mma + 400.0 / cta) * min(1.0, inflight / 49152.0) * (1.0 + 0.05 * (ab - 1));
                if (score > best + 1e-9) {
                    best = score;
                    best_cta = cta;
                    a.stages = st;
                    a.abufs = ab;
                    a.kb = kb;
                }
            }
    }
    (void)packed_kb;
    if (best < 0) return set_error("stream_kernel: tile does not fit in shared memory");
    if (NC % (2 * a.kb)) return set_error("stream_kernel: channel chunks not divisible by the k-block");
    const size_t blk = (size_t)a.ntile * 2 * a.kb * 16;
    const size_t smem = a.abufs * a_bytes + a.stages * blk + (size_t)a.ep_floats * 4;
    const int per_sm = max(1, min(min(want_sm, best_cta), (int)(cap / (smem + 1024))));
    const int nh = per_sm == 1 ? 2 : 1;        // one CTA per SM: twice the epilogue threads
    static SmemCache hw[2];
    if (int e = nh == 2 ? ensure_smem(stream_kernel<2>, smem, &hw[0]) : ensure_smem(stream_kernel<1>, smem, &hw[1])) return e;
    const int slots = sm_count() * per_sm, base_tiles = a.B * ceil_div(a.T, a.nt);
    a.n_split = 1;
    if (a.n_out_par == 1 && base_tiles > slots / 2) {   // even out the last wave: cost ~ waves / n_split (+ an A reload per unit)
        double best_cost = 1e30;
        for (int ns = 1; ns <= a.n_ntiles; ns *= 2) {
            if (a.n_ntiles % ns) break;
            const int waves = ceil_div(base_tiles * ns, slots);
            const double cost = (double)waves / ns + 0.04 * waves;
            if (cost < best_cost - 1e-9) {
                best_cost = cost;
                a.n_split = ns;
            }
        }
    }
    const int tiles = base_tiles * a.n_split;
    const int grid = min(tiles, slots);
    const int nthr = 64 + 128 * nh;
    if (nh == 2) stream_kernel<2><<<grid, nthr, smem, st>>>(a);
    else stream_kernel<1><<<grid, nthr, smem, st>>>(a);
    return check_launch("stream_kernel");
}

// gcrn.py:138-141 conv{i} + bn{i} + ELU, i = 2..5.  xin: SO(Fin); outputs (any may be NULL): out_so = SO(Fo) for
// the next conv, out_ug = UG(Fo) skip for the decoders (elu2: apply ELU once more, gcrn.py:150-153),
// xl0/xl1 = LSTM layer-1 operands (conv5 only).
extern "C" int pdse_gcrn_enc_fwd(const void* xin, void* out_so, void* out_ug, void* xl0, void* xl1, const void* wb,
                                 const float* ep, int B, int T, int Cin, int Cout, int Fin, int elu2, void* stream) {
    if (B <= 0 || T <= 0 || Cin % 16 || Cout % 16) return set_error("pdse_gcrn_enc_fwd: bad shape");
    StreamArgs a{};
    a.src[0] = (const __nv_bfloat16*)xin;
    a.nc[0] = Cin / 8;
    a.npar = 2;
    const int Q = (Fin + 1) / 2;
    a.plane_rows = (long)T * Q;
    a.B = B;
    a.T = T;
    a.P = Q;
    a.nt = max(1, 128 / Q);
    a.n_out_par = 1;
    a.ntap[0] = 3;
    for (int df = 0; df < 3; ++df) {
        a.tap_par[0][df] = df & 1;
        a.tap_shift[0][df] = df >> 1;
    }
    a.w[0] = (const __nv_bfloat16*)wb;
    a.ntile = min(256, 2 * Cout);
    a.n_ntiles = 2 * Cout / a.ntile;
    a.kb = (Cin / 8) % 4 == 0 ? 2 : 1;
    a.ep = ep;
    a.mode = MODE_ENC;
    a.elu2 = elu2;
    a.Fo[0] = (Fin - 3) / 2 + 1;
    a.C_out = Cout;
    a.out_so = (__nv_bfloat16*)out_so;
    a.out_ug = (__nv_bfloat16*)out_ug;
    a.ug_P = a.Fo[0] + 1;
    a.out_xl[0] = (__nv_bfloat16*)xl0;
    a.out_xl[1] = (__nv_bfloat16*)xl1;
    return launch_stream(a, (cudaStream_t)stream);
}

// gcrn.py:150-153 conv{i}_t + bn + ELU on cat(prev, skip), i = 5..2.  Inputs UG(Fin); output UG(Fout).
extern "C" int pdse_gcrn_dec_fwd(const void* prev, const void* skip, void* out_ug, const void* w_even, const void* w_odd,
                                 const float* ep, int B, int T, int C1, int C2, int Cout, int Fin, int Fout,
                                 void* stream) {
    if (B <= 0 || T <= 0 || C1 % 16 || C2 % 16 || Cout % 8) return set_error("pdse_gcrn_dec_fwd: bad shape");
    StreamArgs a{};
    a.src[0] = (const __nv_bfloat16*)prev;
    a.src[1] = (const __nv_bfloat16*)skip;
    a.nc[0] = C1 / 8;
    a.nc[1] = C2 / 8;
    a.npar = 1;
    const int P = Fin + 1;
    a.plane_rows = (long)T * P + 1;
    a.B = B;
    a.T = T;
    a.P = P;
    a.nt = max(1, 128 / P);
    a.n_out_par = 2;
    a.ntap[0] = 2;                    // even outputs: df=0 reads h[j] (row m+1), df=2 reads h[j-1] (row m)
    a.tap_shift[0][0] = 1;
    a.tap_shift[0][1] = 0;
    a.ntap[1] = 1;                    // odd outputs: df=1 reads h[j]
    a.tap_shift[1][0] = 1;
    a.w[0] = (const __nv_bfloat16*)w_even;
    a.w[1] = (const __nv_bfloat16*)w_odd;
    a.ntile = 2 * Cout;
    a.n_ntiles = 1;
    a.kb = ((C1 + C2) / 8) % 4 == 0 ? 2 : 1;
    a.ep = ep;
    a.mode = MODE_DEC;
    a.Fo[0] = (Fout + 1) / 2;         // even outputs f' = 2j <  Fout
    a.Fo[1] = Fout / 2;               // odd  outputs f' = 2j+1 < Fout
    a.C_out = Cout;
    a.out_ug = (__nv_bfloat16*)out_ug;
    a.ug_P = Fout + 1;
    return launch_stream(a, (cudaStream_t)stream);
}

// LSTM input projection (gcrn.py:12-15, the W_ih half of nn.LSTM): x [64][rows][8] (row = t*B + b) ->
// pre [T][2048][Bp] fp32 (+ b_ih + b_hh), gate rows in the recurrence kernel's order.
extern "C" int pdse_lstm_inproj(const void* x, const void* w_ih, const float* bias, float* pre, int B, int Bp, int T,
                                void* stream) {
    if (B <= 0 || T <= 0 || Bp < B) return set_error("pdse_lstm_inproj: bad shape");
    static const bool old_path = getenv("PDSE_LIN_OLD") != nullptr;     // A/B switch: A tile in shared memory (stream_kernel)
    if (!old_path) {
        LinArgs l;
        l.x = (const __nv_bfloat16*)x;
        l.w = (const __nv_bfloat16*)w_ih;
        l.bias = bias;
        l.out = pre;
        l.rows = T * B;
        l.Bl = B;
        l.Bp = Bp;
        if (Bp != 32 && Bp != 64) return set_error("pdse_lstm_inproj: Bp must be 32 or 64");
        static SmemCache hw32, hw64;
        if (int e = Bp == 32 ? ensure_smem(lin_ta_kernel<32>, (size_t)LIN_SMEM, &hw32) : ensure_smem(lin_ta_kernel<64>, (size_t)LIN_SMEM, &hw64))
            return e;
        // schedule (see LinUnit): whole tiles first, the left-over tiles of the last partial round cut into n-tile groups
        const int tiles = ceil_div(l.rows, 128), G = min(tiles, sm_count());
        l.rounds = tiles / G;
        const int left = tiles - l.rounds * G;
        l.left_split = 1;
        while (left > 0 && l.left_split < 16 && left * l.left_split * 2 <= G) l.left_split *= 2;
        l.debug = getenv("PDSE_LIN_DEBUG") ? atoi(getenv("PDSE_LIN_DEBUG")) : 0;
        if (Bp == 32) lin_ta_kernel<32><<<G, LIN_THR, LIN_SMEM, (cudaStream_t)stream>>>(l);
        else lin_ta_kernel<64><<<G, LIN_THR, LIN_SMEM, (cudaStream_t)stream>>>(l);
        return check_launch("pdse_lstm_inproj");
    }
    StreamArgs a{};
    a.src[0] = (const __nv_bfloat16*)x;
    a.nc[0] = 64;
    a.npar = 1;
    a.plane_rows = (long)T * B;
    a.B = 1;
    a.T = T * B;
    a.P = 1;
    a.nt = 128;
    a.n_out_par = 1;
    a.ntap[0] = 1;
    a.w[0] = (const __nv_bfloat16*)w_ih;
    a.ntile = 128;
    a.n_ntiles = 16;
    a.kb = 2;
    a.ep = bias;
    a.mode = MODE_LIN;
    a.out_f32 = pre;
    a.Bl = B;
    a.Bp = Bp;
    a.N_total = 2048;
    return launch_stream(a, (cudaStream_t)stream);
}

static long long* g_lstm_prof = nullptr;
// debug hook: device buffer of 6 int64 cycle counters written by CTA (0,0) of the next recurrence launches
extern "C" int pdse_debug_lstm_prof(void* dev_buf) {
    g_lstm_prof = (long long*)dev_buf;
    return 0;
}

// debug hook: how many 16-CTA clusters of the DSMEM recurrence kernel (bp = 16 or 32 sequences per cluster) the current
// device can hold at once (cudaOccupancyMaxActiveClusters); negative on error
extern "C" int pdse_debug_lstm_clusters(int bp) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(16, 16);
    cfg.blockDim = dim3(LSTM_THR);
    cfg.dynamicSmemBytes = bp == 16 ? LdCfg<16>::SMEM : LdCfg<32>::SMEM;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 16;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const void* fn = bp == 16 ? (const void*)lstm_dsmem_kernel<16> : (const void*)lstm_dsmem_kernel<32>;
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
        cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes) != cudaSuccess)
        return -1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, fn, &cfg) != cudaSuccess) return -2;
    return n;
}

// LSTM recurrence of one layer, both groups (gcrn.py:28 / :33).  sync: 2 zeroed counters.
extern "C" int pdse_lstm_rec(const void* whh0, const void* whh1, const float* pre0, const float* pre1, float* h0,
                             float* h1, void* hbuf, unsigned int* sync, int B, int Bp, int T, void* stream) {
    if (B <= 0 || T <= 0 || Bp < B || (Bp != 32 && Bp != 64)) return set_error("pdse_lstm_rec: Bp must be 32 or 64 and >= B");
    LstmArgs a;
    a.whh[0] = (const __nv_bfloat16*)whh0;
    a.whh[1] = (const __nv_bfloat16*)whh1;
    a.pre[0] = pre0;
    a.pre[1] = pre1;
    a.hout[0] = h0;
    a.hout[1] = h1;
    a.hbuf = (__nv_bfloat16*)hbuf;
    a.sync = sync;
    a.B = B;
    a.Bp = Bp;
    a.T = T;
    a.prof = g_lstm_prof;
    const size_t stage = max((size_t)64 * Bp * 16, (size_t)128 * (Bp + 1) * 4);
    a.stage_bytes = (int)((stage + 127) & ~(size_t)127);
    const size_t smem = 131072 + a.stage_bytes + (size_t)32 * Bp * 4 + (size_t)4 * Bp * 16;
    if (smem > 227 * 1024) return set_error("pdse_lstm_rec: batch chunk too large for shared memory");
    // Preferred: one 16-CTA cluster per (group, 16 or 32 sequences) exchanging h through DSMEM bulk copies; else one
    // cluster per group with the hardware cluster barrier and h through L2; else a cooperative launch.
    // Per-device state (kernel attributes and the cluster occupancy belong to a device): mode + 1, 0 = not probed yet;
    // bit 8 = eight co-resident 16-CTA clusters are available (16 sequences per cluster on 128 CTAs).
    static std::atomic<int> mode_cache[MAX_DEVICES];   // PDSE_LSTM_MODE caps the variant: 3 two-stream DSMEM, 2 DSMEM, 1 cluster barrier, 0 cooperative
    static SmemCache hw[4];
    const int dev = current_device();
    const void* fn_cl = Bp == 32 ? (const void*)lstm_rec_kernel<32, true> : (const void*)lstm_rec_kernel<64, true>;
    const void* fn_co = Bp == 32 ? (const void*)lstm_rec_kernel<32, false> : (const void*)lstm_rec_kernel<64, false>;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(16, 2);
    cfg.blockDim = dim3(LSTM_THR);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 16;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int cached = mode_cache[dev].load(std::memory_order_acquire);
    if (cached == 0) {
        int mode = 0, wide = 0;
        const char* force = getenv("PDSE_LSTM_MODE");
        const char* force_bp = getenv("PDSE_LSTM_BP");
        const bool ok = cudaFuncSetAttribute((const void*)lstm_rec_kernel<32, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
                        cudaFuncSetAttribute((const void*)lstm_rec_kernel<64, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
                        cudaFuncSetAttribute((const void*)lstm_dsmem_kernel<32>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
                        cudaFuncSetAttribute((const void*)lstm_dsmem_kernel<16>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
                        cudaFuncSetAttribute((const void*)lstm_dsmem_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, LdCfg<32>::SMEM) == cudaSuccess &&
                        cudaFuncSetAttribute((const void*)lstm_dsmem_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, LdCfg<16>::SMEM) == cudaSuccess &&
                        cudaFuncSetAttribute((const void*)lstm_dsmem2_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
                        cudaFuncSetAttribute((const void*)lstm_dsmem2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L2_SMEM) == cudaSuccess;
        if (ok) {
            // the fallback kernels' shared-memory attribute must be in place before their occupancy is queried
            (void)cudaFuncSetAttribute(fn_cl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            int nclusters = 0;
            if (cudaOccupancyMaxActiveClusters(&nclusters, fn_cl, &cfg) == cudaSuccess && nclusters >= 2) mode = 1;
            cudaLaunchConfig_t c2 = cfg;
            c2.gridDim = dim3(16, 4);
            c2.dynamicSmemBytes = LdCfg<32>::SMEM;
            if (cudaOccupancyMaxActiveClusters(&nclusters, (const void*)lstm_dsmem_kernel<32>, &c2) == cudaSuccess && nclusters >= 4) mode = 2;
            c2.gridDim = dim3(16, 8);
            c2.dynamicSmemBytes = LdCfg<16>::SMEM;
            if (mode == 2 && cudaOccupancyMaxActiveClusters(&nclusters, (const void*)lstm_dsmem_kernel<16>, &c2) == cudaSuccess && nclusters >= 8) wide = 1;
            // two interleaved 16-sequence streams per cluster (the same 4 clusters as 32 sequences per cluster)
            c2.gridDim = dim3(16, 4);
            c2.blockDim = dim3(L2_THR);
            c2.dynamicSmemBytes = L2_SMEM;
            if (mode == 2 && !wide && cudaOccupancyMaxActiveClusters(&nclusters, (const void*)lstm_dsmem2_kernel, &c2) == cudaSuccess && nclusters >= 4) mode = 3;
        }
        if (force) mode = min(mode, atoi(force));
        if (force_bp) wide = atoi(force_bp) == 16 ? 1 : 0;
        (void)cudaGetLastError();
        cached = (mode + 1) | (wide << 8);
        mode_cache[dev].store(cached, std::memory_order_release);
    }
    const int mode = (cached & 0xff) - 1, wide = cached >> 8;
    if (mode < 2) {
        for (int v = 0; v < 2; ++v)
            if (int e = ensure_smem(v ? fn_cl : fn_co, smem, &hw[(Bp / 32 - 1) * 2 + v])) return e;
    }
    void* params[] = {&a};
    if (mode == 3) {
        cfg.gridDim = dim3(16, 2 * ceil_div(B, 2 * L2_BP));
        cfg.blockDim = dim3(L2_THR);
        cfg.dynamicSmemBytes = L2_SMEM;
        PDSE_CUDA(cudaLaunchKernelExC(&cfg, (const void*)lstm_dsmem2_kernel, params));
    } else if (mode == 2) {
        // clusters are independent of each other (no co-residency requirement), so chunks of 16 sequences are used whenever
        // all of them fit on the device at once; otherwise 32 per cluster
        const int bp = wide ? 16 : 32;
        cfg.gridDim = dim3(16, 2 * ceil_div(B, bp));
        cfg.dynamicSmemBytes = bp == 16 ? LdCfg<16>::SMEM : LdCfg<32>::SMEM;
        PDSE_CUDA(cudaLaunchKernelExC(&cfg, bp == 16 ? (const void*)lstm_dsmem_kernel<16> : (const void*)lstm_dsmem_kernel<32>, params));
    } else if (mode == 1) {
        PDSE_CUDA(cudaLaunchKernelExC(&cfg, fn_cl, params));
    } else {
        PDSE_CUDA(cudaMemsetAsync(sync, 0, 2 * sizeof(unsigned int), (cudaStream_t)stream));
        PDSE_CUDA(cudaLaunchCooperativeKernel(fn_co, dim3(16, 2), dim3(LSTM_THR), params, smem, (cudaStream_t)stream));
    }
    return check_launch("pdse_lstm_rec");
}

// gcrn.py:29-31 (mode 1) / :33-38 (mode 2) LayerNorm(1024) with the group shuffles fused
extern "C" int pdse_gcrn_ln(const float* h0, const float* h1, const float* w, const float* b, void* xl0, void* xl1,
                            void* ug, int B, int T, int mode, void* stream) {
    if (B <= 0 || T <= 0 || (mode != 1 && mode != 2)) return set_error("pdse_gcrn_ln: bad arguments");
    LnArgs a;
    a.h[0] = h0;
    a.h[1] = h1;
    a.w = w;
    a.bia = b;
    a.xl[0] = (__nv_bfloat16*)xl0;
    a.xl[1] = (__nv_bfloat16*)xl1;
    a.ug = (__nv_bfloat16*)ug;
    a.rows = B * T;
    a.B = B;
    a.T = T;
    a.mode = mode;
    ln_kernel<<<ceil_div(a.rows, 8), 256, 0, (cudaStream_t)stream>>>(a);
    return check_launch("pdse_gcrn_ln");
}

// gcrn.py:154/160 conv1_t + bn1_t + ELU, :162-163 fc, trainer :942 (/11).  Writes X_init [B][2][T][161].
extern "C" int pdse_gcrn_out_fwd(const void* d2_1, const void* d2_2, const void* e1_ug, const float* wf1,
                                 const float* wf2, float* xinit, int B, int T, void* stream) {
    if (B <= 0 || T <= 0) return set_error("pdse_gcrn_out_fwd: empty input");
    GOutArgs a;
    a.d2[0] = (const __nv_bfloat16*)d2_1;
    a.d2[1] = (const __nv_bfloat16*)d2_2;
    a.e1 = (const __nv_bfloat16*)e1_ug;
    a.wf[0] = wf1;
    a.wf[1] = wf2;
    a.xinit = xinit;
    a.B = B;
    a.T = T;
    dim3 grid(ceil_div(T, OUT_FR), B, 2);
    const size_t smem = OUT_SMEM;
    static SmemCache hw;
    if (int e = ensure_smem(gout_kernel, smem, &hw)) return e;
    gout_kernel<<<grid, 192, smem, (cudaStream_t)stream>>>(a);
    return check_launch("pdse_gcrn_out_fwd");
}
