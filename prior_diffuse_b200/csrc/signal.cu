// Framed windowed DFT (STFT) with fused sqrt-compression, and its inverse with fused
// decompression + overlap-add + envelope normalisation.
//
// Replaces: torch.stft call at trainer/complex_ddpm_trainer.py:926-930 (batched twin
// utils/dataset.py:61-67) + compression :931-937; decompression :1004-1008 +
// torch.istft :1009-1015; RMS normalisation :922-923.
//
// n_fft = win = 320, hop = 160, periodic Hann, center=True (reflect), onesided (161 bins).
// Arithmetic is fp32 FFMA (the 1e-5 bar rules out bf16/tf32 tensor cores).
// The 320-point real DFT is folded twice (n <-> 320-n, then n <-> 160-n) so that each
// output bin costs 79 cos-MACs + 79 sin-MACs instead of 320+320; twiddles come from a
// table built in float64 from the exact integer (k*n mod 320) (see pdse_signal_tables).
#include "common.cuh"
#include <math_constants.h>

namespace pdse {

constexpr int NFFT = 320, HOP = 160, NF = 161;
constexpr int NSLOT = 192;                       // thread slots: [0,96) even bins, [96,192) odd bins
constexpr int TAB_HANN = 0;                      // [320]
constexpr int TAB_COS = 320;                     // [79][192]
constexpr int TAB_SIN = TAB_COS + 79 * NSLOT;    // [79][192]
constexpr int TAB_FLOATS = TAB_SIN + 79 * NSLOT;

__host__ __device__ inline int slot_to_bin(int slot) { return slot < 96 ? 2 * slot : 2 * (slot - 96) + 1; }

// ------------------------------------------------------------------ RMS
// `lengths` (optional, int32[B]): true sample count of every utterance in a zero-padded ragged batch of width L
// (utils/dataset.py:45-60 pads with pad_sequence); NULL = all utterances are L samples long.
__global__ void rms_kernel(const float* __restrict__ wav, const int* __restrict__ lengths, int Lpitch,
                           float* __restrict__ rms) {
    const float* w = wav + (size_t)blockIdx.x * Lpitch;
    const int L = lengths ? lengths[blockIdx.x] : Lpitch;
    float acc = 0.f;
    for (int i = threadIdx.x; i < L; i += blockDim.x) acc = fmaf(w[i], w[i], acc);
    __shared__ float red[32];
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        acc = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (threadIdx.x == 0) rms[blockIdx.x] = sqrtf(acc / (float)L);
    }
}

// ------------------------------------------------------------------ STFT + compress
constexpr int FT = 16;  // frames per CTA

__global__ void __launch_bounds__(NSLOT)
stft_compress_kernel(const float* __restrict__ wav, const float* __restrict__ rms, const float* __restrict__ tab,
                     const int* __restrict__ lengths, float* __restrict__ out, int Lpitch, int T, int compress) {
    __shared__ __align__(16) float sm[8192];
    float* sx = sm;                 // (FT+1)*160 samples
    float* fold = sm + 2720;        // 4 x [80][FT]: ae, ao, be, bo   (row 0 holds the DC/Nyquist/n=80 terms)
    const int b = blockIdx.y, t0 = blockIdx.x * FT, tid = threadIdx.x;
    const float* w = wav + (size_t)b * Lpitch;
    const float inv = rms ? 1.f / rms[b] : 1.f;
    const int L = lengths ? lengths[b] : Lpitch;        // reflect padding happens at the utterance's own end
    const int Tb = 1 + L / HOP;                          // frames past it are written as zeros

    for (int i = tid; i < (FT + 1) * HOP; i += NSLOT) {
        int g = t0 * HOP + i - HOP;                 // index into the unpadded signal
        if (g < 0) g = -g;                           // reflect (center=True)
        if (g >= L) g = 2 * (L - 1) - g;
        sx[i] = (g >= 0 && g < L) ? w[g] * inv : 0.f;
    }
    __syncthreads();
    const float* hann = tab + TAB_HANN;
    for (int i = tid; i < FT * 80; i += NSLOT) {
        const int f = i / 80, n = i % 80;
        const float* s = sx + f * HOP;
        float ae, ao, be, bo;
        if (n == 0) {
            // row 0 carries the terms outside the n=1..79 sums
            const float x0 = s[0] * hann[0], x160 = s[160] * hann[160];
            const float a80 = s[80] * hann[80] + s[240] * hann[240];
            const float b80 = s[80] * hann[80] - s[240] * hann[240];
            ae = x0 + x160;   // even bins: + (-1)^m a80
            ao = x0 - x160;   // odd bins
            be = a80;
            bo = b80;
        } else {
            const float p = s[n] * hann[n], q = s[320 - n] * hann[320 - n];
            const float r = s[160 - n] * hann[160 - n], u = s[160 + n] * hann[160 + n];
            const float a1 = p + q, a2 = r + u, b1 = p - q, b2 = r - u;
            ae = a1 + a2;
            ao = a1 - a2;
            be = b1 - b2;
            bo = b1 + b2;
        }
        fold[(0 * 80 + n) * FT + f] = ae;
        fold[(1 * 80 + n) * FT + f] = ao;
        fold[(2 * 80 + n) * FT + f] = be;
        fold[(3 * 80 + n) * FT + f] = bo;
    }
    __syncthreads();

    const bool odd = tid >= 96;
    const int m = odd ? tid - 96 : tid;
    const int k = slot_to_bin(tid);
    const bool valid = k < NF;
    const float* A = fold + (odd ? 1 : 0) * 80 * FT;
    const float* Bv = fold + (odd ? 3 : 2) * 80 * FT;
    float re[FT], im[FT];
    const float sgn = (m & 1) ? -1.f : 1.f;
#pragma unroll
    for (int f = 0; f < FT; ++f) {
        if (!odd) {
            re[f] = A[f] + sgn * fold[(2 * 80) * FT + f];   // x0 + x160 + (-1)^m a80
            im[f] = 0.f;
        } else {
            re[f] = A[f];                                    // x0 - x160
            im[f] = -sgn * fold[(3 * 80) * FT + f];          // -(-1)^m b80
        }
    }
    const float* ct = tab + TAB_COS + tid;
    const float* st = tab + TAB_SIN + tid;
#pragma unroll 2
    for (int n = 1; n < 80; ++n) {
        const float c = __ldg(ct + (n - 1) * NSLOT), s = __ldg(st + (n - 1) * NSLOT);
        const float4* a4 = reinterpret_cast<const float4*>(A + n * FT);
        const float4* b4 = reinterpret_cast<const float4*>(Bv + n * FT);
#pragma unroll
        for (int j = 0; j < FT / 4; ++j) {
            const float4 a = a4[j], bb = b4[j];
            re[4 * j + 0] = fmaf(a.x, c, re[4 * j + 0]);
            re[4 * j + 1] = fmaf(a.y, c, re[4 * j + 1]);
            re[4 * j + 2] = fmaf(a.z, c, re[4 * j + 2]);
            re[4 * j + 3] = fmaf(a.w, c, re[4 * j + 3]);
            im[4 * j + 0] = fmaf(-bb.x, s, im[4 * j + 0]);
            im[4 * j + 1] = fmaf(-bb.y, s, im[4 * j + 1]);
            im[4 * j + 2] = fmaf(-bb.z, s, im[4 * j + 2]);
            im[4 * j + 3] = fmaf(-bb.w, s, im[4 * j + 3]);
        }
    }
    __syncthreads();   // everyone is done with sx/fold: reuse as the output staging tile
    float* so = sm;    // [FT][2][161]
    if (valid) {
#pragma unroll
        for (int f = 0; f < FT; ++f) {
            float r = re[f], i = im[f];
            if (compress) {   // z * |z|^(-1/2); 0 where |z| = 0 (atan2(0,0) = 0, mag = 0)
                const float mag = sqrtf(r * r + i * i);
                const float g = mag > 0.f ? 1.f / sqrtf(mag) : 0.f;
                r *= g;
                i *= g;
            }
            so[(f * 2 + 0) * NF + k] = r;
            so[(f * 2 + 1) * NF + k] = i;
        }
    }
    __syncthreads();
    const int nf = min(FT, T - t0);
    for (int i = tid; i < nf * 2 * NF; i += NSLOT) {
        const int f = i / (2 * NF), rem = i % (2 * NF), ch = rem / NF, kk = rem % NF;
        out[(((size_t)b * 2 + ch) * T + t0 + f) * NF + kk] = t0 + f < Tb ? so[(f * 2 + ch) * NF + kk] : 0.f;
    }
}

// ------------------------------------------------------------------ decompress + ISTFT
constexpr int FI = 16;        // hop blocks per CTA  (FI+1 frames are synthesised)
constexpr int FIP = 20;       // frame pitch of the folded arrays (multiple of 4 >= FI+1)
constexpr int XSZ = (2 * (FI + 1) * 161 + 3) / 4 * 4;   // spectra staging, padded so `fold` stays 16-byte aligned

__global__ void __launch_bounds__(NSLOT)
decompress_istft_kernel(const float* __restrict__ spec, const float* __restrict__ rms, const float* __restrict__ tab,
                        const int* __restrict__ lengths, float* __restrict__ wav, int Lpitch, int Tpitch, int decompress) {
    extern __shared__ __align__(16) float smi[];
    float* X = smi;                               // [2][FI+1][161]  decompressed spectra; later frames [FI+1][320]
    float* fold = smi + XSZ;                      // 4 x [80][FIP]: Re, Ro, Ie, Io (row 0: k=0/160/80 terms)
    const int b = blockIdx.y, c0 = blockIdx.x, tid = threadIdx.x;
    const int tf = c0 * FI;                       // first frame synthesised by this CTA
    const int L = lengths ? lengths[b] : Lpitch;  // ragged batch: frames / samples past the utterance's end do not exist
    const int T = lengths ? min(Tpitch, 1 + L / HOP) : Tpitch;

    for (int i = tid; i < (FI + 1) * NF; i += NSLOT) {
        const int lf = i / NF, k = i % NF, t = tf + lf;
        float r = 0.f, im = 0.f;
        if (t < T) {
            r = spec[(((size_t)b * 2 + 0) * Tpitch + t) * NF + k];
            im = spec[(((size_t)b * 2 + 1) * Tpitch + t) * NF + k];
            if (decompress) {   // z * |z|  (mag^2, phase kept)
                const float mag = sqrtf(r * r + im * im);
                r *= mag;
                im *= mag;
            }
        }
        X[lf * NF + k] = r;
        X[(FI + 1) * NF + lf * NF + k] = im;
    }
    __syncthreads();
    const float* Xr = X;
    const float* Xi = X + (FI + 1) * NF;
    for (int i = tid; i < (FI + 1) * 80; i += NSLOT) {
        const int lf = i / 80, k = i % 80;
        float re, ro, ie, io;
        if (k == 0) {
            const float x0 = Xr[lf * NF + 0], x160 = Xr[lf * NF + 160];
            re = x0 + x160;                 // even n: X0 + X160 (+ 2 (-1)^m Xr80)
            ro = x0 - x160;                 // odd n
            ie = 2.f * Xr[lf * NF + 80];
            io = 2.f * Xi[lf * NF + 80];
        } else {
            const float r1 = Xr[lf * NF + k], r2 = Xr[lf * NF + 160 - k];
            const float i1 = Xi[lf * NF + k], i2 = Xi[lf * NF + 160 - k];
            re = 2.f * (r1 + r2);
            ro = 2.f * (r1 - r2);
            ie = 2.f * (i1 - i2);
            io = 2.f * (i1 + i2);
        }
        fold[(0 * 80 + k) * FIP + lf] = re;
        fold[(1 * 80 + k) * FIP + lf] = ro;
        fold[(2 * 80 + k) * FIP + lf] = ie;
        fold[(3 * 80 + k) * FIP + lf] = io;
    }
    __syncthreads();

    const bool odd = tid >= 96;
    const int m = odd ? tid - 96 : tid;
    const int n = slot_to_bin(tid);
    const bool valid = n < NF;
    const float* R = fold + (odd ? 1 : 0) * 80 * FIP;
    const float* I = fold + (odd ? 3 : 2) * 80 * FIP;
    const float sgn = (m & 1) ? -1.f : 1.f;
    float E[FIP], O[FIP];
#pragma unroll
    for (int f = 0; f < FIP; ++f) {
        if (!odd) {
            E[f] = R[f] + sgn * fold[(2 * 80) * FIP + f];    // X0 + X160 + 2 (-1)^m Xr80
            O[f] = 0.f;
        } else {
            E[f] = R[f];
            O[f] = sgn * fold[(3 * 80) * FIP + f];           // 2 (-1)^m Xi80
        }
    }
    const float* ct = tab + TAB_COS + tid;
    const float* st = tab + TAB_SIN + tid;
#pragma unroll 2
    for (int k = 1; k < 80; ++k) {
        const float c = __ldg(ct + (k - 1) * NSLOT), s = __ldg(st + (k - 1) * NSLOT);
        const float4* r4 = reinterpret_cast<const float4*>(R + k * FIP);
        const float4* i4 = reinterpret_cast<const float4*>(I + k * FIP);
#pragma unroll
        for (int j = 0; j < FIP / 4; ++j) {
            const float4 a = r4[j], bb = i4[j];
            E[4 * j + 0] = fmaf(a.x, c, E[4 * j + 0]);
            E[4 * j + 1] = fmaf(a.y, c, E[4 * j + 1]);
            E[4 * j + 2] = fmaf(a.z, c, E[4 * j + 2]);
            E[4 * j + 3] = fmaf(a.w, c, E[4 * j + 3]);
            O[4 * j + 0] = fmaf(bb.x, s, O[4 * j + 0]);
            O[4 * j + 1] = fmaf(bb.y, s, O[4 * j + 1]);
            O[4 * j + 2] = fmaf(bb.z, s, O[4 * j + 2]);
            O[4 * j + 3] = fmaf(bb.w, s, O[4 * j + 3]);
        }
    }
    __syncthreads();   // X is dead: reuse as the windowed-frame buffer [FI+1][320]
    float* fr = smi;
    if (valid) {
        const float wn = tab[TAB_HANN + n] * (1.f / 320.f);
#pragma unroll
        for (int f = 0; f < FI + 1; ++f) {
            fr[f * NFFT + n] = (E[f] - O[f]) * wn;
            if (n >= 1 && n <= 159) fr[f * NFFT + (NFFT - n)] = (E[f] + O[f]) * wn;
        }
    }
    __syncthreads();
    const float scale = rms ? rms[b] : 1.f;
    const float* hann = tab + TAB_HANN;
    for (int i = tid; i < FI * HOP; i += NSLOT) {
        const int jj = i / HOP, r = i % HOP;
        const int hb = 1 + tf + jj;                 // hop block in padded coordinates
        const long o = (long)HOP * (hb - 1) + r;    // output sample (centre pad stripped)
        if (o >= Lpitch) continue;
        if (o >= L) {                               // padding of a ragged batch
            wav[(size_t)b * Lpitch + o] = 0.f;
            continue;
        }
        const int ta = hb - 1, tb = hb;             // frames overlapping this block
        float acc = 0.f, env = 0.f;
        if (ta < T) {
            acc += fr[jj * NFFT + HOP + r];
            env += hann[HOP + r] * hann[HOP + r];
        }
        if (tb < T) {
            acc += fr[(jj + 1) * NFFT + r];
            env += hann[r] * hann[r];
        }
        wav[(size_t)b * Lpitch + o] = env > 1e-11f ? acc / env * scale : 0.f;
    }
}

}  // namespace pdse

// ---------------------------------------------------------------------------- C ABI
extern "C" int pdse_signal_table_floats(void) { return pdse::TAB_FLOATS; }

// Host-side table builder (float64 twiddles from the exact integer k*n mod 320).
extern "C" int pdse_signal_tables(float* host_out) {
    using namespace pdse;
    const double two_pi = 6.283185307179586476925286766559;
    for (int n = 0; n < NFFT; ++n) host_out[TAB_HANN + n] = (float)(0.5 - 0.5 * cos(two_pi * n / NFFT));
    for (int n = 1; n < 80; ++n)
        for (int s = 0; s < NSLOT; ++s) {
            const int k = slot_to_bin(s);
            double c = 0.0, sn = 0.0;
            if (k < NF) {
                const int r = (k * n) % NFFT;
                c = cos(two_pi * r / NFFT);
                sn = sin(two_pi * r / NFFT);
            }
            host_out[TAB_COS + (n - 1) * NSLOT + s] = (float)c;
            host_out[TAB_SIN + (n - 1) * NSLOT + s] = (float)sn;
        }
    return 0;
}

extern "C" int pdse_rms_ragged_f32(const float* wav, const int* lengths, int B, int L, float* rms, void* stream) {
    using namespace pdse;
    if (B <= 0 || L <= 0) return set_error("pdse_rms_f32: empty input");
    rms_kernel<<<B, 512, 0, (cudaStream_t)stream>>>(wav, lengths, L, rms);
    return check_launch("pdse_rms_f32");
}
extern "C" int pdse_rms_f32(const float* wav, int B, int L, float* rms, void* stream) {
    return pdse_rms_ragged_f32(wav, nullptr, B, L, rms, stream);
}

extern "C" int pdse_stft_compress_ragged_f32(const float* wav, const float* rms, const float* tables, const int* lengths,
                                             float* out, int B, int L, int compress, void* stream) {
    using namespace pdse;
    if (B <= 0) return set_error("pdse_stft_compress_f32: B must be > 0");
    if (L <= HOP) return set_error("pdse_stft_compress_f32: need L > 160 samples (reflect padding)");
    const int T = 1 + L / HOP;
    dim3 grid(ceil_div(T, FT), B);
    stft_compress_kernel<<<grid, NSLOT, 0, (cudaStream_t)stream>>>(wav, rms, tables, lengths, out, L, T, compress);
    return check_launch("pdse_stft_compress_f32");
}
extern "C" int pdse_stft_compress_f32(const float* wav, const float* rms, const float* tables, float* out, int B,
                                      int L, int compress, void* stream) {
    return pdse_stft_compress_ragged_f32(wav, rms, tables, nullptr, out, B, L, compress, stream);
}

extern "C" int pdse_decompress_istft_ragged_f32(const float* spec, const float* rms, const float* tables,
                                                const int* lengths, float* wav, int B, int T, int L, int decompress,
                                                void* stream) {
    using namespace pdse;
    if (B <= 0 || T <= 0 || L <= 0) return set_error("pdse_decompress_istft_f32: empty input");
    if (L > HOP * T) return set_error("pdse_decompress_istft_f32: length exceeds the frames' support");
    const int nblocks = ceil_div(L, HOP);
    dim3 grid(ceil_div(nblocks, FI), B);
    const size_t smem = (size_t)(XSZ + 4 * 80 * FIP) * sizeof(float);
    static SmemCache hw;
    if (int e = ensure_smem(decompress_istft_kernel, smem, &hw)) return e;
    decompress_istft_kernel<<<grid, NSLOT, smem, (cudaStream_t)stream>>>(spec, rms, tables, lengths, wav, L, T, decompress);
    return check_launch("pdse_decompress_istft_f32");
}
extern "C" int pdse_decompress_istft_f32(const float* spec, const float* rms, const float* tables, float* wav, int B,
                                         int T, int L, int decompress, void* stream) {
    return pdse_decompress_istft_ragged_f32(spec, rms, tables, nullptr, wav, B, T, L, decompress, stream);
}
