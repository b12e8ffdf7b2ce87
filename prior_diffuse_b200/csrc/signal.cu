// Framed windowed real FFT (STFT) with fused sqrt-compression, and its inverse with fused
// decompression + overlap-add + envelope normalisation (+ optional float -> PCM16).
//
// Replaces: torch.stft call at trainer/complex_ddpm_trainer.py:926-930 (batched twin
// utils/dataset.py:61-67) + compression :931-937; decompression :1004-1008 +
// torch.istft :1009-1015; RMS normalisation :922-923; the writer's conversion :1018.
//
// n_fft = win = 320, hop = 160, periodic Hann, center=True (reflect), onesided (161 bins).
// Arithmetic is fp32 (the 1e-5 bar rules out bf16/tf32 tensor cores).  One WARP transforms one frame:
// the 320 real samples are packed as 160 complex points z[m] = x[2m] + i x[2m+1]; 160 = 5 x 32, so lane j takes
// z[j + 32 m2] (m2 = 0..4), does the radix-5 butterfly in registers, multiplies by W160^(j q), and the five 32-point FFTs
// run ACROSS the lanes with butterfly shuffles (decimation in frequency: natural order in, bit-reversed lane order out).
// The real-input untangling X[k] = E + w_k O, X[160-k] = conj(E - w_k O) goes through a 1.3 KB per-warp staging array so
// that the 161 bins leave the warp in order (coalesced stores).  ~8 kFLOP per frame instead of the 51 kFLOP of the
// direct folded DFT this replaces, no table reads in the loop: the kernels are bound by their 1928 B per frame of HBM
// traffic.  The inverse mirrors it (decimation in time, bit-reversed in, natural out).  Twiddles are 320th roots of
// unity from a table built in float64 on the host (pdse_signal_tables).  tests/emu.py re-runs this lane arithmetic in
// NumPy against numpy.fft (tests/test_pack_emulation.py).
#include "common.cuh"
#include <math_constants.h>

namespace pdse {

constexpr int NFFT = 320, HOP = 160, NF = 161;
constexpr int TAB_HANN = 0;                      // [320] periodic Hann window
constexpr int TAB_TW = 320;                      // [320] x (re, im): W320^r = exp(-2 pi i r / 320)
constexpr int TAB_FLOATS = TAB_TW + 2 * 320;

// ------------------------------------------------------------------ RMS
// `lengths` (optional, int32[B]): true sample count of every utterance in a zero-padded ragged batch of width L
// (utils/dataset.py:45-60 pads with pad_sequence); NULL = all utterances are L samples long.
__global__ void rms_kernel(const float* __restrict__ wav, const int* __restrict__ lengths, int Lpitch,
                           float* __restrict__ rms) {
    const float* w = wav + (size_t)blockIdx.x * Lpitch;
    const int L = lengths ? lengths[blockIdx.x] : Lpitch;
    // eight independent loads / partial sums per thread per trip: one CTA per utterance is latency-bound otherwise (a single
    // dependent chain of 94 loads per thread took 46 us for 64 x 3 s, as long as the whole STFT)
    const int n = blockDim.x;
    float p[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int i = threadIdx.x;
    for (; i + 7 * n < L; i += 8 * n) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = __ldg(w + i + k * n);
#pragma unroll
        for (int k = 0; k < 8; ++k) p[k] = fmaf(v[k], v[k], p[k]);
    }
    for (; i < L; i += n) p[0] = fmaf(w[i], w[i], p[0]);
    float acc = ((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]));
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

// ------------------------------------------------------------------ warp-level 160-point complex FFT pieces
struct cpx {
    float x, y;
};
__device__ __forceinline__ cpx cmul(cpx a, cpx b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ cpx cmulc(cpx a, cpx b) { return {a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y}; }   // a * conj(b)
__device__ __forceinline__ cpx cadd(cpx a, cpx b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ cpx csub(cpx a, cpx b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ cpx cshfl(cpx a, int mask) {
    return {__shfl_xor_sync(0xffffffffu, a.x, mask), __shfl_xor_sync(0xffffffffu, a.y, mask)};
}
// 5-point DFT of x[0..4] in place; INV: exp(+i) kernel (unscaled)
template <bool INV>
__device__ __forceinline__ void radix5(cpx (&x)[5]) {
    constexpr float C1 = 0.30901699437494745f, S1 = 0.9510565162951535f, C2 = -0.8090169943749473f, S2 = 0.5877852522924731f;
    const cpx t1 = cadd(x[1], x[4]), t2 = cadd(x[2], x[3]), t3 = csub(x[1], x[4]), t4 = csub(x[2], x[3]);
    const cpx a1 = {x[0].x + C1 * t1.x + C2 * t2.x, x[0].y + C1 * t1.y + C2 * t2.y};
    const cpx a2 = {x[0].x + C2 * t1.x + C1 * t2.x, x[0].y + C2 * t1.y + C1 * t2.y};
    const cpx b1 = {S1 * t3.x + S2 * t4.x, S1 * t3.y + S2 * t4.y};
    const cpx b2 = {S2 * t3.x - S1 * t4.x, S2 * t3.y - S1 * t4.y};
    x[0] = {x[0].x + t1.x + t2.x, x[0].y + t1.y + t2.y};
    // forward: y1 = a1 - i b1, y4 = a1 + i b1, y2 = a2 - i b2, y3 = a2 + i b2; inverse: the signs swap
    const cpx ib1 = {-b1.y, b1.x}, ib2 = {-b2.y, b2.x};
    if (INV) {
        x[1] = cadd(a1, ib1), x[4] = csub(a1, ib1), x[2] = cadd(a2, ib2), x[3] = csub(a2, ib2);
    } else {
        x[1] = csub(a1, ib1), x[4] = cadd(a1, ib1), x[2] = csub(a2, ib2), x[3] = cadd(a2, ib2);
    }
}
// per-lane constants of the transforms (all 320th roots of unity, from the host-built table)
struct LaneTw {
    cpx tq[4];      // W160^(lane q), q = 1..4
    cpx st[4];      // butterfly twiddles of the stages with half-size 16, 8, 4, 2 (1 on lanes whose stage bit is clear)
    float sg[5];    // butterfly sign of the stages with half-size 16, 8, 4, 2, 1: -1 on lanes whose stage bit is set
    int brev;       // bit reversal of the lane index
};
__device__ __forceinline__ LaneTw lane_twiddles(const float* __restrict__ tab, int lane) {
    LaneTw w;
    const float2* tw = reinterpret_cast<const float2*>(tab + TAB_TW);
#pragma unroll
    for (int q = 1; q < 5; ++q) {
        const float2 v = __ldg(tw + (2 * lane * q) % 320);
        w.tq[q - 1] = {v.x, v.y};
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const int h = 16 >> i;
        w.sg[i] = (lane & h) ? -1.f : 1.f;
        if (i < 4) {
            const float2 v = __ldg(tw + (10 * (16 / h) * (lane & (h - 1))) % 320);
            w.st[i] = (lane & h) ? cpx{v.x, v.y} : cpx{1.f, 0.f};
        }
    }
    w.brev = (int)(__brev((unsigned)lane) >> 27);
    return w;
}
// 32-point FFT across the lanes, decimation in frequency: lane l holds element l on entry, element brev(l) on return.
// Branch-free butterflies: upper lane a + b, lower lane (a - b) w  ==  (partner + sign * own) * twiddle on every lane.
__device__ __forceinline__ cpx fft32_dif(cpx v, const LaneTw& w) {
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const cpx p = cshfl(v, 16 >> i);
        v = {fmaf(w.sg[i], v.x, p.x), fmaf(w.sg[i], v.y, p.y)};
        if (i < 4) v = cmul(v, w.st[i]);
    }
    return v;
}
// inverse: decimation in time, lane l holds element brev(l) on entry, element l on return (unscaled)
__device__ __forceinline__ cpx ifft32_dit(cpx v, const LaneTw& w) {
#pragma unroll
    for (int i = 4; i >= 0; --i) {
        if (i < 4) v = cmulc(v, w.st[i]);
        const cpx p = cshfl(v, 16 >> i);
        v = {fmaf(w.sg[i], v.x, p.x), fmaf(w.sg[i], v.y, p.y)};
    }
    return v;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// ------------------------------------------------------------------ STFT + compress
constexpr int FT = 16;          // frames per CTA
constexpr int SWARPS = 8;       // warps per CTA, one frame at a time each

__global__ void __launch_bounds__(SWARPS * 32)
stft_compress_kernel(const float* __restrict__ wav, const float* __restrict__ rms, const float* __restrict__ tab,
                     const int* __restrict__ lengths, float* __restrict__ out, int Lpitch, int T, int compress) {
    __shared__ __align__(16) float sx[(FT + 1) * HOP];
    __shared__ float sz[SWARPS][2][160];
    const int b = blockIdx.y, t0 = blockIdx.x * FT, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* w = wav + (size_t)b * Lpitch;
    const float inv = rms ? 1.f / rms[b] : 1.f;
    const int L = lengths ? lengths[b] : Lpitch;        // reflect padding happens at the utterance's own end
    const int Tb = 1 + L / HOP;                          // frames past it are written as zeros
    for (int i = tid; i < (FT + 1) * HOP; i += SWARPS * 32) {
        int g = t0 * HOP + i - HOP;                 // index into the unpadded signal
        if (g < 0) g = -g;                           // reflect (center=True)
        if (g >= L) g = 2 * (L - 1) - g;
        sx[i] = (g >= 0 && g < L) ? w[g] * inv : 0.f;
    }
    const LaneTw tw = lane_twiddles(tab, lane);
    float2 win[5];
#pragma unroll
    for (int m = 0; m < 5; ++m) win[m] = __ldg(reinterpret_cast<const float2*>(tab + TAB_HANN) + lane + 32 * m);
    const float2* w320 = reinterpret_cast<const float2*>(tab + TAB_TW);
    __syncthreads();
    float* zr = sz[warp][0];
    float* zi = sz[warp][1];
    for (int f = warp; f < FT && t0 + f < T; f += SWARPS) {
        const int t = t0 + f;
        float* o_re = out + (((size_t)b * 2 + 0) * T + t) * NF;
        float* o_im = out + (((size_t)b * 2 + 1) * T + t) * NF;
        if (t >= Tb) {
            for (int k = lane; k < NF; k += 32) o_re[k] = 0.f, o_im[k] = 0.f;
            continue;
        }
        // z[m] = x[2m] + i x[2m+1] (windowed); lane j holds m = j + 32 m2
        cpx x[5];
#pragma unroll
        for (int m = 0; m < 5; ++m) {
            const float2 v = *reinterpret_cast<const float2*>(sx + f * HOP + 2 * (lane + 32 * m));
            x[m] = {v.x * win[m].x, v.y * win[m].y};
        }
        radix5<false>(x);
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            cpx v = q ? cmul(x[q], tw.tq[q - 1]) : x[q];
            v = fft32_dif(v, tw);             // Z[5 k1 + q] with k1 = brev(lane)
            zr[5 * tw.brev + q] = v.x;
            zi[5 * tw.brev + q] = v.y;
        }
        __syncwarp();
        // real-input untangling: A = Z[k], B = conj(Z[160-k]); E = (A+B)/2, O = (A-B)/(2i); X[k] = E + w O, X[160-k] = conj(E - w O)
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int k = lane + 32 * i;
            if (k <= 80) {
                const int k2 = k ? 160 - k : 0;
                const float ar = zr[k], ai = zi[k], br = zr[k2], bi = -zi[k2];
                const float er = 0.5f * (ar + br), ei = 0.5f * (ai + bi);
                const float orr = 0.5f * (ai - bi), oi = -0.5f * (ar - br);
                const float2 wk = __ldg(w320 + k);
                const float pr = wk.x * orr - wk.y * oi, pi = wk.x * oi + wk.y * orr;
                float x1r = er + pr, x1i = ei + pi, x2r = er - pr, x2i = -(ei - pi);
                if (compress) {   // z * |z|^(-1/2); 0 where |z| = 0 (atan2(0,0) = 0, mag = 0)
                    const float m1 = sqrt_approx(x1r * x1r + x1i * x1i), m2 = sqrt_approx(x2r * x2r + x2i * x2i);
                    const float g1 = m1 > 0.f ? rsqrtf(m1) : 0.f, g2 = m2 > 0.f ? rsqrtf(m2) : 0.f;
                    x1r *= g1, x1i *= g1, x2r *= g2, x2i *= g2;
                }
                o_re[k] = x1r;
                o_im[k] = x1i;
                if (k != 80) {
                    o_re[160 - k] = x2r;
                    o_im[160 - k] = x2i;
                }
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------ decompress + ISTFT
constexpr int FI = 16;        // hop blocks per CTA  (FI+1 frames are synthesised)

__global__ void __launch_bounds__(SWARPS * 32)
decompress_istft_kernel(const float* __restrict__ spec, const float* __restrict__ rms, const float* __restrict__ tab,
                        const int* __restrict__ lengths, float* __restrict__ wav, short* __restrict__ pcm, int Lpitch,
                        int Tpitch, int decompress, int pcm_clip) {
    __shared__ __align__(16) float fr[(FI + 1) * NFFT];      // windowed synthesis frames
    __shared__ float sz[SWARPS][2][160];
    const int b = blockIdx.y, c0 = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tf = c0 * FI;                       // first frame synthesised by this CTA
    const int L = lengths ? lengths[b] : Lpitch;  // ragged batch: frames / samples past the utterance's end do not exist
    const int T = lengths ? min(Tpitch, 1 + L / HOP) : Tpitch;
    const LaneTw tw = lane_twiddles(tab, lane);
    float2 win[5];
#pragma unroll
    for (int m = 0; m < 5; ++m) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(tab + TAB_HANN) + lane + 32 * m);
        win[m] = make_float2(v.x * (1.f / 160.f), v.y * (1.f / 160.f));     // synthesis window and the 1/N of the inverse
    }
    const float2* w320 = reinterpret_cast<const float2*>(tab + TAB_TW);
    float* zr = sz[warp][0];
    float* zi = sz[warp][1];
    for (int lf = warp; lf < FI + 1 && tf + lf < T; lf += SWARPS) {
        const int t = tf + lf;
        const float* s_re = spec + (((size_t)b * 2 + 0) * Tpitch + t) * NF;
        const float* s_im = spec + (((size_t)b * 2 + 1) * Tpitch + t) * NF;
        // Z[k] = E + i O with E = (X[k] + conj X[160-k]) / 2, O = conj(w_k) (X[k] - conj X[160-k]) / 2
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int k = lane + 32 * i;
            if (k <= 80) {
                float ar = s_re[k], ai = s_im[k], br = s_re[160 - k], bi = s_im[160 - k];
                if (decompress) {   // z * |z|  (mag^2, phase kept)
                    const float m1 = sqrt_approx(ar * ar + ai * ai), m2 = sqrt_approx(br * br + bi * bi);
                    ar *= m1, ai *= m1, br *= m2, bi *= m2;
                }
                if (k == 0) ai = 0.f, bi = 0.f;     // irfft ignores the imaginary parts of the DC and Nyquist bins
                bi = -bi;                            // B = conj X[160-k]
                const float er = 0.5f * (ar + br), ei = 0.5f * (ai + bi);
                const float dr = 0.5f * (ar - br), di = 0.5f * (ai - bi);
                const float2 wk = __ldg(w320 + k);
                const float orr = wk.x * dr + wk.y * di, oi = wk.x * di - wk.y * dr;     // conj(w) * d
                zr[k] = er - oi;
                zi[k] = ei + orr;
                if (k != 0 && k != 80) {             // Z[160-k] = conj(E) + i conj(O)
                    zr[160 - k] = er + oi;
                    zi[160 - k] = -ei + orr;
                }
            }
        }
        __syncwarp();
        cpx y[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            cpx v = {zr[5 * tw.brev + q], zi[5 * tw.brev + q]};
            v = ifft32_dit(v, tw);
            y[q] = q ? cmulc(v, tw.tq[q - 1]) : v;
        }
        radix5<true>(y);     // y[m2] = 160 z[lane + 32 m2]
#pragma unroll
        for (int m = 0; m < 5; ++m)
            *reinterpret_cast<float2*>(fr + lf * NFFT + 2 * (lane + 32 * m)) = make_float2(y[m].x * win[m].x, y[m].y * win[m].y);
        __syncwarp();
    }
    __syncthreads();
    const float scale = rms ? rms[b] : 1.f;
    const float* hann = tab + TAB_HANN;
    for (int i = tid; i < FI * HOP; i += SWARPS * 32) {
        const int jj = i / HOP, r = i % HOP;
        const int hb = 1 + tf + jj;                 // hop block in padded coordinates
        const long o = (long)HOP * (hb - 1) + r;    // output sample (centre pad stripped)
        if (o >= Lpitch) continue;
        float v = 0.f;
        if (o < L) {                                // (else: padding of a ragged batch)
            const int ta = hb - 1, tb = hb;         // frames overlapping this block
            float acc = 0.f, env = 0.f;
            if (ta < T) {
                acc += fr[jj * NFFT + HOP + r];
                env += hann[HOP + r] * hann[HOP + r];
            }
            if (tb < T) {
                acc += fr[(jj + 1) * NFFT + r];
                env += hann[r] * hann[r];
            }
            v = env > 1e-11f ? acc / env * scale : 0.f;
        }
        wav[(size_t)b * Lpitch + o] = v;
        if (pcm) {   // the writer's float -> PCM_16 conversion (see pdse_f32_to_pcm16)
            int q;
            if (pcm_clip) {
                const float sc = v * 32768.f;
                q = sc >= 32767.f ? 32767 : sc <= -32768.f ? -32768 : __float2int_rn(sc);
            } else {
                q = (int)(short)__float2int_rn(v * 32767.f);
            }
            pcm[(size_t)b * Lpitch + o] = (short)q;
        }
    }
}

}  // namespace pdse

// ---------------------------------------------------------------------------- C ABI
extern "C" int pdse_signal_table_floats(void) { return pdse::TAB_FLOATS; }

// Host-side table builder: Hann window and the 320th roots of unity, both evaluated in float64.
extern "C" int pdse_signal_tables(float* host_out) {
    using namespace pdse;
    const double two_pi = 6.283185307179586476925286766559;
    for (int n = 0; n < NFFT; ++n) host_out[TAB_HANN + n] = (float)(0.5 - 0.5 * cos(two_pi * n / NFFT));
    for (int r = 0; r < 320; ++r) {
        host_out[TAB_TW + 2 * r] = (float)cos(two_pi * r / 320);
        host_out[TAB_TW + 2 * r + 1] = (float)(-sin(two_pi * r / 320));
    }
    return 0;
}

extern "C" int pdse_rms_ragged_f32(const float* wav, const int* lengths, int B, int L, float* rms, void* stream) {
    using namespace pdse;
    if (B <= 0 || L <= 0) return set_error("pdse_rms_f32: empty input");
    rms_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(wav, lengths, L, rms);
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
    stft_compress_kernel<<<grid, SWARPS * 32, 0, (cudaStream_t)stream>>>(wav, rms, tables, lengths, out, L, T, compress);
    return check_launch("pdse_stft_compress_f32");
}
extern "C" int pdse_stft_compress_f32(const float* wav, const float* rms, const float* tables, float* out, int B,
                                      int L, int compress, void* stream) {
    return pdse_stft_compress_ragged_f32(wav, rms, tables, nullptr, out, B, L, compress, stream);
}

// pcm (optional, int16 [B][L]): the same samples converted as the reference's writer does (:1018), fused into the store
extern "C" int pdse_decompress_istft_pcm16_f32(const float* spec, const float* rms, const float* tables, const int* lengths,
                                               float* wav, short* pcm, int B, int T, int L, int decompress, int pcm_clip,
                                               void* stream) {
    using namespace pdse;
    if (B <= 0 || T <= 0 || L <= 0) return set_error("pdse_decompress_istft_f32: empty input");
    if (L > HOP * T) return set_error("pdse_decompress_istft_f32: length exceeds the frames' support");
    const int nblocks = ceil_div(L, HOP);
    dim3 grid(ceil_div(nblocks, FI), B);
    decompress_istft_kernel<<<grid, SWARPS * 32, 0, (cudaStream_t)stream>>>(spec, rms, tables, lengths, wav, pcm, L, T, decompress,
                                                                            pcm_clip);
    return check_launch("pdse_decompress_istft_f32");
}
extern "C" int pdse_decompress_istft_ragged_f32(const float* spec, const float* rms, const float* tables,
                                                const int* lengths, float* wav, int B, int T, int L, int decompress,
                                                void* stream) {
    return pdse_decompress_istft_pcm16_f32(spec, rms, tables, lengths, wav, nullptr, B, T, L, decompress, 0, stream);
}
extern "C" int pdse_decompress_istft_f32(const float* spec, const float* rms, const float* tables, float* wav, int B,
                                         int T, int L, int decompress, void* stream) {
    return pdse_decompress_istft_ragged_f32(spec, rms, tables, nullptr, wav, B, T, L, decompress, stream);
}
