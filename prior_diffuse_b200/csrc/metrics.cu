// Segmental SNR on the device (utils/metrics.py:36-55, SNRseg): the scalar the reference's evaluation tracks for
// every enhanced utterance, computed where the waveforms already are so that a multi-GPU gather carries one
// float per utterance instead of the waveforms.
#include "common.cuh"

namespace pdse {
namespace {

constexpr int SS_WIN = 480;     // round(0.03 * 16000)
constexpr int SS_HOP = 120;     // floor(0.25 * 0.03 * 16000)

// one CTA per utterance, one warp per 480-sample frame; frame k starts at 120 k; the last frame is dropped (:53)
__global__ void __launch_bounds__(256) ssnr_kernel(const float* __restrict__ clean, const float* __restrict__ proc,
                                                   const int* __restrict__ lengths, int Lpitch, float* __restrict__ out) {
    __shared__ float win[SS_WIN];
    __shared__ double part[8];
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = lengths ? lengths[b] : Lpitch;
    const int nfr = (L - (SS_WIN - SS_HOP)) / SS_HOP - 1;          // frames that enter the mean
    for (int i = tid; i < SS_WIN; i += 256)
        win[i] = (float)(0.5 * (1.0 - cos(6.283185307179586476925286766559 * (double)(i + 1) / (double)(SS_WIN + 1))));
    __syncthreads();
    const float* c = clean + (size_t)b * Lpitch;
    const float* p = proc + (size_t)b * Lpitch;
    double acc = 0.0;
    for (int k = warp; k < nfr; k += 8) {
        float se = 0.f, ne = 0.f;
        for (int i = lane; i < SS_WIN; i += 32) {
            const float w = win[i], x = c[k * SS_HOP + i] * w, y = p[k * SS_HOP + i] * w;
            se = fmaf(x, x, se);
            ne = fmaf(x - y, x - y, ne);
        }
        for (int o = 16; o; o >>= 1) {
            se += __shfl_xor_sync(0xffffffffu, se, o);
            ne += __shfl_xor_sync(0xffffffffu, ne, o);
        }
        const double eps = 2.220446049250313e-16;
        double snr = 10.0 * log10((double)se / ((double)ne + eps) + eps);
        snr = fmin(fmax(snr, -10.0), 35.0);
        acc += snr;
    }
    if (lane == 0) part[warp] = acc;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int i = 0; i < 8; ++i) s += part[i];
        out[b] = nfr > 0 ? (float)(s / nfr) : 0.f;
    }
}

}  // namespace
}  // namespace pdse

extern "C" int pdse_ssnr_f32(const float* clean, const float* processed, const int* lengths, int B, int L, float* out,
                             void* stream) {
    using namespace pdse;
    if (B <= 0 || L < 2 * SS_WIN) return set_error("pdse_ssnr_f32: need at least two 30 ms frames");
    ssnr_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(clean, processed, lengths, L, out);
    return check_launch("pdse_ssnr_f32");
}
