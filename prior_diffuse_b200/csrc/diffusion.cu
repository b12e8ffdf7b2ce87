// Element-wise pieces of the reverse loop (trainer/complex_ddpm_trainer.py:947-998):
//   x_T = N(0, I) [* sqrt(mask)]                         :950-956
//   x   = c1 * (x - c2 * eps) [+ sigma * z * sqrt(mask)] :977-992  (sigma == 0 in the reference)
//   S   = (x + X0) * c                                   :995-997
// All HBM-bound: float4 accesses, one pass, Philox4x32-10 noise generated on device
// (seed, element offset) so a run is reproducible without any host traffic.
#include "common.cuh"

namespace pdse {

struct Philox {
    static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    // counter words (ctr lo, ctr hi, subsequence lo, subsequence hi), key = seed: Philox4x32-10 as in Random123 / cuRAND
    __device__ static uint4 rand4(uint64_t seed, uint64_t ctr, uint64_t subseq = 0) {
        uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
        uint4 c = make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)subseq, (uint32_t)(subseq >> 32));
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
            const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
            c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
            k0 += W0;
            k1 += W1;
        }
        return c;
    }
    // four standard normals from one counter (Box-Muller on 2x2 uniforms)
    __device__ static float4 normal4(uint64_t seed, uint64_t ctr) {
        const uint4 r = rand4(seed, ctr);
        const float s = 2.3283064365386963e-10f;   // 2^-32
        const float u0 = fmaf((float)r.x, s, 0.5f * s), u1 = (float)r.y * s;
        const float u2 = fmaf((float)r.z, s, 0.5f * s), u3 = (float)r.w * s;
        const float r0 = sqrtf(-2.f * __logf(u0)), r1 = sqrtf(-2.f * __logf(u2));
        float s0, c0, s1, c1;
        __sincosf(6.283185307179586f * u1, &s0, &c0);
        __sincosf(6.283185307179586f * u3, &s1, &c1);
        return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
    }
};

__device__ __forceinline__ float mask_sqrt(float x0, float inv_max) {
    return sqrtf(fmaf(0.5f * fabsf(x0), inv_max, 0.5f));   // sqrt(0.5 + 0.5 |X0| / max)
}

// per (b, ch) max |x| over T*F elements           (:951-952)
// ragged batch: only the utterance's own 1 + len/160 frames (a prefix of its [T][161] plane) take part
__global__ void absmax_kernel(const float* __restrict__ x, const int* __restrict__ lengths, int n, float* __restrict__ out) {
    const float* p = x + (size_t)blockIdx.x * n;
    const int nv = lengths ? min(n, (1 + lengths[blockIdx.x >> 1] / 160) * 161) : n;
    float m = 0.f;
    for (int i = threadIdx.x; i < nv; i += blockDim.x) m = fmaxf(m, fabsf(p[i]));
    __shared__ float red[32];
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) out[blockIdx.x] = m;
    }
}

// x[i] = ((gen ? N(0,1) : x[i]) + (add ? add[i] : 0)) * (mask ? sqrt(0.5 + 0.5|x0|/max) : 1)
__global__ void init_state_kernel(float* __restrict__ x, const float* __restrict__ x0, const float* __restrict__ amax,
                                  const float* __restrict__ add, long n4, int plane, long last, int gen, uint64_t seed,
                                  uint64_t offset) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        float4 v = gen ? Philox::normal4(seed, offset + (uint64_t)i) : reinterpret_cast<float4*>(x)[i];
        if (add) {   // deltamu: x_T = z + X_init (:947-948)
            const float4 a4 = reinterpret_cast<const float4*>(add)[i];
            v.x += a4.x; v.y += a4.y; v.z += a4.z; v.w += a4.w;
        }
        if (x0) {
            const float4 z = reinterpret_cast<const float4*>(x0)[i];
            const long e = i * 4;
            // plane % 4 != 0 in general, so the four lanes may straddle a (b,ch) boundary
            v.x *= mask_sqrt(z.x, 1.f / amax[min((e + 0) / plane, last)]);
            v.y *= mask_sqrt(z.y, 1.f / amax[min((e + 1) / plane, last)]);
            v.z *= mask_sqrt(z.z, 1.f / amax[min((e + 2) / plane, last)]);
            v.w *= mask_sqrt(z.w, 1.f / amax[min((e + 3) / plane, last)]);
        }
        reinterpret_cast<float4*>(x)[i] = v;
    }
}

// x = c1 * (x - c2 * eps) + sigma * z [* sqrt(mask)];  finalize: out = (x + x0) * scale
__global__ void ddpm_update_kernel(float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ x0,
                                   const float* __restrict__ amax, float* __restrict__ out, long n4, int plane, long last,
                                   float c1, float c2, float sigma, int use_mask, int finalize, float scale,
                                   uint64_t seed, uint64_t offset) {
    pdl_trigger();
    pdl_wait();     // eps is the last decoder block's output: only this kernel's launch overlaps that block's tail
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        float4 v = reinterpret_cast<const float4*>(x)[i];
        const float4 e = reinterpret_cast<const float4*>(eps)[i];
        v.x = c1 * fmaf(-c2, e.x, v.x);
        v.y = c1 * fmaf(-c2, e.y, v.y);
        v.z = c1 * fmaf(-c2, e.z, v.z);
        v.w = c1 * fmaf(-c2, e.w, v.w);
        float4 z0 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (finalize == 1 || use_mask) z0 = reinterpret_cast<const float4*>(x0)[i];
        if (sigma != 0.f) {
            float4 z = Philox::normal4(seed, offset + (uint64_t)i);
            if (use_mask) {
                const long el = i * 4;
                z.x *= mask_sqrt(z0.x, 1.f / amax[min((el + 0) / plane, last)]);
                z.y *= mask_sqrt(z0.y, 1.f / amax[min((el + 1) / plane, last)]);
                z.z *= mask_sqrt(z0.z, 1.f / amax[min((el + 2) / plane, last)]);
                z.w *= mask_sqrt(z0.w, 1.f / amax[min((el + 3) / plane, last)]);
            }
            v.x = fmaf(sigma, z.x, v.x);
            v.y = fmaf(sigma, z.y, v.y);
            v.z = fmaf(sigma, z.z, v.z);
            v.w = fmaf(sigma, z.w, v.w);
        }
        if (finalize) {
            if (finalize == 2) z0 = make_float4(0.f, 0.f, 0.f, 0.f);   // deltamu / condition branches: no "+ X_init" (:993-994)
            v.x = (v.x + z0.x) * scale;
            v.y = (v.y + z0.y) * scale;
            v.z = (v.z + z0.z) * scale;
            v.w = (v.w + z0.w) * scale;
            reinterpret_cast<float4*>(out)[i] = v;
        } else {
            reinterpret_cast<float4*>(x)[i] = v;
        }
    }
}

// ---------------------------------------------------------------- ATen-compatible N(0, 1) stream (validation mode)
// torch.randn_like on a CUDA tensor (trainer/complex_ddpm_trainer.py:950, :987) = at::native::normal_ ->
// distribution_elementwise_grid_stride_kernel<float, 4> (aten/src/ATen/native/cuda/DistributionTemplates.h): thread idx
// of a (grid x 256) launch owns Philox subsequence idx, starts at counter offset/4, and its i-th curand_normal4 call
// fills elements idx + stride*(4i + 0..3), stride = 256*grid.  curand_normal4 = two Box-Muller pairs computed exactly as
// curand_normal.h:70-87 does on the device (logf, sqrtf, __sincosf; constants 2.3283064e-10f and * 6.2831855f).
__device__ __forceinline__ float2 box_muller_curand(uint32_t x, uint32_t y) {
    const float u = x * 2.3283064e-10f + (2.3283064e-10f / 2);
    const float v = y * (2.3283064e-10f * 6.2831855f) + ((2.3283064e-10f * 6.2831855f) / 2);
    const float s = sqrtf(-2.0f * logf(u));
    float2 r;
    __sincosf(v, &r.x, &r.y);
    r.x *= s;
    r.y *= s;
    return r;
}
__global__ void __launch_bounds__(256) randn_aten_kernel(float* __restrict__ out, long numel, uint64_t seed, uint64_t offset) {
    const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
    const long stride = (long)blockDim.x * gridDim.x;
    const long rounded = ((numel - 1) / (stride * 4) + 1) * stride * 4;
    uint64_t ctr = offset >> 2;     // ATen keeps the offset a multiple of 4: the cuRAND state index stays 0
    for (long li0 = idx; li0 < rounded; li0 += stride * 4, ++ctr) {
        const uint4 r = Philox::rand4(seed, ctr, (uint64_t)idx);
        const float2 a = box_muller_curand(r.x, r.y), b = box_muller_curand(r.z, r.w);
        const float v[4] = {a.x, a.y, b.x, b.y};
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            const long li = li0 + stride * ii;
            if (li < numel) out[li] = v[ii];
        }
    }
}

// ---------------------------------------------------------------- float -> 16-bit PCM (the writer's conversion)
// trainer/complex_ddpm_trainer.py:1018 sf.write(path, wav, 16000): python-soundfile's default subtype for WAV is PCM_16
// and libsndfile (third-party, absent here; src/pcm.c f2s_array / f2s_clip_array) converts normalised floats as
//   clip = 0 (libsndfile default): (short) lrintf(x * 32767.f)                       -- wraps past full scale
//   clip = 1 (SFC_SET_CLIPPING):   x * 32768.f saturated to [-32768, 32767], lrintf
// lrintf rounds to nearest even = cvt.rni.  Parity unpinned (no libsndfile here): the rule above is the published one.
__global__ void pcm16_kernel(const float* __restrict__ wav, short* __restrict__ out, long n, int clip) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const float x = wav[i];
        int v;
        if (clip) {
            const float sc = x * 32768.f;
            v = sc >= 32767.f ? 32767 : sc <= -32768.f ? -32768 : __float2int_rn(sc);
        } else {
            v = (int)(short)__float2int_rn(x * 32767.f);
        }
        out[i] = (short)v;
    }
}

__global__ void scale_kernel(float* __restrict__ x, long n, float s) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) x[i] *= s;
}

inline int ew_grid(long n4) {
    long g = (n4 + 255) / 256;
    const long cap = 148L * 8;
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace pdse

extern "C" int pdse_absmax_ragged_f32(const float* x, const int* lengths, int rows, int n, float* out, void* stream) {
    using namespace pdse;
    if (rows <= 0 || n <= 0) return set_error("pdse_absmax_f32: empty input");
    absmax_kernel<<<rows, 512, 0, (cudaStream_t)stream>>>(x, lengths, n, out);
    return check_launch("pdse_absmax_f32");
}
extern "C" int pdse_absmax_f32(const float* x, int rows, int n, float* out, void* stream) {
    return pdse_absmax_ragged_f32(x, nullptr, rows, n, out, stream);
}

// Buffers are processed as float4: every pointer must be 16-byte aligned and have capacity for
// n rounded up to a multiple of 4 floats (the tail lanes are computed and stored, never read back).
extern "C" int pdse_init_state_add_f32(float* x, const float* x0, const float* amax, const float* add, long n, int plane,
                                       int generate, unsigned long long seed, unsigned long long offset, void* stream);
extern "C" int pdse_init_state_f32(float* x, const float* x0, const float* amax, long n, int plane, int generate,
                                   unsigned long long seed, unsigned long long offset, void* stream) {
    return pdse_init_state_add_f32(x, x0, amax, nullptr, n, plane, generate, seed, offset, stream);
}
extern "C" int pdse_init_state_add_f32(float* x, const float* x0, const float* amax, const float* add, long n, int plane,
                                       int generate, unsigned long long seed, unsigned long long offset, void* stream) {
    using namespace pdse;
    if (n <= 0) return set_error("pdse_init_state_f32: empty input");
    if (x0 && (!amax || plane <= 0)) return set_error("pdse_init_state_f32: mask needs amax and plane");
    const long last = plane > 0 ? (n - 1) / plane : 0;
    init_state_kernel<<<ew_grid((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>(x, x0, amax, add, (n + 3) / 4, plane, last,
                                                                              generate, seed, offset);
    return check_launch("pdse_init_state_f32");
}

extern "C" int pdse_ddpm_update_f32(float* x, const float* eps, const float* x0, const float* amax, float* out,
                                    long n, int plane, float c1, float c2, float sigma, int use_mask, int finalize,
                                    float scale, unsigned long long seed, unsigned long long offset, void* stream) {
    using namespace pdse;
    if (n <= 0) return set_error("pdse_ddpm_update_f32: empty input");
    if ((finalize == 1 || use_mask) && !x0) return set_error("pdse_ddpm_update_f32: x0 required");
    if (finalize && !out) return set_error("pdse_ddpm_update_f32: out required when finalize=1");
    if (use_mask && (!amax || plane <= 0)) return set_error("pdse_ddpm_update_f32: mask needs amax and plane");
    const long last = plane > 0 ? (n - 1) / plane : 0;
    PDSE_CUDA(launch_pdl(ddpm_update_kernel, dim3(ew_grid((n + 3) / 4)), dim3(256), 0, (cudaStream_t)stream, x, eps, x0, amax, out,
                         (n + 3) / 4, plane, last, c1, c2, sigma, use_mask, finalize, scale, (uint64_t)seed, (uint64_t)offset));
    return check_launch("pdse_ddpm_update_f32");
}

extern "C" int pdse_scale_f32(float* x, long n, float s, void* stream) {
    using namespace pdse;
    if (n <= 0) return set_error("pdse_scale_f32: empty input");
    scale_kernel<<<ew_grid((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>(x, n, s);
    return check_launch("pdse_scale_f32");
}

// ATen's launch policy for the normal kernel (calc_execution_policy, DistributionTemplates.h): 256 threads, grid =
// min(SMs * (max threads per SM / 256), ceil(n / 256)); the generator's offset then advances by
// ((n - 1) / (256 * grid * 4) + 1) * 4.
extern "C" int pdse_randn_aten_policy(long n, int* grid_x, unsigned long long* offset_increment) {
    using namespace pdse;
    if (n <= 0 || !grid_x || !offset_increment) return set_error("pdse_randn_aten_policy: bad arguments");
    int dev = 0, tpm = 2048;
    PDSE_CUDA(cudaGetDevice(&dev));
    PDSE_CUDA(cudaDeviceGetAttribute(&tpm, cudaDevAttrMaxThreadsPerMultiProcessor, dev));
    const long cap = (long)sm_count() * (tpm / 256), want = (n + 255) / 256;
    const long g = want < cap ? want : cap;
    *grid_x = (int)g;
    *offset_increment = (unsigned long long)(((n - 1) / (256 * g * 4) + 1) * 4);
    return PDSE_OK;
}
// out[n] = the values torch.randn(n, device="cuda") yields on this device for a generator at (seed, philox_offset)
extern "C" int pdse_randn_aten_f32(float* out, long n, unsigned long long seed, unsigned long long philox_offset, int grid_x,
                                   void* stream) {
    using namespace pdse;
    if (n <= 0 || grid_x <= 0) return set_error("pdse_randn_aten_f32: empty input");
    if (philox_offset & 3) return set_error("pdse_randn_aten_f32: the Philox offset must be a multiple of 4");
    randn_aten_kernel<<<grid_x, 256, 0, (cudaStream_t)stream>>>(out, n, seed, philox_offset);
    return check_launch("pdse_randn_aten_f32");
}

extern "C" int pdse_f32_to_pcm16(const float* wav, short* out, long n, int clip, void* stream) {
    using namespace pdse;
    if (n <= 0) return set_error("pdse_f32_to_pcm16: empty input");
    pcm16_kernel<<<ew_grid((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>(wav, out, n, clip);
    return check_launch("pdse_f32_to_pcm16");
}
