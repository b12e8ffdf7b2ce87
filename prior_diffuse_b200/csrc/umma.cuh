// Device-side primitives for sm_100a: mbarrier, bulk async copy (UBLKCP), TMEM
// allocation, tcgen05.mma / commit / ld, and the shared-memory matrix descriptor
// for the K-major, no-swizzle ("interleaved") canonical layout.
//
// Operand convention used by every GEMM-shaped kernel in this library ("CP8"):
//   a matrix X[rows][K] (bf16) is stored as K/8 *chunk planes*; plane kc holds,
//   for every row r, the 8 consecutive K-elements 8kc..8kc+7 as one 16-byte unit:
//       addr(r, kc) = base + kc * plane_stride + r * 16
//   With SBO = 128 B (eight 16-byte rows form one contiguous core matrix) and
//   LBO = plane_stride this IS the tcgen05 canonical K-major SWIZZLE_NONE layout,
//   and -- because rows are a dense 16-byte sequence -- a window that starts at an
//   arbitrary row r0 is again canonical (start address += 16*r0).  Convolution
//   taps are therefore plain start-address shifts: implicit GEMM with no im2col.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "opfmt.h"

namespace pdse {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t addr, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    while (!mbar_try_wait(addr, parity)) {
    }
}

// --------------------------------------------------------- bulk async copies
// global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP).
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// generic-proxy writes to smem -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------------- TMEM
// Called by one full warp.  ncols: power of two in [32, 512].
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---------------------------------------------------------------- descriptors
// K-major, SWIZZLE_NONE shared-memory matrix descriptor (see header comment).
//   bits [0,14)  start address >> 4        bits [16,30) LBO >> 4 (K-chunk stride)
//   bits [32,46) SBO >> 4 (8-row stride)   bits [46,48) version = 1 (sm_100)
//   bits [61,64) layout type = 0 (no swizzle)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFFu);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    return d;
}
// Descriptors that differ only in the start address: add the byte offset (a multiple of 16) to the address field.
// Shared-memory addresses stay below 2^18, so the 14-bit field never carries into LBO.
__device__ __forceinline__ uint64_t dadd(uint64_t desc, uint32_t byte_off) { return desc + (uint64_t)(byte_off >> 4); }
// kind::f16 instruction descriptor: bf16 x bf16 -> fp32, both operands K-major.
//   [4,6) D fmt = 1 (f32)  [7,10) A fmt = 1 (bf16)  [10,13) B fmt = 1 (bf16)
//   [17,23) N>>3           [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t M, uint32_t N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// the same for the library's operand format (opfmt.h): A / B format code 0 = fp16, 1 = bf16
__host__ __device__ constexpr uint32_t make_idesc_op(uint32_t M, uint32_t N) {
    return PDSE_OP_FP16 ? (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24) : make_idesc_bf16(M, N);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; one thread issues for the whole CTA.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Same with the A operand in tensor memory: row m = TMEM lane m, 32-bit column c of the operand holds the K-elements
// (2c, 2c+1) as a bf16 pair (lower index in the low half) -- 8 columns per K = 16 instruction.
__device__ __forceinline__ void umma_bf16_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrives once every tcgen05 op issued so far by this thread has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// --------------------------------------------------------- TMEM -> registers
// 32 lanes x 32-bit, N consecutive columns: thread i of warp w reads lane 32*(w%4)+i.
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
// registers -> TMEM (same lane / column mapping as tmem_ld8)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------------- helpers
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
// 8 fp32 -> one 16-byte CP8 unit
// ---- operand format (opfmt.h): fp32 -> operand pair / scalar, operand pair -> fp32
#if PDSE_OP_FP16
typedef __half op_t;
constexpr uint32_t OP_ONE_PAIR = 0x3C003C00u;          // (1.0, 1.0)
__device__ __forceinline__ uint32_t pack_op(float lo, float hi) {      // saturates at +-65504 (no inf / NaN from a large activation)
    uint32_t p;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(hi), "f"(lo));
    return p;
}
__device__ __forceinline__ op_t to_op(float x) {
    const uint32_t p = pack_op(x, 0.f);
    const unsigned short lo = (unsigned short)(p & 0xffffu);
    return *reinterpret_cast<const op_t*>(&lo);
}
__device__ __forceinline__ float op_to_float(op_t x) { return __half2float(x); }
__device__ __forceinline__ unsigned short op_bits(op_t x) { return __half_as_ushort(x); }
__device__ __forceinline__ float2 op_pair_to_float2(uint32_t p) { return __half22float2(*reinterpret_cast<const __half2*>(&p)); }
#else
typedef __nv_bfloat16 op_t;
constexpr uint32_t OP_ONE_PAIR = 0x3F803F80u;
__device__ __forceinline__ uint32_t pack_op(float lo, float hi) { return pack_bf16(lo, hi); }
__device__ __forceinline__ op_t to_op(float x) { return __float2bfloat16(x); }
__device__ __forceinline__ float op_to_float(op_t x) { return __bfloat162float(x); }
__device__ __forceinline__ unsigned short op_bits(op_t x) { return __bfloat16_as_ushort(x); }
__device__ __forceinline__ float2 op_pair_to_float2(uint32_t p) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&p)); }
#endif
// 8 fp32 -> one 16-byte CP8 unit
__device__ __forceinline__ uint4 pack8(const float* v) {
    return make_uint4(pack_op(v[0], v[1]), pack_op(v[2], v[3]), pack_op(v[4], v[5]), pack_op(v[6], v[7]));
}
__device__ __forceinline__ float fast_sigmoid(float x) {
    // sigmoid(x) = 0.5 * tanh(0.5 x) + 0.5 : one MUFU op
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
    return fmaf(0.5f, t, 0.5f);
}
__device__ __forceinline__ float prelu(float x, float a) { return x >= 0.f ? x : a * x; }

// ------------------------------------------------ legacy warp-level MMA (small fp32-accurate GEMMs only)
// mma.sync m16n8k8 TF32: fragments (g = lane/4, t = lane%4): A a0 (g,t) a1 (g+8,t) a2 (g,t+4) a3 (g+8,t+4);
// B b0 (k=t, n=g) b1 (k=t+4, n=g); C/D c0 (g,2t) c1 (g,2t+1) c2 (g+8,2t) c3 (g+8,2t+1).
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const float (&a)[4], float b0, float b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
                   "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}

// One GEMM "phase" helper: the CTA-wide protocol around a batch of MMAs issued by
// thread 0.  Callers: (1) all threads finish writing operands / reading TMEM, then
// call phase_begin(); (2) thread 0 issues umma_bf16(...) calls; (3) all call
// phase_end(bar, parity) which commits and waits for the MMAs to land in TMEM.
__device__ __forceinline__ void phase_begin() {
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
}
__device__ __forceinline__ void phase_end(uint64_t* bar, uint32_t& parity) {
    if (threadIdx.x == 0) umma_commit(bar);
    mbar_wait(bar, parity);
    parity ^= 1u;
    __syncwarp();   // re-converge before the .sync.aligned tcgen05.ld of the epilogue
    tc_fence_after();
}
// same, when thread 0 has already committed (work was placed between the MMA issue and the wait)
__device__ __forceinline__ void phase_wait(uint64_t* bar, uint32_t& parity) {
    mbar_wait(bar, parity);
    parity ^= 1u;
    __syncwarp();
    tc_fence_after();
}
// mma.sync m16n8k16 bf16 (fragments: A a0 (g, 2t..2t+1) a1 (g+8, ..) a2 (g, 2t+8..) a3 (g+8, 2t+8..); B b0 (k = 2t..2t+1, n = g)
// b1 (k = 2t+8.., n = g); C/D as m16n8k8)
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                               uint32_t b1) {
#if PDSE_OP_FP16
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
#else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
#endif
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// ------------------------------------------------------------------- bias MMAs
// A per-channel bias is stored as a B-operand block [2][N][8] = (hi, lo, 0...) and added by one K = 16 MMA against a
// constant "ones plane" A operand (1, 1, 0, ...): no epilogue has to load or add a per-channel bias.
// D (+)= bias: one K=16 MMA of the ones plane against a bias block
__device__ __forceinline__ void umma_bias(uint32_t d, uint32_t ones, uint32_t block, uint32_t N, uint32_t accumulate) {
    umma_bf16(d, make_smem_desc(ones, 2048, 128), make_smem_desc(block, N * 16, 128), make_idesc_op(128, N), accumulate);
}
__device__ __forceinline__ void init_ones_plane(uint8_t* ones, int tid, int nthr) {
    for (int i = tid; i < 256; i += nthr)
        reinterpret_cast<uint4*>(ones)[i] = i < 128 ? make_uint4(OP_ONE_PAIR, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
}
__device__ __forceinline__ float tanh_fast(float x) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x));
    return t;
}


}  // namespace pdse
