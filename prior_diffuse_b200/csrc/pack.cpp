// Weight packing on the HOST, behind the C ABI: reference state_dict tensors (fp32) -> the operand blobs the sm_100a
// kernels read (bf16 tcgen05 B operands in the K-major no-swizzle "CP8" layout [K/8][N][8], bias blocks, folded
// BatchNorm affines, composed gate / time projections).  Same arithmetic, in float64, as prior_diffuse_b200/pack.py,
// which stays as the NumPy statement of the layouts that tests/test_pack_emulation.py pins against the oracle;
// tests/test_pack_c.py checks these blobs against it element for element.
//
// A packed network is ONE contiguous blob with a fixed directory (the architectures are fixed): pdse_pack_layout()
// lists {name, dtype, byte offset, elements} of every section, so a host in any language can upload the blob with one
// copy and hand section pointers to the kernels (or call the composite forward entry points in forward.cu).
// Reference layouts: model/diff3.py (DiffUNet1), model/gcrn.py (GCRN); SURVEY.md C.1.
#include "../../include/pdse.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "opfmt.h"

namespace pdse {
namespace {

using vec = std::vector<double>;

struct StateDict {
    std::unordered_map<std::string, std::pair<const float*, long>> m;
    std::string missing;
    StateDict(const pdse_tensor* sd, int n) {
        for (int i = 0; i < n; ++i) m[sd[i].name] = {sd[i].data, sd[i].numel};
    }
    // tensor as float64 (row-major, as stored); numel must match when given
    vec get(const std::string& key, long numel = -1) {
        auto it = m.find(key);
        if (it == m.end() || (numel >= 0 && it->second.second != numel)) {
            if (missing.empty()) missing = key;
            return vec((size_t)(numel > 0 ? numel : 1), 0.0);
        }
        vec v((size_t)it->second.second);
        for (long i = 0; i < it->second.second; ++i) v[(size_t)i] = (double)it->second.first[i];
        return v;
    }
};

inline uint16_t f32_to_bf16(float f) {      // round to nearest even, as torch's .to(torch.bfloat16)
    uint32_t u;
    std::memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);   // NaN
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
// fp32 -> IEEE fp16, round to nearest even (as torch's .to(torch.float16)): inf for |f| >= 65520, subnormals kept
inline uint16_t f32_to_f16(float f) {
    uint32_t x;
    std::memcpy(&x, &f, 4);
    const uint16_t sign = (uint16_t)((x >> 16) & 0x8000u);
    x &= 0x7fffffffu;
    if (x > 0x7f800000u) return (uint16_t)(sign | 0x7e00u);          // NaN
    if (x >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);         // rounds to infinity
    if (x < 0x38800000u) {                                           // below 2^-14: subnormal result, unit 2^-24
        if (x < 0x33000000u) return sign;                            // below 2^-25 (a tie at exactly 2^-25 goes to even = 0)
        const int shift = 126 - (int)(x >> 23);                      // 14 .. 24
        const uint32_t m = (x & 0x7fffffu) | 0x800000u, half = 1u << (shift - 1), rem = m & ((1u << shift) - 1u);
        uint32_t h = m >> shift;
        if (rem > half || (rem == half && (h & 1u))) ++h;
        return (uint16_t)(sign | h);
    }
    uint32_t h = (((x >> 23) - 112u) << 10) | ((x >> 13) & 0x3ffu);
    const uint32_t rem = x & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;          // a carry into the exponent is the right result
    return (uint16_t)(sign | h);
}
inline float f16_to_f32(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
    float f;
    if (e == 0) {
        f = std::ldexp((float)m, -24);
        if (sign) f = -f;
        return f;
    }
    const uint32_t u = sign | (e == 31 ? 0x7f800000u | (m << 13) : ((e + 112u) << 23) | (m << 13));
    std::memcpy(&f, &u, 4);
    return f;
}
// the library's 16-bit operand format (opfmt.h)
inline uint16_t f32_to_op(float f) { return PDSE_OP_FP16 ? f32_to_f16(f) : f32_to_bf16(f); }
inline double op_round(double x) {
    if (PDSE_OP_FP16) return (double)f16_to_f32(f32_to_f16((float)x));
    const uint16_t h = f32_to_bf16((float)x);
    uint32_t u = (uint32_t)h << 16;
    float f;
    std::memcpy(&f, &u, 4);
    return (double)f;
}
// set when a finite weight does not fit the operand format (fp16: |w| >= 65520): the pack call fails instead of shipping inf
static thread_local bool g_op_overflow = false;

// W[N][K] -> [K/8][N][8], K zero-padded to a multiple of 8 (pack.cp8)
vec cp8(const vec& w, int N, int K) {
    const int kp = (K + 7) / 8 * 8;
    vec out((size_t)N * kp, 0.0);
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) out[((size_t)(k / 8) * N + n) * 8 + (k % 8)] = w[(size_t)n * K + k];
    return out;
}
// bias as a B operand [2][N][8]: row n = (hi, lo, 0...) of chunk 0 (pack.bias_block)
vec bias_block(const vec& b) {
    const size_t N = b.size();
    vec out(2 * N * 8, 0.0);
    for (size_t n = 0; n < N; ++n) {
        const double hi = op_round(b[n]);
        out[n * 8 + 0] = hi;
        out[n * 8 + 1] = b[n] - hi;
    }
    return out;
}
// C[M][N] = A[M][K] * B[K][N]
vec matmul(const vec& A, const vec& B, int M, int K, int N) {
    vec C((size_t)M * N, 0.0);
    for (int m = 0; m < M; ++m)
        for (int k = 0; k < K; ++k) {
            const double a = A[(size_t)m * K + k];
            for (int n = 0; n < N; ++n) C[(size_t)m * N + n] += a * B[(size_t)k * N + n];
        }
    return C;
}
vec transpose(const vec& A, int M, int N) {
    vec T((size_t)M * N);
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) T[(size_t)n * M + m] = A[(size_t)m * N + n];
    return T;
}
void bn_affine(StateDict& sd, const std::string& key, int C, vec& s, vec& sh) {
    const vec w = sd.get(key + ".weight", C), b = sd.get(key + ".bias", C);
    const vec mean = sd.get(key + ".running_mean", C), var = sd.get(key + ".running_var", C);
    s.resize(C);
    sh.resize(C);
    for (int c = 0; c < C; ++c) {
        s[c] = w[c] / std::sqrt(var[c] + 1e-5);
        sh[c] = b[c] - mean[c] * s[c];
    }
}
// [64][K] rows (l | r) -> [128][K] rows (l | r | 0.5 Wlc l | 0.5 Wrc r)   (pack.with_gates)
vec with_gates(const vec& wlr, int K, const vec& wlc, const vec& wrc) {
    vec out((size_t)128 * K);
    std::copy(wlr.begin(), wlr.end(), out.begin());
    const vec l(wlr.begin(), wlr.begin() + (size_t)32 * K), r(wlr.begin() + (size_t)32 * K, wlr.end());
    const vec gl = matmul(wlc, l, 32, 32, K), gr = matmul(wrc, r, 32, 32, K);
    for (size_t i = 0; i < gl.size(); ++i) {
        out[(size_t)64 * K + i] = 0.5 * gl[i];
        out[(size_t)96 * K + i] = 0.5 * gr[i];
    }
    return out;
}

struct Out {                      // sequential writer into the blob
    uint8_t* base;
    size_t off = 0;
    void h(const vec& v) {        // -> 16-bit operand format
        uint16_t* p = reinterpret_cast<uint16_t*>(base + off);
        for (size_t i = 0; i < v.size(); ++i) {
            p[i] = f32_to_op((float)v[i]);
            if (PDSE_OP_FP16 && (p[i] & 0x7fffu) == 0x7c00u && std::isfinite(v[i])) g_op_overflow = true;
        }
        off += v.size() * 2;
    }
    void f(const vec& v) {        // -> fp32
        float* p = reinterpret_cast<float*>(base + off);
        for (size_t i = 0; i < v.size(); ++i) p[i] = (float)v[i];
        off += v.size() * 4;
    }
    void align() { off = (off + 127) & ~(size_t)127; }
};

// gate / out matrices of a BiConv(Trans)GLU block as [out][in]
void gate_mats(StateDict& sd, const std::string& p, bool transposed, int cout, vec& wlc, vec& wrc, vec& w2) {
    wlc = sd.get(p + ".l_conv.weight", 32 * 32);
    wrc = sd.get(p + ".r_conv.weight", 32 * 32);
    w2 = sd.get(p + ".conv2.weight", 32 * cout);           // conv: [cout][32]; convT: [32][cout]
    if (transposed) {
        wlc = transpose(wlc, 32, 32);
        wrc = transpose(wrc, 32, 32);
        w2 = transpose(w2, 32, cout);
    }
}

struct Tail {
    vec w2, b_lr4, b_out, f;      // bf16: [w2] b_lr4 [b_out]; fp32: slope[4] | w2vec[32] b2[4]
};
// pack._glu_tail
Tail glu_tail(StateDict& sd, const std::string& p, bool transposed, const std::string& bn_key, const std::string& prelu_key,
              int cout, const vec* blr_extra) {
    Tail t;
    vec wlc, wrc, w2;
    gate_mats(sd, p, transposed, cout, wlc, wrc, w2);
    const vec b2 = sd.get(p + ".conv2.bias", cout);
    vec blr = sd.get(p + ".l.bias", 32);
    const vec rb = sd.get(p + ".r.bias", 32);
    blr.insert(blr.end(), rb.begin(), rb.end());
    if (blr_extra)
        for (int i = 0; i < 64; ++i) blr[i] += (*blr_extra)[i];
    vec s, sh;
    if (cout == 64) {
        bn_affine(sd, bn_key, 64, s, sh);
        vec w(64 * 32);
        for (int o = 0; o < 64; ++o)
            for (int k = 0; k < 32; ++k) w[o * 32 + k] = 0.5 * s[o] * w2[o * 32 + k];
        t.w2 = cp8(w, 64, 32);
    }
    const vec lcb = sd.get(p + ".l_conv.bias", 32), rcb = sd.get(p + ".r_conv.bias", 32);
    const vec bl(blr.begin(), blr.begin() + 32), br(blr.begin() + 32, blr.end());
    const vec gl = matmul(wlc, bl, 32, 32, 1), gr = matmul(wrc, br, 32, 32, 1);
    vec b4 = blr;
    for (int i = 0; i < 32; ++i) b4.push_back(0.5 * (gl[i] + lcb[i]));
    for (int i = 0; i < 32; ++i) b4.push_back(0.5 * (gr[i] + rcb[i]));
    t.b_lr4 = bias_block(b4);
    if (cout == 64) {
        vec bo(64);
        for (int o = 0; o < 64; ++o) bo[o] = b2[o] * s[o] + sh[o];
        t.b_out = bias_block(bo);
        const vec slope = sd.get(prelu_key + ".weight", 1);
        t.f = {slope[0], 0.0, 0.0, 0.0};
    } else {
        t.f.resize(36, 0.0);
        for (int k = 0; k < 32; ++k) t.f[k] = 0.5 * w2[k];
        t.f[32] = b2[0];
    }
    return t;
}
void emit_tail(Out& wb, Out& wf, const Tail& t) {
    if (!t.w2.empty()) wb.h(t.w2);
    wb.h(t.b_lr4);
    if (!t.b_out.empty()) wb.h(t.b_out);
    wf.f(t.f);
}

// l | r conv weights of a block as [64 out][32 in][kh=2][kw]
vec conv_lr(StateDict& sd, const std::string& p, int kw, bool transposed) {
    const vec wl = sd.get(p + ".l.weight", 32 * 32 * 2 * kw), wr = sd.get(p + ".r.weight", 32 * 32 * 2 * kw);
    vec out((size_t)64 * 32 * 2 * kw);
    for (int half = 0; half < 2; ++half) {
        const vec& w = half ? wr : wl;
        for (int o = 0; o < 32; ++o)
            for (int i = 0; i < 32; ++i)
                for (int t = 0; t < 2 * kw; ++t) {
                    const size_t src = transposed ? ((size_t)i * 32 + o) * 2 * kw + t : ((size_t)o * 32 + i) * 2 * kw + t;
                    out[(((size_t)(half * 32 + o)) * 32 + i) * 2 * kw + t] = w[src];
                }
    }
    return out;
}
// one tap [64][32] of conv_lr
vec tap(const vec& wlr, int kw, int dt, int df) {
    vec out(64 * 32);
    for (int o = 0; o < 64; ++o)
        for (int i = 0; i < 32; ++i) out[o * 32 + i] = wlr[(((size_t)o * 32 + i) * 2 + dt) * kw + df];
    return out;
}

// ---------------------------------------------------------------------------------------------- DiffUNet1 sections
constexpr int BIAS_ROW = 452;
const uint32_t kTimeTableBits[50 * 128] = {
#include "time_table.inc"
};
inline int bias_off_enc(int i) { return 2 + 32 * (i - 2); }
inline int bias_off_dec(int br, int i) { return 130 + 160 * br + 32 * (5 - i); }

void pack_enc1(StateDict& sd, Out& wb, Out& wf) {
    const std::string p = "en.conv1";
    const vec w1 = sd.get(p + ".conv1.weight", 32 * 2), b1 = sd.get(p + ".conv1.bias", 32);
    const vec wlr = conv_lr(sd, p, 5, false);                        // [64][32][2][5]
    vec w((size_t)64 * 20, 0.0), extra(64, 0.0);                     // Wf[o][c*10 + dt*5 + df]
    for (int o = 0; o < 64; ++o)
        for (int k = 0; k < 32; ++k)
            for (int t = 0; t < 10; ++t) {
                const double v = wlr[((size_t)o * 32 + k) * 10 + t];
                for (int c = 0; c < 2; ++c) w[(size_t)o * 20 + c * 10 + t] += v * w1[k * 2 + c];
                extra[o] += v * b1[k];
            }
    vec wlc, wrc, w2;
    gate_mats(sd, p, false, 64, wlc, wrc, w2);
    const vec g = with_gates(w, 20, wlc, wrc);                       // [128][20]
    vec padded((size_t)128 * 32, 0.0);
    for (int o = 0; o < 128; ++o)
        for (int k = 0; k < 20; ++k) padded[(size_t)o * 32 + k] = g[(size_t)o * 20 + k];
    wb.h(cp8(padded, 128, 32));
    emit_tail(wb, wf, glu_tail(sd, p, false, "en.en1.0", "en.en1.1", 64, &extra));
    wf.f(sd.get("preprocess.conv.weight", 8));
    const vec bp = sd.get("preprocess.conv.bias", 2);
    wf.f({bp[0], bp[1], 0.0, 0.0});
}
void pack_enc(StateDict& sd, int i, Out& wb, Out& wf) {
    const std::string p = "en.conv" + std::to_string(i);
    wb.h(cp8(sd.get(p + ".conv1.weight", 32 * 64), 32, 64));
    const vec wlr = conv_lr(sd, p, 3, false);
    vec wlc, wrc, w2;
    gate_mats(sd, p, false, 64, wlc, wrc, w2);
    for (int dt = 0; dt < 2; ++dt)
        for (int df = 0; df < 3; ++df) wb.h(cp8(with_gates(tap(wlr, 3, dt, df), 32, wlc, wrc), 128, 32));
    const std::string n = std::to_string(i);
    emit_tail(wb, wf, glu_tail(sd, p, false, "en.en" + n + ".0", "en.en" + n + ".1", 64, nullptr));
}
void pack_dec(StateDict& sd, const std::string& br, int i, Out& wb, Out& wf) {
    const std::string p = br + ".de" + std::to_string(i) + ".0";
    const int kw = i == 1 ? 5 : 3, g = (kw - 1) / 2;
    wb.h(cp8(transpose(sd.get(p + ".conv1.weight", 128 * 32), 128, 32), 32, 128));
    const vec wlr = conv_lr(sd, p, kw, true);
    vec wlc, wrc, w2;
    gate_mats(sd, p, true, i == 1 ? 1 : 64, wlc, wrc, w2);
    for (int dt = 0; dt < 2; ++dt)
        for (int a = 0; a <= g; ++a) wb.h(cp8(with_gates(tap(wlr, kw, dt, 2 * a), 32, wlc, wrc), 128, 32));
    for (int dt = 0; dt < 2; ++dt)
        for (int a = 0; a < g; ++a) wb.h(cp8(with_gates(tap(wlr, kw, dt, 2 * a + 1), 32, wlc, wrc), 128, 32));
    const std::string n = std::to_string(i);
    if (i == 1) emit_tail(wb, wf, glu_tail(sd, p, true, "", "", 1, nullptr));
    else emit_tail(wb, wf, glu_tail(sd, p, true, br + ".de" + n + ".2", br + ".de" + n + ".3", 64, nullptr));
}
void pack_tcm(StateDict& sd, int m, int r, Out& wb, Out& wf) {
    const std::string p = "TCMs." + std::to_string(m) + ".residual" + std::to_string(r);
    int perm[256];
    for (int kk = 0; kk < 256; ++kk) perm[kk] = (kk % 64) * 4 + kk / 64;       // kernel channel kk -> reference channel
    const vec c1 = sd.get(p + ".conv1.weight", 64 * 256);
    vec w1(64 * 256);
    for (int o = 0; o < 64; ++o)
        for (int kk = 0; kk < 256; ++kk) w1[o * 256 + kk] = c1[o * 256 + perm[kk]];
    wb.h(cp8(w1, 64, 256));
    const vec wm = sd.get(p + ".mainbranch.2.weight", 64 * 64 * 5), wk = sd.get(p + ".maskbranch.2.weight", 64 * 64 * 5);
    for (int which = 0; which < 2; ++which)
        for (int t = 0; t < 5; ++t) {
            vec w(64 * 64);
            for (int o = 0; o < 64; ++o)
                for (int i = 0; i < 64; ++i) w[o * 64 + i] = which ? 0.5 * wk[(o * 64 + i) * 5 + t] : wm[(o * 64 + i) * 5 + t];
            wb.h(cp8(w, 64, 64));
        }
    const vec c2 = sd.get(p + ".conv2.2.weight", 256 * 64);
    vec w3(256 * 64);
    for (int kk = 0; kk < 256; ++kk)
        for (int i = 0; i < 64; ++i) w3[kk * 64 + i] = c2[perm[kk] * 64 + i];
    wb.h(cp8(w3, 256, 64));
    wb.h(bias_block(sd.get(p + ".conv1.bias", 64)));
    wb.h(bias_block(sd.get(p + ".mainbranch.2.bias", 64)));
    vec bk = sd.get(p + ".maskbranch.2.bias", 64);
    for (auto& v : bk) v *= 0.5;
    wb.h(bias_block(bk));
    const vec b3r = sd.get(p + ".conv2.2.bias", 256);
    vec b3(256);
    for (int kk = 0; kk < 256; ++kk) b3[kk] = b3r[perm[kk]];
    wb.h(bias_block(b3));
    vec s, sh;
    bn_affine(sd, p + ".mainbranch.1", 64, s, sh);
    wf.f(s), wf.f(sh);
    bn_affine(sd, p + ".maskbranch.1", 64, s, sh);
    wf.f(s), wf.f(sh);
    bn_affine(sd, p + ".conv2.1", 64, s, sh);
    for (auto& v : s) v *= 0.5;
    wf.f(s), wf.f(sh);
    wf.f({sd.get(p + ".mainbranch.0.weight", 1)[0], sd.get(p + ".maskbranch.0.weight", 1)[0], sd.get(p + ".conv2.0.weight", 1)[0], 0.0});
}
void pack_time(StateDict& sd, Out& wf) {
    // diff3.py:89-95 builds the sinusoid table [50][128] in float32 with arguments up to 4.9e5, where one ulp of the
    // argument moves sin / cos by percents: the table is therefore not recomputed here but shipped as the exact float32
    // bit patterns torch produces for that expression (time_table.inc, written by tests/golden/make_time_table.py and
    // pinned by tests/test_pack_c.py)
    vec table(50 * 128);
    for (int i = 0; i < 50 * 128; ++i) {
        float f;
        std::memcpy(&f, &kTimeTableBits[i], 4);
        table[i] = (double)f;
    }
    wf.f(table);
    wf.f(sd.get("time_embedding.projection1.weight", 512 * 128));
    wf.f(sd.get("time_embedding.projection1.bias", 512));
    wf.f(sd.get("time_embedding.projection2.weight", 512 * 512));
    wf.f(sd.get("time_embedding.projection2.bias", 512));
    vec rows((size_t)BIAS_ROW * 512, 0.0), bias(BIAS_ROW, 0.0);
    {
        const vec w = sd.get("en.tp1.weight", 2 * 512), b = sd.get("en.tp1.bias", 2);
        std::copy(w.begin(), w.end(), rows.begin());
        bias[0] = b[0], bias[1] = b[1];
    }
    auto compose = [&](const vec& w1 /*[32][C]*/, int C, const std::string& tp, const vec& b1, int o) {
        const vec tw = sd.get(tp + ".weight", (long)C * 512), tb = sd.get(tp + ".bias", C);
        const vec r = matmul(w1, tw, 32, C, 512), bb = matmul(w1, tb, 32, C, 1);
        std::copy(r.begin(), r.end(), rows.begin() + (size_t)o * 512);
        for (int i = 0; i < 32; ++i) bias[o + i] = bb[i] + b1[i];
    };
    for (int i = 2; i <= 5; ++i) {
        const std::string n = std::to_string(i);
        compose(sd.get("en.conv" + n + ".conv1.weight", 32 * 64), 64, "en.tp" + n, sd.get("en.conv" + n + ".conv1.bias", 32), bias_off_enc(i));
    }
    const char* brs[2] = {"de_real", "de_imag"};
    for (int bi = 0; bi < 2; ++bi)
        for (int i = 5; i >= 1; --i) {
            const std::string p = std::string(brs[bi]) + ".de" + std::to_string(i) + ".0";
            compose(transpose(sd.get(p + ".conv1.weight", 128 * 32), 128, 32), 128, p + ".tp", sd.get(p + ".conv1.bias", 32), bias_off_dec(bi, i));
        }
    wf.f(rows);
    wf.f(bias);
}

// ---------------------------------------------------------------------------------------------- directory
struct Section {
    std::string name;
    int dtype;          // 0 = bf16, 1 = fp32
    long elems;
};
std::vector<Section> diffunet1_sections() {
    std::vector<Section> s;
    auto blk = [&](const std::string& n, long h, long f) {
        s.push_back({n + ".wb", 0, h});
        s.push_back({n + ".wf", 1, f});
    };
    blk("enc1", 9216, 16);
    for (int i = 2; i <= 5; ++i) blk("enc" + std::to_string(i), 31744, 4);
    for (int k = 0; k < 18; ++k) blk("tcm" + std::to_string(k), 80896, 388);
    for (int br = 0; br < 2; ++br)
        for (int i = 5; i >= 1; --i) blk("dec" + std::to_string(br) + "_" + std::to_string(i), i == 1 ? 47104 : 33792, i == 1 ? 36 : 4);
    const std::pair<const char*, long> tm[] = {{"time.table", 6400}, {"time.p1w", 65536}, {"time.p1b", 512}, {"time.p2w", 262144},
                                              {"time.p2b", 512}, {"time.rows", (long)BIAS_ROW * 512}, {"time.bias", BIAS_ROW}};
    for (auto& t : tm) s.push_back({t.first, 1, t.second});
    return s;
}

// GCRN (model/gcrn.py:87-166)
const int GCRN_CH[6] = {2, 16, 32, 64, 128, 256};
struct GDec {
    int cin, cout, fin, fout;
};
const GDec GCRN_DEC[6] = {{0, 0, 0, 0}, {32, 1, 80, 161}, {64, 16, 39, 80}, {128, 32, 19, 39}, {256, 64, 9, 19}, {512, 128, 4, 9}};
inline int gcrn_kb(int cin) { return (cin / 8) % 4 == 0 ? 2 : 1; }

std::vector<Section> gcrn_sections() {
    std::vector<Section> s;
    s.push_back({"conv1.wb", 0, 2 * 32 * 8});
    s.push_back({"conv1.wf", 1, 4 * 16});
    for (int i = 2; i <= 5; ++i) {
        const long cin = GCRN_CH[i - 1], cout = GCRN_CH[i];
        s.push_back({"conv" + std::to_string(i) + ".wb", 0, 3 * 2 * cout * cin});
        s.push_back({"conv" + std::to_string(i) + ".wf", 1, 4 * cout});
    }
    for (int layer = 1; layer <= 2; ++layer)
        for (int g = 0; g < 2; ++g) {
            const std::string n = "lstm" + std::to_string(layer) + "_" + std::to_string(g);
            s.push_back({n + ".w_ih", 0, 2048 * 512});
            s.push_back({n + ".w_hh", 0, 2048 * 512});
            s.push_back({n + ".wf", 1, 2048});
        }
    s.push_back({"ln.wf", 1, 4096});
    for (int br = 1; br <= 2; ++br) {
        for (int i = 5; i >= 2; --i) {
            const long cin = GCRN_DEC[i].cin, cout = GCRN_DEC[i].cout;
            const std::string n = "dec" + std::to_string(br) + "_" + std::to_string(i);
            s.push_back({n + ".w_even", 0, 2 * 2 * cout * cin});
            s.push_back({n + ".w_odd", 0, 2 * cout * cin});
            s.push_back({n + ".wf", 1, 4 * cout});
        }
        s.push_back({"out" + std::to_string(br) + ".wf", 1, 96 + 96 + 4 + 161 * 161 + 164});
    }
    return s;
}

long layout_of(const std::vector<Section>& secs, pdse_blob_entry* out, int cap) {
    size_t off = 0;
    for (size_t i = 0; i < secs.size(); ++i) {
        off = (off + 127) & ~(size_t)127;
        if (out && (int)i < cap) {
            std::snprintf(out[i].name, sizeof(out[i].name), "%s", secs[i].name.c_str());
            out[i].dtype = secs[i].dtype;
            out[i].offset = (long)off;
            out[i].elems = secs[i].elems;
        }
        off += (size_t)secs[i].elems * (secs[i].dtype ? 4 : 2);
    }
    return (long)((off + 127) & ~(size_t)127);
}

// weights -> the order the streaming GEMM consumes them (pack._stream): per n-tile, per tap: cp8 of the tile's rows
void stream(Out& wb, const std::vector<vec>& taps, int N, int K, int ntile) {
    for (int j = 0; j < N / ntile; ++j)
        for (const vec& w : taps) {
            const vec rows(w.begin() + (size_t)j * ntile * K, w.begin() + (size_t)(j + 1) * ntile * K);
            wb.h(cp8(rows, ntile, K));
        }
}
// interleave value / gate output channels per n-tile: [val(ct) | gate(ct)] blocks (pack.glu_rows); wv, wg [C][K]
vec glu_rows(const vec& wv, const vec& wg, int C, int K, int ntile) {
    const int ct = ntile / 2;
    vec out;
    out.reserve((size_t)2 * C * K);
    for (int j = 0; j < C / ct; ++j) {
        out.insert(out.end(), wv.begin() + (size_t)j * ct * K, wv.begin() + (size_t)(j + 1) * ct * K);
        out.insert(out.end(), wg.begin() + (size_t)j * ct * K, wg.begin() + (size_t)(j + 1) * ct * K);
    }
    return out;
}
vec glu_ep(const vec& bv, const vec& bg, const vec& s, const vec& sh, int ntile) {
    const int ct = ntile / 2;
    vec out;
    for (size_t j = 0; j < bv.size() / ct; ++j)
        for (const vec* v : {&bv, &bg, &s, &sh}) out.insert(out.end(), v->begin() + j * ct, v->begin() + (j + 1) * ct);
    return out;
}
// tap df of a [A][B][1][3] conv weight as [rows][cols]: conv (transposed = false): rows = A (out), cols = B (in);
// transposed conv ([Cin][Cout][1][3]): rows = B (out), cols = A (in)
vec tap3(const vec& w, int A, int B, int df, bool transposed) {
    vec out((size_t)A * B);
    for (int a = 0; a < A; ++a)
        for (int b = 0; b < B; ++b) {
            const double v = w[((size_t)a * B + b) * 3 + df];
            if (transposed) out[(size_t)b * A + a] = v;
            else out[(size_t)a * B + b] = v;
        }
    return out;
}

}  // namespace
}  // namespace pdse

using namespace pdse;

extern "C" long pdse_pack_layout(int net, pdse_blob_entry* out, int capacity, int* count) {
    std::vector<Section> secs;
    if (net == PDSE_NET_DIFFUNET1) secs = diffunet1_sections();
    else if (net == PDSE_NET_GCRN) secs = gcrn_sections();
    else {
        set_error("pdse_pack_layout: unknown network");
        return -1;
    }
    if (count) *count = (int)secs.size();
    return layout_of(secs, out, capacity);
}

extern "C" int pdse_pack_diffunet1(const pdse_tensor* sd_in, int n, void* blob_host) {
    if (!sd_in || n <= 0 || !blob_host) return set_error("pdse_pack_diffunet1: bad arguments");
    StateDict sd(sd_in, n);
    g_op_overflow = false;
    const std::vector<Section> secs = diffunet1_sections();
    std::vector<pdse_blob_entry> dir(secs.size());
    const long total = layout_of(secs, dir.data(), (int)dir.size());
    std::memset(blob_host, 0, (size_t)total);
    auto at = [&](const std::string& name) {
        for (auto& e : dir)
            if (name == e.name) return Out{(uint8_t*)blob_host, (size_t)e.offset};
        return Out{(uint8_t*)blob_host, 0};
    };
    auto check = [&](Out& o, const std::string& name) {
        for (auto& e : dir)
            if (name == e.name) return o.off == (size_t)e.offset + (size_t)e.elems * (e.dtype ? 4 : 2);
        return false;
    };
    bool ok = true;
    {
        Out wb = at("enc1.wb"), wf = at("enc1.wf");
        pack_enc1(sd, wb, wf);
        ok = ok && check(wb, "enc1.wb") && check(wf, "enc1.wf");
    }
    for (int i = 2; i <= 5; ++i) {
        const std::string nme = "enc" + std::to_string(i);
        Out wb = at(nme + ".wb"), wf = at(nme + ".wf");
        pack_enc(sd, i, wb, wf);
        ok = ok && check(wb, nme + ".wb") && check(wf, nme + ".wf");
    }
    for (int m = 0; m < 3; ++m)
        for (int r = 1; r <= 6; ++r) {
            const std::string nme = "tcm" + std::to_string(m * 6 + r - 1);
            Out wb = at(nme + ".wb"), wf = at(nme + ".wf");
            pack_tcm(sd, m, r, wb, wf);
            ok = ok && check(wb, nme + ".wb") && check(wf, nme + ".wf");
        }
    const char* brs[2] = {"de_real", "de_imag"};
    for (int bi = 0; bi < 2; ++bi)
        for (int i = 5; i >= 1; --i) {
            const std::string nme = "dec" + std::to_string(bi) + "_" + std::to_string(i);
            Out wb = at(nme + ".wb"), wf = at(nme + ".wf");
            pack_dec(sd, brs[bi], i, wb, wf);
            ok = ok && check(wb, nme + ".wb") && check(wf, nme + ".wf");
        }
    {
        Out wf = at("time.table");
        // the time sections are contiguous up to alignment: write them one by one at their own offsets
        struct Cap : Out {
        };
        std::vector<std::string> names = {"time.table", "time.p1w", "time.p1b", "time.p2w", "time.p2b", "time.rows", "time.bias"};
        // pack_time emits the seven arrays in this order; redirect each to its section
        std::vector<uint8_t> tmp((size_t)(6400 + 65536 + 512 + 262144 + 512 + (long)BIAS_ROW * 512 + BIAS_ROW) * 4);
        Out t{tmp.data(), 0};
        pack_time(sd, t);
        size_t src = 0;
        for (auto& nm : names)
            for (auto& e : dir)
                if (nm == e.name) {
                    std::memcpy((uint8_t*)blob_host + e.offset, tmp.data() + src, (size_t)e.elems * 4);
                    src += (size_t)e.elems * 4;
                }
        ok = ok && src == tmp.size() && t.off == tmp.size();
        (void)wf;
    }
    if (!sd.missing.empty()) {
        std::snprintf(error_buffer(), 512, "pdse_pack_diffunet1: state_dict entry '%s' is missing or has the wrong size", sd.missing.c_str());
        return PDSE_EINVAL;
    }
    if (!ok) return set_error("pdse_pack_diffunet1: internal layout mismatch");
    if (g_op_overflow) return set_error("pdse_pack_diffunet1: a folded weight exceeds the fp16 operand range (|w| >= 65520); build with PDSE_OPERANDS=bf16");
    return PDSE_OK;
}

extern "C" int pdse_pack_gcrn(const pdse_tensor* sd_in, int n, void* blob_host) {
    if (!sd_in || n <= 0 || !blob_host) return set_error("pdse_pack_gcrn: bad arguments");
    StateDict sd(sd_in, n);
    g_op_overflow = false;
    const std::vector<Section> secs = gcrn_sections();
    std::vector<pdse_blob_entry> dir(secs.size());
    const long total = layout_of(secs, dir.data(), (int)dir.size());
    std::memset(blob_host, 0, (size_t)total);
    bool ok = true;
    auto at = [&](const std::string& name) {
        for (auto& e : dir)
            if (name == e.name) return Out{(uint8_t*)blob_host, (size_t)e.offset};
        ok = false;
        return Out{(uint8_t*)blob_host, 0};
    };
    auto check = [&](Out& o, const std::string& name) {
        for (auto& e : dir)
            if (name == e.name) return o.off == (size_t)e.offset + (size_t)e.elems * (e.dtype ? 4 : 2);
        return false;
    };
    {   // conv1: k = c*3 + df (K = 6 -> 16), rows [val(16) | gate(16)]
        Out wb = at("conv1.wb"), wf = at("conv1.wf");
        const vec wv = sd.get("conv1.conv1.weight", 16 * 2 * 3), wg = sd.get("conv1.conv2.weight", 16 * 2 * 3);
        vec w((size_t)32 * 16, 0.0);
        for (int o = 0; o < 16; ++o)
            for (int k = 0; k < 6; ++k) {
                w[(size_t)o * 16 + k] = wv[o * 6 + k];
                w[(size_t)(16 + o) * 16 + k] = wg[o * 6 + k];
            }
        wb.h(cp8(w, 32, 16));
        vec s, sh;
        bn_affine(sd, "bn1", 16, s, sh);
        wf.f(glu_ep(sd.get("conv1.conv1.bias", 16), sd.get("conv1.conv2.bias", 16), s, sh, 32));
        ok = ok && check(wb, "conv1.wb") && check(wf, "conv1.wf");
    }
    for (int i = 2; i <= 5; ++i) {
        const std::string nme = "conv" + std::to_string(i);
        const int cin = GCRN_CH[i - 1], cout = GCRN_CH[i], ntile = std::min(256, 2 * cout);
        Out wb = at(nme + ".wb"), wf = at(nme + ".wf");
        const vec wv = sd.get(nme + ".conv1.weight", (long)cout * cin * 3), wg = sd.get(nme + ".conv2.weight", (long)cout * cin * 3);
        std::vector<vec> taps;
        for (int df = 0; df < 3; ++df) taps.push_back(glu_rows(tap3(wv, cout, cin, df, false), tap3(wg, cout, cin, df, false), cout, cin, ntile));
        stream(wb, taps, 2 * cout, cin, ntile);
        vec s, sh;
        bn_affine(sd, "bn" + std::to_string(i), cout, s, sh);
        wf.f(glu_ep(sd.get(nme + ".conv1.bias", cout), sd.get(nme + ".conv2.bias", cout), s, sh, ntile));
        ok = ok && check(wb, nme + ".wb") && check(wf, nme + ".wf");
        (void)gcrn_kb(cin);
    }
    for (int layer = 1; layer <= 2; ++layer)
        for (int g = 0; g < 2; ++g) {
            const std::string nme = "lstm" + std::to_string(layer) + "_" + std::to_string(g);
            const std::string p = "glstm.lstm_list" + std::to_string(layer) + "." + std::to_string(g);
            // gate-row order of the recurrence kernel: n = cta*128 + lane, lane = gate*32 + unit%32 -> reference row gate*512 + unit
            std::vector<int> rows(2048);
            for (int nn = 0; nn < 2048; ++nn) rows[nn] = ((nn % 128) / 32) * 512 + (nn / 128) * 32 + nn % 32;
            const vec wih = sd.get(p + ".weight_ih_l0", 2048 * 512), whh = sd.get(p + ".weight_hh_l0", 2048 * 512);
            vec wi((size_t)2048 * 512), wh((size_t)2048 * 512);
            for (int nn = 0; nn < 2048; ++nn)
                for (int kk = 0; kk < 512; ++kk) {
                    // layer-1 K order: kk = f*128 + cl -> reference feature cl*4 + f
                    const int col = layer == 1 ? (kk % 128) * 4 + kk / 128 : kk;
                    wi[(size_t)nn * 512 + kk] = wih[(size_t)rows[nn] * 512 + col];
                    wh[(size_t)nn * 512 + kk] = whh[(size_t)rows[nn] * 512 + kk];
                }
            Out a = at(nme + ".w_ih"), b = at(nme + ".w_hh"), f = at(nme + ".wf");
            stream(a, {wi}, 2048, 512, 128);
            for (int c = 0; c < 16; ++c) b.h(cp8(vec(wh.begin() + (size_t)c * 128 * 512, wh.begin() + (size_t)(c + 1) * 128 * 512), 128, 512));
            const vec bi = sd.get(p + ".bias_ih_l0", 2048), bh = sd.get(p + ".bias_hh_l0", 2048);
            vec bias(2048);
            for (int nn = 0; nn < 2048; ++nn) bias[nn] = bi[rows[nn]] + bh[rows[nn]];
            f.f(bias);
            ok = ok && check(a, nme + ".w_ih") && check(b, nme + ".w_hh") && check(f, nme + ".wf");
        }
    {
        Out f = at("ln.wf");
        for (int i = 1; i <= 2; ++i) {
            f.f(sd.get("glstm.ln" + std::to_string(i) + ".weight", 1024));
            f.f(sd.get("glstm.ln" + std::to_string(i) + ".bias", 1024));
        }
        ok = ok && check(f, "ln.wf");
    }
    for (int br = 1; br <= 2; ++br) {
        for (int i = 5; i >= 2; --i) {
            const int cin = GCRN_DEC[i].cin, cout = GCRN_DEC[i].cout, ntile = 2 * cout;
            const std::string nme = "dec" + std::to_string(br) + "_" + std::to_string(i);
            const std::string p = "conv" + std::to_string(i) + "_t_" + std::to_string(br);
            const vec wv = sd.get(p + ".conv1.weight", (long)cin * cout * 3), wg = sd.get(p + ".conv2.weight", (long)cin * cout * 3);
            auto tp = [&](int df) { return glu_rows(tap3(wv, cin, cout, df, true), tap3(wg, cin, cout, df, true), cout, cin, ntile); };
            Out we = at(nme + ".w_even"), wo = at(nme + ".w_odd"), f = at(nme + ".wf");
            stream(we, {tp(0), tp(2)}, 2 * cout, cin, ntile);
            stream(wo, {tp(1)}, 2 * cout, cin, ntile);
            vec s, sh;
            bn_affine(sd, "bn" + std::to_string(i) + "_t_" + std::to_string(br), cout, s, sh);
            f.f(glu_ep(sd.get(p + ".conv1.bias", cout), sd.get(p + ".conv2.bias", cout), s, sh, ntile));
            ok = ok && check(we, nme + ".w_even") && check(wo, nme + ".w_odd") && check(f, nme + ".wf");
        }
        {   // conv1_t (32 -> 1 GLU, k3 s2) + bn1_t + ELU + fc, scaled by 1/11 (trainer :942)
            const std::string nme = "out" + std::to_string(br), p = "conv1_t_" + std::to_string(br);
            Out f = at(nme + ".wf");
            f.f(sd.get(p + ".conv1.weight", 32 * 3));
            f.f(sd.get(p + ".conv2.weight", 32 * 3));
            vec s, sh;
            bn_affine(sd, "bn1_t_" + std::to_string(br), 1, s, sh);
            f.f({sd.get(p + ".conv1.bias", 1)[0], sd.get(p + ".conv2.bias", 1)[0], s[0], sh[0]});
            const vec fw = sd.get("fc" + std::to_string(br) + ".weight", 161 * 161);
            vec fcw((size_t)161 * 161);
            for (int o = 0; o < 161; ++o)
                for (int i = 0; i < 161; ++i) fcw[(size_t)i * 161 + o] = fw[(size_t)o * 161 + i] / 11.0;
            f.f(fcw);
            const vec fb = sd.get("fc" + std::to_string(br) + ".bias", 161);
            vec fcb(164, 0.0);
            for (int i = 0; i < 161; ++i) fcb[i] = fb[i] / 11.0;
            f.f(fcb);
            ok = ok && check(f, nme + ".wf");
        }
    }
    if (!sd.missing.empty()) {
        std::snprintf(error_buffer(), 512, "pdse_pack_gcrn: state_dict entry '%s' is missing or has the wrong size", sd.missing.c_str());
        return PDSE_EINVAL;
    }
    if (!ok) return set_error("pdse_pack_gcrn: internal layout mismatch");
    if (g_op_overflow) return set_error("pdse_pack_gcrn: a folded weight exceeds the fp16 operand range (|w| >= 65520); build with PDSE_OPERANDS=bf16");
    return PDSE_OK;
}
