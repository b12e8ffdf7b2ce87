// Whole-network entry points of the C ABI: workspace sizing and the launch sequences of one DiffUNet1 evaluation
// (model/diff3.py:37-57) and one GCRN evaluation (model/gcrn.py:136-166) over a blob packed by pack.cpp.  These are the
// same sequences prior_diffuse_b200/denoiser.py and gcrn.py issue through the per-op entry points (tests compare them
// bit for bit); they exist so that a host in any language can drive the path without re-deriving workspace shapes.
#include "../../include/pdse.h"

#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace pdse {
namespace {

constexpr int NF = 161;
const int ENC_F[6] = {161, 79, 39, 19, 9, 4};                 // F after encoder block i
const int GCRN_CH[6] = {2, 16, 32, 64, 128, 256};
const int GCRN_F[6] = {161, 80, 39, 19, 9, 4};
struct GDec {
    int cin, cout, fin, fout;
};
const GDec GCRN_DEC[6] = {{0, 0, 0, 0}, {32, 1, 80, 161}, {64, 16, 39, 80}, {128, 32, 19, 39}, {256, 64, 9, 19}, {512, 128, 4, 9}};
const int TCM_DIL[18] = {1, 2, 4, 8, 16, 32, 1, 2, 4, 8, 16, 32, 1, 2, 4, 8, 16, 32};

struct Dir {
    std::vector<pdse_blob_entry> e;
    long bytes = 0;
    long off(const std::string& name) const {
        for (const auto& x : e)
            if (name == x.name) return x.offset;
        return -1;
    }
};
const Dir& dir_of(int net) {
    static Dir d[3];
    static std::once_flag once[3];
    std::call_once(once[net], [net] {
        int n = 0;
        pdse_pack_layout(net, nullptr, 0, &n);
        d[net].e.resize((size_t)n);
        d[net].bytes = pdse_pack_layout(net, d[net].e.data(), n, &n);
    });
    return d[net];
}

// bump allocator over the workspace (256-byte aligned sections); the same walk sizes and carves it
struct Carve {
    uint8_t* base;
    size_t off = 0;
    void* take(size_t bytes) {
        off = (off + 255) & ~(size_t)255;
        void* p = base ? base + off : nullptr;
        off += bytes;
        return p;
    }
};

struct DnWs {
    int* status;
    void** tcm_table;
    int* flags;
    void* e[6];
    float* x;
    void *am0, *ak0, *am1, *ak1, *dec_in;
    void* d[2][6];
    void* h1;
};
size_t carve_diffunet1(Carve& c, int B, int T, DnWs& w) {
    auto cp8 = [&](size_t npos) { return c.take((size_t)B * 8 * npos * 8 * 2); };
    w.status = (int*)c.take(32);
    w.tcm_table = (void**)c.take(36 * sizeof(void*));
    w.flags = (int*)c.take((size_t)(32 + 19 * B * ((T + 127) / 128)) * 4);
    for (int i = 1; i <= 5; ++i) w.e[i] = cp8((size_t)T * 2 * ((ENC_F[i] + 1) / 2));
    w.x = (float*)c.take((size_t)B * 32 * T * 8 * 4);
    w.am0 = cp8(T), w.ak0 = cp8(T), w.am1 = cp8(T), w.ak1 = cp8(T);
    w.dec_in = cp8((size_t)T * 4);
    for (int br = 0; br < 2; ++br)
        for (int i = 5; i >= 2; --i) {
            const int fo = 2 * ENC_F[i] + 1;
            w.d[br][i] = cp8((size_t)T * 2 * ((fo + 1) / 2));
        }
    w.h1 = c.take((size_t)B * 2 * 4 * ((size_t)(T + 1) * (ENC_F[1] + 2) + 2) * 8 * 2);   // split de1: G = 2
    return (c.off + 255) & ~(size_t)255;
}

struct GcWs {
    int* status;
    void *so[5], *ug[6], *lstm_ug, *xl1[2], *xl2[2], *hbuf, *d[3][6];
    float *pre[2], *h[2];
    unsigned int* sync;
    int Bp;
};
size_t carve_gcrn(Carve& c, int B, int T, GcWs& w) {
    w.status = (int*)c.take(32);
    w.Bp = B <= 32 ? 32 : 64;
    for (int i = 1; i <= 4; ++i) {
        const int ch = GCRN_CH[i], f = GCRN_F[i];
        w.so[i] = c.take((size_t)B * (ch / 8) * 2 * T * ((f + 1) / 2) * 16);
        w.ug[i] = c.take((size_t)B * (ch / 8) * ((size_t)T * (f + 1) + 1) * 16);
    }
    w.ug[5] = c.take((size_t)B * 32 * ((size_t)T * 5 + 1) * 16);
    w.lstm_ug = c.take((size_t)B * 32 * ((size_t)T * 5 + 1) * 16);
    for (int g = 0; g < 2; ++g) {
        w.xl1[g] = c.take((size_t)64 * T * B * 16);
        w.xl2[g] = c.take((size_t)64 * T * B * 16);
        w.pre[g] = (float*)c.take((size_t)T * 2048 * w.Bp * 4);
        w.h[g] = (float*)c.take((size_t)T * B * 512 * 4);
    }
    w.hbuf = c.take((size_t)2 * 2 * 64 * w.Bp * 16);
    w.sync = (unsigned int*)c.take(8);
    for (int br = 1; br <= 2; ++br)
        for (int i = 5; i >= 2; --i) w.d[br][i] = c.take((size_t)B * (GCRN_DEC[i].cout / 8) * ((size_t)T * (GCRN_DEC[i].fout + 1) + 1) * 16);
    return (c.off + 255) & ~(size_t)255;
}

struct PtrTable {
    long off[36];
};
__global__ void fill_ptr_table(void** table, const uint8_t* blob, PtrTable t) {
    if (threadIdx.x < 36) table[threadIdx.x] = (void*)(blob + t.off[threadIdx.x]);
}

inline int enc_nt(int Fin) { return std::max(1, 256 / ((Fin + 1) / 2)); }
inline int dec_nt(int Fin, int kw) { return std::max(1, 384 / (Fin + (kw - 1) / 2)); }
inline int bias_off_enc(int i) { return 2 + 32 * (i - 2); }
inline int bias_off_dec(int br, int i) { return 130 + 160 * br + 32 * (5 - i); }

}  // namespace
}  // namespace pdse

using namespace pdse;

extern "C" long pdse_workspace_bytes(int net, int B, int T) {
    if (B <= 0 || T <= 0) {
        set_error("pdse_workspace_bytes: empty input");
        return -1;
    }
    Carve c{nullptr};
    if (net == PDSE_NET_DIFFUNET1) {
        DnWs w;
        return (long)carve_diffunet1(c, B, T, w);
    }
    if (net == PDSE_NET_GCRN) {
        if (B > 64) {
            set_error("pdse_workspace_bytes: GCRN takes at most 64 utterances per call");
            return -1;
        }
        GcWs w;
        return (long)carve_gcrn(c, B, T, w);
    }
    set_error("pdse_workspace_bytes: unknown network");
    return -1;
}

extern "C" int pdse_diffunet1_time_bias(const void* blob_dev, const float* t, int n, float* rows, void* stream) {
    const Dir& d = dir_of(PDSE_NET_DIFFUNET1);
    const uint8_t* b = (const uint8_t*)blob_dev;
    auto f = [&](const char* name) { return (const float*)(b + d.off(name)); };
    return pdse_time_embed(t, n, f("time.table"), f("time.p1w"), f("time.p1b"), f("time.p2w"), f("time.p2b"), f("time.rows"),
                           f("time.bias"), rows, stream);
}

extern "C" int pdse_diffunet1_fwd(const void* blob_dev, void* workspace, const float* x, const float* x0, const float* rows,
                                  int bias_stride, const int* lengths, float* eps, int B, int T, void* stream) {
    if (!blob_dev || !workspace || !x || !x0 || !rows || !eps || B <= 0 || T <= 0) return set_error("pdse_diffunet1_fwd: bad arguments");
    const Dir& d = dir_of(PDSE_NET_DIFFUNET1);
    const uint8_t* blob = (const uint8_t*)blob_dev;
    Carve c{(uint8_t*)workspace};
    DnWs w;
    carve_diffunet1(c, B, T, w);
    auto wb = [&](const std::string& n) { return (const void*)(blob + d.off(n + ".wb")); };
    auto wf = [&](const std::string& n) { return (const float*)(blob + d.off(n + ".wf")); };
    if (int e = pdse_enc1_fwd(x, x0, w.e[1], wb("enc1"), wf("enc1"), rows, bias_stride, B, T, stream)) return e;
    for (int i = 2; i <= 5; ++i) {
        const std::string n = "enc" + std::to_string(i);
        const int Fin = ENC_F[i - 1];
        if (int e = pdse_enc_fwd(w.e[i - 1], w.e[i], wb(n), wf(n), rows, bias_stride, bias_off_enc(i), B, T, Fin, enc_nt(Fin), stream)) return e;
    }
    PtrTable pt;
    for (int k = 0; k < 18; ++k) {
        pt.off[2 * k] = d.off("tcm" + std::to_string(k) + ".wb");
        pt.off[2 * k + 1] = d.off("tcm" + std::to_string(k) + ".wf");
    }
    fill_ptr_table<<<1, 64, 0, (cudaStream_t)stream>>>(w.tcm_table, blob, pt);
    if (int e = check_launch("pdse_diffunet1_fwd (table)")) return e;
    if (int e = pdse_tcm_flow(w.e[5], w.am0, w.ak0, w.am1, w.ak1, w.x, w.dec_in, w.tcm_table, w.flags, TCM_DIL, lengths, w.status, B, T,
                              stream))
        return e;
    for (int i = 5; i >= 1; --i) {
        const int Fin = ENC_F[i], kw = i == 1 ? 5 : 3;
        const std::string n0 = "dec0_" + std::to_string(i), n1 = "dec1_" + std::to_string(i);
        const void* xa0 = i == 5 ? w.dec_in : w.d[0][i + 1];
        const void* xa1 = i == 5 ? w.dec_in : w.d[1][i + 1];
        if (int e = pdse_dec_fwd(xa0, xa1, w.e[i], i == 1 ? nullptr : w.d[0][i], i == 1 ? nullptr : w.d[1][i], i == 1 ? eps : nullptr,
                                 wb(n0), wb(n1), wf(n0), wf(n1), rows, bias_stride, bias_off_dec(0, i), bias_off_dec(1, i), B, T, Fin, kw,
                                 dec_nt(Fin, kw), i == 1 ? 1 : 0, i == 1 ? w.h1 : nullptr, stream))
            return e;
    }
    return PDSE_OK;
}

extern "C" int pdse_gcrn_fwd(const void* blob_dev, void* workspace, const float* y, float* xinit, int B, int T, void* stream) {
    if (!blob_dev || !workspace || !y || !xinit || B <= 0 || T <= 0) return set_error("pdse_gcrn_fwd: bad arguments");
    if (B > 64) return set_error("pdse_gcrn_fwd: at most 64 utterances per call");
    const Dir& d = dir_of(PDSE_NET_GCRN);
    const uint8_t* blob = (const uint8_t*)blob_dev;
    Carve c{(uint8_t*)workspace};
    GcWs w;
    carve_gcrn(c, B, T, w);
    auto sec = [&](const std::string& n) { return (const void*)(blob + d.off(n)); };
    auto secf = [&](const std::string& n) { return (const float*)(blob + d.off(n)); };
    if (int e = pdse_gcrn_conv1_fwd(y, w.so[1], w.ug[1], sec("conv1.wb"), secf("conv1.wf"), B, T, stream)) return e;
    for (int i = 2; i <= 5; ++i) {
        const bool last = i == 5;
        const std::string n = "conv" + std::to_string(i);
        if (int e = pdse_gcrn_enc_fwd(w.so[i - 1], last ? nullptr : w.so[i], w.ug[i], last ? w.xl1[0] : nullptr, last ? w.xl1[1] : nullptr,
                                      sec(n + ".wb"), secf(n + ".wf"), B, T, GCRN_CH[i - 1], GCRN_CH[i], GCRN_F[i - 1], last ? 0 : 1, stream))
            return e;
    }
    const float* ln = secf("ln.wf");
    for (int layer = 1; layer <= 2; ++layer) {
        const std::string l0 = "lstm" + std::to_string(layer) + "_0", l1 = "lstm" + std::to_string(layer) + "_1";
        for (int g = 0; g < 2; ++g) {
            const std::string n = g ? l1 : l0;
            if (int e = pdse_lstm_inproj(layer == 1 ? w.xl1[g] : w.xl2[g], sec(n + ".w_ih"), secf(n + ".wf"), w.pre[g], B, w.Bp, T, stream)) return e;
        }
        if (int e = pdse_lstm_rec(sec(l0 + ".w_hh"), sec(l1 + ".w_hh"), w.pre[0], w.pre[1], w.h[0], w.h[1], w.hbuf, w.sync, B, w.Bp, T, stream))
            return e;
        if (int e = pdse_gcrn_ln(w.h[0], w.h[1], ln + (layer == 1 ? 0 : 2048), ln + (layer == 1 ? 1024 : 3072), w.xl2[0], w.xl2[1], w.lstm_ug,
                                 B, T, layer, stream))
            return e;
    }
    for (int br = 1; br <= 2; ++br) {
        const void* prev = w.lstm_ug;
        for (int i = 5; i >= 2; --i) {
            const std::string n = "dec" + std::to_string(br) + "_" + std::to_string(i);
            if (int e = pdse_gcrn_dec_fwd(prev, w.ug[i], w.d[br][i], sec(n + ".w_even"), sec(n + ".w_odd"), secf(n + ".wf"), B, T,
                                          GCRN_DEC[i].cin / 2, GCRN_DEC[i].cin / 2, GCRN_DEC[i].cout, GCRN_DEC[i].fin, GCRN_DEC[i].fout, stream))
                return e;
            prev = w.d[br][i];
        }
    }
    return pdse_gcrn_out_fwd(w.d[1][2], w.d[2][2], w.ug[1], secf("out1.wf"), secf("out2.wf"), xinit, B, T, stream);
}
