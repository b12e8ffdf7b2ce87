/* C-only host of libpdse.so: one DiffUNet1 evaluation (model/diff3.py:37-57) and one GCRN evaluation
 * (model/gcrn.py:136-166) driven through include/pdse.h alone -- no Python, no pack.py, no torch.
 *
 *   abi_host <in.bin> <out.bin> [--pack-only]
 *
 * in.bin  (written by tests/test_abi.py): int32 n_tensors, then per tensor {int32 name_len, name, int64 numel, fp32 data};
 *         tensors named "@x", "@x0", "@t", "@y" are inputs ([B][2][T][161], [B][2][T][161], [1], [B][2][T][161]),
 *         "@shape" holds (B, T) as two floats; every other tensor is a state_dict entry ("ddpm/<key>" or "gcrn/<key>").
 * out.bin: eps [B][2][T][161] then X_init [B][2][T][161] (fp32); with --pack-only (no GPU needed) the two packed blobs.
 */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pdse.h"

#define CHECK(expr)                                                        \
    do {                                                                   \
        int rc_ = (expr);                                                  \
        if (rc_ != 0) {                                                    \
            fprintf(stderr, "%s failed: %s\n", #expr, pdse_last_error()); \
            return 2;                                                      \
        }                                                                  \
    } while (0)
#define CUDA(expr)                                                                 \
    do {                                                                           \
        cudaError_t e_ = (expr);                                                   \
        if (e_ != cudaSuccess) {                                                   \
            fprintf(stderr, "%s failed: %s\n", #expr, cudaGetErrorString(e_));    \
            return 3;                                                              \
        }                                                                          \
    } while (0)

typedef struct {
    char* name;
    float* data;
    long numel;
} entry;

static entry* find(entry* e, int n, const char* name) {
    for (int i = 0; i < n; ++i)
        if (strcmp(e[i].name, name) == 0) return &e[i];
    return NULL;
}

int main(int argc, char** argv) {
    if (argc < 3) return 1;
    const int pack_only = argc > 3 && strcmp(argv[3], "--pack-only") == 0;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 1;
    int n = 0;
    if (fread(&n, 4, 1, f) != 1) return 1;
    entry* all = (entry*)calloc((size_t)n, sizeof(entry));
    pdse_tensor* ddpm = (pdse_tensor*)calloc((size_t)n, sizeof(pdse_tensor));
    pdse_tensor* gcrn = (pdse_tensor*)calloc((size_t)n, sizeof(pdse_tensor));
    int n_ddpm = 0, n_gcrn = 0;
    for (int i = 0; i < n; ++i) {
        int len = 0;
        long long numel = 0;
        if (fread(&len, 4, 1, f) != 1) return 1;
        all[i].name = (char*)calloc((size_t)len + 1, 1);
        if (fread(all[i].name, 1, (size_t)len, f) != (size_t)len || fread(&numel, 8, 1, f) != 1) return 1;
        all[i].numel = (long)numel;
        all[i].data = (float*)malloc((size_t)numel * 4 + 16);
        if (fread(all[i].data, 4, (size_t)numel, f) != (size_t)numel) return 1;
        if (strncmp(all[i].name, "ddpm/", 5) == 0) ddpm[n_ddpm++] = (pdse_tensor){all[i].name + 5, all[i].data, all[i].numel};
        if (strncmp(all[i].name, "gcrn/", 5) == 0) gcrn[n_gcrn++] = (pdse_tensor){all[i].name + 5, all[i].data, all[i].numel};
    }
    fclose(f);

    /* ---- pack both networks on the host ---- */
    const long ddpm_bytes = pdse_pack_layout(PDSE_NET_DIFFUNET1, NULL, 0, NULL);
    const long gcrn_bytes = pdse_pack_layout(PDSE_NET_GCRN, NULL, 0, NULL);
    if (ddpm_bytes <= 0 || gcrn_bytes <= 0) return 2;
    void* ddpm_blob = malloc((size_t)ddpm_bytes);
    void* gcrn_blob = malloc((size_t)gcrn_bytes);
    CHECK(pdse_pack_diffunet1(ddpm, n_ddpm, ddpm_blob));
    CHECK(pdse_pack_gcrn(gcrn, n_gcrn, gcrn_blob));
    FILE* out = fopen(argv[2], "wb");
    if (!out) return 1;
    if (pack_only) {
        fwrite(ddpm_blob, 1, (size_t)ddpm_bytes, out);
        fwrite(gcrn_blob, 1, (size_t)gcrn_bytes, out);
        fclose(out);
        printf("packed %ld + %ld bytes\n", ddpm_bytes, gcrn_bytes);
        return 0;
    }

    /* ---- device side ---- */
    CHECK(pdse_check_device());
    entry *ex = find(all, n, "@x"), *ex0 = find(all, n, "@x0"), *et = find(all, n, "@t"), *ey = find(all, n, "@y"), *es = find(all, n, "@shape");
    if (!ex || !ex0 || !et || !ey || !es) return 1;
    const int B = (int)es->data[0], T = (int)es->data[1];
    const size_t nel = (size_t)B * 2 * T * 161, cap = (nel + 3) / 4 * 4 * sizeof(float);
    cudaStream_t stream;
    CUDA(cudaStreamCreate(&stream));
    void *d_ddpm, *d_gcrn, *ws_d, *ws_g;
    float *x, *x0, *t, *y, *rows, *eps, *xinit;
    const long wsd = pdse_workspace_bytes(PDSE_NET_DIFFUNET1, B, T), wsg = pdse_workspace_bytes(PDSE_NET_GCRN, B, T);
    if (wsd <= 0 || wsg <= 0) return 2;
    CUDA(cudaMalloc(&d_ddpm, (size_t)ddpm_bytes));
    CUDA(cudaMalloc(&d_gcrn, (size_t)gcrn_bytes));
    CUDA(cudaMalloc(&ws_d, (size_t)wsd));
    CUDA(cudaMalloc(&ws_g, (size_t)wsg));
    CUDA(cudaMemset(ws_d, 0, (size_t)wsd));         /* workspaces are zeroed ONCE (guard rows, status block) */
    CUDA(cudaMemset(ws_g, 0, (size_t)wsg));
    CUDA(cudaMalloc((void**)&x, cap));
    CUDA(cudaMalloc((void**)&x0, cap));
    CUDA(cudaMalloc((void**)&y, cap));
    CUDA(cudaMalloc((void**)&eps, cap));
    CUDA(cudaMalloc((void**)&xinit, cap));
    CUDA(cudaMalloc((void**)&t, 4));
    CUDA(cudaMalloc((void**)&rows, (size_t)pdse_bias_row_floats() * 4));
    CUDA(cudaMemcpy(d_ddpm, ddpm_blob, (size_t)ddpm_bytes, cudaMemcpyHostToDevice));
    CUDA(cudaMemcpy(d_gcrn, gcrn_blob, (size_t)gcrn_bytes, cudaMemcpyHostToDevice));
    CUDA(cudaMemcpy(x, ex->data, nel * 4, cudaMemcpyHostToDevice));
    CUDA(cudaMemcpy(x0, ex0->data, nel * 4, cudaMemcpyHostToDevice));
    CUDA(cudaMemcpy(y, ey->data, nel * 4, cudaMemcpyHostToDevice));
    CUDA(cudaMemcpy(t, et->data, 4, cudaMemcpyHostToDevice));

    CHECK(pdse_diffunet1_time_bias(d_ddpm, t, 1, rows, stream));
    CHECK(pdse_diffunet1_fwd(d_ddpm, ws_d, x, x0, rows, 0, NULL, eps, B, T, stream));
    CHECK(pdse_gcrn_fwd(d_gcrn, ws_g, y, xinit, B, T, stream));
    int status[8];
    CUDA(cudaMemcpyAsync(status, ws_d, sizeof(status), cudaMemcpyDeviceToHost, stream));
    CUDA(cudaStreamSynchronize(stream));
    CHECK(pdse_status_check(status));              /* kernel-side errors of the persistent TCM kernel */

    float* h = (float*)malloc(nel * 4);
    CUDA(cudaMemcpy(h, eps, nel * 4, cudaMemcpyDeviceToHost));
    fwrite(h, 4, nel, out);
    CUDA(cudaMemcpy(h, xinit, nel * 4, cudaMemcpyDeviceToHost));
    fwrite(h, 4, nel, out);
    fclose(out);
    printf("ok B=%d T=%d\n", B, T);
    return 0;
}
