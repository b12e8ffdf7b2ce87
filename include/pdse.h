/* pdse.h -- C ABI of libpdse.so, the sm_100a (B200) implementation of the Prior-DiffuSE
 * inference hot path (ishine/Prior-DiffuSE: STFT -> prior -> DiffUNet1 reverse loop -> ISTFT).
 *
 * The reference has no FFI of its own (pure PyTorch): its seam is the
 * nn.Module call convention used by trainer/complex_ddpm_trainer.py.  Each entry point below
 * names the reference call site it replaces (file:line relative to the reference root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name says "host"; the caller owns all memory;
 *   - no allocation, no host synchronisation, no default-stream use: every call only enqueues
 *     work on `stream` (a cudaStream_t passed as void*), so all of them are legal under CUDA
 *     graph capture;
 *   - return 0 on success, negative on error; pdse_last_error() returns a thread-local message;
 *   - entry points may be called concurrently from several host threads on distinct streams / workspaces / devices:
 *     the only process-wide state (kernel attributes, SM counts) is kept per device in atomics;
 *   - a persistent kernel cannot return a status: it records failures in a caller-owned STICKY status block
 *     (see pdse_status_check), which the caller reads at its own synchronisation point;
 *   - "bf16" in the comments below names the 16-bit tensor-core OPERAND format of the build: IEEE fp16 by default, bf16 with
 *     -DPDSE_OP_BF16 (pdse_operand_format()); the DB-AIAT entry points (pdse_db_*, pdse_aia_*) always use fp16
 *   - "CP8 split" = bf16 activation layout [B][C/8][T*2Q][8], Q = (F+1)/2,
 *     position(t, f) = t*2Q + (f&1)*Q + (f>>1)   (see DESIGN.md "Data layout");
 *   - weight blobs (`wb` bf16, `wf` fp32) are produced by prior_diffuse_b200/pack.py.
 */
#ifndef PDSE_H
#define PDSE_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ---------------------------------------------------------------------------- */
const char* pdse_last_error(void);
int pdse_abi_version(void);
/* 16-bit tensor-core operand format of the packed weight blobs and activation planes of the GCRN / DiffUNet1 / DiffWave
 * kernels: 1 = IEEE fp16 (default; conversions saturate, the packers refuse |w| >= 65520), 0 = bf16 (-DPDSE_OP_BF16).
 * The DB-AIAT kernels (pdse_db_*, pdse_aia_*) always take fp16. */
int pdse_operand_format(void);
int pdse_check_device(void);        /* 0 iff the current device is sm_100 */
int pdse_sm_count(void);
/* Kernel-side errors: status_host = HOST copy of a status block int32[8] = {code, detail0, detail1, count, timeout_us,
 * 0, 0, 0} that the caller allocated on the device, zeroed ONCE, and passed to the entry points that take `status` (they
 * never clear it).  Word 4 is the caller's knob: wall-clock bound of one in-kernel dependency wait in microseconds
 * (0 = 2 s default, < 0 = fail on the first unsatisfied poll); it is read on the device, so captured graphs follow it.
 * Returns 0 when clean, else negative with the text in pdse_last_error().  code 1 = a dependency wait of
 * pdse_tcm_flow timed out (detail0 = launch, detail1 = tile): the output of that call is invalid. */
int pdse_status_check(const int* status_host);

/* ---- a1/a2/a9: STFT + sqrt-compression, decompression + ISTFT ---------------------------- */
/* window + twiddle tables (Hann[320], then the 320th roots of unity as (re, im) pairs): pdse_signal_table_floats()
 * floats, built on the HOST in float64 */
int pdse_signal_table_floats(void);
int pdse_signal_tables(float* host_out);
/* trainer/complex_ddpm_trainer.py:922-923  c = sqrt(mean(x^2)) per utterance; wav [B][L] */
int pdse_rms_f32(const float* wav, int B, int L, float* rms, void* stream);
/* :926-930 torch.stft(n_fft=320, hop=160, hann, center, onesided) + :931-937 sqrt compression.
 * wav [B][L] (divided by rms[b] when rms != NULL) -> out [B][2][T][161], T = 1 + L/160 */
int pdse_stft_compress_f32(const float* wav, const float* rms, const float* tables, float* out,
                           int B, int L, int compress, void* stream);
/* :1004-1008 decompression + :1009-1015 torch.istft(length=L) (* rms[b] when rms != NULL) */
int pdse_decompress_istft_f32(const float* spec, const float* rms, const float* tables, float* wav,
                              int B, int T, int L, int decompress, void* stream);

/* the same with the reference writer's float -> PCM_16 conversion (:1018 sf.write; see pdse_f32_to_pcm16) fused into
 * the store: pcm (optional) int16 [B][L]; lengths optional (ragged batch, below) */
int pdse_decompress_istft_pcm16_f32(const float* spec, const float* rms, const float* tables, const int* lengths,
                                    float* wav, short* pcm, int B, int T, int L, int decompress, int pcm_clip,
                                    void* stream);

/* Ragged batches (utils/dataset.py:45-60 zero-pads a batch to its longest utterance and carries
 * wav_len / frame_num beside it): the same three ops with `lengths` = int32[B] true sample counts (NULL = all L).
 * RMS is taken over the utterance's own samples, the STFT reflects at its own end and writes zeros for frames
 * past 1 + len/160, the ISTFT uses only those frames and writes zeros past len.  The networks' convolutions are causal
 * in time except the TCM's dilated ones, which take `lengths` themselves (pdse_tcm_flow), so each utterance's valid
 * frames equal the result of running it alone (tests/test_gpu_parity.py). */
int pdse_rms_ragged_f32(const float* wav, const int* lengths, int B, int L, float* rms, void* stream);
int pdse_stft_compress_ragged_f32(const float* wav, const float* rms, const float* tables,
                                  const int* lengths, float* out, int B, int L, int compress, void* stream);
int pdse_decompress_istft_ragged_f32(const float* spec, const float* rms, const float* tables,
                                     const int* lengths, float* wav, int B, int T, int L, int decompress,
                                     void* stream);

/* ---- a8: reverse-loop element-wise steps ------------------------------------------------- */
/* :951-952 per-(b,ch) max |X0| ; x [rows][n] */
int pdse_absmax_f32(const float* x, int rows, int n, float* out, void* stream);
/* same over the valid frames only: row r belongs to utterance r/2, (1 + lengths[r/2]/160) * 161 leading elements */
int pdse_absmax_ragged_f32(const float* x, const int* lengths, int rows, int n, float* out, void* stream);
/* :950-956 x_T = N(0,I) (generate=1, Philox4x32-10(seed, offset)) or the caller's x (generate=0),
 * times sqrt(0.5 + 0.5|X0|/max) when x0 != NULL.  Buffers hold n rounded up to 4 floats. */
int pdse_init_state_f32(float* x, const float* x0, const float* amax, long n, int plane, int generate,
                        unsigned long long seed, unsigned long long offset, void* stream);
/* the `deltamu` branch (:947-948): x_T = (z + add) [* sqrt(mask)] */
int pdse_init_state_add_f32(float* x, const float* x0, const float* amax, const float* add, long n, int plane,
                            int generate, unsigned long long seed, unsigned long long offset, void* stream);
/* :977-992 x = c1*(x - c2*eps) + sigma*z[*sqrt(mask)] ; :993-997 finalize = 1: out = (x + x0)*scale (pirorgrad),
 * finalize = 2: out = x*scale (deltamu / noisy-feature-conditioned branches) */
int pdse_ddpm_update_f32(float* x, const float* eps, const float* x0, const float* amax, float* out,
                         long n, int plane, float c1, float c2, float sigma, int use_mask, int finalize,
                         float scale, unsigned long long seed, unsigned long long offset, void* stream);
int pdse_scale_f32(float* x, long n, float s, void* stream);
/* Validation mode of the noise draw (:950, :987 torch.randn_like on the CUDA generator): out[n] = the values ATen's
 * normal kernel writes for a generator at (seed, philox_offset) -- Philox4x32-10, subsequence = thread index of a
 * (grid_x x 256) launch, curand_normal4's Box-Muller.  pdse_randn_aten_policy returns ATen's grid for n elements on the
 * current device and by how much the generator's offset advances (the next draw starts there). */
int pdse_randn_aten_policy(long n, int* grid_x, unsigned long long* offset_increment);
int pdse_randn_aten_f32(float* out, long n, unsigned long long seed, unsigned long long philox_offset, int grid_x,
                        void* stream);
/* :1018 sf.write(path, wav, 16000) -- the writer's float -> PCM_16 conversion (libsndfile src/pcm.c):
 * clip = 0: (short) lrintf(x * 32767) (libsndfile's default, wraps past full scale); clip = 1: x * 32768 saturated */
int pdse_f32_to_pcm16(const float* wav, short* out, long n, int clip, void* stream);

/* ---- 8f-2: evaluation scalar on the device ----------------------------------------------- */
/* utils/metrics.py:36-55 SNRseg(clean, processed, fs = 16000): 30 ms Hann frames, 75 % overlap, per-frame SNR
 * clamped to [-10, 35] dB, last frame dropped, mean -> out[B].  lengths (optional int32[B]) as above. */
int pdse_ssnr_f32(const float* clean, const float* processed, const int* lengths, int B, int L, float* out,
                  void* stream);

/* ---- a6: DiffUNet1 (model/diff3.py:37-57) ------------------------------------------------ */
int pdse_bias_row_floats(void);
/* diff3.py:69-87 TimeEmbedding + every block's time projection (en.tp*, de*.tp) composed with the
 * block's 1x1 conv.  t [n] float (fractional or integral) -> out [n][pdse_bias_row_floats()] */
int pdse_time_embed(const float* t, int n, const float* table, const float* p1w, const float* p1b,
                    const float* p2w, const float* p2b, const float* rows, const float* rbias,
                    float* out, void* stream);
/* diff3.py:38 Preprocess + :146-149 pad/time-bias/en.conv1/en1.  x, x0 [B][2][T][161] fp32 ->
 * out CP8 split F=79.  bias: time-bias table, row b*bias_stride (bias_stride 0 = one shared row) */
int pdse_enc1_fwd(const float* x, const float* x0, void* out, const void* wb, const float* wf,
                  const float* bias, int bias_stride, int B, int T, void* stream);
/* diff3.py:150-165 encoder block i=2..5 (pad, time bias, BiConvGLU (2,3)/(1,2), BN, PReLU) */
int pdse_enc_fwd(const void* xin, void* out, const void* wb, const float* wf, const float* bias,
                 int bias_stride, int bias_off, int B, int T, int Fin, int nt, void* stream);
/* diff3.py:249-277 TCM residual stack, launch k = 0..18 (see csrc/denoiser.cu).
 * lengths (optional int32[B], sample counts of a zero-padded ragged batch): the dilated convs are symmetric in time
 * (diff3.py:224-243), so frames >= 1 + lengths[b]/160 are treated as the convs' own zero padding -- every utterance's
 * valid frames then equal the result of running it alone.  NULL = every utterance has T frames. */
int pdse_tcm_fwd(const void* e5, const void* am_in, const void* ak_in, void* am_out, void* ak_out,
                 float* x, void* dec_in, const void* wA, const float* fA, const void* wB,
                 const float* fB, const int* lengths, int B, int T, int dilation, void* stream);
/* the same 19 launches as ONE persistent dataflow kernel (per-tile dependency flags instead of launch boundaries).
 * wtab: device table [18][2] of {bf16 blob, fp32 blob} pointers; flags: int32[32 + 19*B*ceil(T/128)] scratch;
 * dilations_host: 18 ints on the HOST; status: sticky status block (pdse_status_check), required.
 * A dependency wait polls tightly, then sleeps between polls, bounded by wall-clock time (status[4], 2 s by default);
 * when it expires the failure is recorded in `status`, every other wait gives up at once and the kernel drains. */
int pdse_tcm_flow(const void* e5, void* am0, void* ak0, void* am1, void* ak1, float* x, void* dec_in,
                  const void* wtab, int* flags, const int* dilations_host, const int* lengths, int* status,
                  int B, int T, void* stream);
/* diff3.py:206-212 decoder block de{i} of BOTH branches (BiConvTransGLU, Chomp_T, BN, PReLU);
 * last=1 (de1, kw=5): writes eps [B][2][T][161] fp32 (channel 0 = de_real, 1 = de_imag).
 * hws = NULL: one fused launch (nt time rows per tile, nt * (Fin + (kw-1)/2) <= 384).
 * hws != NULL: split path (two launches; nt ignored): the 1x1 conv output goes through the caller-owned bf16 workspace
 *   hws[B][2][4][(T+1)*(Fin+G)+G][8] (G = (kw-1)/2), which must be ZERO before its first use with a given shape
 *   (guard slots are never written) and must not be shared between blocks of different shape. */
int pdse_dec_fwd(const void* xa_re, const void* xa_im, const void* skip, void* out_re, void* out_im,
                 float* eps, const void* wb_re, const void* wb_im, const float* wf_re,
                 const float* wf_im, const float* bias, int bias_stride, int bias_off_re,
                 int bias_off_im, int B, int T, int Fin, int kw, int nt, int last, void* hws, void* stream);

/* ---- a3: GCRN prior (model/gcrn.py:136-166) ---------------------------------------------- */
/* layouts: SO = [B][C/8][2][T*Q][8] (strided-conv input), UG = [B][C/8][T*(F+1)+1][8] (transposed-conv
 * input, zero guard row before every frame); see csrc/gcrn.cu */
/* gcrn.py:137 conv1+bn1+ELU: y [B][2][T][161] fp32 -> SO(F=80) for conv2, UG(F=80, ELU twice) decoder skip */
int pdse_gcrn_conv1_fwd(const float* y, void* out_so, void* out_ug, const void* wb, const float* ep,
                        int B, int T, void* stream);
/* gcrn.py:138-141 conv{i}+bn{i}+ELU, i=2..5 (any output may be NULL; xl0/xl1 = LSTM layer-1 operands) */
int pdse_gcrn_enc_fwd(const void* xin, void* out_so, void* out_ug, void* xl0, void* xl1, const void* wb,
                      const float* ep, int B, int T, int Cin, int Cout, int Fin, int elu2, void* stream);
/* gcrn.py:150-153 / :156-159 conv{i}_t+bn+ELU on cat(prev, skip), i=5..2 */
int pdse_gcrn_dec_fwd(const void* prev, const void* skip, void* out_ug, const void* w_even,
                      const void* w_odd, const float* ep, int B, int T, int C1, int C2, int Cout, int Fin,
                      int Fout, void* stream);
/* gcrn.py:12-15 nn.LSTM(512,512): input projection (W_ih x + b_ih + b_hh) and the recurrence of one
 * layer for both groups.  x [64][T*B][8] bf16 (row = t*B+b); pre [T][2048][Bp] fp32; h [T*B][512] fp32 */
int pdse_lstm_inproj(const void* x, const void* w_ih, const float* bias, float* pre, int B, int Bp, int T,
                     void* stream);
/* debug hook: 6 int64 cycle counters (wait, h load, MMA, gates, cell, tail) of CTA (0,0); NULL disables */
int pdse_debug_lstm_prof(void* dev_buf);
/* debug hook: 12 int64 cycle counters of CTA (0,0) of the next decoder launches (producer 0..5, tiles 6, consumer 8..11); NULL disables */
int pdse_debug_dec_prof(void* dev_buf);
/* debug hook: 12 int64 cycle counters of CTA 0 of the next persistent TCM launches (tile phases 0..8, dependency wait 9, hand-over 10, tasks 11) */
int pdse_debug_tcm_prof(void* dev_buf);
/* debug hook: number of co-resident 16-CTA clusters of the DSMEM recurrence kernel (bp = 16 | 32 sequences per cluster) */
int pdse_debug_lstm_clusters(int bp);
int pdse_lstm_rec(const void* whh0, const void* whh1, const float* pre0, const float* pre1, float* h0,
                  float* h1, void* hbuf, unsigned int* sync, int B, int Bp, int T, void* stream);
/* gcrn.py:29-31 (mode 1: stack/flatten interleave + ln1 -> layer-2 operands) and :33-38 (mode 2: cat +
 * ln2 -> UG 256-channel F=4 decoder input) */
int pdse_gcrn_ln(const float* h0, const float* h1, const float* w, const float* b, void* xl0, void* xl1,
                 void* ug, int B, int T, int mode, void* stream);
/* gcrn.py:154/160 conv1_t+bn1_t+ELU, :162-163 fc1/fc2, trainer :942 (/11): X_init [B][2][T][161] fp32 */
int pdse_gcrn_out_fwd(const void* d2_1, const void* d2_2, const void* e1_ug, const float* wf1,
                      const float* wf2, float* xinit, int B, int T, void* stream);

/* ---- a4: aia_complex_trans_ri prior (model/dbaiat.py:450-478) ---------------------------- */
/* layouts (csrc/dbaiat.cu): dense skip tensors are CP8 planes [B][planes][(T+G)*(F+1)+1][8] with row
 * (t+G)*(F+1)+1+f, G = pdse_db_guard_frames() zero frames in front (causal padding) and one zero guard
 * row per frame (frequency padding); conv outputs before LayerNorm are fp32 [B][T*(F+1)][N];
 * transformer tensors are fp32 [B][T*80][32], sequence-major (row: n = b*T+t, col: n = b*80+w) */
int pdse_db_guard_frames(void);
/* DenseBlock conv{i} (dbaiat.py:614-621, dil = 2^(i-1), taps (t-dil | t) x (f-1..f+1)), or with dil = 0 the
 * (1 x 3) convs enc_conv1 (:491, center = 0: taps f..f+2, stride 2 taken by the LayerNorm pass) and
 * SPConvTranspose2d.conv (:592, center = 1, N = 128).  chunk c (64 input channels) = planes
 * chunk_plane[c].. of source chunk_src[c]; w = [chunk][tap][8][N][8] bf16 */
int pdse_db_conv_fwd(const void* src0, const void* src1, int ppb0, int ppb1, const int* chunk_src,
                     const int* chunk_plane, int nchunks, int B, int T, int pitch, int dil, int center,
                     const void* w, const float* bias, int N, void* pre, void* stream);
/* LayerNorm over frequency + PReLU and what follows it.  mode 0: DenseBlock norm{i}/prelu{i} -> CP8 planes
 * plane0.. of out_planes; 1: enc_norm1/enc_prelu1 (:499-500) + dual_trans.input (:115-118) -> state fp32;
 * 2: sub-pixel shuffle + pad1 + dec_norm1/dec_prelu1/out_conv (:545-547) -> out_f32[B][2][T][161] channel ch;
 * 3: inp_conv/inp_norm/inp_prelu (:498) from x [B][2][T][161] -> CP8 planes */
int pdse_db_ln_fwd(int mode, const void* pre, const float* x, const float* gamma, const float* beta,
                   const float* slope, const float* cw, void* out_planes, int ppb, int plane0,
                   float* out_f32, int ch, int B, int T, int pitch, int F, void* stream);
/* TransformerEncoderLayer.forward first half (:74-79): y1 = norm1(src + self_attn(norm3(src))) per sequence;
 * Y1 fp32 [nseq][L][32], XG = the GRU's bf16 operand [ceil(nseq/128)][L][4][128][8] */
int pdse_aia_attn_fwd(const float* S, const float* w, float* Y1, void* XG, int B, int T, int is_row,
                      void* stream);
/* :80-83 bidirectional GRU(32 -> 64) and linear2 split per direction: P [2][nseq][L][32] fp32 */
int pdse_aia_gru_fwd(const void* XG, const void* w, const float* bias, float* P, int L, int nseq,
                     void* stream);
/* :84-87 Z = norm2(y1 + P_fwd + P_bwd + b2); stats[b] += (sum Z, sum Z^2) for the GroupNorm (:145,150) */
int pdse_aia_post_fwd(const float* Y1, const float* P0, const float* P1, const float* w, float* Z,
                      double* stats, int B, int npos, void* stream);
/* AIA_Transformer.forward :145-152: S += k1 GN(Zr) + k2 GN(Zc); O = output(S) bf16 [B][T*80][64];
 * pool[b][c] += sum over (t, f) of O */
int pdse_aia_combine_fwd(float* S, const float* Zr, const float* Zc, const double* st_r, const double* st_c,
                         const float* w, void* O, double* pool, int B, int T, void* stream);
/* AHAM.forward (:268-288): softmax over the four layer outputs' pooled 1x1 conv -> decoder input planes */
int pdse_aia_aham_fwd(const void* O0, const void* O1, const void* O2, const void* O3, const double* pool,
                      const float* w, void* xbuf, int B, int T, void* stream);

/* ---- 8f-4: diff2.DiffWave (model/diff2.py:12-158), the time-domain gated-tanh residual stack --------------------
 * layouts (csrc/diffwave.cu): residual stream / skip sum fp32 planes [B][16][L][4]; bf16 operands (x + e_i, cond)
 * [B][8][L + 2G][8] with G = pdse_dw_guard_rows() ZERO rows in front of and behind every utterance (the convolutions'
 * padding; buffers are zeroed once by the caller, guards are never written).  64 residual channels. */
int pdse_dw_guard_rows(void);
/* DiffusionEmbedding :71-95 (table lookup / lerp, two SiLU linears) + every layer's diffusion_projection :132:
 * t [B] -> dtab [B][n_rows]; rows [n_rows][512] / rbias [n_rows] = the layers' projections stacked (n_rows = 64 layers) */
int pdse_dw_embed(const float* t, int B, const float* table, const float* p1w, const float* p1b, const float* p2w,
                  const float* p2b, const float* rows, const float* rbias, int n_rows, float* dtab, void* stream);
/* :29-31, :38-40  x = relu(w audio + b), cond = relu(w audio_init + b), y0 = bf16(x + e_0); win = w[64] | b[64] */
int pdse_dw_pre_fwd(const float* audio, const float* init, const float* win, const float* dtab, int dstride, float* x,
                    void* y, void* cond, int B, int L, void* stream);
/* ResidualBlock.forward :131-158 of layer i as one kernel: z = dilated_conv(y_in) + conditioner_projection(cond) as one
 * K = 384 implicit GEMM, gate, output projection, x = (x + residual) / sqrt 2, skip += s, y_out = bf16(x + e_{i+1}).
 * wb: W_cat [48][128][8] | W_o [8][128][8] | bias block (conv pair) | bias block (output); dnext = dtab + 64 (i + 1) */
int pdse_dw_layer_fwd(const void* y_in, void* y_out, const void* cond, float* x, float* skip, const void* wb,
                      const float* dnext, int dstride, int B, int L, int dilation, int first, int last, void* stream);
/* :50-55  out[B][L] = w_out . relu(W_skip (skip * scale) + b_skip) + b_out; wb: W_skip [8][64][8] | bias block;
 * wout = w[64] | b */
int pdse_dw_post_fwd(const float* skip, const void* wb, const float* wout, float scale, float* out, int B, int L,
                     void* stream);

/* ---- b: weight packing on the HOST, workspaces and whole-network entry points ----------------------------------
 * Everything a non-Python host needs: pack the reference's state_dict tensors into ONE blob per network,
 * upload it with one copy, size and zero a workspace, call the forward.  The per-op entry points above remain the
 * fine-grained interface (sections of the blob are their `wb` / `wf` arguments, see pdse_pack_layout). */
typedef struct {
    const char* name;       /* state_dict key, e.g. "en.conv2.l.weight" (model/diff3.py, model/gcrn.py key names) */
    const float* data;      /* HOST pointer, fp32, row-major as torch stores it */
    long numel;
} pdse_tensor;
typedef struct {
    char name[40];          /* "<block>.wb" (bf16 operand blob) / "<block>.wf" (fp32 blob) / "time.*" / "lstm*.w_ih" ... */
    int dtype;              /* 0 = 16-bit operand format (pdse_operand_format()), 1 = fp32 */
    long offset;            /* byte offset in the blob (128-byte aligned) */
    long elems;
} pdse_blob_entry;
enum { PDSE_NET_DIFFUNET1 = 1, PDSE_NET_GCRN = 2 };
/* directory of a packed network (fixed: the architectures are fixed).  Returns the blob size in bytes (negative on
 * error); fills up to `capacity` entries and *count with the number of sections. */
long pdse_pack_layout(int net, pdse_blob_entry* out, int capacity, int* count);
/* model/diff3.py DiffUNet1 / model/gcrn.py GCRN state_dict (every floating-point entry; BatchNorm running statistics
 * included, num_batches_tracked not needed) -> blob_host[pdse_pack_layout(net) bytes].  A Nocon / DiffUNet checkpoint
 * packs through pdse_pack_diffunet1 after the key mapping of pack.diffunet_as_diffunet1. */
int pdse_pack_diffunet1(const pdse_tensor* sd, int n, void* blob_host);
int pdse_pack_gcrn(const pdse_tensor* sd, int n, void* blob_host);
/* device workspace of one network evaluation at batch B, T frames: bytes (negative on error).  The workspace must be
 * ZEROED ONCE before its first use (guard rows / dummy slots are never written) and belongs to one stream at a time.
 * Its first 32 bytes are the sticky status block of pdse_status_check. */
long pdse_workspace_bytes(int net, int B, int T);
/* diff3.py:39 + every block's time projection: t [n] (device) -> rows [n][pdse_bias_row_floats()] (device) */
int pdse_diffunet1_time_bias(const void* blob_dev, const float* t, int n, float* rows, void* stream);
/* eps = DiffUNet1(x, x0, t) (model/diff3.py:37-57): x, x0, eps [B][2][T][161] fp32 on the device (eps capacity rounded
 * up to a multiple of 4 floats); rows from pdse_diffunet1_time_bias (row b * bias_stride; 0 = one shared row);
 * lengths optional (ragged batch).  Enqueues 17 kernels on `stream`; legal under graph capture after one eager call. */
int pdse_diffunet1_fwd(const void* blob_dev, void* workspace, const float* x, const float* x0, const float* rows,
                       int bias_stride, const int* lengths, float* eps, int B, int T, void* stream);
/* X_init = GCRN(y) / 11 (model/gcrn.py:136-166 + trainer :942): y, xinit [B][2][T][161] fp32 on the device; B <= 64
 * per call (the recurrence holds one chunk of 64 sequences) */
int pdse_gcrn_fwd(const void* blob_dev, void* workspace, const float* y, float* xinit, int B, int T, void* stream);

/* ---- test hook: one 128xNxK tcgen05 GEMM on CP8 operands with a row-shifted A window ----- */
int pdse_probe_gemm(const void* A, const void* B, float* D, int a_rows, int N, int K, int row_shift,
                    int swap_lbo_sbo, void* stream);

/* measurement hook (tests/gpu_probe_tmem.py): tcgen05.ld drain rate with / without a concurrent MMA stream, and the
 * bulk-copy rate of one lane.  mode bits: 1 MMA stream, 2 TMEM drain, 4 bulk copies.  out: int64[4 * ctas] */
int pdse_probe_tmem(long long* out, const void* src, int mode, int iters, int mma_n, int ld_cols, int copy_bytes,
                    int ctas, long cta_stride, int nblk, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PDSE_H */
