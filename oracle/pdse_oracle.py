"""CPU oracle for the Prior-DiffuSE inference path  --  TEST INFRASTRUCTURE ONLY.

This file restates, in plain functional PyTorch (CPU, fp32), the algorithm of the
reference's generate/eval path.  It exists so the CUDA path can be checked on a
GPU box where ``/root/reference`` does not exist.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import it; the product path (``prior_diffuse_b200``) never does.

Parity pinning: the reference ships no tests or golden vectors.
This restatement is pinned against the reference's own ``nn.Module``s, imported
from ``/root/reference`` by ``tests/golden/make_golden.py`` with shared weights
and inputs; the outputs of that run are committed under ``tests/golden/`` and
re-checked by ``tests/test_oracle.py`` on every run (and the live comparison is
re-run whenever ``/root/reference`` is present).

Every function cites the reference file:line it follows (paths relative to the
reference root).  Weights are passed as a ``state_dict``-style mapping with the
reference's own key names, so a reference checkpoint drops in unchanged.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

N_FFT = 320      # conf/diff.yml: fft_num / win_size
HOP = 160        # conf/diff.yml: win_shift
N_FREQ = 161
FEAT_SCALE = 11.0  # trainer/complex_ddpm_trainer.py:30  (self.c = 11)
BN_EPS = 1e-5
LN_EPS = 1e-5

NOISE_SCHEDULE = np.linspace(1e-4, 0.05, 50).tolist()               # utils/params.py:40
INFERENCE_NOISE_SCHEDULE = [0.0001, 0.001, 0.01, 0.05, 0.2, 0.5]    # utils/params.py:41


# ---------------------------------------------------------------------------
# a1 / a2 / a9: STFT + sqrt compression and the inverse
# ---------------------------------------------------------------------------
def rms_normalize(wav: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """trainer/complex_ddpm_trainer.py:922-923 : c = sqrt(sum(x^2)/len); x /= c."""
    c = torch.sqrt(torch.sum(wav * wav, dim=-1, keepdim=True) / wav.shape[-1])
    return wav / c, c


def stft(wav: torch.Tensor) -> torch.Tensor:
    """trainer/complex_ddpm_trainer.py:926-930 (batched twin utils/dataset.py:61-67).

    Legacy real-view call -> [.., F, T, 2]; the trainer permutes to [2, T, F]
    (batched: [B, 2, T, F]).  center=True, reflect pad, periodic Hann, onesided.
    """
    z = torch.stft(wav, n_fft=N_FFT, hop_length=HOP, win_length=N_FFT,
                   window=torch.hann_window(N_FFT, device=wav.device), return_complex=True)
    z = torch.view_as_real(z)                       # [B, F, T, 2]
    return z.permute(0, 3, 2, 1).contiguous()       # [B, 2, T, F]


def compress_sqrt(feat: torch.Tensor) -> torch.Tensor:
    """trainer/complex_ddpm_trainer.py:931-937 : mag**0.5 with the phase kept."""
    phase = torch.atan2(feat[:, 1], feat[:, 0])
    mag = torch.norm(feat, dim=1) ** 0.5
    return torch.stack((mag * torch.cos(phase), mag * torch.sin(phase)), dim=1)


def decompress_sqrt(est: torch.Tensor) -> torch.Tensor:
    """trainer/complex_ddpm_trainer.py:1004-1008 : mag**2 with the phase kept."""
    mag = torch.norm(est, dim=1)
    phase = torch.atan2(est[:, 1], est[:, 0])
    mag = mag ** 2
    return torch.stack((mag * torch.cos(phase), mag * torch.sin(phase)), dim=1)


def istft(spec: torch.Tensor, length: int) -> torch.Tensor:
    """trainer/complex_ddpm_trainer.py:1009-1015 : [B,2,T,F] -> wav [B, length]."""
    z = torch.complex(spec[:, 0], spec[:, 1]).permute(0, 2, 1)   # [B, F, T]
    return torch.istft(z, n_fft=N_FFT, hop_length=HOP, win_length=N_FFT,
                       window=torch.hann_window(N_FFT, device=spec.device), length=length)


def stft_compress(wav: torch.Tensor) -> torch.Tensor:
    return compress_sqrt(stft(wav))


def decompress_istft(est: torch.Tensor, length: int) -> torch.Tensor:
    return istft(decompress_sqrt(est), length)


# ---------------------------------------------------------------------------
# a7: schedule
# ---------------------------------------------------------------------------
def inference_schedule(fast_sampling: bool = True,
                       noise_schedule: Sequence[float] = NOISE_SCHEDULE,
                       inference_noise_schedule: Sequence[float] = INFERENCE_NOISE_SCHEDULE):
    """trainer/complex_ddpm_trainer.py:105-156 (incl. the sigmas[0] wrap-around at :128)."""
    training = np.array(noise_schedule)
    inference = np.array(inference_noise_schedule) if fast_sampling else training
    talpha = 1 - training
    talpha_cum = np.cumprod(talpha)
    beta = inference
    alpha = 1 - beta
    alpha_cum = np.cumprod(alpha)
    sigmas = [0 for _ in alpha]
    for n in range(len(alpha) - 1, -1, -1):
        sigmas[n] = ((1.0 - alpha_cum[n - 1]) / (1.0 - alpha_cum[n]) * beta[n]) ** 0.5
    T = []
    for s in range(len(inference)):
        for t in range(len(training) - 1):
            if talpha_cum[t + 1] <= alpha_cum[s] <= talpha_cum[t]:
                tw = (talpha_cum[t] ** 0.5 - alpha_cum[s] ** 0.5) / (
                    talpha_cum[t] ** 0.5 - talpha_cum[t + 1] ** 0.5)
                T.append(t + tw)
                break
    return alpha, beta, alpha_cum, sigmas, np.array(T, dtype=np.float32)


# ---------------------------------------------------------------------------
# small helpers
# ---------------------------------------------------------------------------
def _bn(x: torch.Tensor, sd: SD, key: str) -> torch.Tensor:
    """eval-mode BatchNorm (both nets in .eval(), trainer :400-401)."""
    shape = [1, -1] + [1] * (x.dim() - 2)
    scale = sd[key + ".weight"] / torch.sqrt(sd[key + ".running_var"] + BN_EPS)
    shift = sd[key + ".bias"] - sd[key + ".running_mean"] * scale
    return x * scale.view(shape) + shift.view(shape)


def _prelu(x: torch.Tensor, a: torch.Tensor) -> torch.Tensor:
    return torch.where(x >= 0, x, a * x)


# ---------------------------------------------------------------------------
# a6: DiffUNet1
# ---------------------------------------------------------------------------
def time_embedding_table(max_steps: int = 50) -> torch.Tensor:
    """model/diff3.py:89-95."""
    steps = torch.arange(max_steps).unsqueeze(1)
    dims = torch.arange(64).unsqueeze(0)
    table = steps * 10.0 ** (dims * 4.0 / 63.0)
    return torch.cat([torch.sin(table), torch.cos(table)], dim=1)


def time_embedding(sd: SD, t: torch.Tensor) -> torch.Tensor:
    """model/diff3.py:69-87 : table lookup (int) or lerp (float), then 2x (Linear, SiLU)."""
    table = time_embedding_table(50).to(t.device)
    if t.dtype in (torch.int32, torch.int64):
        x = table[t]
    else:
        lo = torch.floor(t).long()
        hi = torch.ceil(t).long()
        x = table[lo] + (table[hi] - table[lo]) * (t - lo).unsqueeze(1)
    x = F.linear(x, sd["time_embedding.projection1.weight"], sd["time_embedding.projection1.bias"])
    x = x * torch.sigmoid(x)
    x = F.linear(x, sd["time_embedding.projection2.weight"], sd["time_embedding.projection2.bias"])
    return x * torch.sigmoid(x)


def _biconvglu(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """model/diff3.py:318-326."""
    x = F.conv2d(x, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"])
    left = F.conv2d(x, sd[p + ".l.weight"], sd[p + ".l.bias"], stride=(1, 2))
    right = F.conv2d(x, sd[p + ".r.weight"], sd[p + ".r.bias"], stride=(1, 2))
    lm = torch.sigmoid(F.conv2d(left, sd[p + ".l_conv.weight"], sd[p + ".l_conv.bias"]))
    rm = torch.sigmoid(F.conv2d(right, sd[p + ".r_conv.weight"], sd[p + ".r_conv.bias"]))
    return F.conv2d(left * rm + right * lm, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"])


def _biconvtransglu(sd: SD, p: str, x: torch.Tensor, temb: torch.Tensor) -> torch.Tensor:
    """model/diff3.py:341-351."""
    tb = F.linear(temb, sd[p + ".tp.weight"], sd[p + ".tp.bias"])
    x = F.conv_transpose2d(x + tb[:, :, None, None], sd[p + ".conv1.weight"], sd[p + ".conv1.bias"])
    left = F.conv_transpose2d(x, sd[p + ".l.weight"], sd[p + ".l.bias"], stride=(1, 2))
    right = F.conv_transpose2d(x, sd[p + ".r.weight"], sd[p + ".r.bias"], stride=(1, 2))
    lm = torch.sigmoid(F.conv_transpose2d(left, sd[p + ".l_conv.weight"], sd[p + ".l_conv.bias"]))
    rm = torch.sigmoid(F.conv_transpose2d(right, sd[p + ".r_conv.weight"], sd[p + ".r_conv.bias"]))
    return F.conv_transpose2d(left * rm + right * lm, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"])


def diffunet1_encoder(sd: SD, x: torch.Tensor, temb: torch.Tensor):
    """model/diff3.py:144-166 : causal pad row, time bias (added to the pad row too),
    BiConvGLU, BN, PReLU; returns the five skip tensors."""
    skips = []
    for i in range(1, 6):
        x = F.pad(x, (0, 0, 1, 0))
        tb = F.linear(temb, sd[f"en.tp{i}.weight"], sd[f"en.tp{i}.bias"])
        x = _biconvglu(sd, f"en.conv{i}", x + tb[:, :, None, None])
        x = _prelu(_bn(x, sd, f"en.en{i}.0"), sd[f"en.en{i}.1.weight"])
        skips.append(x)
    return x, skips


def tcm_residual(sd: SD, p: str, x: torch.Tensor, dilation: int) -> torch.Tensor:
    """model/diff3.py:249-257 (branches :221-247): BN sits before the zero-padded conv."""
    y = F.conv1d(x, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"])

    def branch(b):
        h = _bn(_prelu(y, sd[f"{p}.{b}.0.weight"]), sd, f"{p}.{b}.1")
        return F.conv1d(h, sd[f"{p}.{b}.2.weight"], sd[f"{p}.{b}.2.bias"],
                        padding=2 * dilation, dilation=dilation)

    g = branch("mainbranch") * torch.sigmoid(branch("maskbranch"))
    h = _bn(_prelu(g, sd[p + ".conv2.0.weight"]), sd, p + ".conv2.1")
    return x + F.conv1d(h, sd[p + ".conv2.2.weight"], sd[p + ".conv2.2.bias"])


def diffunet1_tcms(sd: SD, x: torch.Tensor) -> torch.Tensor:
    """model/diff3.py:49-53, 260-277."""
    b, c, t, f = x.shape
    x = x.permute(0, 2, 1, 3).reshape(b, t, c * f).permute(0, 2, 1)
    for m in range(3):
        for r in range(1, 7):
            x = tcm_residual(sd, f"TCMs.{m}.residual{r}", x, 2 ** (r - 1))
    x = x.permute(0, 2, 1).reshape(b, t, 64, 4).permute(0, 2, 1, 3)
    return x


def diffunet1_decoder(sd: SD, br: str, x: torch.Tensor, skips, temb) -> torch.Tensor:
    """model/diff3.py:206-212 ; Chomp_T drops the LAST row (:298-304)."""
    for i in range(5, 0, -1):
        x = _biconvtransglu(sd, f"{br}.de{i}.0", torch.cat((x, skips[i - 1]), dim=1), temb)
        x = x[:, :, :-1, :]
        if i != 1:
            x = _prelu(_bn(x, sd, f"{br}.de{i}.2"), sd[f"{br}.de{i}.3.weight"])
    return x


def diffunet1_forward(sd: SD, x: torch.Tensor, x_init: Optional[torch.Tensor], t: torch.Tensor,
                      taps: Optional[dict] = None) -> torch.Tensor:
    """model/diff3.py:37-57; with ``x_init=None`` the unconditioned twin Nocon.forward (model/piror_grad.py:28-40),
    which is the same network without Preprocess."""
    if x_init is None:
        h = x
    else:
        h = F.conv2d(torch.cat((x, x_init), dim=1), sd["preprocess.conv.weight"], sd["preprocess.conv.bias"])
    temb = time_embedding(sd, t)
    h, skips = diffunet1_encoder(sd, h, temb)
    if taps is not None:
        taps["temb"] = temb
        taps["skips"] = skips
    h = diffunet1_tcms(sd, h)
    if taps is not None:
        taps["tcm"] = h
    re = diffunet1_decoder(sd, "de_real", h, skips, temb)
    im = diffunet1_decoder(sd, "de_imag", h, skips, temb)
    return torch.cat((re, im), dim=1)


def diffunet_forward(sd: SD, x: torch.Tensor) -> torch.Tensor:
    """model/diff.py:23-33 : DiffUNet1 without Preprocess / time conditioning (the pad row stays 0)."""
    skips = []
    h = x
    for i in range(1, 6):
        h = _biconvglu(sd, f"en.conv{i}", F.pad(h, (0, 0, 1, 0)))                  # model/diff.py:68-90
        h = _prelu(_bn(h, sd, f"en.en{i}.0"), sd[f"en.en{i}.1.weight"])
        skips.append(h)
    h = diffunet1_tcms(sd, h)
    outs = []
    for br in ("de_real", "de_imag"):
        d = h
        for i in range(5, 0, -1):                                                  # model/diff.py:130-136, 264-272
            p = f"{br}.de{i}.0"
            z = F.conv_transpose2d(torch.cat((d, skips[i - 1]), dim=1), sd[p + ".conv1.weight"], sd[p + ".conv1.bias"])
            left = F.conv_transpose2d(z, sd[p + ".l.weight"], sd[p + ".l.bias"], stride=(1, 2))
            right = F.conv_transpose2d(z, sd[p + ".r.weight"], sd[p + ".r.bias"], stride=(1, 2))
            lm = torch.sigmoid(F.conv_transpose2d(left, sd[p + ".l_conv.weight"], sd[p + ".l_conv.bias"]))
            rm = torch.sigmoid(F.conv_transpose2d(right, sd[p + ".r_conv.weight"], sd[p + ".r_conv.bias"]))
            d = F.conv_transpose2d(left * rm + right * lm, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"])[:, :, :-1, :]
            if i != 1:
                d = _prelu(_bn(d, sd, f"{br}.de{i}.2"), sd[f"{br}.de{i}.3.weight"])
        outs.append(d)
    return torch.cat(outs, dim=1)


# ---------------------------------------------------------------------------
# a3: GCRN
# ---------------------------------------------------------------------------
def _lstm(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """nn.LSTM(512, 512, 1, batch_first=True) forward with zero initial state
    (model/gcrn.py:12-15); gate order i, f, g, o."""
    w_ih, w_hh = sd[p + ".weight_ih_l0"], sd[p + ".weight_hh_l0"]
    bias = sd[p + ".bias_ih_l0"] + sd[p + ".bias_hh_l0"]
    B, T, _ = x.shape
    H = w_hh.shape[1]
    pre = F.linear(x, w_ih, bias)
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    out = []
    for t in range(T):
        g = pre[:, t] + F.linear(h, w_hh)
        i, f, gg, o = g.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        out.append(h)
    return torch.stack(out, dim=1)


def gcrn_glstm(sd: SD, x: torch.Tensor) -> torch.Tensor:
    """model/gcrn.py:23-40 : stack+flatten interleaves the groups after layer 1, cat after layer 2."""
    b, c, t, f = x.shape
    out = x.transpose(1, 2).contiguous().view(b, t, -1)
    parts = torch.chunk(out, 2, dim=-1)
    out = torch.stack([_lstm(sd, f"glstm.lstm_list1.{i}", parts[i]) for i in range(2)], dim=-1)
    out = torch.flatten(out, start_dim=-2, end_dim=-1)
    out = F.layer_norm(out, (1024,), sd["glstm.ln1.weight"], sd["glstm.ln1.bias"], LN_EPS)
    parts = torch.chunk(out, 2, dim=-1)
    out = torch.cat([_lstm(sd, f"glstm.lstm_list2.{i}", parts[i]) for i in range(2)], dim=-1)
    out = F.layer_norm(out, (1024,), sd["glstm.ln2.weight"], sd["glstm.ln2.bias"], LN_EPS)
    return out.view(b, t, c, -1).transpose(1, 2).contiguous()


def _glu_conv(sd: SD, p: str, x):
    """model/gcrn.py:57-61."""
    a = F.conv2d(x, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], stride=(1, 2))
    g = F.conv2d(x, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], stride=(1, 2))
    return a * torch.sigmoid(g)


def _glu_convT(sd: SD, p: str, x, output_padding=(0, 0)):
    """model/gcrn.py:80-84."""
    a = F.conv_transpose2d(x, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], stride=(1, 2),
                           output_padding=output_padding)
    g = F.conv_transpose2d(x, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], stride=(1, 2),
                           output_padding=output_padding)
    return a * torch.sigmoid(g)


def gcrn_forward(sd: SD, x: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
    """model/gcrn.py:136-166.  ELU is in-place in the reference (:131), so the skip
    tensors e1..e4 are ELU'd a second time inside the decoder concat (:150-153) and,
    because the same tensor object is modified, the SECOND decoder sees skips that
    were already ELU'd twice and applies a third ELU (see note below)."""
    e = []
    h = x
    for i in range(1, 6):
        h = F.elu(_bn(_glu_conv(sd, f"conv{i}", h), sd, f"bn{i}"))
        e.append(h)
    out = gcrn_glstm(sd, e[4])
    if taps is not None:
        taps["enc"] = [t.clone() for t in e]
        taps["glstm"] = out
    out = torch.cat((out, e[4]), dim=1)

    # torch.cat copies, so the in-place ELU acts on the concatenated copy only; the
    # encoder tensors e1..e4 themselves are NOT modified and both decoders see the
    # same skips (single extra ELU each).  Verified against the reference module
    # in tests/golden/make_golden.py.
    outs = []
    for br in (1, 2):
        d = out
        for i in range(5, 1, -1):
            op = (0, 1) if i == 2 else (0, 0)
            d = _bn(_glu_convT(sd, f"conv{i}_t_{br}", d, op), sd, f"bn{i}_t_{br}")
            d = F.elu(torch.cat((d, e[i - 2]), dim=1))
        d = F.elu(_bn(_glu_convT(sd, f"conv1_t_{br}", d), sd, f"bn1_t_{br}"))
        outs.append(F.linear(d, sd[f"fc{br}.weight"], sd[f"fc{br}.bias"]))
    return torch.cat(outs, dim=1)


# ---------------------------------------------------------------------------
# a4: aia_complex_trans_ri (DB-AIAT prior)
# ---------------------------------------------------------------------------
GN_EPS = 1e-8        # nn.GroupNorm(1, 32, eps=1e-8), model/dbaiat.py:128-129
AIA_HEADS = 4        # model/dbaiat.py:124
AIA_LAYERS = 4       # model/dbaiat.py:457


def _chan_prelu(x: torch.Tensor, a: torch.Tensor) -> torch.Tensor:
    """nn.PReLU(C) on [B, C, T, F] (or the 1-parameter form)."""
    return torch.where(x >= 0, x, a.view(1, -1, 1, 1) * x)


def dense_block(sd: SD, p: str, x: torch.Tensor, depth: int = 4, taps: Optional[dict] = None) -> torch.Tensor:
    """DenseBlock.forward (model/dbaiat.py:623-631): causal (2 x 3) convs with time dilation 2^i over the
    concatenation [newest, ..., oldest, input]; LayerNorm over the frequency axis; per-channel PReLU."""
    skip = x
    out = x
    for i in range(depth):
        dil = 2 ** i
        out = F.pad(skip, (1, 1, dil, 0))                                    # pad_length = dil (twidth = 2)
        out = F.conv2d(out, sd[f"{p}.conv{i + 1}.weight"], sd[f"{p}.conv{i + 1}.bias"], dilation=(dil, 1))
        out = F.layer_norm(out, (out.shape[-1],), sd[f"{p}.norm{i + 1}.weight"], sd[f"{p}.norm{i + 1}.bias"], LN_EPS)
        out = _chan_prelu(out, sd[f"{p}.prelu{i + 1}.weight"])
        if taps is not None:
            taps[f"{p}.{i}"] = out
        skip = torch.cat([out, skip], dim=1)
    return out


def dense_encoder(sd: SD, x: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
    """dense_encoder.forward (model/dbaiat.py:497-501): [B,2,T,161] -> [B,64,T,80]."""
    p = "en_ri"
    out = F.conv2d(x, sd[p + ".inp_conv.weight"], sd[p + ".inp_conv.bias"])
    out = F.layer_norm(out, (N_FREQ,), sd[p + ".inp_norm.weight"], sd[p + ".inp_norm.bias"], LN_EPS)
    out = _chan_prelu(out, sd[p + ".inp_prelu.weight"])
    if taps is not None:
        taps["enc_in"] = out
    out = dense_block(sd, p + ".enc_dense1", out, taps=taps)
    out = F.conv2d(out, sd[p + ".enc_conv1.weight"], sd[p + ".enc_conv1.bias"], stride=(1, 2))
    out = F.layer_norm(out, (80,), sd[p + ".enc_norm1.weight"], sd[p + ".enc_norm1.bias"], LN_EPS)
    return _chan_prelu(out, sd[p + ".enc_prelu1.weight"])


def _mha(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """nn.MultiheadAttention(d, 4) self-attention, x [L, N, d] (model/dbaiat.py:75-77)."""
    L, N, d = x.shape
    hd = d // AIA_HEADS
    qkv = F.linear(x, sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"])
    q, k, v = (z.reshape(L, N * AIA_HEADS, hd).transpose(0, 1) for z in qkv.chunk(3, dim=-1))
    att = torch.softmax(torch.bmm(q, k.transpose(1, 2)) / math.sqrt(hd), dim=-1)
    o = torch.bmm(att, v).transpose(0, 1).reshape(L, N, d)
    return F.linear(o, sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"])


def _gru_dir(x: torch.Tensor, w_ih, w_hh, b_ih, b_hh, reverse: bool) -> torch.Tensor:
    """One direction of nn.GRU (gate order r, z, n; zero initial state), x [L, N, d] -> [L, N, H]."""
    L, N, _ = x.shape
    H = w_hh.shape[1]
    pre = F.linear(x, w_ih, b_ih)
    h = x.new_zeros(N, H)
    out = [None] * L
    for t in (range(L - 1, -1, -1) if reverse else range(L)):
        gh = F.linear(h, w_hh, b_hh)
        r = torch.sigmoid(pre[t, :, :H] + gh[:, :H])
        z = torch.sigmoid(pre[t, :, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(pre[t, :, 2 * H:] + r * gh[:, 2 * H:])
        h = (1 - z) * n + z * h
        out[t] = h
    return torch.stack(out)


def aia_encoder_layer(sd: SD, p: str, src: torch.Tensor) -> torch.Tensor:
    """TransformerEncoderLayer.forward (model/dbaiat.py:64-88), src [L, N, 32]; dropout = 0."""
    d = src.shape[-1]

    def ln(x, n):
        return F.layer_norm(x, (d,), sd[f"{p}.norm{n}.weight"], sd[f"{p}.norm{n}.bias"], LN_EPS)

    src = src + _mha(sd, p + ".self_attn", ln(src, 3))
    src = ln(src, 1)
    g = p + ".gru."
    out = torch.cat([_gru_dir(src, sd[g + "weight_ih_l0" + s], sd[g + "weight_hh_l0" + s], sd[g + "bias_ih_l0" + s],
                              sd[g + "bias_hh_l0" + s], s != "") for s in ("", "_reverse")], dim=-1)
    src = src + F.linear(torch.relu(out), sd[p + ".linear2.weight"], sd[p + ".linear2.bias"])
    return ln(src, 2)


def aia_transformer(sd: SD, x: torch.Tensor, taps: Optional[dict] = None) -> List[torch.Tensor]:
    """AIA_Transformer.forward (model/dbaiat.py:136-154): returns the per-layer outputs [B,64,T,F']."""
    p = "dual_trans"
    b, _, T, Fq = x.shape
    out = _prelu(F.conv2d(x, sd[p + ".input.0.weight"], sd[p + ".input.0.bias"]), sd[p + ".input.1.weight"])
    if taps is not None:
        taps["aia_in"] = out
    outs = []
    for i in range(AIA_LAYERS):
        row = out.permute(3, 0, 2, 1).reshape(Fq, b * T, -1)                 # attention / GRU along frequency
        row = aia_encoder_layer(sd, f"{p}.row_trans.{i}", row).view(Fq, b, T, -1).permute(1, 3, 2, 0)
        row = F.group_norm(row, 1, sd[f"{p}.row_norm.{i}.weight"], sd[f"{p}.row_norm.{i}.bias"], GN_EPS)
        col = out.permute(2, 0, 3, 1).reshape(T, b * Fq, -1)                 # ... along time
        col = aia_encoder_layer(sd, f"{p}.col_trans.{i}", col).view(T, b, Fq, -1).permute(1, 3, 0, 2)
        col = F.group_norm(col, 1, sd[f"{p}.col_norm.{i}.weight"], sd[f"{p}.col_norm.{i}.bias"], GN_EPS)
        if taps is not None:
            taps[f"aia_row{i}"], taps[f"aia_col{i}"] = row, col
        out = out + sd[p + ".k1"] * row + sd[p + ".k2"] * col
        if taps is not None:
            taps[f"aia_state{i}"] = out
        outs.append(F.conv2d(_prelu(out, sd[p + ".output.0.weight"]), sd[p + ".output.1.weight"],
                             sd[p + ".output.1.bias"]))
    return outs



def aham(sd: SD, outs: List[torch.Tensor]) -> torch.Tensor:
    """AHAM.forward (model/dbaiat.py:268-288): softmax over the 4 layer outputs of a 1x1 conv of their global
    average pools; result = last + sum_i a_i out_i."""
    y = torch.stack([F.conv2d(o.mean(dim=(2, 3), keepdim=True), sd["aham.conv1.weight"], sd["aham.conv1.bias"])
                     for o in outs], dim=-1)                                  # [B,1,1,1,4]
    a = torch.softmax(y, dim=-1)
    return outs[-1] + (torch.stack(outs, dim=-1) * a).sum(dim=-1)


def dense_decoder(sd: SD, p: str, x: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
    """dense_decoder.forward (model/dbaiat.py:541-548) with SPConvTranspose2d (:587-602): [B,64,T,80] -> [B,1,T,161]."""
    out = dense_block(sd, p + ".dec_dense1", x, taps=taps)
    out = F.conv2d(F.pad(out, (1, 1, 0, 0)), sd[p + ".dec_conv1.conv.weight"], sd[p + ".dec_conv1.conv.bias"])
    B, C2, T, Wd = out.shape
    out = out.view(B, 2, C2 // 2, T, Wd).permute(0, 2, 3, 4, 1).reshape(B, C2 // 2, T, 2 * Wd)   # sub-pixel, r = 2
    out = F.pad(out, (1, 0, 0, 0))
    out = F.layer_norm(out, (N_FREQ,), sd[p + ".dec_norm1.weight"], sd[p + ".dec_norm1.bias"], LN_EPS)
    out = _chan_prelu(out, sd[p + ".dec_prelu1.weight"])
    return F.conv2d(out, sd[p + ".out_conv.weight"], sd[p + ".out_conv.bias"])


def dbaiat_forward(sd: SD, x: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
    """aia_complex_trans_ri.forward (model/dbaiat.py:461-478): x [B,2,T,161] -> [B,2,T,161]."""
    e = dense_encoder(sd, x, taps)
    outs = aia_transformer(sd, e, taps)
    m = aham(sd, outs)
    if taps is not None:
        taps.update(enc=e, aia=outs, aham=m)
    return torch.cat([dense_decoder(sd, "de1", m, taps), dense_decoder(sd, "de2", m, taps)], dim=1)



# ---------------------------------------------------------------------------
# a8: the reverse loop  (trainer/complex_ddpm_trainer.py:941-998 ; batched twin :439-495)
# ---------------------------------------------------------------------------
# ---------------------------------------------------------------------------------------------- diff2.DiffWave (8f-4)
def diffwave_forward(sd: SD, audio: torch.Tensor, audio_init: torch.Tensor, t: torch.Tensor,
                     cycle: int = 10) -> torch.Tensor:
    """model/diff2.py:28-56 (DiffWave.forward) with ResidualBlock.forward :131-158 (default branch) and
    DiffusionEmbedding :71-95.  audio, audio_init [B, L]; t [B] (int64: table lookup, float: lerp) -> [B, 1, L]."""
    layers = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("residual_layers."))
    x = F.relu(F.conv1d(audio.unsqueeze(1), sd["input_projection.weight"], sd["input_projection.bias"]))      # :29-31
    cond = F.relu(F.conv1d(audio_init.unsqueeze(1), sd["input_projection.weight"], sd["input_projection.bias"]))   # :38-40
    table = time_embedding_table(50)                                   # :89-95, the same sinusoid table as diff3.py
    if t.dtype in (torch.int32, torch.int64):
        e = table[t]
    else:
        lo, hi = torch.floor(t).long(), torch.ceil(t).long()
        e = table[lo] + (table[hi] - table[lo]) * (t - lo).unsqueeze(-1)
    e = F.linear(e, sd["diffusion_embedding.projection1.weight"], sd["diffusion_embedding.projection1.bias"])
    e = e * torch.sigmoid(e)
    e = F.linear(e, sd["diffusion_embedding.projection2.weight"], sd["diffusion_embedding.projection2.bias"])
    e = e * torch.sigmoid(e)
    skip = None
    for i in range(layers):
        p, d = f"residual_layers.{i}", 2 ** (i % cycle)
        dstep = F.linear(e, sd[p + ".diffusion_projection.weight"], sd[p + ".diffusion_projection.bias"]).unsqueeze(-1)
        c = F.conv1d(cond, sd[p + ".conditioner_projection.weight"], sd[p + ".conditioner_projection.bias"], padding=d, dilation=d)
        y = F.conv1d(x + dstep, sd[p + ".dilated_conv.weight"], sd[p + ".dilated_conv.bias"], padding=d, dilation=d) + c
        gate, filt = torch.chunk(y, 2, dim=1)
        y = torch.sigmoid(gate) * torch.tanh(filt)
        y = F.conv1d(y, sd[p + ".output_projection.weight"], sd[p + ".output_projection.bias"])
        residual, s = torch.chunk(y, 2, dim=1)
        x = (x + residual) / math.sqrt(2.0)
        skip = s if skip is None else skip + s
    x = skip / math.sqrt(layers)
    x = F.relu(F.conv1d(x, sd["skip_projection.weight"], sd["skip_projection.bias"]))
    return F.conv1d(x, sd["output_projection.weight"], sd["output_projection.bias"])



def diffwave_enhance(sd: SD, noisy: torch.Tensor, x_T: torch.Tensor, fast: bool = True, cycle: int = 10) -> torch.Tensor:
    """The reverse update of trainer/complex_ddpm_trainer.py:967-992 (x = c1 (x - c2 eps); the noise coefficient is 0
    there) around diff2.DiffWave on waveforms, conditioned on the noisy waveform.  The reference never wires diff2 into
    a loop: this restates ITS loop with the other network, it is not pinned against a reference run.
    noisy, x_T [B, L] -> [B, L]"""
    alpha, beta, alpha_cum, _, T = inference_schedule(fast)
    x = x_T.clone()
    for n in range(len(alpha) - 1, -1, -1):
        t = torch.full((noisy.shape[0],), float(T[n]))
        # (the reference's _lerp_embedding only broadcasts for B = 1: evaluate per utterance)
        eps = torch.cat([diffwave_forward(sd, x[i:i + 1], noisy[i:i + 1], t[i:i + 1], cycle) for i in range(noisy.shape[0])])[:, 0]
        c1, c2 = 1.0 / alpha[n] ** 0.5, beta[n] / (1.0 - alpha_cum[n]) ** 0.5
        x = float(c1) * (x - float(c2) * eps)
    return x

def sigma_mask(x_init: torch.Tensor) -> torch.Tensor:
    """trainer/complex_ddpm_trainer.py:951-955 : 0.5 + 0.5*|X0| / max_{T,F}|X0| per (b, ch)."""
    tmp = torch.flatten(torch.abs(x_init), start_dim=2)
    tmp = tmp / torch.max(tmp, dim=2, keepdim=True).values
    tmp = tmp / 2 + 0.5
    return tmp.view(x_init.shape)


def reverse_loop(sd_ddpm: SD, x_init: torch.Tensor, x_T: torch.Tensor, fast: bool = True,
                 use_sigma_mask: bool = False, trace: Optional[list] = None, mode: str = "priorgrad",
                 cond: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Steps 5-8 of SURVEY.md Appendix A.  ``x_init`` is already divided by c=11.

    The additive noise coefficient ``newsigma = max(0, sigma - c1*gamma[n])`` is
    identically 0 (:986-992), so the loop is deterministic given ``x_T``.

    ``mode`` selects the trainer's three branches (:945-948, :967-975, :993-994): "priorgrad" (params.pirorgrad, the
    shipped configuration), "deltamu" (Nocon denoiser, x_T = z + X_init, no final add) and "condition" (DiffUNet1
    conditioned on the noisy features ``cond`` = batch_feat / 11, no final add).
    """
    alpha, beta, alpha_cum, sigmas, Tn = inference_schedule(fast)
    audio = x_T.clone()
    if mode == "deltamu":
        audio = audio + x_init
    if use_sigma_mask:
        audio = audio * (sigma_mask(x_init) ** 0.5)
    N = audio.shape[0]
    for n in range(len(alpha) - 1, -1, -1):
        c1 = 1 / alpha[n] ** 0.5
        c2 = beta[n] / (1 - alpha_cum[n]) ** 0.5
        t = torch.tensor([Tn[n]], device=audio.device).repeat(N)   # the reference's per-step scalar copy (:969)
        second = {"priorgrad": x_init, "deltamu": None, "condition": cond}[mode]
        eps = diffunet1_forward(sd_ddpm, audio, second, t)
        audio = c1 * (audio - c2 * eps)
        if trace is not None:
            trace.append((eps, audio.clone()))
    if mode == "priorgrad":
        audio = audio + x_init
    return audio * FEAT_SCALE


def enhance(sd_prior: SD, sd_ddpm: SD, wav: torch.Tensor, x_T: torch.Tensor, fast: bool = True,
            use_sigma_mask: bool = False, prior: str = "GCRN", stages: Optional[dict] = None, mode: str = "priorgrad"):
    """wav [B, L] -> enhanced wav [B, L]; SURVEY.md Appendix A steps 1-10
    (trainer/complex_ddpm_trainer.py:921-1016 with eval-mode BN on both nets)."""
    w, c = rms_normalize(wav)
    feat = stft_compress(w)
    prior_fn = {"GCRN": gcrn_forward, "DiffUNet": diffunet_forward, "aia_complex_trans_ri": dbaiat_forward}[prior]
    x_init = prior_fn(sd_prior, feat) / FEAT_SCALE
    spec = reverse_loop(sd_ddpm, x_init, x_T, fast, use_sigma_mask, mode=mode, cond=feat / FEAT_SCALE)
    out = decompress_istft(spec, wav.shape[-1]) * c
    if stages is not None:
        stages.update(feat=feat, x_init=x_init, spec=spec)
    return out


# ---------------------------------------------------------------------------
# 8f-2: segmental SNR (the scalar the reference's evaluation tracks)
# ---------------------------------------------------------------------------
def snr_seg(clean: np.ndarray, processed: np.ndarray, fs: int = 16000, frame_len: float = 0.03,
            overlap: float = 0.75) -> float:
    """utils/metrics.py:36-55 (SNRseg) with extractOverlappedWindows (:22-33) written out."""
    eps = np.finfo(np.float64).eps
    win = round(frame_len * fs)
    skip = int(np.floor((1 - overlap) * frame_len * fs))
    hann = 0.5 * (1 - np.cos(2 * np.pi * np.arange(1, win + 1) / (win + 1)))
    n = (len(clean) - (win - skip)) // skip
    idx = skip * np.arange(n)[:, None] + np.arange(win)[None, :]
    c = hann * np.asarray(clean, dtype=np.float64)[idx]
    p = hann * np.asarray(processed, dtype=np.float64)[idx]
    snr = 10 * np.log10((c ** 2).sum(-1) / (((c - p) ** 2).sum(-1) + eps) + eps)
    return float(np.mean(np.clip(snr, -10, 35)[:-1]))


# ---------------------------------------------------------------------------
# explicit-formula STFT / ISTFT (used to pin the kernels' arithmetic, fp64)
# ---------------------------------------------------------------------------
def stft_direct_f64(wav: np.ndarray) -> np.ndarray:
    """Definition of the transform torch.stft computes for the call at
    trainer/complex_ddpm_trainer.py:926-930, written out as a DFT in float64.
    wav [L] -> [2, T, F]."""
    L = wav.shape[0]
    x = np.pad(wav.astype(np.float64), (N_FFT // 2, N_FFT // 2), mode="reflect")
    T = 1 + L // HOP
    n = np.arange(N_FFT)
    win = 0.5 - 0.5 * np.cos(2 * np.pi * n / N_FFT)
    frames = np.stack([x[t * HOP:t * HOP + N_FFT] * win for t in range(T)])   # [T, 320]
    k = np.arange(N_FREQ)
    ang = -2 * np.pi * np.outer(n, k) / N_FFT
    return np.stack([frames @ np.cos(ang), frames @ np.sin(ang)])
