"""Weight tables for the two networks on the path.

The reference keeps its weights in ``nn.Module`` trees; the checkpoint format
(`trainer/complex_ddpm_trainer.py:616-622`, loaded at `:906-913`) is a list of
``state_dict``s.  The B200 path keeps weights as a flat *table*: a list of
``(key, shape, kind, fan)`` rows that (a) reproduces the reference's
``state_dict`` key names / shapes / registration order exactly (SURVEY.md
C.1), (b) drives random initialisation with the same distributions torch's
default ``reset_parameters`` uses, and (c) is what the packer walks when it
lays the weights out for the kernels.

Nothing here touches the GPU.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import List, Tuple

import torch

Row = Tuple[str, Tuple[int, ...], str, int]

N_FREQ = 161
N_TRAIN_STEPS = 50  # len(params.noise_schedule), utils/params.py:40


# ---------------------------------------------------------------------------
# row builders
# ---------------------------------------------------------------------------
def _affine(rows: List[Row], key: str, w_shape, fan_in: int, bias_len: int):
    """weight+bias pair initialised like nn.Conv*/nn.Linear (U(+-1/sqrt(fan_in)))."""
    rows.append((key + ".weight", tuple(w_shape), "uniform", fan_in))
    rows.append((key + ".bias", (bias_len,), "uniform", fan_in))


def _conv(rows, key, cout, cin, kh, kw):
    _affine(rows, key, (cout, cin, kh, kw), cin * kh * kw, cout)


def _convT(rows, key, cin, cout, kh, kw):
    # nn.ConvTranspose2d stores [Cin, Cout, kh, kw]; torch computes fan_in from dim 1.
    _affine(rows, key, (cin, cout, kh, kw), cout * kh * kw, cout)


def _conv1d(rows, key, cout, cin, k):
    _affine(rows, key, (cout, cin, k), cin * k, cout)


def _linear(rows, key, cout, cin):
    _affine(rows, key, (cout, cin), cin, cout)


def _bn(rows, key, c):
    rows.append((key + ".weight", (c,), "ones", 0))
    rows.append((key + ".bias", (c,), "zeros", 0))
    rows.append((key + ".running_mean", (c,), "zeros", 0))
    rows.append((key + ".running_var", (c,), "ones", 0))
    rows.append((key + ".num_batches_tracked", (), "count", 0))


def _prelu(rows, key):
    rows.append((key + ".weight", (1,), "prelu", 0))


# ---------------------------------------------------------------------------
# DiffUNet1  (model/diff3.py:14-57; key layout SURVEY.md C.1)
# ---------------------------------------------------------------------------
def diffunet1_table() -> List[Row]:
    rows: List[Row] = []
    _conv(rows, "preprocess.conv", 2, 4, 1, 1)                       # diff3.py:98-103
    _linear(rows, "time_embedding.projection1", 512, 128)            # diff3.py:62-67
    _linear(rows, "time_embedding.projection2", 512, 512)
    # encoder: five BiConvGLU blocks (diff3.py:105-143, 307-326)
    for i in range(1, 6):
        cin, kw = (2, 5) if i == 1 else (64, 3)
        p = f"en.conv{i}"
        _conv(rows, p + ".conv1", 32, cin, 1, 1)
        _conv(rows, p + ".l", 32, 32, 2, kw)
        _conv(rows, p + ".l_conv", 32, 32, 1, 1)
        _conv(rows, p + ".r", 32, 32, 2, kw)
        _conv(rows, p + ".r_conv", 32, 32, 1, 1)
        _conv(rows, p + ".conv2", 64, 32, 1, 1)
    for i in range(1, 6):
        _linear(rows, f"en.tp{i}", 2 if i == 1 else 64, 512)
    for i in range(1, 6):
        _bn(rows, f"en.en{i}.0", 64)
        _prelu(rows, f"en.en{i}.1")
    # two decoders of five BiConvTransGLU blocks (diff3.py:169-212, 329-351)
    for br in ("de_real", "de_imag"):
        for i in range(5, 0, -1):
            cout, kw = (1, 5) if i == 1 else (64, 3)
            p = f"{br}.de{i}.0"
            _linear(rows, p + ".tp", 128, 512)
            _convT(rows, p + ".conv1", 128, 32, 1, 1)
            _convT(rows, p + ".l", 32, 32, 2, kw)
            _convT(rows, p + ".l_conv", 32, 32, 1, 1)
            _convT(rows, p + ".r_conv", 32, 32, 1, 1)
            _convT(rows, p + ".r", 32, 32, 2, kw)
            _convT(rows, p + ".conv2", 32, cout, 1, 1)
            if i != 1:
                _bn(rows, f"{br}.de{i}.2", 64)
                _prelu(rows, f"{br}.de{i}.3")
    # 3 x 6 dilated residual blocks (diff3.py:215-277)
    for m in range(3):
        for r in range(1, 7):
            p = f"TCMs.{m}.residual{r}"
            _conv1d(rows, p + ".conv1", 64, 256, 1)
            for br in ("mainbranch", "maskbranch"):
                _prelu(rows, f"{p}.{br}.0")
                _bn(rows, f"{p}.{br}.1", 64)
                _conv1d(rows, f"{p}.{br}.2", 64, 64, 5)
            _prelu(rows, p + ".conv2.0")
            _bn(rows, p + ".conv2.1", 64)
            _conv1d(rows, p + ".conv2.2", 256, 64, 1)
    return rows


TCM_DILATIONS = [1, 2, 4, 8, 16, 32] * 3  # diff3.py:263-268


def diffunet_table() -> List[Row]:
    """DiffUNet (model/diff.py:13-33), the default prior of conf/diff.yml: DiffUNet1 without Preprocess,
    TimeEmbedding and the per-block time projections; same registration order otherwise."""
    drop = ("preprocess.", "time_embedding.", "en.tp")
    return [r for r in diffunet1_table() if not r[0].startswith(drop) and ".tp." not in r[0]]


# ---------------------------------------------------------------------------
# GCRN  (model/gcrn.py:87-134)
# ---------------------------------------------------------------------------
GCRN_ENC_CH = [2, 16, 32, 64, 128, 256]


def gcrn_table() -> List[Row]:
    rows: List[Row] = []
    for i in range(1, 6):
        cin, cout = GCRN_ENC_CH[i - 1], GCRN_ENC_CH[i]
        _conv(rows, f"conv{i}.conv1", cout, cin, 1, 3)
        _conv(rows, f"conv{i}.conv2", cout, cin, 1, 3)
    for lst in ("lstm_list1", "lstm_list2"):
        for g in range(2):
            p = f"glstm.{lst}.{g}"
            rows.append((p + ".weight_ih_l0", (2048, 512), "uniform", 512))
            rows.append((p + ".weight_hh_l0", (2048, 512), "uniform", 512))
            rows.append((p + ".bias_ih_l0", (2048,), "uniform", 512))
            rows.append((p + ".bias_hh_l0", (2048,), "uniform", 512))
    for i in (1, 2):
        rows.append((f"glstm.ln{i}.weight", (1024,), "ones", 0))
        rows.append((f"glstm.ln{i}.bias", (1024,), "zeros", 0))
    dec_ch = {5: (512, 128), 4: (256, 64), 3: (128, 32), 2: (64, 16), 1: (32, 1)}
    for br in (1, 2):
        for i in range(5, 0, -1):
            cin, cout = dec_ch[i]
            _convT(rows, f"conv{i}_t_{br}.conv1", cin, cout, 1, 3)
            _convT(rows, f"conv{i}_t_{br}.conv2", cin, cout, 1, 3)
    for i in range(1, 6):
        _bn(rows, f"bn{i}", GCRN_ENC_CH[i])
    for br in (1, 2):
        for i in range(5, 0, -1):
            _bn(rows, f"bn{i}_t_{br}", dec_ch[i][1])
    _linear(rows, "fc1", N_FREQ, N_FREQ)
    _linear(rows, "fc2", N_FREQ, N_FREQ)
    return rows


# ---------------------------------------------------------------------------
# aia_complex_trans_ri  (DB-AIAT prior, model/dbaiat.py:450-478)
# ---------------------------------------------------------------------------
def _ln(rows, key, n):
    rows.append((key + ".weight", (n,), "ones", 0))
    rows.append((key + ".bias", (n,), "zeros", 0))


def _dense_block(rows, p, width, depth=4, ch=64):
    """DenseBlock (dbaiat.py:605-631): conv{i} [64, 64*i, 2, 3] (time dilation 2^(i-1)), LayerNorm over the
    frequency axis, per-channel PReLU."""
    for i in range(1, depth + 1):
        _conv(rows, f"{p}.conv{i}", ch, ch * i, 2, 3)
        _ln(rows, f"{p}.norm{i}", width)
        rows.append((f"{p}.prelu{i}.weight", (ch,), "prelu", 0))


def _aia_layer(rows, p, d=32):
    """TransformerEncoderLayer (dbaiat.py:41-88): MHA(4 heads) + bidirectional GRU(d -> 2d) + Linear(4d -> d)."""
    rows.append((p + ".self_attn.in_proj_weight", (3 * d, d), "xavier", d + 3 * d))
    rows.append((p + ".self_attn.in_proj_bias", (3 * d,), "zeros", 0))
    rows.append((p + ".self_attn.out_proj.weight", (d, d), "uniform", d))
    rows.append((p + ".self_attn.out_proj.bias", (d,), "zeros", 0))
    for sfx in ("", "_reverse"):
        rows.append((p + ".gru.weight_ih_l0" + sfx, (6 * d, d), "uniform", 2 * d))
        rows.append((p + ".gru.weight_hh_l0" + sfx, (6 * d, 2 * d), "uniform", 2 * d))
        rows.append((p + ".gru.bias_ih_l0" + sfx, (6 * d,), "uniform", 2 * d))
        rows.append((p + ".gru.bias_hh_l0" + sfx, (6 * d,), "uniform", 2 * d))
    _linear(rows, p + ".linear2", d, 4 * d)
    for n in (1, 2, 3):
        _ln(rows, f"{p}.norm{n}", d)


AIA_LAYERS = 4


def dbaiat_table() -> List[Row]:
    rows: List[Row] = []
    # dense_encoder (dbaiat.py:481-501)
    _conv(rows, "en_ri.inp_conv", 64, 2, 1, 1)
    _ln(rows, "en_ri.inp_norm", N_FREQ)
    rows.append(("en_ri.inp_prelu.weight", (64,), "prelu", 0))
    _dense_block(rows, "en_ri.enc_dense1", N_FREQ)
    _conv(rows, "en_ri.enc_conv1", 64, 64, 1, 3)
    _ln(rows, "en_ri.enc_norm1", 80)
    rows.append(("en_ri.enc_prelu1.weight", (64,), "prelu", 0))
    # AIA_Transformer(64, 64, num_layers=4) (dbaiat.py:91-154)
    rows.append(("dual_trans.k1", (1,), "ones", 0))
    rows.append(("dual_trans.k2", (1,), "ones", 0))
    _conv(rows, "dual_trans.input.0", 32, 64, 1, 1)
    rows.append(("dual_trans.input.1.weight", (1,), "prelu", 0))
    for kind in ("row_trans", "col_trans"):
        for i in range(AIA_LAYERS):
            _aia_layer(rows, f"dual_trans.{kind}.{i}")
    for kind in ("row_norm", "col_norm"):
        for i in range(AIA_LAYERS):
            _ln(rows, f"dual_trans.{kind}.{i}", 32)          # GroupNorm(1, 32)
    rows.append(("dual_trans.output.0.weight", (1,), "prelu", 0))
    _conv(rows, "dual_trans.output.1", 64, 32, 1, 1)
    # AHAM (dbaiat.py:249-288); k3 is registered but never used
    rows.append(("aham.k3", (1,), "zeros", 0))
    _conv(rows, "aham.conv1", 1, 64, 1, 1)
    # two dense_decoders (dbaiat.py:527-548)
    for de in ("de1", "de2"):
        _dense_block(rows, f"{de}.dec_dense1", 80)
        _conv(rows, f"{de}.dec_conv1.conv", 128, 64, 1, 3)     # SPConvTranspose2d, r = 2
        _ln(rows, f"{de}.dec_norm1", N_FREQ)
        rows.append((f"{de}.dec_prelu1.weight", (64,), "prelu", 0))
        _conv(rows, f"{de}.out_conv", 1, 64, 1, 1)
    return rows


def nocon_table() -> List[Row]:
    """Nocon (model/piror_grad.py:15-40), the ``deltamu`` denoiser: DiffUNet1 without Preprocess."""
    return [r for r in diffunet1_table() if not r[0].startswith("preprocess.")]


# ---------------------------------------------------------------------------
# DiffWave  (model/diff2.py:12-158) -- SURVEY 8(f) item 4.  utils/params.py has no entries for it, so the
# hyper-parameters are the DiffWave base configuration the module was written for.
# ---------------------------------------------------------------------------
DIFFWAVE_CHANNELS = 64          # params.residual_channels
DIFFWAVE_LAYERS = 30            # params.residual_layers
DIFFWAVE_CYCLE = 10             # params.dilation_cycle_length  (dilations 1 .. 512)


def diffwave_table(layers: int = DIFFWAVE_LAYERS, channels: int = DIFFWAVE_CHANNELS) -> List[Row]:
    """registration order of DiffWave.__init__ (diff2.py:15-26) and ResidualBlock.__init__ (:112-129, the default
    ``fix_in=False, split=False`` branch: one 2C-wide output_projection)"""
    rows: List[Row] = []
    c = channels
    _conv1d(rows, "input_projection", c, 1, 1)
    _linear(rows, "diffusion_embedding.projection1", 512, 128)
    _linear(rows, "diffusion_embedding.projection2", 512, 512)
    _convT(rows, "spectrogram_upsampler.conv1", 1, 1, 3, 32)          # registered but unused by forward (:38 commented out)
    _convT(rows, "spectrogram_upsampler.conv2", 1, 1, 3, 32)
    for i in range(layers):
        p = f"residual_layers.{i}"
        _conv1d(rows, p + ".dilated_conv", 2 * c, c, 3)
        _linear(rows, p + ".diffusion_projection", c, 512)
        _conv1d(rows, p + ".conditioner_projection", 2 * c, c, 3)
        _conv1d(rows, p + ".output_projection", 2 * c, c, 1)
    _conv1d(rows, "skip_projection", c, c, 1)
    _conv1d(rows, "output_projection", 1, c, 1)
    return rows


TABLES = {"DiffWave": diffwave_table, "Nocon": nocon_table, "DiffUNet1": diffunet1_table, "GCRN": gcrn_table, "DiffUNet": diffunet_table,
          "aia_complex_trans_ri": dbaiat_table}


# ---------------------------------------------------------------------------
# initialisation
# ---------------------------------------------------------------------------
def init_state_dict(name: str, seed: int = 1234) -> "OrderedDict[str, torch.Tensor]":
    """Random-init weights with torch's default distributions (seeded, CPU, fp32).

    ``main.py:23`` seeds the reference with 1234; the draw order differs from
    ``nn.Module.__init__`` so the values are not the reference's, only the
    distributions are.
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for key, shape, kind, fan in TABLES[name]():
        if kind == "uniform":
            bound = 1.0 / math.sqrt(fan)
            sd[key] = (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound
        elif kind == "xavier":   # nn.init.xavier_uniform_ (MultiheadAttention.in_proj_weight); fan = fan_in + fan_out
            bound = math.sqrt(6.0 / fan)
            sd[key] = (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound
        elif kind == "ones":
            sd[key] = torch.ones(shape, dtype=torch.float32)
        elif kind == "zeros":
            sd[key] = torch.zeros(shape, dtype=torch.float32)
        elif kind == "count":
            sd[key] = torch.zeros((), dtype=torch.int64)
        elif kind == "prelu":
            sd[key] = torch.full(shape, 0.25, dtype=torch.float32)
        else:  # pragma: no cover
            raise ValueError(kind)
    return sd


def randomize_norm_stats(sd, seed: int = 4321):
    """Seeded perturbation of BN running stats / affine, LayerNorm affine and PReLU
    slopes so that folding mistakes show up in parity runs."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    out = OrderedDict()
    for k, v in sd.items():
        if k.endswith("running_mean"):
            v = 0.2 * torch.randn(v.shape, generator=g)
        elif k.endswith("running_var"):
            v = 0.5 + torch.rand(v.shape, generator=g)
        elif k.endswith("num_batches_tracked"):
            v = v.clone()
        elif _is_norm_affine(k, sd):
            if k.endswith(".weight"):
                v = 0.75 + 0.5 * torch.rand(v.shape, generator=g)
            else:
                v = 0.1 * torch.randn(v.shape, generator=g)
        elif _is_prelu_key(k, sd):
            v = 0.05 + 0.4 * torch.rand(v.shape, generator=g)
        elif k in ("dual_trans.k1", "dual_trans.k2"):
            v = 0.6 + 0.8 * torch.rand(v.shape, generator=g)
        elif k.endswith(("in_proj_bias", "out_proj.bias")):   # zero-initialised by torch; make them visible
            v = 0.05 * torch.randn(v.shape, generator=g)
        else:
            v = v.clone()
        out[k] = v.to(sd[k].dtype)
    return out


def _is_norm_affine(k: str, sd) -> bool:
    stem = k.rsplit(".", 1)[0]
    if stem + ".running_mean" in sd:
        return k.endswith(".weight") or k.endswith(".bias")
    return stem.startswith("glstm.ln") or "norm" in stem.rsplit(".", 2)[-1] or "_norm." in stem


def _is_prelu_key(k: str, sd) -> bool:
    # PReLU rows: 1-element ".weight" rows without a sibling ".bias", or the per-channel dbaiat "prelu" modules.
    if not k.endswith(".weight") or (k[:-7] + ".bias") in sd:
        return False
    return sd[k].numel() == 1 or "prelu" in k


def table_shapes(name: str):
    return OrderedDict((k, tuple(s)) for k, s, _, _ in TABLES[name]())
