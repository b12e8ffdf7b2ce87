"""Weight packing: reference ``state_dict`` -> the operand layouts the sm_100a kernels read.

Everything GEMM-shaped is stored as tcgen05 B operands in the K-major no-swizzle
canonical layout ("CP8", see csrc/umma.cuh): a matrix W[N][K] becomes
``[K/8][N][8]`` so that one 16-byte unit holds 8 consecutive K values of one
output channel.  Convolutions are stored tap by tap; each tap is its own
``[Cin/8][N][8]`` block because a tap is just a shifted A window in the kernels.

Packing happens once per checkpoint on the host (numpy, fp32 -> bf16 at the end);
``tests/test_pack_emulation.py`` re-runs the kernels' arithmetic in numpy from these
very arrays and checks it against the oracle, so the layouts are pinned on CPU.

Reference layouts: SURVEY.md C.1 (model/diff3.py, model/gcrn.py).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict

import numpy as np
import torch

BN_EPS = 1e-5
ENC_F = [161, 79, 39, 19, 9, 4]          # F after encoder block i (index 0 = input)
N_BIAS_ROW = 452                          # floats per row of the time-bias table

# offsets into one row of the time-bias table produced by pdse_time_embed
BIAS_TB1 = 0


def bias_off_enc(i: int) -> int:          # i = 2..5
    return 2 + 32 * (i - 2)


def bias_off_dec(branch: int, i: int) -> int:   # branch 0 = de_real, 1 = de_imag ; i = 5..1
    return 130 + 160 * branch + 32 * (5 - i)


def _np(t) -> np.ndarray:
    return t.detach().cpu().double().numpy() if isinstance(t, torch.Tensor) else np.asarray(t, dtype=np.float64)


def cp8(w: np.ndarray) -> np.ndarray:
    """W[N][K] -> [K/8][N][8] (K zero-padded to a multiple of 8)."""
    n, k = w.shape
    kp = (k + 7) // 8 * 8
    out = np.zeros((n, kp), dtype=np.float64)
    out[:, :k] = w
    return np.ascontiguousarray(out.reshape(n, kp // 8, 8).transpose(1, 0, 2))


def bn_affine(sd, key):
    s = _np(sd[key + ".weight"]) / np.sqrt(_np(sd[key + ".running_var"]) + BN_EPS)
    return s, _np(sd[key + ".bias"]) - _np(sd[key + ".running_mean"]) * s


class Blob:
    """Named arrays concatenated into one 16-bit operand blob (fp16 / bf16, csrc/opfmt.h) and one fp32 blob."""

    def __init__(self):
        self.h: "OrderedDict[str, np.ndarray]" = OrderedDict()   # -> 16-bit operand format
        self.f: "OrderedDict[str, np.ndarray]" = OrderedDict()   # -> fp32

    def offsets(self, which: str) -> Dict[str, int]:
        d, off, out = (self.h if which == "h" else self.f), 0, {}
        for k, v in d.items():
            out[k] = off
            off += v.size
        out["_total"] = off
        return out

    def flat(self, which: str) -> np.ndarray:
        d = self.h if which == "h" else self.f
        return np.concatenate([v.reshape(-1) for v in d.values()]).astype(np.float32)


def _pad4(v: np.ndarray, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.float64)
    out[:v.size] = v.reshape(-1)
    return out


# ----------------------------------------------------------------------------
# GLU tail shared by every BiConv(Trans)GLU block: gates, 1x1 out, BN, PReLU
# ----------------------------------------------------------------------------
def bias_block(b: np.ndarray) -> np.ndarray:
    """Bias as a tcgen05 B operand: [2 chunks][N][8] with row n = (hi, lo, 0, ...), hi + lo = b to ~2^-17.
    Multiplied by the constant A chunk (1, 1, 0, ...) ("ones plane") it adds b[n] to every accumulator row,
    so no epilogue ever has to load or add a per-channel bias."""
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    from . import lib as _lib       # (lazy: the operand format is a property of the built library, csrc/opfmt.h)
    hi = torch.tensor(b, dtype=torch.float32).to(_lib.op_dtype()).double().numpy()
    out = np.zeros((2, b.size, 8))
    out[0, :, 0] = hi
    out[0, :, 1] = b - hi
    return out


def _gate_mats(sd, p: str, transposed: bool):
    def mat(name):   # -> [out][in]
        w = _np(sd[f"{p}.{name}.weight"])[:, :, 0, 0]
        return w.T if transposed else w
    return mat("l_conv"), mat("r_conv"), mat("conv2")


def with_gates(wlr: np.ndarray, wlc: np.ndarray, wrc: np.ndarray) -> np.ndarray:
    """[64][K] rows (l | r) -> [128][K] rows (l | r | lm' | rm'): the gate 1x1 convs act linearly on l / r, so they
    are composed into the conv that produces l / r (lm' = 0.5 Wlc l, rm' = 0.5 Wrc r; the 0.5 is the sigmoid's)."""
    return np.concatenate([wlr, 0.5 * wlc @ wlr[:32], 0.5 * wrc @ wlr[32:]])


def _glu_tail(blob: Blob, sd, p: str, transposed: bool, bn_key, prelu_key, cout: int, blr_extra=None):
    """cross-gating / 1x1 / BN / PReLU of a BiConv(Trans)GLU block (diff3.py:321-326).

    sigmoid(z) = 0.5 tanh(z/2) + 0.5: the inner 1/2 is folded into the (composed) gate weights and biases, the
    outer affine into g' = l (tanh_r + 1) + r (tanh_l + 1) = 2 g and from there into the next linear layer
    (w2 * 0.5); the BN scale is folded into w2's rows and every bias rides on a bias MMA."""
    wlc, wrc, w2 = _gate_mats(sd, p, transposed)
    b2 = _np(sd[p + ".conv2.bias"])
    blr = np.concatenate([_np(sd[p + ".l.bias"]), _np(sd[p + ".r.bias"])])
    if blr_extra is not None:
        blr = blr + blr_extra
    if cout == 64:
        s, sh = bn_affine(sd, bn_key)
        blob.h["w2"] = cp8(0.5 * s[:, None] * w2)           # [4][64][8]
    blob.h["b_lr4"] = bias_block(np.concatenate([blr, 0.5 * (wlc @ blr[:32] + _np(sd[p + ".l_conv.bias"])),
                                                 0.5 * (wrc @ blr[32:] + _np(sd[p + ".r_conv.bias"]))]))   # [2][128][8]
    if cout == 64:
        blob.h["b_out"] = bias_block(b2 * s + sh)           # (D + b2) * s + sh
        blob.f["slope"] = _pad4(_np(sd[prelu_key + ".weight"]), 4)
    else:                                                   # de1: 32 -> 1, no BN / PReLU
        blob.f["w2vec"] = 0.5 * w2.reshape(32)
        blob.f["b2"] = _pad4(b2, 4)


def pack_enc1(sd) -> Blob:
    """en.conv1 with Preprocess folded in front (model/diff3.py:38, 146-147, 318-320).

    conv1 (1x1, 2->32) is linear and feeds l/r directly, so it is composed into the
    (2x5) taps: Wf[o][c][dt][df] = sum_k Wlr[o][k][dt][df] * W1[k][c]; K = 2*2*5 = 20
    (k = c*10 + dt*5 + df, zero-padded to 32).  The pad row and the time bias enter
    through the A operand u = preprocess(x, x_init) + tb (u = tb on the pad row)."""
    b = Blob()
    p = "en.conv1"
    w1 = _np(sd[p + ".conv1.weight"])[:, :, 0, 0]           # [32][2]
    b1 = _np(sd[p + ".conv1.bias"])
    wlr = np.concatenate([_np(sd[p + ".l.weight"]), _np(sd[p + ".r.weight"])])    # [64][32][2][5]
    wf = np.einsum("okdf,kc->ocdf", wlr, w1).reshape(64, 20)
    wlc, wrc, _ = _gate_mats(sd, p, False)
    b.h["wf"] = cp8(np.pad(with_gates(wf, wlc, wrc), ((0, 0), (0, 12))))          # [4][128][8]
    _glu_tail(b, sd, p, False, "en.en1.0", "en.en1.1", 64, blr_extra=np.einsum("okdf,k->o", wlr, b1))
    b.f["wp"] = _np(sd["preprocess.conv.weight"])[:, :, 0, 0].reshape(8)          # [2][4]
    b.f["bp"] = _pad4(_np(sd["preprocess.conv.bias"]), 4)
    return b


def pack_enc(sd, i: int) -> Blob:
    """en.conv{i}, i = 2..5 (kernel (2,3), stride (1,2))."""
    b = Blob()
    p = f"en.conv{i}"
    b.h["w1"] = cp8(_np(sd[p + ".conv1.weight"])[:, :, 0, 0])                     # [8][32][8]
    wlr = np.concatenate([_np(sd[p + ".l.weight"]), _np(sd[p + ".r.weight"])])    # [64][32][2][3]
    wlc, wrc, _ = _gate_mats(sd, p, False)
    b.h["wlr"] = np.stack([cp8(with_gates(wlr[:, :, dt, df], wlc, wrc)) for dt in range(2) for df in range(3)])   # [6][4][128][8]
    _glu_tail(b, sd, p, False, f"en.en{i}.0", f"en.en{i}.1", 64)
    return b


def pack_dec(sd, br: str, i: int) -> Blob:
    """{br}.de{i}.0 BiConvTransGLU (model/diff3.py:341-351).  ConvTranspose2d weights are
    [Cin][Cout][kh][kw]; out[t'][f'] gathers h[t'-dt][(f'-df)/2], so even outputs f'=2j
    use df=2a with h[j-a] and odd outputs f'=2j+1 use df=2a+1 with h[j-a]."""
    b = Blob()
    p = f"{br}.de{i}.0"
    kw = 5 if i == 1 else 3
    g = (kw - 1) // 2
    b.h["w1"] = cp8(_np(sd[p + ".conv1.weight"])[:, :, 0, 0].T)                   # [16][32][8]
    wl, wr = _np(sd[p + ".l.weight"]), _np(sd[p + ".r.weight"])                  # [32 in][32 out][2][kw]
    wlr = np.concatenate([wl.transpose(1, 0, 2, 3), wr.transpose(1, 0, 2, 3)])   # [64 out][32 in][2][kw]
    wlc, wrc, _ = _gate_mats(sd, p, True)
    b.h["wlr_even"] = np.stack([cp8(with_gates(wlr[:, :, dt, 2 * a], wlc, wrc)) for dt in range(2) for a in range(g + 1)])
    b.h["wlr_odd"] = np.stack([cp8(with_gates(wlr[:, :, dt, 2 * a + 1], wlc, wrc)) for dt in range(2) for a in range(g)])
    if i == 1:
        _glu_tail(b, sd, p, True, None, None, 1)
    else:
        _glu_tail(b, sd, p, True, f"{br}.de{i}.2", f"{br}.de{i}.3", 64)
    return b


# TCM channel order: the reference flattens [B,64,T,4] to channel c*4+f (diff3.py:49-50).
# The kernels keep kk = f*64 + c (chunk kc = f*8 + c/8) so that the encoder's CP8 output
# converts with pure 16-byte moves.
def tcm_perm() -> np.ndarray:
    kk = np.arange(256)
    return (kk % 64) * 4 + kk // 64          # kk -> reference channel


def pack_tcm(sd, m: int, r: int) -> Blob:
    """Residual block (diff3.py:215-257).  bf16: w1 | wm | wk | w3 | b_1 | b_m | b_k | b_3 ; every conv bias is a
    bias block (bias MMA); sigmoid(z) = 0.5 tanh(z/2) + 0.5 is folded: wk, bk carry the inner 1/2 and the BN scale
    after the (positively homogeneous) PReLU carries the outer 1/2:  g = 0.5 * main * (tanh + 1)."""
    b = Blob()
    p = f"TCMs.{m}.residual{r}"
    perm = tcm_perm()
    b.h["w1"] = cp8(_np(sd[p + ".conv1.weight"])[:, perm, 0])                     # [32][64][8]
    wm, wk = _np(sd[f"{p}.mainbranch.2.weight"]), 0.5 * _np(sd[f"{p}.maskbranch.2.weight"])   # [64][64][5]
    b.h["wm"] = np.stack([cp8(wm[:, :, tap]) for tap in range(5)])                # [5][8][64][8]
    b.h["wk"] = np.stack([cp8(wk[:, :, tap]) for tap in range(5)])
    b.h["w3"] = cp8(_np(sd[p + ".conv2.2.weight"])[perm, :, 0])                   # [8][256][8]
    b.h["b_1"] = bias_block(_np(sd[p + ".conv1.bias"]))
    b.h["b_m"] = bias_block(_np(sd[f"{p}.mainbranch.2.bias"]))
    b.h["b_k"] = bias_block(0.5 * _np(sd[f"{p}.maskbranch.2.bias"]))
    b.h["b_3"] = bias_block(_np(sd[p + ".conv2.2.bias"])[perm])                   # [2][256][8]
    for tag, br in (("m", "mainbranch"), ("k", "maskbranch")):
        s, sh = bn_affine(sd, f"{p}.{br}.1")
        b.f["s" + tag] = s
        b.f["sh" + tag] = sh
    s, sh = bn_affine(sd, p + ".conv2.1")
    b.f["sc"] = 0.5 * s
    b.f["shc"] = sh
    b.f["slopes"] = np.array([_np(sd[p + ".mainbranch.0.weight"])[0], _np(sd[p + ".maskbranch.0.weight"])[0],
                              _np(sd[p + ".conv2.0.weight"])[0], 0.0])
    return b


def pack_time(sd) -> Dict[str, np.ndarray]:
    """TimeEmbedding MLP (diff3.py:62-87) plus every per-block time projection composed with
    the block's 1x1 conv: hb = W1 (Wtp temb + btp) + b1 = (W1 Wtp) temb + (W1 btp + b1)."""
    rows, bias = np.zeros((N_BIAS_ROW, 512)), np.zeros(N_BIAS_ROW)
    rows[0:2] = _np(sd["en.tp1.weight"])
    bias[0:2] = _np(sd["en.tp1.bias"])
    for i in range(2, 6):
        w1 = _np(sd[f"en.conv{i}.conv1.weight"])[:, :, 0, 0]                      # [32][64]
        o = bias_off_enc(i)
        rows[o:o + 32] = w1 @ _np(sd[f"en.tp{i}.weight"])
        bias[o:o + 32] = w1 @ _np(sd[f"en.tp{i}.bias"]) + _np(sd[f"en.conv{i}.conv1.bias"])
    for bi, br in enumerate(("de_real", "de_imag")):
        for i in range(5, 0, -1):
            p = f"{br}.de{i}.0"
            w1 = _np(sd[p + ".conv1.weight"])[:, :, 0, 0].T                       # [32][128]
            o = bias_off_dec(bi, i)
            rows[o:o + 32] = w1 @ _np(sd[p + ".tp.weight"])
            bias[o:o + 32] = w1 @ _np(sd[p + ".tp.bias"]) + _np(sd[p + ".conv1.bias"])
    steps = np.arange(50, dtype=np.float64)[:, None]
    dims = np.arange(64, dtype=np.float64)[None, :]
    # diff3.py:89-95 builds the table in float32; keep the float32 rounding of the argument
    arg = (torch.arange(50).unsqueeze(1) * 10.0 ** (torch.arange(64).unsqueeze(0) * 4.0 / 63.0))
    table = torch.cat([torch.sin(arg), torch.cos(arg)], dim=1).double().numpy()
    del steps, dims
    return {
        "table": table.astype(np.float32),                                        # [50][128]
        "p1w": _np(sd["time_embedding.projection1.weight"]).astype(np.float32),   # [512][128]
        "p1b": _np(sd["time_embedding.projection1.bias"]).astype(np.float32),
        "p2w": _np(sd["time_embedding.projection2.weight"]).astype(np.float32),   # [512][512]
        "p2b": _np(sd["time_embedding.projection2.bias"]).astype(np.float32),
        "rows": rows.astype(np.float32),                                          # [452][512]
        "bias": bias.astype(np.float32),
    }


def pack_diffunet1(sd):
    out = {"enc1": pack_enc1(sd), "time": pack_time(sd)}
    for i in range(2, 6):
        out[f"enc{i}"] = pack_enc(sd, i)
    for bi, br in enumerate(("de_real", "de_imag")):
        for i in range(5, 0, -1):
            out[f"dec{bi}_{i}"] = pack_dec(sd, br, i)
    for m in range(3):
        for r in range(1, 7):
            out[f"tcm{m * 6 + r - 1}"] = pack_tcm(sd, m, r)
    return out


# =============================================================================
# GCRN prior (model/gcrn.py:87-166)
# =============================================================================
GCRN_CH = [2, 16, 32, 64, 128, 256]
GCRN_F = [161, 80, 39, 19, 9, 4]             # F after encoder layer i
GCRN_DEC = {5: (512, 128, 4, 9), 4: (256, 64, 9, 19), 3: (128, 32, 19, 39), 2: (64, 16, 39, 80), 1: (32, 1, 80, 161)}
#            i: (Cin, Cout, Fin, Fout)


def _stream(w_nk_taps, ntile: int, kb: int) -> np.ndarray:
    """weights -> the order the streaming GEMM consumes them.

    w_nk_taps: list over taps of W[N][K].  Returns [n_tile][tap][k_block][2*kb planes][ntile][8]
    flattened; one (tap, k_block) block is one bulk copy of ntile*2*kb*16 bytes."""
    n, k = w_nk_taps[0].shape
    assert n % ntile == 0 and k % (16 * kb) == 0, (n, k, ntile, kb)
    out = []
    for j in range(n // ntile):
        for w in w_nk_taps:
            planes = cp8(w[j * ntile:(j + 1) * ntile])            # [K/8][ntile][8]
            out.append(planes.reshape(-1))
    return np.concatenate(out)


def glu_rows(wv: np.ndarray, wg: np.ndarray, ntile: int) -> np.ndarray:
    """interleave value/gate output channels per n-tile: [val(Ct) | gate(Ct)] blocks."""
    ct = ntile // 2
    blocks = []
    for j in range(wv.shape[0] // ct):
        blocks += [wv[j * ct:(j + 1) * ct], wg[j * ct:(j + 1) * ct]]
    return np.concatenate(blocks)


def _glu_ep(blob: Blob, bv, bg, scale, shift, ntile):
    ct = ntile // 2
    rows = []
    for j in range(bv.size // ct):
        s = slice(j * ct, (j + 1) * ct)
        rows += [bv[s], bg[s], scale[s], shift[s]]
    blob.f["ep"] = np.concatenate(rows)


def gcrn_kb(cin: int) -> int:
    return 2 if (cin // 8) % 4 == 0 else 1


def pack_gcrn_conv1(sd) -> Blob:
    b = Blob()
    wv = _np(sd["conv1.conv1.weight"])[:, :, 0, :].reshape(16, 6)       # k = c*3 + df
    wg = _np(sd["conv1.conv2.weight"])[:, :, 0, :].reshape(16, 6)
    b.h["w"] = cp8(np.pad(np.concatenate([wv, wg]), ((0, 0), (0, 10))))  # [2][32][8]
    s, sh = bn_affine(sd, "bn1")
    _glu_ep(b, _np(sd["conv1.conv1.bias"]), _np(sd["conv1.conv2.bias"]), s, sh, 32)
    return b


def pack_gcrn_enc(sd, i: int) -> Blob:
    """conv{i}, i = 2..5: GluConv2d k=(1,3) stride (1,2) + BN (+ELU in the epilogue)."""
    b = Blob()
    cin, cout = GCRN_CH[i - 1], GCRN_CH[i]
    ntile = min(256, 2 * cout)
    wv, wg = _np(sd[f"conv{i}.conv1.weight"])[:, :, 0, :], _np(sd[f"conv{i}.conv2.weight"])[:, :, 0, :]
    b.h["w"] = _stream([glu_rows(wv[:, :, df], wg[:, :, df], ntile) for df in range(3)], ntile, gcrn_kb(cin))
    s, sh = bn_affine(sd, f"bn{i}")
    _glu_ep(b, _np(sd[f"conv{i}.conv1.bias"]), _np(sd[f"conv{i}.conv2.bias"]), s, sh, ntile)
    return b


def pack_gcrn_dec(sd, br: int, i: int) -> Blob:
    """conv{i}_t_{br}, i = 5..2: even outputs use df=0 (h[j]) and df=2 (h[j-1]); odd use df=1 (h[j])."""
    b = Blob()
    cin, cout, _, _ = GCRN_DEC[i]
    ntile = 2 * cout
    p = f"conv{i}_t_{br}"
    wv = _np(sd[p + ".conv1.weight"])[:, :, 0, :].transpose(1, 0, 2)    # [Cout][Cin][3]
    wg = _np(sd[p + ".conv2.weight"])[:, :, 0, :].transpose(1, 0, 2)
    kb = gcrn_kb(cin)
    b.h["w_even"] = _stream([glu_rows(wv[:, :, df], wg[:, :, df], ntile) for df in (0, 2)], ntile, kb)
    b.h["w_odd"] = _stream([glu_rows(wv[:, :, 1], wg[:, :, 1], ntile)], ntile, kb)
    s, sh = bn_affine(sd, f"bn{i}_t_{br}")
    _glu_ep(b, _np(sd[p + ".conv1.bias"]), _np(sd[p + ".conv2.bias"]), s, sh, ntile)
    return b


def lstm_row_order() -> np.ndarray:
    """gate-row order of the recurrence kernel: n = cta*128 + lane, lane = gate*32 + (unit % 32),
    unit = cta*32 + lane%32  ->  reference row gate*512 + unit (gate order i,f,g,o)."""
    n = np.arange(2048)
    cta, lane = n // 128, n % 128
    return (lane // 32) * 512 + cta * 32 + lane % 32


def lstm1_col_order() -> np.ndarray:
    """layer-1 K order: kk = f*128 + cl  ->  reference feature cl*4 + f (within the group)."""
    kk = np.arange(512)
    return (kk % 128) * 4 + kk // 128


def pack_gcrn_lstm(sd, layer: int, g: int) -> Blob:
    b = Blob()
    p = f"glstm.lstm_list{layer}.{g}"
    rows = lstm_row_order()
    w_ih = _np(sd[p + ".weight_ih_l0"])[rows]
    if layer == 1:
        w_ih = w_ih[:, lstm1_col_order()]
    b.h["w_ih"] = _stream([w_ih], 128, 2)                                # 16 n-tiles x 16 k-blocks of 4 planes
    b.h["w_hh"] = np.stack([cp8(_np(sd[p + ".weight_hh_l0"])[rows][c * 128:(c + 1) * 128]) for c in range(16)])
    b.f["bias"] = (_np(sd[p + ".bias_ih_l0"]) + _np(sd[p + ".bias_hh_l0"]))[rows]
    return b


def pack_gcrn_out(sd, br: int) -> Blob:
    """conv1_t_{br} (32 -> 1 GLU, k3 s2) + bn1_t + ELU + fc{br}, scaled by 1/11 (trainer :942)."""
    b = Blob()
    p = f"conv1_t_{br}"
    b.f["wv"] = _np(sd[p + ".conv1.weight"])[:, 0, 0, :].reshape(-1)     # [32][3]
    b.f["wg"] = _np(sd[p + ".conv2.weight"])[:, 0, 0, :].reshape(-1)
    s, sh = bn_affine(sd, f"bn1_t_{br}")
    b.f["misc"] = np.array([_np(sd[p + ".conv1.bias"])[0], _np(sd[p + ".conv2.bias"])[0], s[0], sh[0]])
    b.f["fcw"] = (_np(sd[f"fc{br}.weight"]).T / 11.0).reshape(-1)        # [f_in][f_out]
    b.f["fcb"] = _pad4(_np(sd[f"fc{br}.bias"]) / 11.0, 164)
    return b


def pack_gcrn(sd):
    out = {"conv1": pack_gcrn_conv1(sd)}
    for i in range(2, 6):
        out[f"conv{i}"] = pack_gcrn_enc(sd, i)
    for layer in (1, 2):
        for g in range(2):
            out[f"lstm{layer}_{g}"] = pack_gcrn_lstm(sd, layer, g)
    ln = Blob()
    for i in (1, 2):
        ln.f[f"w{i}"] = _np(sd[f"glstm.ln{i}.weight"])
        ln.f[f"b{i}"] = _np(sd[f"glstm.ln{i}.bias"])
    out["ln"] = ln
    for br in (1, 2):
        for i in range(5, 1, -1):
            out[f"dec{br}_{i}"] = pack_gcrn_dec(sd, br, i)
        out[f"out{br}"] = pack_gcrn_out(sd, br)
    return out


def diffunet_as_diffunet1(sd, out_scale: float = 1.0):
    """Express the time-independent prior DiffUNet (model/diff.py) as a DiffUNet1 with an identity Preprocess on
    the first input and all-zero time paths, so the very same kernels run it.  ``out_scale``
    folds the trainer's division by c = 11 (:942) into the last 1x1 convs."""
    from .weights import diffunet1_table
    out = OrderedDict()
    for key, shape, _, _ in diffunet1_table():
        if key in sd:
            out[key] = sd[key].detach().clone().float() if sd[key].is_floating_point() else sd[key].clone()
        else:
            out[key] = torch.zeros(shape, dtype=torch.float32)
    out["preprocess.conv.weight"][0, 0, 0, 0] = 1.0
    out["preprocess.conv.weight"][1, 1, 0, 0] = 1.0
    if out_scale != 1.0:
        for br in ("de_real", "de_imag"):
            out[f"{br}.de1.0.conv2.weight"] *= out_scale
            out[f"{br}.de1.0.conv2.bias"] *= out_scale
    return out
