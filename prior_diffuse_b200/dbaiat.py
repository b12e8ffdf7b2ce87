"""Host side of the DB-AIAT prior ``aia_complex_trans_ri`` (model/dbaiat.py:450-478): weight packing for
csrc/dbaiat.cu, per-shape workspaces and the launch sequence.  Output: X_init = prior(y) / 11
(trainer/complex_ddpm_trainer.py:941-942; the 1/11 is folded into the decoders' out_conv).

Layouts (csrc/dbaiat.cu header): a DenseBlock's growing ``torch.cat`` is one CP8 buffer of 8-plane groups
(group 0 = block input, group i+1 = conv{i+1} output); the transformer state is fp32 ``[B][T][80][32]``;
per-sequence tensors are sequence-major (row: n = b*T + t, col: n = b*80 + w)."""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict

import numpy as np
import torch

from . import lib as _lib
from .pack import _np, cp8

N_FREQ = 161
F2 = 80
LAYERS = 4
LN_DENSE, LN_ENC_OUT, LN_DEC_OUT, LN_IN = 0, 1, 2, 3
OUT_SCALE = 1.0 / 11.0


# ---------------------------------------------------------------------------------------------- packing
def _conv_taps(w: np.ndarray, groups: int) -> np.ndarray:
    """conv weight [N][64*groups][kh][3] -> [chunk][tap = kt*3 + kf][8][N][8].  Chunk g holds the channels of plane
    group g; torch.cat([out, skip]) puts the newest map first, so group g is reference block ``groups-1-g``."""
    n, cin, kh, kw = w.shape
    assert cin == 64 * groups and kw == 3
    out = np.zeros((groups, kh * 3, 8, n, 8))
    for g in range(groups):
        blk = w[:, (groups - 1 - g) * 64:(groups - g) * 64]
        for kt in range(kh):
            for kf in range(3):
                out[g, kt * 3 + kf] = cp8(blk[:, :, kt, kf])
    return out


def pack_dense_block(sd, p: str) -> Dict[str, np.ndarray]:
    d = {}
    for i in range(4):
        d[f"w{i}"] = _conv_taps(_np(sd[f"{p}.conv{i + 1}.weight"]), i + 1)
        d[f"b{i}"] = _np(sd[f"{p}.conv{i + 1}.bias"])
        d[f"g{i}"], d[f"be{i}"] = _np(sd[f"{p}.norm{i + 1}.weight"]), _np(sd[f"{p}.norm{i + 1}.bias"])
        d[f"s{i}"] = _np(sd[f"{p}.prelu{i + 1}.weight"])
    return d


def pack_aia_layer(sd, p: str) -> Dict[str, np.ndarray]:
    """TransformerEncoderLayer (dbaiat.py:41-88) -> the flat fp32 block of aia_attn_kernel, the GRU B operands
    (x | h rows against r | z | n_x | n_h columns) and the linear2 / norm2 vector of aia_post_kernel."""
    sc = math.log2(math.e) / math.sqrt(8.0)      # softmax in base 2: exp(x) = 2^(x log2 e)
    win, bi = _np(sd[p + ".self_attn.in_proj_weight"]).copy(), _np(sd[p + ".self_attn.in_proj_bias"]).copy()
    win[:32] *= sc
    bi[:32] *= sc
    attn = np.concatenate([_np(sd[p + ".norm3.weight"]), _np(sd[p + ".norm3.bias"]), win.T.reshape(-1), bi,
                           _np(sd[p + ".self_attn.out_proj.weight"]).T.reshape(-1), _np(sd[p + ".self_attn.out_proj.bias"]),
                           _np(sd[p + ".norm1.weight"]), _np(sd[p + ".norm1.bias"])])
    w2 = _np(sd[p + ".linear2.weight"])
    gw, gb = [], []
    for d, sfx in enumerate(("", "_reverse")):
        wih, whh = _np(sd[p + ".gru.weight_ih_l0" + sfx]), _np(sd[p + ".gru.weight_hh_l0" + sfx])
        bih, bhh = _np(sd[p + ".gru.bias_ih_l0" + sfx]), _np(sd[p + ".gru.bias_hh_l0" + sfx])
        m = np.zeros((256, 96))
        m[:128, :32], m[:128, 32:] = wih[:128], whh[:128]          # r, z: both operands
        m[128:192, :32] = wih[128:]                                 # n_x = W_in x
        m[192:, 32:] = whh[128:]                                    # n_h = W_hn h
        gw.append(np.concatenate([cp8(m).reshape(-1), cp8(w2[:, d * 64:(d + 1) * 64]).reshape(-1)]))
        gb.append(np.concatenate([bih[:128] + bhh[:128], bih[128:], bhh[128:]]))
    post = np.concatenate([_np(sd[p + ".linear2.bias"]), _np(sd[p + ".norm2.weight"]), _np(sd[p + ".norm2.bias"])])
    return {"attn": attn, "gru_w": np.concatenate(gw), "gru_b": np.concatenate(gb), "post": post}


def pack_dbaiat(sd) -> Dict[str, np.ndarray]:
    """arrays whose name starts with ``h:`` go to the device as fp16 (the operand format of csrc/dbaiat.cu), the rest as fp32"""
    out: Dict[str, np.ndarray] = {}
    e = "en_ri"
    out["in_cw"] = np.concatenate([_np(sd[e + ".inp_conv.weight"]).reshape(64, 2).reshape(-1), _np(sd[e + ".inp_conv.bias"])])
    out["in_g"], out["in_b"], out["in_s"] = (_np(sd[e + ".inp_norm.weight"]), _np(sd[e + ".inp_norm.bias"]),
                                             _np(sd[e + ".inp_prelu.weight"]))
    for k, v in pack_dense_block(sd, e + ".enc_dense1").items():
        out[("h:" if k[0] == "w" else "") + "enc_" + k] = v
    out["h:enc_out_w"] = _conv_taps(_np(sd[e + ".enc_conv1.weight"]), 1)
    out["enc_out_b"] = _np(sd[e + ".enc_conv1.bias"])
    out["enc_out_g"], out["enc_out_be"], out["enc_out_s"] = (_np(sd[e + ".enc_norm1.weight"]), _np(sd[e + ".enc_norm1.bias"]),
                                                             _np(sd[e + ".enc_prelu1.weight"]))
    t = "dual_trans"
    out["enc_out_cw"] = np.concatenate([_np(sd[t + ".input.0.weight"]).reshape(32, 64).T.reshape(-1), _np(sd[t + ".input.0.bias"]),
                                        _np(sd[t + ".input.1.weight"])])
    for kind in ("row", "col"):
        for i in range(LAYERS):
            for k, v in pack_aia_layer(sd, f"{t}.{kind}_trans.{i}").items():
                out[("h:" if k == "gru_w" else "") + f"{kind}{i}_{k}"] = v
    for i in range(LAYERS):
        out[f"combine{i}"] = np.concatenate([
            _np(sd[t + ".k1"]), _np(sd[t + ".k2"]), _np(sd[f"{t}.row_norm.{i}.weight"]), _np(sd[f"{t}.row_norm.{i}.bias"]),
            _np(sd[f"{t}.col_norm.{i}.weight"]), _np(sd[f"{t}.col_norm.{i}.bias"]), _np(sd[t + ".output.0.weight"]), np.zeros(1),
            _np(sd[t + ".output.1.weight"]).reshape(64, 32).T.reshape(-1), _np(sd[t + ".output.1.bias"])])
    out["aham"] = np.concatenate([_np(sd["aham.conv1.weight"]).reshape(-1), _np(sd["aham.conv1.bias"])])
    for d, de in enumerate(("de1", "de2")):
        for k, v in pack_dense_block(sd, de + ".dec_dense1").items():
            out[("h:" if k[0] == "w" else "") + f"dec{d}_" + k] = v
        out[f"h:dec{d}_out_w"] = _conv_taps(_np(sd[de + ".dec_conv1.conv.weight"]), 1)
        out[f"dec{d}_out_b"] = _np(sd[de + ".dec_conv1.conv.bias"])
        out[f"dec{d}_out_g"], out[f"dec{d}_out_be"], out[f"dec{d}_out_s"] = (
            _np(sd[de + ".dec_norm1.weight"]), _np(sd[de + ".dec_norm1.bias"]), _np(sd[de + ".dec_prelu1.weight"]))
        out[f"dec{d}_out_cw"] = np.concatenate([_np(sd[de + ".out_conv.weight"]).reshape(-1),
                                                _np(sd[de + ".out_conv.bias"])]) * OUT_SCALE
    return out


# ---------------------------------------------------------------------------------------------- engine
class DBAIATEngine:
    def __init__(self, state_dict, device):
        self.lib = _lib.load(require_device=True)
        self.device = torch.device(device)
        self.hg = int(self.lib.pdse_db_guard_frames())
        self.w: Dict[str, torch.Tensor] = {}
        for name, arr in pack_dbaiat(state_dict).items():
            t = torch.from_numpy(np.ascontiguousarray(arr.reshape(-1), dtype=np.float32)).to(self.device)
            if name.startswith("h:"):
                if not bool(torch.isfinite(t).all()) or float(t.abs().max()) > 6.0e4:
                    raise ValueError(f"DB-AIAT weight block {name[2:]} does not fit the fp16 operand format (|w| must stay below 6e4)")
                self.w[name[2:]] = t.to(torch.float16).contiguous()
            else:
                self.w[name] = t.contiguous()
        self._ws: Dict[tuple, Dict[str, torch.Tensor]] = {}
        self.timing = None

    def _timed(self, name, rc):
        if self.timing is None:
            _lib.check(rc())
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(rc())
        e1.record()
        self.timing.append((name, e0, e1))

    def workspace(self, B: int, T: int) -> Dict[str, torch.Tensor]:
        ws = self._ws.get((B, T))
        if ws is None:
            dev = self.device
            bf = dict(dtype=torch.float16, device=dev)     # 16-bit operand planes (fp16)
            f32 = dict(dtype=torch.float32, device=dev)
            rows_e = (T + self.hg) * (N_FREQ + 1) + 1
            rows_d = (T + self.hg) * (F2 + 1) + 1
            npos = T * F2
            ws = {
                "ebuf": torch.zeros(B, 40, rows_e, 8, **bf),                 # guards / lead-in frames stay zero
                "xbuf": torch.zeros(B, 8, rows_d, 8, **bf),
                "dbuf0": torch.zeros(B, 32, rows_d, 8, **bf), "dbuf1": torch.zeros(B, 32, rows_d, 8, **bf),
                "pre": torch.zeros(B, T * (N_FREQ + 1), 64, **f32),
                "S": torch.zeros(B, npos, 32, **f32),
                "acc": torch.zeros(LAYERS * 2 * B * 2 + LAYERS * B * 64, dtype=torch.float64, device=dev),
            }
            for k in ("r", "c"):
                nseq, L = (B * T, F2) if k == "r" else (B * F2, T)
                ws["Y1" + k] = torch.zeros(B, npos, 32, **f32)
                ws["P" + k] = torch.zeros(2, B, npos, 32, **f32)
                ws["Z" + k] = torch.zeros(B, npos, 32, **f32)
                ws["XG" + k] = torch.zeros((nseq + 127) // 128, L, 4, 128, 8, **bf)   # tail-group rows stay zero
            for i in range(LAYERS):
                ws[f"O{i}"] = torch.zeros(B, npos, 64, **bf)
            self._ws[(B, T)] = ws
        return ws

    # ------------------------------------------------------------------ pieces
    def _dense_block(self, tag, own, own_ppb, own_first, xsrc, x_ppb, B, T, F, s):
        """four (conv -> LayerNorm -> PReLU) layers; layer i reads plane groups 0..i and writes group i+1.
        ``xsrc`` holds group 0; groups >= 1 live in ``own`` starting at plane ``own_first``."""
        L, p, ws = self.lib, _lib.ptr, self._cur
        for i in range(4):
            src = (C.c_int * 4)(*([1] + [0] * 3))
            plane = (C.c_int * 4)(*([0] + [own_first + 8 * g for g in range(3)]))
            self._timed(f"{tag}_conv{i}", lambda: L.pdse_db_conv_fwd(
                p(own), p(xsrc), own_ppb, x_ppb, src, plane, i + 1, B, T, F + 1, 2 ** i, 1, p(self.w[f"{tag}_w{i}"]),
                p(self.w[f"{tag}_b{i}"]), 64, p(ws["pre"]), s))
            self._timed(f"{tag}_ln{i}", lambda: L.pdse_db_ln_fwd(
                LN_DENSE, p(ws["pre"]), None, p(self.w[f"{tag}_g{i}"]), p(self.w[f"{tag}_be{i}"]), p(self.w[f"{tag}_s{i}"]),
                None, p(own), own_ppb, own_first + 8 * i, None, 0, B, T, F + 1, F, s))

    def forward(self, y: torch.Tensor, out: torch.Tensor = None, stream=None, upto: str = None) -> torch.Tensor:
        """y [B,2,T,161] fp32 (compressed STFT) -> X_init [B,2,T,161] fp32 (already divided by 11)."""
        B, _, T, F = y.shape
        assert F == N_FREQ and y.is_contiguous() and y.dtype == torch.float32
        if out is None:
            out = torch.empty_like(y)
        L, ws, s = self.lib, self.workspace(B, T), _lib.stream_ptr(stream)
        self._cur = ws
        p, w = _lib.ptr, self.w
        run = self._timed
        eb = ws["ebuf"]
        # ---- dense_encoder (dbaiat.py:497-501); ebuf: group 0 = inp, groups 1..4 = dense outputs
        run("enc_in", lambda: L.pdse_db_ln_fwd(LN_IN, None, p(y), p(w["in_g"]), p(w["in_b"]), p(w["in_s"]), p(w["in_cw"]),
                                               p(eb), 40, 0, None, 0, B, T, N_FREQ + 1, N_FREQ, s))
        self._dense_block("enc", eb, 40, 8, eb, 40, B, T, N_FREQ, s)
        if upto == "enc_dense":
            return None
        src = (C.c_int * 4)(0, 0, 0, 0)
        plane = (C.c_int * 4)(32, 0, 0, 0)
        run("enc_out_conv", lambda: L.pdse_db_conv_fwd(p(eb), None, 40, 0, src, plane, 1, B, T, N_FREQ + 1, 0, 0,
                                                       p(w["enc_out_w"]), p(w["enc_out_b"]), 64, p(ws["pre"]), s))
        run("enc_out_ln", lambda: L.pdse_db_ln_fwd(LN_ENC_OUT, p(ws["pre"]), None, p(w["enc_out_g"]), p(w["enc_out_be"]),
                                                   p(w["enc_out_s"]), p(w["enc_out_cw"]), None, 0, 0, p(ws["S"]), 0, B, T,
                                                   N_FREQ + 1, F2, s))
        if upto == "enc":
            return None
        # ---- AIA_Transformer (dbaiat.py:136-154)
        ws["acc"].zero_()
        acc = ws["acc"]
        npos = T * F2

        def stats(i, k):
            return C.c_void_p(acc.data_ptr() + 8 * ((i * 2 + k) * B * 2))

        def pool(i):
            return C.c_void_p(acc.data_ptr() + 8 * (LAYERS * 2 * B * 2 + i * B * 64))

        for i in range(LAYERS):
            for k, kind in enumerate(("row", "col")):
                sfx = kind[0]
                nseq, Ls = (B * T, F2) if k == 0 else (B * F2, T)
                run(f"attn_{kind}", lambda: L.pdse_aia_attn_fwd(p(ws["S"]), p(w[f"{kind}{i}_attn"]), p(ws["Y1" + sfx]),
                                                                p(ws["XG" + sfx]), B, T, 1 - k, s))
                run(f"gru_{kind}", lambda: L.pdse_aia_gru_fwd(p(ws["XG" + sfx]), p(w[f"{kind}{i}_gru_w"]), p(w[f"{kind}{i}_gru_b"]),
                                                              p(ws["P" + sfx]), Ls, nseq, s))
                run(f"post_{kind}", lambda: L.pdse_aia_post_fwd(p(ws["Y1" + sfx]), p(ws["P" + sfx][0]), p(ws["P" + sfx][1]),
                                                                p(w[f"{kind}{i}_post"]), p(ws["Z" + sfx]), stats(i, k), B, npos, s))
            run("combine", lambda: L.pdse_aia_combine_fwd(p(ws["S"]), p(ws["Zr"]), p(ws["Zc"]), stats(i, 0), stats(i, 1),
                                                          p(w[f"combine{i}"]), p(ws[f"O{i}"]), pool(i), B, T, s))
            if upto == f"aia{i}":
                return None
        # ---- AHAM (dbaiat.py:268-288) -> decoder input planes
        run("aham", lambda: L.pdse_aia_aham_fwd(p(ws["O0"]), p(ws["O1"]), p(ws["O2"]), p(ws["O3"]), pool(0), p(w["aham"]),
                                                p(ws["xbuf"]), B, T, s))
        if upto == "aham":
            return None
        # ---- two dense_decoders (dbaiat.py:541-548)
        for d in range(2):
            db = ws[f"dbuf{d}"]
            self._dense_block(f"dec{d}", db, 32, 0, ws["xbuf"], 8, B, T, F2, s)
            src = (C.c_int * 4)(0, 0, 0, 0)
            plane = (C.c_int * 4)(24, 0, 0, 0)
            run("dec_out_conv", lambda: L.pdse_db_conv_fwd(p(db), None, 32, 0, src, plane, 1, B, T, F2 + 1, 0, 1,
                                                           p(w[f"dec{d}_out_w"]), p(w[f"dec{d}_out_b"]), 128, p(ws["pre"]), s))
            run("dec_out_ln", lambda: L.pdse_db_ln_fwd(LN_DEC_OUT, p(ws["pre"]), None, p(w[f"dec{d}_out_g"]), p(w[f"dec{d}_out_be"]),
                                                       p(w[f"dec{d}_out_s"]), p(w[f"dec{d}_out_cw"]), None, 0, 0, p(out), d, B, T,
                                                       F2 + 1, N_FREQ, s))
        return out
