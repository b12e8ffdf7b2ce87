"""Host side of ``diff2.DiffWave`` (model/diff2.py:12-158): weight packing, workspaces and the launch
sequence of one evaluation -- embedding, input projection, one fused kernel per residual layer, skip / output projections
(csrc/diffwave.cu).  The module wrapper is ``modules.DiffWave``."""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict

import numpy as np
import torch

from . import lib as _lib
from .pack import _np, bias_block, cp8

CHANNELS = 64


def _bf16(a: np.ndarray, device) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32).reshape(-1)).to(device).to(_lib.op_dtype()).contiguous()


def pack_diffwave(sd, cycle: int):
    """per layer one bf16 blob: W_cat [48 planes][128][8] (K = (phase, tap, channel): phase 0 = dilated_conv on x + e,
    phase 1 = conditioner_projection on cond; taps t-d, t, t+d) | W_o [8][128][8] | bias block of the conv pair | bias
    block of the output projection"""
    layers = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("residual_layers."))
    if sd["input_projection.weight"].shape[0] != CHANNELS:
        raise ValueError("the DiffWave kernels are built for residual_channels = 64")
    blobs, rows, rbias = [], [], []
    for i in range(layers):
        p = f"residual_layers.{i}"
        wd, wc = _np(sd[p + ".dilated_conv.weight"]), _np(sd[p + ".conditioner_projection.weight"])      # [128][64][3]
        wcat = np.concatenate([w[:, :, tap] for w in (wd, wc) for tap in range(3)], axis=1)                # [128][384]
        wo = _np(sd[p + ".output_projection.weight"])[:, :, 0]                                             # [128][64]
        blobs.append(np.concatenate([cp8(wcat).reshape(-1), cp8(wo).reshape(-1),
                                     bias_block(_np(sd[p + ".dilated_conv.bias"]) + _np(sd[p + ".conditioner_projection.bias"])).reshape(-1),
                                     bias_block(_np(sd[p + ".output_projection.bias"])).reshape(-1)]))
        rows.append(_np(sd[p + ".diffusion_projection.weight"]))
        rbias.append(_np(sd[p + ".diffusion_projection.bias"]))
    post = np.concatenate([cp8(_np(sd["skip_projection.weight"])[:, :, 0]).reshape(-1),
                           bias_block(_np(sd["skip_projection.bias"])).reshape(-1)])
    return {"layers": layers, "dilations": [2 ** (i % cycle) for i in range(layers)], "blobs": blobs,
            "rows": np.concatenate(rows), "rbias": np.concatenate(rbias), "post": post,
            "win": np.concatenate([_np(sd["input_projection.weight"]).reshape(-1), _np(sd["input_projection.bias"])]),
            "wout": np.concatenate([_np(sd["output_projection.weight"]).reshape(-1), _np(sd["output_projection.bias"])])}


class DiffWaveEngine:
    def __init__(self, state_dict, device, cycle: int = 10):
        self.lib = _lib.load(require_device=True)
        self.device = torch.device(device)
        pk = pack_diffwave(state_dict, cycle)
        self.layers, self.dil = pk["layers"], pk["dilations"]
        if max(self.dil) > 512:
            raise ValueError("dilations above 512 are not supported (guard rows)")
        f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(self.device)   # noqa: E731
        self.wb = [_bf16(b, self.device) for b in pk["blobs"]]
        self.post = _bf16(pk["post"], self.device)
        self.rows, self.rbias, self.win, self.wout = f32(pk["rows"]), f32(pk["rbias"]), f32(pk["win"]), f32(pk["wout"])
        # DiffusionEmbedding (:71-95): the same sinusoid table and two SiLU linears as diff3.TimeEmbedding
        arg = torch.arange(50).unsqueeze(1) * 10.0 ** (torch.arange(64).unsqueeze(0) * 4.0 / 63.0)
        self.table = torch.cat([torch.sin(arg), torch.cos(arg)], dim=1).contiguous().to(self.device)
        self.emb = {k: state_dict[f"diffusion_embedding.{k}"].detach().float().contiguous().to(self.device)
                    for k in ("projection1.weight", "projection1.bias", "projection2.weight", "projection2.bias")}
        self.guard = self.lib.pdse_dw_guard_rows()
        self._ws: Dict[tuple, Dict[str, torch.Tensor]] = {}
        self.timing = None

    def workspace(self, B: int, L: int):
        ws = self._ws.get((B, L))
        if ws is None:
            dev, Lg = self.device, L + 2 * self.guard
            ws = {"x": torch.zeros(B, 16, L, 4, device=dev), "skip": torch.zeros(B, 16, L, 4, device=dev),
                  # operand planes with zero guard rows (= the convolutions' zero padding; never written)
                  "y0": torch.zeros(B, 8, Lg, 8, dtype=_lib.op_dtype(), device=dev),
                  "y1": torch.zeros(B, 8, Lg, 8, dtype=_lib.op_dtype(), device=dev),
                  "cond": torch.zeros(B, 8, Lg, 8, dtype=_lib.op_dtype(), device=dev),
                  "dtab": torch.zeros(B, self.layers * CHANNELS, device=dev)}
            self._ws[(B, L)] = ws
        return ws

    def forward(self, audio: torch.Tensor, audio_init: torch.Tensor, t: torch.Tensor, stream=None) -> torch.Tensor:
        """audio, audio_init [B, L] fp32 on the device; t [B] or [1] (integral or fractional step) -> [B, 1, L]"""
        B, L = audio.shape
        assert audio_init.shape == (B, L) and audio.is_contiguous() and audio_init.is_contiguous()
        Lh, p, chk, s = self.lib, _lib.ptr, _lib.check, _lib.stream_ptr(stream)
        ws = self.workspace(B, L)
        tt = t.to(device=self.device, dtype=torch.float32).reshape(-1)
        if tt.numel() == 1:
            tt = tt.expand(B)
        tt = tt.contiguous()
        n_rows = self.layers * CHANNELS
        run = self._timed
        e = self.emb
        run("dw_embed", lambda: Lh.pdse_dw_embed(p(tt), B, p(self.table), p(e["projection1.weight"]), p(e["projection1.bias"]),
                                                 p(e["projection2.weight"]), p(e["projection2.bias"]), p(self.rows), p(self.rbias),
                                                 n_rows, p(ws["dtab"]), s))
        run("dw_pre", lambda: Lh.pdse_dw_pre_fwd(p(audio), p(audio_init), p(self.win), p(ws["dtab"]), n_rows, p(ws["x"]), p(ws["y0"]),
                                                 p(ws["cond"]), B, L, s))
        for i in range(self.layers):
            y_in, y_out = (ws["y0"], ws["y1"]) if i % 2 == 0 else (ws["y1"], ws["y0"])
            last = i == self.layers - 1
            dnext = None if last else C.c_void_p(ws["dtab"].data_ptr() + 4 * (i + 1) * CHANNELS)
            run("dw_layer", lambda: Lh.pdse_dw_layer_fwd(p(y_in), None if last else p(y_out), p(ws["cond"]), p(ws["x"]), p(ws["skip"]),
                                                         p(self.wb[i]), dnext, n_rows, B, L, self.dil[i], int(i == 0), int(last), s))
        out = torch.empty(B, 1, L, device=self.device)
        run("dw_post", lambda: Lh.pdse_dw_post_fwd(p(ws["skip"]), p(self.post), p(self.wout), 1.0 / math.sqrt(self.layers), p(out), B, L, s))
        return out

    def _timed(self, name, rc_fn):
        if self.timing is None:
            _lib.check(rc_fn())
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(rc_fn())
        e1.record()
        self.timing.append((name, e0, e1))


class DiffWaveSampler:
    """The trainer's reverse loop (trainer/complex_ddpm_trainer.py:967-992: eps = D(x, cond, t_n); x = c1 (x - c2 eps), the
    noise coefficient being 0 there) around ``diff2.DiffWave`` on waveforms, conditioned on the noisy waveform -- the
    "dilated Conv1d residual stack" denoiser north_star describes, as ONE CUDA graph per (B, L).  The reference defines
    the network (model/diff2.py) but never wires it into a loop, so the schedule, the update and the
    x_T = N(0, I) start are the ones of its DiffUNet1 path (``pipeline.inference_schedule``)."""

    def __init__(self, state_dict, device="cuda:0", fast_sampling: bool = True, cycle: int = 10, seed: int = 7, rank: int = 0):
        from .pipeline import inference_schedule
        self.engine = DiffWaveEngine(state_dict, device, cycle)
        self.device = self.engine.device
        self.lib = self.engine.lib
        alpha, beta, alpha_cum, _, T = inference_schedule(fast_sampling)
        self.n_steps = len(alpha)
        self.c1 = [float(1.0 / alpha[n] ** 0.5) for n in range(self.n_steps)]
        self.c2 = [float(beta[n] / (1.0 - alpha_cum[n]) ** 0.5) for n in range(self.n_steps)]
        self.t = [torch.full((1,), float(T[n]), device=self.device) for n in range(self.n_steps)]
        self.seed, self.rank, self._offset = seed, rank, 0
        self._plans: Dict[tuple, dict] = {}

    def _loop(self, p):
        Lh, ptr, s = self.lib, _lib.ptr, _lib.stream_ptr()
        B, L = p["cond"].shape
        n = B * L
        for k in range(self.n_steps - 1, -1, -1):
            eps = self.engine.forward(p["x"][:n].view(B, L), p["cond"], self.t[k].expand(B))
            _lib.check(Lh.pdse_ddpm_update_f32(ptr(p["x"]), ptr(eps), None, None, None, n, 0, self.c1[k], self.c2[k], 0.0, 0, 0, 1.0, 0, 0, s))

    def enhance(self, noisy: torch.Tensor, x_T: torch.Tensor = None) -> torch.Tensor:
        """noisy [B, L] fp32 on the device -> enhanced [B, L] (a view of a static buffer, valid until the next call)"""
        B, L = noisy.shape
        n = B * L
        with torch.cuda.device(self.device):
            p = self._plans.get((B, L))
            if p is None:
                p = {"x": torch.zeros((n + 3) // 4 * 4, device=self.device), "cond": torch.zeros(B, L, device=self.device), "graph": None}
                self._plans[(B, L)] = p
            p["cond"].copy_(noisy)
            if x_T is not None:
                p["x"][:n].copy_(x_T.reshape(-1))
            else:      # x_T ~ N(0, I): the library's Philox stream keyed by (seed, rank), counter = running element offset
                _lib.check(self.lib.pdse_init_state_f32(_lib.ptr(p["x"]), None, None, n, 0, 1, self.seed + 7919 * self.rank, self._offset,
                                                        _lib.stream_ptr()))
                self._offset += (n + 3) // 4 * 4
            if p["graph"] is None:
                self._loop(p)                       # warm-up: workspaces, shared-memory attributes
                torch.cuda.synchronize()
                if x_T is not None:
                    p["x"][:n].copy_(x_T.reshape(-1))
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._loop(p)
                p["graph"] = g
                if x_T is None:                     # the warm-up consumed the drawn state: draw again for the captured run
                    _lib.check(self.lib.pdse_init_state_f32(_lib.ptr(p["x"]), None, None, n, 0, 1, self.seed + 7919 * self.rank,
                                                            self._offset, _lib.stream_ptr()))
                    self._offset += (n + 3) // 4 * 4
            p["graph"].replay()
            return p["x"][:n].view(B, L)
