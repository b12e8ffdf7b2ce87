"""Host side of the GCRN prior kernels (model/gcrn.py:136-166): packed weights on the device,
per-shape workspaces and the launch sequence.  The output is X_init = GCRN(y) / 11
(trainer/complex_ddpm_trainer.py:941-942)."""
from __future__ import annotations

import ctypes as C
from typing import Dict

import numpy as np
import torch

from . import lib as _lib
from . import pack as P

N_FREQ = 161
MAX_CHUNK = 64          # utterances per LSTM launch (h operand must fit beside W_hh in shared memory)


class GCRNEngine:
    def __init__(self, state_dict, device):
        self.lib = _lib.load(require_device=True)
        self.device = torch.device(device)
        # weights: packed on the host by the C ABI (pdse_pack_gcrn, csrc/pack.cpp) into ONE blob, uploaded once; every
        # section of its directory (pdse_pack_layout) is a typed view: "<block>.wb" / ".wf" / ".w_ih" / ".w_hh" / ".w_even" ...
        blob, directory = _lib.pack_state_dict(_lib.NET_GCRN, state_dict)
        self.blob = torch.from_numpy(blob).to(self.device)
        self.sec: Dict[str, torch.Tensor] = {}
        for name, (dtype, off, n) in directory.items():
            self.sec[name] = self.blob[off:off + (4 if dtype else 2) * n].view(torch.float32 if dtype else _lib.op_dtype())
        self.wb = {k[:-3]: v for k, v in self.sec.items() if k.endswith(".wb")}
        self.wf = {k[:-3]: v for k, v in self.sec.items() if k.endswith(".wf")}
        self._ws: Dict[tuple, Dict[str, torch.Tensor]] = {}
        self.timing = None     # list -> record (name, start_event, end_event) per launch

    def _timed(self, name, rc):
        """rc: zero-arg callable returning the ABI status"""
        if self.timing is None:
            _lib.check(rc())
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(rc())
        e1.record()
        self.timing.append((name, e0, e1))

    def _side_stream(self):
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=self.device)
        return self._side

    def _sub(self, name: str, key: str):
        return _lib.ptr(self.sec[f"{name}.{key}"])

    def workspace(self, B: int, T: int) -> Dict[str, torch.Tensor]:
        ws = self._ws.get((B, T))
        if ws is None:
            dev = self.device
            bf = dict(dtype=_lib.op_dtype(), device=dev)
            Bp = 32 if B <= 32 else 64      # batch columns of the recurrence MMA (N = 32 or 64)
            ws = {"Bp": Bp}
            for i in range(1, 5):
                c, f = P.GCRN_CH[i], P.GCRN_F[i]
                ws[f"e{i}_so"] = torch.zeros(B, c // 8, 2, T * ((f + 1) // 2), 8, **bf)
                ws[f"e{i}_ug"] = torch.zeros(B, c // 8, T * (f + 1) + 1, 8, **bf)
            ws["e5_ug"] = torch.zeros(B, 32, T * 5 + 1, 8, **bf)
            ws["lstm_ug"] = torch.zeros(B, 32, T * 5 + 1, 8, **bf)
            for g in range(2):
                ws[f"xl1_{g}"] = torch.zeros(64, T * B, 8, **bf)
                ws[f"xl2_{g}"] = torch.zeros(64, T * B, 8, **bf)
                ws[f"pre_{g}"] = torch.zeros(T, 2048, Bp, dtype=torch.float32, device=dev)
                ws[f"h_{g}"] = torch.zeros(T * B, 512, dtype=torch.float32, device=dev)
            ws["hbuf"] = torch.zeros(2, 2, 64, Bp, 8, **bf)
            ws["sync"] = torch.zeros(2, dtype=torch.int32, device=dev)
            for br in (1, 2):
                for i in range(5, 1, -1):
                    _, cout, _, fout = P.GCRN_DEC[i]
                    ws[f"d{br}_{i}"] = torch.zeros(B, cout // 8, T * (fout + 1) + 1, 8, **bf)
            self._ws[(B, T)] = ws
        return ws

    def forward(self, y: torch.Tensor, out: torch.Tensor = None, stream=None) -> torch.Tensor:
        """y [B,2,T,161] fp32 (compressed STFT) -> X_init [B,2,T,161] fp32 (already divided by 11)."""
        B, _, T, F = y.shape
        assert F == N_FREQ and y.is_contiguous() and y.dtype == torch.float32
        if out is None:
            out = torch.empty_like(y)
        for b0 in range(0, B, MAX_CHUNK):
            b1 = min(B, b0 + MAX_CHUNK)
            self._forward_chunk(y[b0:b1], out[b0:b1], stream)
        return out

    def _forward_chunk(self, y, out, stream):
        B, _, T, _ = y.shape
        L, ws, s = self.lib, self.workspace(B, T), _lib.stream_ptr(stream)
        p, chk = _lib.ptr, _lib.check
        Bp = ws["Bp"]
        self._timed("gcrn_conv1_fwd", lambda: L.pdse_gcrn_conv1_fwd(p(y), p(ws["e1_so"]), p(ws["e1_ug"]), p(self.wb["conv1"]), p(self.wf["conv1"]),
                                  B, T, s))
        for i in range(2, 6):
            cin, cout, fin = P.GCRN_CH[i - 1], P.GCRN_CH[i], P.GCRN_F[i - 1]
            last = i == 5
            self._timed("gcrn_enc_fwd", lambda: L.pdse_gcrn_enc_fwd(p(ws[f"e{i - 1}_so"]), None if last else p(ws[f"e{i}_so"]), p(ws[f"e{i}_ug"]),
                                    p(ws["xl1_0"]) if last else None, p(ws["xl1_1"]) if last else None,
                                    p(self.wb[f"conv{i}"]), p(self.wf[f"conv{i}"]), B, T, cin, cout, fin,
                                    0 if last else 1, s))
        for layer in (1, 2):
            for g in range(2):
                name = f"lstm{layer}_{g}"
                self._timed("lstm_inproj", lambda: L.pdse_lstm_inproj(p(ws[f"xl{layer}_{g}"]), self._sub(name, "w_ih"), p(self.wf[name]),
                                       p(ws[f"pre_{g}"]), B, Bp, T, s))
            self._timed("lstm_rec", lambda: L.pdse_lstm_rec(self._sub(f"lstm{layer}_0", "w_hh"), self._sub(f"lstm{layer}_1", "w_hh"),
                                p(ws["pre_0"]), p(ws["pre_1"]), p(ws["h_0"]), p(ws["h_1"]), p(ws["hbuf"]),
                                p(ws["sync"]), B, Bp, T, s))
            ln = self.wf["ln"]
            w = C.c_void_p(ln.data_ptr() + 4 * (0 if layer == 1 else 2048))
            b = C.c_void_p(ln.data_ptr() + 4 * (1024 if layer == 1 else 3072))
            self._timed("gcrn_ln", lambda: L.pdse_gcrn_ln(p(ws["h_0"]), p(ws["h_1"]), w, b, p(ws["xl2_0"]), p(ws["xl2_1"]), p(ws["lstm_ug"]),
                               B, T, layer, s))
        # the two decoder branches (gcrn.py:150-160) are independent chains of four launches: run them on two streams
        # (fork / join; captured into the graph as parallel branches) so each hides the other's load latencies
        cur = stream if stream is not None else torch.cuda.current_stream()
        fork = self.timing is None
        if fork:
            side = self._side_stream()
            side.wait_stream(cur)
        for br in (1, 2):
            sbr = _lib.stream_ptr(side) if (fork and br == 2) else s
            prev = ws["lstm_ug"]
            for i in range(5, 1, -1):
                cin, cout, fin, fout = P.GCRN_DEC[i]
                skip = ws["e5_ug"] if i == 5 else ws[f"e{i}_ug"]
                name = f"dec{br}_{i}"
                self._timed("gcrn_dec_fwd", lambda: L.pdse_gcrn_dec_fwd(p(prev), p(skip), p(ws[f"d{br}_{i}"]), self._sub(name, "w_even"),
                                        self._sub(name, "w_odd"), p(self.wf[name]), B, T, cin // 2, cin // 2, cout,
                                        fin, fout, sbr))
                prev = ws[f"d{br}_{i}"]
        if fork:
            cur.wait_stream(side)
        self._timed("gcrn_out_fwd", lambda: L.pdse_gcrn_out_fwd(p(ws["d1_2"]), p(ws["d2_2"]), p(ws["e1_ug"]), p(self.wf["out1"]), p(self.wf["out2"]),
                                p(out), B, T, s))
