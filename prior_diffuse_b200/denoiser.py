"""Host side of the DiffUNet1 kernels: packed weights on the device, per-shape workspaces,
and the launch sequence of one denoiser evaluation (model/diff3.py:37-57).

``DenoiserEngine`` is what the reverse loop drives; ``modules.DiffUNet1`` wraps it behind the
reference's ``nn.Module`` signature.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import os

import torch

from . import lib as _lib
from . import pack as P
from .weights import TCM_DILATIONS

N_FREQ = 161


def _enc_nt(Fin: int) -> int:
    """time rows per encoder tile: two 128-row sub-tiles (one per warpgroup chain)"""
    return max(1, 256 // ((Fin + 1) // 2))


def _dec_nt(Fin: int, kw: int) -> int:
    """time rows per decoder tile: three 128-row M-tiles (one per consumer warpgroup)"""
    return max(1, 384 // (Fin + (kw - 1) // 2))


class DenoiserEngine:
    def __init__(self, state_dict, device):
        self.lib = _lib.load(require_device=True)
        self.device = torch.device(device)
        # weights: packed on the host by the C ABI (pdse_pack_diffunet1, csrc/pack.cpp) into ONE blob, uploaded once;
        # wb / wf / time are typed views of its sections (pdse_pack_layout)
        blob, directory = _lib.pack_state_dict(_lib.NET_DIFFUNET1, state_dict)
        self.blob = torch.from_numpy(blob).to(self.device)
        self.wb: Dict[str, torch.Tensor] = {}
        self.wf: Dict[str, torch.Tensor] = {}
        self.time: Dict[str, torch.Tensor] = {}
        for name, (dtype, off, n) in directory.items():
            view = self.blob[off:off + (4 if dtype else 2) * n].view(torch.float32 if dtype else _lib.op_dtype())
            block, kind = name.rsplit(".", 1)
            if block == "time":
                self.time[kind] = view
            elif kind == "wb":
                self.wb[block] = view
            else:
                self.wf[block] = view
        expect = {"enc1": (9216, 16), "enc": (31744, 4), "dec": (33792, 4), "dec_last": (47104, 36),
                  "tcm": (80896, 388)}
        for name in self.wb:
            kind = "enc1" if name == "enc1" else "enc" if name.startswith("enc") else "tcm" if name.startswith("tcm") \
                else "dec_last" if name.endswith("_1") else "dec"
            got = (self.wb[name].numel(), self.wf[name].numel())
            assert got == expect[kind], (name, got, expect[kind])
        # device table of {bf16 blob, fp32 blob} pointers of the 18 residual blocks for the persistent TCM kernel
        self.tcm_table = torch.tensor([[self.wb[f"tcm{k}"].data_ptr(), self.wf[f"tcm{k}"].data_ptr()] for k in range(18)],
                                      dtype=torch.int64, device=self.device)
        self.tcm_dil = (C.c_int * 18)(*TCM_DILATIONS)
        self.tcm_persistent = True     # False: one launch per residual-block boundary (19 launches)
        # decoder blocks that take the split path (1x1 conv to HBM, then conv + tail from it) instead of the fused
        # kernel: measured faster only for de1 (0.36 vs 0.42 ms at 64 x 3 s); True / False force it for all / none
        self.dec_split = {"1": True, "0": False}.get(os.environ.get("PDSE_DEC_SPLIT", ""), (1,))
        self._ws: Dict[tuple, Dict[str, torch.Tensor]] = {}
        self.timing = None     # set to a list to record (name, start_event, end_event) per launch (eager runs)
        # sticky kernel-side status {code, launch, tile, count}: written by the persistent TCM kernel when a dependency
        # wait times out, never cleared by the library; check_status() reads it at a synchronisation point
        self.status = torch.zeros(8, dtype=torch.int32, device=self.device)
        self._status_host = torch.zeros(8, dtype=torch.int32).pin_memory()
        self._status_event = None

    def set_wait_timeout_us(self, us: int):
        """wall-clock bound of one dependency wait of the persistent TCM kernel: 0 = 2 s (default), < 0 = fail on the first
        unsatisfied poll (forces the error path in tests).  Lives in the status block on the device: graphs follow it."""
        self.status[4] = int(us)

    def post_status(self):
        """enqueue an asynchronous copy of the status word to the host on the current stream (no synchronisation);
        poll_status() / check_status(synchronize=False) look at it once the stream has got there"""
        self._status_host.copy_(self.status, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._status_event = ev

    def poll_status(self):
        """non-blocking: raise if an already finished call recorded a kernel-side error"""
        if self._status_event is not None and self._status_event.query():
            self._status_event = None
            self.check_status(synchronize=False)

    def check_status(self, synchronize: bool = True):
        """Raise if a kernel recorded an error since the last check (synchronises the current stream unless the caller
        already has).  The status word is cleared after it has been reported, so the engine stays usable."""
        if synchronize:
            self._status_host.copy_(self.status, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        if int(self._status_host[0]) != 0:
            rc = self.lib.pdse_status_check(C.c_void_p(self._status_host.data_ptr()))
            msg = self.lib.pdse_last_error().decode() if rc != 0 else "kernel-side error"
            self.status[:4].zero_()
            self._status_host.zero_()
            raise RuntimeError("libpdse: " + msg)

    def _timed(self, name, rc_fn):
        if self.timing is None:
            _lib.check(rc_fn())
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(rc_fn())
        e1.record()
        self.timing.append((name, e0, e1))

    # ------------------------------------------------------------------ workspaces
    def _dec_split(self, i: int) -> bool:
        return self.dec_split if isinstance(self.dec_split, bool) else i in self.dec_split

    def kernel_launches(self) -> int:
        """kernels per evaluation: 5 encoder blocks, the TCM stack (1 persistent kernel or 19), 5 decoder blocks
        (+ 1 for every block on the split path)"""
        return 5 + (1 if self.tcm_persistent else 19) + 5 + sum(self._dec_split(i) for i in range(1, 6))

    def _dec_h(self, ws, i: int, B: int, T: int):
        """split decoder path: workspace of the 1x1 conv output of block i, unsplit guarded layout (allocated zeroed on
        first use, outside graph capture: the guard slots are never written)"""
        if not self._dec_split(i):
            return None
        if f"h_{i}" not in ws:
            G = 2 if i == 1 else 1
            ws[f"h_{i}"] = torch.zeros(B, 2, 4, (T + 1) * (P.ENC_F[i] + G) + G, 8, dtype=_lib.op_dtype(), device=self.device)
        return _lib.ptr(ws[f"h_{i}"])

    def workspace(self, B: int, T: int) -> Dict[str, torch.Tensor]:
        key = (B, T)
        ws = self._ws.get(key)
        if ws is None:
            dev = self.device

            def cp8(npos):   # zero-initialised: dummy slots / guards are never written
                return torch.zeros(B, 8, npos, 8, dtype=_lib.op_dtype(), device=dev)

            ws = {}
            for i in range(1, 6):
                ws[f"e{i}"] = cp8(T * 2 * ((P.ENC_F[i] + 1) // 2))
            ws["x"] = torch.zeros(B, 32, T, 8, dtype=torch.float32, device=dev)
            for n in ("am0", "ak0", "am1", "ak1"):
                ws[n] = cp8(T)
            ws["dec_in"] = cp8(T * 4)
            ws["tcm_flags"] = torch.zeros(32 + 19 * B * ((T + 127) // 128), dtype=torch.int32, device=dev)
            for br in (0, 1):
                for i in range(5, 1, -1):
                    fo = 2 * P.ENC_F[i] + 1
                    ws[f"d{br}_{i}"] = cp8(T * 2 * ((fo + 1) // 2))
            n = B * 2 * T * N_FREQ
            ws["eps"] = torch.zeros((n + 3) // 4 * 4, dtype=torch.float32, device=dev)
            self._ws[key] = ws
        return ws

    # ------------------------------------------------------------------ time bias table
    def time_bias(self, t: torch.Tensor, out: Optional[torch.Tensor] = None, stream=None) -> torch.Tensor:
        """t [n] (float32 / int64) -> bias rows [n][452] (diff3.py:39 + every block's tp)."""
        t = t.to(device=self.device, dtype=torch.float32).contiguous()
        n = t.numel()
        if out is None:
            out = torch.empty(n, P.N_BIAS_ROW, dtype=torch.float32, device=self.device)
        tm = self.time
        _lib.check(self.lib.pdse_time_embed(_lib.ptr(t), n, _lib.ptr(tm["table"]), _lib.ptr(tm["p1w"]),
                                            _lib.ptr(tm["p1b"]), _lib.ptr(tm["p2w"]), _lib.ptr(tm["p2b"]),
                                            _lib.ptr(tm["rows"]), _lib.ptr(tm["bias"]), _lib.ptr(out),
                                            _lib.stream_ptr(stream)))
        return out

    # ------------------------------------------------------------------ one evaluation
    def forward(self, x: torch.Tensor, x0: torch.Tensor, bias_rows: torch.Tensor, bias_stride: int,
                stream=None, upto: Optional[str] = None, lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
        """eps = D(x, x0, t).  x, x0: [B,2,T,161] fp32 contiguous on the device; ``bias_rows`` from
        time_bias() (row b*bias_stride).  Returns a VIEW of the workspace eps buffer [B,2,T,161].

        ``lengths`` (int32[B] on the device, sample counts): zero-padded ragged batch.  The encoder and decoder blocks are
        causal in time, the TCM's dilated convs are not (diff3.py:224-243, padding = 2*dilation on both sides), so the
        TCM treats frames >= 1 + lengths[b]//160 as its own zero padding: every utterance's valid frames then equal the
        result of running it alone."""
        B, _, T, F = x.shape
        assert F == N_FREQ and x.is_contiguous() and x0.is_contiguous()
        assert x.dtype == torch.float32 and x0.dtype == torch.float32
        L, ws, s = self.lib, self.workspace(B, T), _lib.stream_ptr(stream)
        p, chk = _lib.ptr, _lib.check
        bias = p(bias_rows)
        run = self._timed
        eager = stream is None and not torch.cuda.is_current_stream_capturing()
        if eager:
            self.poll_status()
        run("enc1", lambda: L.pdse_enc1_fwd(p(x), p(x0), p(ws["e1"]), p(self.wb["enc1"]), p(self.wf["enc1"]), bias,
                                            bias_stride, B, T, s))
        for i in range(2, 6):
            Fin = P.ENC_F[i - 1]
            run(f"enc{i}", lambda: L.pdse_enc_fwd(p(ws[f"e{i - 1}"]), p(ws[f"e{i}"]), p(self.wb[f"enc{i}"]),
                                                  p(self.wf[f"enc{i}"]), bias, bias_stride, P.bias_off_enc(i), B, T, Fin,
                                                  _enc_nt(Fin), s))
        if upto == "enc":
            return None
        if self.tcm_persistent:
            run("tcm_flow", lambda: L.pdse_tcm_flow(p(ws["e5"]), p(ws["am0"]), p(ws["ak0"]), p(ws["am1"]), p(ws["ak1"]),
                                                    p(ws["x"]), p(ws["dec_in"]), p(self.tcm_table), p(ws["tcm_flags"]),
                                                    self.tcm_dil, p(lengths), p(self.status), B, T, s))
        else:
            for k in range(19):
                # launch k reads the activated maps launch k-1 wrote (ping-pong buffers)
                a_in, k_in, a_out, k_out = ("am0", "ak0", "am1", "ak1") if k % 2 else ("am1", "ak1", "am0", "ak0")
                wA = fA = wB = fB = None
                if k >= 1:
                    wA = C.c_void_p(self.wb[f"tcm{k - 1}"].data_ptr() + 16384 * 2)   # skip w1: wm | wk | w3
                    fA = p(self.wf[f"tcm{k - 1}"])
                if k <= 17:
                    wB = p(self.wb[f"tcm{k}"])
                    fB = p(self.wf[f"tcm{k}"])
                run("tcm", lambda: L.pdse_tcm_fwd(p(ws["e5"]), p(ws[a_in]), p(ws[k_in]), p(ws[a_out]), p(ws[k_out]),
                                                  p(ws["x"]), p(ws["dec_in"]), wA, fA, wB, fB, p(lengths), B, T,
                                                  TCM_DILATIONS[k - 1] if k >= 1 else 1, s))
        if upto == "tcm":
            return None
        for i in range(5, 0, -1):
            Fin = P.ENC_F[i]
            kw = 5 if i == 1 else 3
            xa = (ws["dec_in"], ws["dec_in"]) if i == 5 else (ws[f"d0_{i + 1}"], ws[f"d1_{i + 1}"])
            out = (None, None) if i == 1 else (ws[f"d0_{i}"], ws[f"d1_{i}"])
            run(f"dec{i}", lambda: L.pdse_dec_fwd(
                p(xa[0]), p(xa[1]), p(ws[f"e{i}"]), p(out[0]), p(out[1]), p(ws["eps"]) if i == 1 else None,
                p(self.wb[f"dec0_{i}"]), p(self.wb[f"dec1_{i}"]), p(self.wf[f"dec0_{i}"]), p(self.wf[f"dec1_{i}"]),
                bias, bias_stride, P.bias_off_dec(0, i), P.bias_off_dec(1, i), B, T, Fin, kw, _dec_nt(Fin, kw),
                1 if i == 1 else 0, self._dec_h(ws, i, B, T), s))
        if eager and self.tcm_persistent:
            self.post_status()
        return ws["eps"][:B * 2 * T * N_FREQ].view(B, 2, T, N_FREQ)


class DiffUNetPriorEngine:
    """Prior ``DiffUNet`` (model/diff.py:13-33) on the DiffUNet1 kernels: X_init = DiffUNet(y) / 11."""

    def __init__(self, state_dict, device, out_scale: float = 1.0 / 11.0):
        self.engine = DenoiserEngine(P.diffunet_as_diffunet1(state_dict, out_scale), device)
        self.rows = self.engine.time_bias(torch.zeros(1))      # time paths are all-zero: any t gives hb = b1
        self._zeros: Dict[tuple, torch.Tensor] = {}

    @property
    def timing(self):
        return self.engine.timing

    @timing.setter
    def timing(self, v):
        self.engine.timing = v

    def check_status(self, synchronize: bool = True):
        self.engine.check_status(synchronize)

    def set_wait_timeout_us(self, us: int):
        """wall-clock bound of one dependency wait of the persistent TCM kernel: 0 = 2 s (default), < 0 = fail on the first
        unsatisfied poll (forces the error path in tests).  Lives in the status block on the device: graphs follow it."""
        self.status[4] = int(us)

    def post_status(self):
        self.engine.post_status()

    def set_wait_timeout_us(self, us: int):
        self.engine.set_wait_timeout_us(us)

    def poll_status(self):
        self.engine.poll_status()

    def forward(self, y: torch.Tensor, out: Optional[torch.Tensor] = None, stream=None,
                lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
        z = self._zeros.get(tuple(y.shape))
        if z is None:
            z = self._zeros[tuple(y.shape)] = torch.zeros_like(y)
        eps = self.engine.forward(y, z, self.rows, 0, stream=stream, lengths=lengths)
        if out is None:
            return eps.clone()
        out.copy_(eps)
        return out
