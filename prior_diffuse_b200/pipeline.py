"""The generate / eval hot path of trainer/complex_ddpm_trainer.py (:903-1018, batched twin
:408-495) as one CUDA graph per input shape:

    wav -> RMS normalise -> STFT + sqrt-compress -> prior -> X_init/11
        -> x_T [* sqrt(mask)] -> N x {eps = D(x, X_init, t_n); x = c1 (x - c2 eps)}
        -> (x + X_init) * 11 -> decompress -> ISTFT -> * rms

Both networks run with eval-mode BatchNorm (as at :400-401), so utterances are
independent and a batch shards over GPUs with no collective on the path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import lib as _lib
from . import pack as P
from . import signal as S
from .dbaiat import DBAIATEngine
from .denoiser import DenoiserEngine, DiffUNetPriorEngine
from .gcrn import GCRNEngine

FEAT_SCALE = 11.0   # trainer/complex_ddpm_trainer.py:30
NOISE_SCHEDULE = np.linspace(1e-4, 0.05, 50).tolist()               # utils/params.py:40
INFERENCE_NOISE_SCHEDULE = [0.0001, 0.001, 0.01, 0.05, 0.2, 0.5]    # utils/params.py:41


def inference_schedule(fast_sampling: bool = True, noise_schedule: Sequence[float] = NOISE_SCHEDULE,
                       inference_noise_schedule: Sequence[float] = INFERENCE_NOISE_SCHEDULE):
    """Same contract as ComplexDDPMTrainer.inference_schedule (:105-156):
    returns (alpha, beta, alpha_cum, sigmas, T) with T the fractional training-step index of
    every inference step (float32)."""
    train = np.asarray(noise_schedule, dtype=np.float64)
    beta = np.asarray(inference_noise_schedule, dtype=np.float64) if fast_sampling else train
    train_cum = np.cumprod(1.0 - train)
    alpha = 1.0 - beta
    alpha_cum = np.cumprod(alpha)
    n = len(alpha)
    # sqrt(beta~_t); index n-1 wraps around for n = 0 exactly like the reference (:128)
    sigmas = [float(((1.0 - alpha_cum[i - 1]) / (1.0 - alpha_cum[i]) * beta[i]) ** 0.5) for i in range(n)]
    root = np.sqrt(train_cum)
    T = []
    for s in range(n):
        for t in range(len(train) - 1):
            if train_cum[t + 1] <= alpha_cum[s] <= train_cum[t]:
                T.append(t + (root[t] - alpha_cum[s] ** 0.5) / (root[t] - root[t + 1]))
                break
    return alpha, beta, alpha_cum, sigmas, np.asarray(T, dtype=np.float32)


class _Plan:
    """static buffers + captured graph for one (B, L)"""

    def __init__(self):
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.buf: Dict[str, torch.Tensor] = {}
        self.ragged = False
        # host-pipeline bookkeeping (Enhancer.enhance_host_pipelined): last graph completion / last D2H of this plan
        self.done: Optional[torch.cuda.Event] = None
        self.copied: Optional[torch.cuda.Event] = None


class Enhancer:
    """B200 replacement for the reference's generate path (prior = GCRN, DiffUNet or aia_complex_trans_ri)."""

    MODES = ("priorgrad", "deltamu", "condition")

    def __init__(self, prior_state_dict, ddpm_state_dict, device="cuda:0", fast_sampling: bool = True,
                 sigma_mask: bool = False, use_graph: bool = True, prior: str = "GCRN", mode: str = "priorgrad",
                 rng: str = "philox", rank: int = 0):
        """``mode``: which of the trainer's three reverse-loop branches runs (:945-948, :967-975, :993-994):
        "priorgrad" -- DiffUNet1(x, X_init, t), final x + X_init (params.pirorgrad, the shipped setting);
        "deltamu"   -- Nocon(x, t) (``ddpm_state_dict`` in model/piror_grad.py layout), x_T = z + X_init;
        "condition" -- DiffUNet1(x, noisy features / 11, t).

        ``rng``: "philox" -- x_T from the library's own counter-based generator (seed, rank, running element offset);
        "aten" -- x_T is, value for value, what ``torch.manual_seed(seed); torch.randn_like(x)`` yields on this device
        (the reference's noise stream at :950; validation mode).  ``rank`` separates the noise of data-parallel shards."""
        if mode not in self.MODES:
            raise ValueError(f"mode must be one of {self.MODES}")
        if rng not in ("philox", "aten"):
            raise ValueError("rng must be 'philox' or 'aten'")
        self.mode = mode
        self.rng = rng
        self.rank = int(rank)
        self.device = torch.device(device)
        self.lib = _lib.load(require_device=True)
        engines = {"GCRN": GCRNEngine, "DiffUNet": DiffUNetPriorEngine, "aia_complex_trans_ri": DBAIATEngine}
        if prior not in engines:
            raise ValueError(f"prior {prior!r} is not built (available: {', '.join(engines)})")
        self.prior_name = prior
        with torch.cuda.device(self.device):
            self.prior = engines[prior](prior_state_dict, self.device)
            if mode == "deltamu":     # Nocon = DiffUNet1 with an identity Preprocess on x
                ddpm_state_dict = P.diffunet_as_diffunet1(ddpm_state_dict)
            self.ddpm = DenoiserEngine(ddpm_state_dict, self.device)
        self.fast = fast_sampling
        self.sigma_mask = sigma_mask
        self.use_graph = use_graph
        alpha, beta, alpha_cum, _, T = inference_schedule(fast_sampling)
        self.n_steps = len(alpha)
        self.c1 = [float(1.0 / alpha[n] ** 0.5) for n in range(self.n_steps)]
        self.c2 = [float(beta[n] / (1.0 - alpha_cum[n]) ** 0.5) for n in range(self.n_steps)]
        self.t_index = torch.from_numpy(T.copy())
        with torch.cuda.device(self.device):
            # every time-dependent bias of every step, once (diff3.py:39 + all tp projections)
            self.bias_rows = self.ddpm.time_bias(self.t_index)
        self._plans: Dict[tuple, _Plan] = {}
        self._rng_offset = 0          # philox: float4 units drawn so far; aten: the generator's Philox offset
        self._rng_seed = None
        self.kernels_per_call = 0
        self._copy_in: Optional[torch.cuda.Stream] = None
        self._copy_out: Optional[torch.cuda.Stream] = None

    # ------------------------------------------------------------------ kernel-side errors
    def _status_engines(self):
        return [self.ddpm] + ([self.prior] if self.prior_name == "DiffUNet" else [])

    def check(self, synchronize: bool = True):
        """Raise if a kernel recorded an error (e.g. a dependency wait of the persistent TCM kernel timed out).
        ``enhance_host`` calls this at its synchronisation point; after ``enhance`` (asynchronous) the error surfaces at
        the next call or at an explicit ``check()``."""
        for e in self._status_engines():
            e.check_status(synchronize)

    # ------------------------------------------------------------------ one pass (eager or under capture)
    def _run(self, pl: _Plan, stream=None, trace: Optional[dict] = None):
        b, L, lib = pl.buf, self.lib, self.lib
        p, chk, s = _lib.ptr, _lib.check, _lib.stream_ptr(stream)
        B, n = b["wav"].shape
        T = S.n_frames(n)
        nel, plane = B * 2 * T * S.N_FREQ, T * S.N_FREQ
        ln = b["len"] if pl.ragged else None
        shape = (B, 2, T, S.N_FREQ)
        feat, xinit = b["feat"][:nel].view(shape), b["xinit"][:nel].view(shape)
        S.rms(b["wav"], out=b["rms"], stream=stream, lengths=ln)
        S.stft_compress(b["wav"], b["rms"], out=feat, stream=stream, lengths=ln)
        if self.prior_name == "DiffUNet":
            self.prior.forward(feat, out=xinit, stream=stream, lengths=ln)
        else:
            self.prior.forward(feat, out=xinit, stream=stream)
        launches = 2 + {"GCRN": self._prior_launches(B), "DiffUNet": self.ddpm.kernel_launches() + 1, "aia_complex_trans_ri": 63}[self.prior_name]
        if self.sigma_mask:
            chk(lib.pdse_absmax_ragged_f32(p(b["xinit"]), p(ln), B * 2, plane, p(b["amax"]), s))
            launches += 1
        if self.sigma_mask or self.mode == "deltamu":
            chk(lib.pdse_init_state_add_f32(p(b["x"]), p(b["xinit"]) if self.sigma_mask else None, p(b["amax"]),
                                            p(b["xinit"]) if self.mode == "deltamu" else None, nel, plane, 0, 0, 0, s))
            launches += 1
        second = xinit
        if self.mode == "condition":      # batch_feat /= c (:943)
            b["cond"].copy_(b["feat"])
            chk(lib.pdse_scale_f32(p(b["cond"]), nel, 1.0 / FEAT_SCALE, s))
            second = b["cond"][:nel].view(shape)
            launches += 2
        fin = 1 if self.mode == "priorgrad" else 2
        x = b["x"][:nel].view(shape)
        for n_ in range(self.n_steps - 1, -1, -1):
            eps = self.ddpm.forward(x, second, self.bias_rows[n_:n_ + 1], 0, stream=stream, lengths=ln)
            last = n_ == 0
            # newsigma == 0 for every step in the reference (:986-992)
            chk(lib.pdse_ddpm_update_f32(p(b["x"]), p(eps), p(b["xinit"]), None, p(b["spec"]) if last else None, nel,
                                         plane, self.c1[n_], self.c2[n_], 0.0, 0, fin if last else 0, FEAT_SCALE, 0, 0, s))
            launches += self.ddpm.kernel_launches() + 1
            if trace is not None:
                trace.setdefault("eps", []).append(eps.clone())
                trace.setdefault("x", []).append((b["spec"] if last else b["x"])[:nel].view(shape).clone())
        spec = b["spec"][:nel].view(shape)
        # "pcm": the writer's float -> int16 conversion (:1018 sf.write, PCM_16), fused into the overlap-add store
        S.decompress_istft(spec, n, b["rms"], out=b["out"], stream=stream, lengths=ln, pcm=b.get("pcm"))
        launches += 1
        self.kernels_per_call = launches

    @staticmethod
    def _prior_launches(B: int) -> int:
        chunks = (B + 63) // 64
        return chunks * (5 + 2 * (2 + 1 + 1) + 8 + 1)

    def _plan(self, B: int, n: int, ragged: bool = False, slot: int = 0, pcm16: bool = False) -> _Plan:
        key = (B, n) if not (ragged or slot or pcm16) else (B, n, ragged, slot, pcm16)
        pl = self._plans.get(key)
        if pl is None:
            pl = _Plan()
            pl.ragged = ragged
            dev = self.device
            T = S.n_frames(n)
            # every [B,2,T,161] tensor is processed as float4 by the element-wise kernels: capacity rounded up to 4 floats
            nel = (B * 2 * T * S.N_FREQ + 3) // 4 * 4
            f32 = dict(dtype=torch.float32, device=dev)
            pl.buf = {
                "wav": torch.zeros(B, n, **f32), "rms": torch.zeros(B, **f32),
                "feat": torch.zeros(nel, **f32), "xinit": torch.zeros(nel, **f32),
                "cond": torch.zeros(nel, **f32) if self.mode == "condition" else None,
                "x": torch.zeros(nel, **f32), "spec": torch.zeros(nel, **f32), "amax": torch.zeros(B * 2, **f32),
                "out": torch.zeros(B, n, **f32), "len": torch.full((B,), n, dtype=torch.int32, device=dev),
            }
            if pcm16:
                pl.buf["pcm"] = torch.zeros(B, n, dtype=torch.int16, device=dev)
            self._plans[key] = pl
        return pl

    def xinit(self, B: int, n: int) -> torch.Tensor:
        """X_init [B,2,T,161] of the last plain (non-ragged) call with this shape (a view of the plan's buffer)"""
        T = S.n_frames(n)
        return self._plans[(B, n)].buf["xinit"][:B * 2 * T * S.N_FREQ].view(B, 2, T, S.N_FREQ)

    def _draw_x_T(self, b, nel: int, seed: int):
        """x_T ~ N(0, I) on the device (:950)"""
        if seed != self._rng_seed:
            self._rng_seed, self._rng_offset = seed, 0
        if self.rng == "aten":
            grid, inc = C.c_int(0), C.c_ulonglong(0)
            _lib.check(self.lib.pdse_randn_aten_policy(nel, C.byref(grid), C.byref(inc)))
            _lib.check(self.lib.pdse_randn_aten_f32(_lib.ptr(b["x"]), nel, seed, self._rng_offset, grid.value, _lib.stream_ptr()))
            self._rng_offset += inc.value
        else:
            # counter = running float4 offset (never reused across calls or shapes); the shard's rank is folded into the key
            key = (seed + 0x9E3779B97F4A7C15 * self.rank) & 0xFFFFFFFFFFFFFFFF
            _lib.check(self.lib.pdse_init_state_f32(_lib.ptr(b["x"]), None, None, nel, 0, 1, key, self._rng_offset, _lib.stream_ptr()))
            self._rng_offset += (nel + 3) // 4

    def _launch(self, pl: _Plan, trace: Optional[dict] = None):
        """run the pass on the current stream: eager (trace / use_graph=False) or by graph replay (captured on first use)"""
        b = pl.buf
        if trace is not None or not self.use_graph:
            self._run(pl, trace=trace)
        else:
            if pl.graph is None:
                keep = b["x"].clone()
                self._run(pl)                      # warm-up: sets shared-memory attributes, fills workspaces
                torch.cuda.current_stream().synchronize()
                b["x"].copy_(keep)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._run(pl)
                pl.graph = g
                b["x"].copy_(keep)
            pl.graph.replay()
        for e in self._status_engines():
            e.post_status()

    # ------------------------------------------------------------------ public API
    @torch.no_grad()
    def enhance(self, wav: torch.Tensor, x_T: Optional[torch.Tensor] = None, seed: int = 7,
                trace: Optional[dict] = None, lengths: Optional[torch.Tensor] = None, _slot: int = 0,
                _pcm16: bool = False) -> torch.Tensor:
        """wav [B, L] fp32 (device tensor) -> enhanced wav [B, L] (a view of a static buffer that the
        next call overwrites).  ``x_T`` [B,2,T,161] reproduces a given noise draw (parity runs);
        otherwise x_T comes from the on-device generator (``rng``, seed, running offset).

        ``lengths`` [B] (ints): ragged batch -- ``wav`` is zero-padded to the longest utterance
        (utils/dataset.py:45-60) and every utterance comes out exactly as if it had been enhanced alone
        (own RMS, reflect padding, sigma-mask maximum and TCM zero padding; the other layers are causal in time);
        samples past its length are zero.  GCRN / DiffUNet priors only (the DB-AIAT attention and GroupNorm span the
        whole sequence).

        The call is asynchronous; a kernel-side error of an earlier call raises here (or at ``check()``)."""
        assert wav.dim() == 2 and wav.dtype == torch.float32
        B, n = wav.shape
        if lengths is not None and self.prior_name == "aia_complex_trans_ri":
            raise ValueError("ragged batches need a time-causal prior (GCRN or DiffUNet)")
        with torch.cuda.device(self.device):
            for e in self._status_engines():
                e.poll_status()
            pl = self._plan(B, n, lengths is not None, _slot, _pcm16)
            b = pl.buf
            b["wav"].copy_(wav, non_blocking=True)
            if lengths is not None:
                lt = torch.as_tensor(lengths, dtype=torch.int32)
                if lt.numel() != B or int(lt.min()) <= S.HOP or int(lt.max()) > n:
                    raise ValueError("lengths: one entry per utterance, 160 < len <= wav.shape[1]")
                b["len"].copy_(lt, non_blocking=True)
            nel = B * 2 * S.n_frames(n) * S.N_FREQ
            if x_T is not None:
                b["x"][:nel].copy_(x_T.reshape(-1), non_blocking=True)
            else:
                self._draw_x_T(b, nel, seed)
            self._launch(pl, trace)
            return b["pcm"] if _pcm16 else b["out"]

    @torch.no_grad()
    def enhance_host(self, wav_host: torch.Tensor, out_host: Optional[torch.Tensor] = None, pcm16: bool = False,
                     **kw) -> torch.Tensor:
        """host (pinned) wav [B, L] -> host wav [B, L]: the call a user of the reference's generate path
        makes (file in, file out); H2D and D2H copies are part of it.  ``pcm16``: the result is the int16 sample
        stream the reference's writer puts into the file (:1018), converted on the device (half the D2H bytes)."""
        with torch.cuda.device(self.device):
            dev_in = wav_host.to(self.device, non_blocking=True)
            out = self.enhance(dev_in, _pcm16=pcm16, **kw)
            if out_host is None:
                out_host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
            out_host.copy_(out, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            self.check(synchronize=False)      # the status copy was enqueued before the synchronisation point
        return out_host

    @torch.no_grad()
    def enhance_host_pipelined(self, batches, outs=None, pcm16: bool = False, seed: int = 7, post=None):
        """A stream of host batches (pinned [B, L] tensors of one shape) -> list of host results, with the H2D copy of
        batch i+1 and the D2H copy of batch i-1 running on copy streams underneath batch i's graph (two plan slots).
        Same results as calling ``enhance_host`` once per batch.  ``post(out_dev)`` (optional) is called on the compute
        stream after every batch's pass (e.g. the data-parallel gather of the shards)."""
        batches = list(batches)
        if not batches:
            return []
        with torch.cuda.device(self.device):
            if self._copy_in is None:
                self._copy_in, self._copy_out = torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)
            main = torch.cuda.current_stream()
            B, n = batches[0].shape
            dt = torch.int16 if pcm16 else torch.float32
            if outs is None:
                outs = [torch.empty((B, n), dtype=dt, pin_memory=True) for _ in batches]
            nel = B * 2 * S.n_frames(n) * S.N_FREQ
            slots = [self._plan(B, n, False, k + 1, pcm16) for k in range(2)]
            for pl in slots:                      # capture outside the pipeline (needs a quiet device)
                if pl.graph is None and self.use_graph:
                    self._draw_x_T(pl.buf, nel, seed)
                    self._launch(pl)
            main.synchronize()
            start = torch.cuda.Event()
            start.record(main)
            self._copy_in.wait_event(start)
            self._copy_out.wait_event(start)
            for i, (src, dst) in enumerate(zip(batches, outs)):
                pl = slots[i & 1]
                b = pl.buf
                with torch.cuda.stream(self._copy_in):
                    if pl.done is not None:
                        self._copy_in.wait_event(pl.done)       # the graph that last read this slot's wav has finished
                    b["wav"].copy_(src, non_blocking=True)
                    h2d = torch.cuda.Event()
                    h2d.record()
                main.wait_event(h2d)
                if pl.copied is not None:
                    main.wait_event(pl.copied)                  # this slot's previous result has left the device
                self._draw_x_T(b, nel, seed)
                self._launch(pl)
                if post is not None:
                    post(b["pcm"] if pcm16 else b["out"])
                pl.done = torch.cuda.Event()
                pl.done.record(main)
                with torch.cuda.stream(self._copy_out):
                    self._copy_out.wait_event(pl.done)
                    dst.copy_(b["pcm"] if pcm16 else b["out"], non_blocking=True)
                    pl.copied = torch.cuda.Event()
                    pl.copied.record()
            self._copy_out.synchronize()
            main.synchronize()
            self.check(synchronize=False)
        return outs


def write_wav(path: str, pcm: torch.Tensor, sample_rate: int = 16000):
    """int16 samples [L] (host) -> a RIFF/WAVE PCM_16 mono file, what ``sf.write(path, wav, 16000)`` produces at
    trainer/complex_ddpm_trainer.py:1018 (standard library only)."""
    import wave
    data = pcm.detach().cpu().contiguous().numpy().astype("<i2", copy=False)
    with wave.open(path, "wb") as f:
        f.setnchannels(1)
        f.setsampwidth(2)
        f.setframerate(sample_rate)
        f.writeframes(data.tobytes())
