"""Drop-in ``nn.Module`` classes for the reference's prior and denoiser.

Same constructor arguments, ``forward`` signatures, ``state_dict`` key names / shapes / order
as ``model/gcrn.py:GCRN`` and ``model/diff3.py:DiffUNet1``, so
``trainer/complex_ddpm_trainer.py`` can build them (``:69-75``), load its checkpoints into them
(``:91-97``, ``:906-913``) and call them (``:941``, ``:968-969``) unchanged.  The forward pass runs
the sm_100a kernels with eval-mode BatchNorm; there is no CPU path and no autograd.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import weights as W
from .denoiser import DenoiserEngine, DiffUNetPriorEngine
from .dbaiat import DBAIATEngine
from .gcrn import GCRNEngine
from .pack import N_BIAS_ROW, diffunet_as_diffunet1

_BUFFER_LEAVES = ("running_mean", "running_var", "num_batches_tracked")


class _Node(nn.Module):
    """parameter container (one level of the reference's module tree)"""


class _TableModule(nn.Module):
    TABLE = ""

    def __init__(self, rows=None):
        super().__init__()
        for key, shape, kind, fan in (rows if rows is not None else W.TABLES[self.TABLE]()):
            *path, leaf = key.split(".")
            node = self
            for name in path:
                if name not in node._modules:
                    node.add_module(name, _Node())
                node = node._modules[name]
            value = self._init(shape, kind, fan)
            if leaf in _BUFFER_LEAVES:
                node.register_buffer(leaf, value)
            else:
                node.register_parameter(leaf, nn.Parameter(value, requires_grad=False))
        self._engine = None

    @staticmethod
    def _init(shape, kind, fan):
        if kind == "uniform":                      # nn.Conv*/Linear/LSTM default init
            bound = 1.0 / math.sqrt(fan)
            return (torch.rand(shape) * 2 - 1) * bound
        if kind == "xavier":                       # nn.MultiheadAttention.in_proj_weight
            return (torch.rand(shape) * 2 - 1) * math.sqrt(6.0 / fan)
        if kind == "ones":
            return torch.ones(shape)
        if kind == "zeros":
            return torch.zeros(shape)
        if kind == "count":
            return torch.zeros((), dtype=torch.int64)
        if kind == "prelu":
            return torch.full(shape, 0.25)
        raise ValueError(kind)

    # weights changed -> repack lazily
    def _apply(self, fn, recurse=True):
        self._engine = None
        return super()._apply(fn, recurse)

    def load_state_dict(self, state_dict, strict=True, assign=False):
        self._engine = None
        return super().load_state_dict(state_dict, strict=strict, assign=assign)

    def _device(self):
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError(f"{type(self).__name__}: the sm_100a kernels need the module on a CUDA device "
                               "(call .cuda()); there is no CPU path")
        return dev

    def check_status(self):
        """raise if a kernel of an earlier forward recorded an error (synchronises); see DenoiserEngine.check_status"""
        if self._engine is not None and hasattr(self._engine, "check_status"):
            with torch.cuda.device(self._device()):
                self._engine.check_status()

    def _check_mode(self):
        if self.training:
            raise RuntimeError(f"{type(self).__name__} implements inference only (eval-mode BatchNorm); "
                               "call .eval() first (trainer/complex_ddpm_trainer.py:400-401)")


class GCRN(_TableModule):
    """model/gcrn.py:87-166.  forward(x [B,2,T,161]) -> [B,2,T,161] (new tensor)."""
    TABLE = "GCRN"

    @torch.no_grad()
    def forward(self, x):
        self._check_mode()
        dev = self._device()
        with torch.cuda.device(dev):   # launches go to the module's device, whichever one is current in the caller
            if self._engine is None:
                self._engine = GCRNEngine(self.state_dict(), dev)
            y = self._engine.forward(x.to(dev, torch.float32).contiguous())
            return y * 11.0   # the engine folds the trainer's /11 (:942) into fc; undo it for the module contract


class aia_complex_trans_ri(_TableModule):   # noqa: N801  (the reference's class name, model/__init__.py:2)
    """model/dbaiat.py:450-478, the DB-AIAT prior.  forward(x [B,2,T,161]) -> [B,2,T,161]."""
    TABLE = "aia_complex_trans_ri"

    @torch.no_grad()
    def forward(self, x):
        self._check_mode()
        dev = self._device()
        with torch.cuda.device(dev):
            if self._engine is None:
                self._engine = DBAIATEngine(self.state_dict(), dev)
            return self._engine.forward(x.to(dev, torch.float32).contiguous()) * 11.0


class DiffUNet(_TableModule):
    """model/diff.py:13-33 (the default ``model.name`` of conf/diff.yml).  forward(x) -> [B,2,T,161]."""
    TABLE = "DiffUNet"

    @torch.no_grad()
    def forward(self, x):
        self._check_mode()
        dev = self._device()
        with torch.cuda.device(dev):
            if self._engine is None:
                self._engine = DiffUNetPriorEngine(self.state_dict(), dev, out_scale=1.0)
            return self._engine.forward(x.to(dev, torch.float32).contiguous())


class Nocon(_TableModule):
    """model/piror_grad.py:15-40, the ``deltamu`` denoiser.  forward(x, t[B]) -> eps [B,2,T,161]."""
    TABLE = "Nocon"

    def __init__(self, params=None):
        super().__init__()
        self.params = params
        if params is not None and len(params.noise_schedule) != W.N_TRAIN_STEPS:
            raise ValueError("the denoiser kernels are built for a 50-step training schedule")

    @torch.no_grad()
    def forward(self, x, t):
        self._check_mode()
        dev = self._device()
        with torch.cuda.device(dev):
            if self._engine is None:
                self._engine = DenoiserEngine(diffunet_as_diffunet1(self.state_dict()), dev)
            rows = self._engine.time_bias(t.reshape(-1))
            if rows.shape[0] not in (1, x.shape[0]):
                raise ValueError("t must have one entry per batch element")
            x = x.to(dev, torch.float32).contiguous()
            return self._engine.forward(x, x, rows, N_BIAS_ROW if rows.shape[0] > 1 else 0).clone()


class DiffWave(_TableModule):
    """model/diff2.py:12-56.  ``DiffWave(args, params)``; forward(audio [B, L], audio_init [B, L],
    diffusion_step [B]) -> [B, 1, L].  ``params`` may carry residual_layers / dilation_cycle_length (defaults: the
    DiffWave base configuration 30 / 10; residual_channels must be 64).  Fractional steps are lerped per utterance (the
    reference's ``_lerp_embedding`` only broadcasts for B = 1)."""
    TABLE = "DiffWave"

    def __init__(self, args=None, params=None):
        def get(key, default):      # attribute- or AttrDict-style params; utils/params.py has none of the entries
            try:
                return getattr(params, key)
            except (AttributeError, KeyError):
                return default
        layers, cycle = get("residual_layers", W.DIFFWAVE_LAYERS), get("dilation_cycle_length", W.DIFFWAVE_CYCLE)
        ch = get("residual_channels", W.DIFFWAVE_CHANNELS)
        if ch != W.DIFFWAVE_CHANNELS:
            raise ValueError("the DiffWave kernels are built for residual_channels = 64")
        super().__init__(W.diffwave_table(layers, ch))
        self.params, self.cycle = params, cycle

    @torch.no_grad()
    def forward(self, audio, audio_init, diffusion_step):
        self._check_mode()
        dev = self._device()
        with torch.cuda.device(dev):
            if self._engine is None:
                from .diffwave import DiffWaveEngine
                self._engine = DiffWaveEngine(self.state_dict(), dev, self.cycle)
            return self._engine.forward(audio.to(dev, torch.float32).contiguous(), audio_init.to(dev, torch.float32).contiguous(),
                                        diffusion_step.reshape(-1))


class DiffUNet1(_TableModule):
    """model/diff3.py:14-57.  forward(x, x_init, t[B]) -> eps [B,2,T,161]."""
    TABLE = "DiffUNet1"

    def __init__(self, params=None):
        super().__init__()
        self.params = params
        if params is not None and len(params.noise_schedule) != W.N_TRAIN_STEPS:
            raise ValueError("DiffUNet1 kernels are built for a 50-step training schedule")

    @torch.no_grad()
    def forward(self, x, x_init, t):
        self._check_mode()
        dev = self._device()
        with torch.cuda.device(dev):
            if self._engine is None:
                self._engine = DenoiserEngine(self.state_dict(), dev)
            rows = self._engine.time_bias(t.reshape(-1))
            stride = N_BIAS_ROW if rows.shape[0] > 1 else 0
            if rows.shape[0] not in (1, x.shape[0]):
                raise ValueError("t must have one entry per batch element")
            eps = self._engine.forward(x.to(dev, torch.float32).contiguous(), x_init.to(dev, torch.float32).contiguous(),
                                       rows, stride)
            return eps.clone()
