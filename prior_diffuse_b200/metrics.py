"""Evaluation scalars computed on the device: segmental SNR of utils/metrics.py:36-55."""
from __future__ import annotations

from typing import Optional

import torch

from . import lib as _lib


def snr_seg(clean: torch.Tensor, processed: torch.Tensor, lengths: Optional[torch.Tensor] = None,
            out: Optional[torch.Tensor] = None, stream=None) -> torch.Tensor:
    """SNRseg(clean, processed, 16000) per utterance: clean / processed [B, L] fp32 on the device -> [B] fp32 (dB).
    ``lengths`` int32 [B] (device) restricts every utterance to its own samples (ragged batch)."""
    L = _lib.load(require_device=True)
    assert clean.shape == processed.shape and clean.dim() == 2
    assert clean.is_cuda and clean.dtype == torch.float32 and clean.is_contiguous() and processed.is_contiguous()
    B, n = clean.shape
    if out is None:
        out = torch.empty(B, dtype=torch.float32, device=clean.device)
    _lib.check(L.pdse_ssnr_f32(_lib.ptr(clean), _lib.ptr(processed), _lib.ptr(lengths), B, n, _lib.ptr(out),
                               _lib.stream_ptr(stream)))
    return out
