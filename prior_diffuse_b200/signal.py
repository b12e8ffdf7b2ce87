"""STFT / ISTFT front and back end on the device (replaces the torch.stft / torch.istft calls
at trainer/complex_ddpm_trainer.py:926-930 and :1009-1015 plus the (de)compression around them).

Layout contract is the reference's: spectrogram ``[B, 2, T, 161]`` fp32 (ch 0 = re, 1 = im).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import lib as _lib

N_FFT, HOP, N_FREQ = 320, 160, 161
_TABLES = {}


def tables(device) -> torch.Tensor:
    device = torch.device(device)
    t = _TABLES.get(device)
    if t is None:
        L = _lib.load(require_device=True)
        host = np.zeros(L.pdse_signal_table_floats(), dtype=np.float32)
        _lib.check(L.pdse_signal_tables(host.ctypes.data_as(C.c_void_p)))
        t = torch.from_numpy(host).to(device)
        _TABLES[device] = t
    return t


def n_frames(length: int) -> int:
    return 1 + length // HOP


def rms(wav: torch.Tensor, out: Optional[torch.Tensor] = None, stream=None,
        lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
    """per-utterance sqrt(mean(x^2))  (:922-923).  wav [B, L] fp32 on the device; ``lengths`` int32 [B] on the
    device = true sample counts of a zero-padded ragged batch."""
    L = _lib.load(require_device=True)
    B, n = wav.shape
    if out is None:
        out = torch.empty(B, dtype=torch.float32, device=wav.device)
    _lib.check(L.pdse_rms_ragged_f32(_lib.ptr(wav), _lib.ptr(lengths), B, n, _lib.ptr(out), _lib.stream_ptr(stream)))
    return out


def stft_compress(wav: torch.Tensor, rms_: Optional[torch.Tensor] = None, compress: bool = True,
                  out: Optional[torch.Tensor] = None, stream=None, lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
    """wav [B, L] -> [B, 2, T, 161]; divides by rms_[b] first when given.  With ``lengths`` every utterance is
    reflected at its own end and frames past 1 + len/160 are zero."""
    L = _lib.load(require_device=True)
    assert wav.is_cuda and wav.dtype == torch.float32 and wav.is_contiguous() and wav.dim() == 2
    B, n = wav.shape
    T = n_frames(n)
    if out is None:
        out = torch.empty(B, 2, T, N_FREQ, dtype=torch.float32, device=wav.device)
    _lib.check(L.pdse_stft_compress_ragged_f32(_lib.ptr(wav), _lib.ptr(rms_), _lib.ptr(tables(wav.device)), _lib.ptr(lengths),
                                               _lib.ptr(out), B, n, int(compress), _lib.stream_ptr(stream)))
    return out


def decompress_istft(spec: torch.Tensor, length: int, rms_: Optional[torch.Tensor] = None, decompress: bool = True,
                     out: Optional[torch.Tensor] = None, stream=None, lengths: Optional[torch.Tensor] = None,
                     pcm: Optional[torch.Tensor] = None, pcm_clip: bool = False) -> torch.Tensor:
    """[B, 2, T, 161] -> wav [B, length]; multiplies by rms_[b] at the end when given.  ``pcm`` (int16 [B, length]):
    also receives the samples as the reference's writer converts them (sf.write PCM_16, :1018)."""
    L = _lib.load(require_device=True)
    assert spec.is_cuda and spec.dtype == torch.float32 and spec.is_contiguous()
    B, _, T, F = spec.shape
    assert F == N_FREQ
    if out is None:
        out = torch.empty(B, length, dtype=torch.float32, device=spec.device)
    _lib.check(L.pdse_decompress_istft_pcm16_f32(_lib.ptr(spec), _lib.ptr(rms_), _lib.ptr(tables(spec.device)),
                                                 _lib.ptr(lengths), _lib.ptr(out), _lib.ptr(pcm), B, T, length,
                                                 int(decompress), int(pcm_clip), _lib.stream_ptr(stream)))
    return out


def stft(wav: torch.Tensor, **kw) -> torch.Tensor:
    """drop-in for the reference's torch.stft call + permute: [B, L] -> [B, 2, T, 161]."""
    return stft_compress(wav, None, compress=False, **kw)


def istft(spec: torch.Tensor, length: int, **kw) -> torch.Tensor:
    return decompress_istft(spec, length, None, decompress=False, **kw)
