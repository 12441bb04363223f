"""Host-side constants of the Whisper front end: slaney mel filterbank and sinusoidal positions.

Restated from the published librosa / openai-whisper formulas (the reference calls
whisper.log_mel_spectrogram at datasets/speech_dataset.py:103, which loads librosa-generated
`mel_filters.npz`; whisper itself is not vendored in /root/reference).
"""
from __future__ import annotations

import math

import numpy as np
import torch

SAMPLE_RATE = 16000
N_FFT = 400
HOP_LENGTH = 160
N_SAMPLES = 480000
N_FRAMES = 3000


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_filterbank(n_mels: int, sr: int = SAMPLE_RATE, n_fft: int = N_FFT, fmin: float = 0.0, fmax: float | None = None) -> torch.Tensor:
    """librosa.filters.mel(sr, n_fft, n_mels, htk=False, norm='slaney') -> f32 [n_mels, n_fft//2+1]."""
    fmax = sr / 2.0 if fmax is None else fmax
    fftfreqs = np.linspace(0.0, sr / 2.0, n_fft // 2 + 1)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    weights = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights = weights * enorm[:, None]
    return torch.from_numpy(weights.astype(np.float32))


def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> torch.Tensor:
    """whisper.model.sinusoids: positional embedding buffer of AudioEncoder (f32 [length, channels])."""
    assert channels % 2 == 0
    log_timescale_increment = math.log(max_timescale) / (channels // 2 - 1)
    inv_timescales = torch.exp(-log_timescale_increment * torch.arange(channels // 2, dtype=torch.float32))
    scaled_time = torch.arange(length, dtype=torch.float32)[:, None] * inv_timescales[None, :]
    return torch.cat([torch.sin(scaled_time), torch.cos(scaled_time)], dim=1)


def pad_or_trim(wav: torch.Tensor, length: int = N_SAMPLES) -> torch.Tensor:
    """whisper.pad_or_trim on the last axis."""
    n = wav.shape[-1]
    if n > length:
        return wav[..., :length]
    if n < length:
        return torch.nn.functional.pad(wav, (0, length - n))
    return wav
