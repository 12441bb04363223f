"""Stand-in for the three openai-whisper audio helpers the reference datasets call by name
(datasets/speech_dataset.py:93,101,103): load_audio, pad_or_trim, log_mel_spectrogram.
Installed as `whisper` ONLY when openai-whisper is not importable.

load_audio reads 16 kHz PCM WAV through scipy/wave (ffmpeg is not in the image).  log_mel_spectrogram is the
dataset-side (DataLoader worker) CPU preprocessing the reference performs; the B200 step computes log-mel on the
GPU (slam_logmel) when the collator is left in its default `audio_pcm` mode, so this function is only reached
when a recipe explicitly asks for CPU mel features (dataset_config.b200_gpu_frontend=false)."""
from __future__ import annotations

import numpy as np
import torch

from slam_llm_b200.frontend import HOP_LENGTH, N_FFT, N_SAMPLES, SAMPLE_RATE, mel_filterbank  # noqa: F401


def load_audio(path: str, sr: int = SAMPLE_RATE) -> np.ndarray:
    from scipy.io import wavfile
    rate, data = wavfile.read(path)
    if data.ndim > 1:
        data = data.mean(axis=1)
    if data.dtype == np.int16:
        data = data.astype(np.float32) / 32768.0
    elif data.dtype == np.int32:
        data = data.astype(np.float32) / 2147483648.0
    else:
        data = data.astype(np.float32)
    if rate != sr:
        from scipy.signal import resample_poly
        data = resample_poly(data, sr, rate).astype(np.float32)
    return data


def pad_or_trim(array, length: int = N_SAMPLES, *, axis: int = -1):
    if torch.is_tensor(array):
        n = array.shape[axis]
        if n > length:
            array = array.index_select(axis, torch.arange(length, device=array.device))
        if n < length:
            pad = [(0, 0)] * array.ndim
            pad[axis] = (0, length - n)
            array = torch.nn.functional.pad(array, [p for sizes in pad[::-1] for p in sizes])
        return array
    n = array.shape[axis]
    if n > length:
        array = array.take(indices=range(length), axis=axis)
    if n < length:
        pad = [(0, 0)] * array.ndim
        pad[axis] = (0, length - n)
        array = np.pad(array, pad)
    return array


def log_mel_spectrogram(audio, n_mels: int = 80, padding: int = 0, device=None) -> torch.Tensor:
    if not torch.is_tensor(audio):
        audio = torch.from_numpy(np.asarray(audio, dtype=np.float32))
    if device is not None:
        audio = audio.to(device)
    if padding > 0:
        audio = torch.nn.functional.pad(audio, (0, padding))
    window = torch.hann_window(N_FFT).to(audio.device)
    stft = torch.stft(audio, N_FFT, HOP_LENGTH, window=window, return_complex=True)
    magnitudes = stft[..., :-1].abs() ** 2
    mel_spec = mel_filterbank(n_mels).to(audio.device) @ magnitudes
    log_spec = torch.clamp(mel_spec, min=1e-10).log10()
    log_spec = torch.maximum(log_spec, log_spec.max() - 8.0)
    return (log_spec + 4.0) / 4.0
