"""Shared pieces of the two speech datasets (jsonl map-style, multitask dynamic-frame): audio front end, token / label layout and
the collation of the audio side.  The CONTRACT they implement is the reference's (src/slam_llm/datasets/speech_dataset.py:17-298,
speech_dataset_large.py:23-275): items are [audio(-1) x L, prompt, answer, eos] with labels -100 outside answer + eos, and batches carry
input_ids / labels / attention_mask / modality_mask plus exactly one of audio (raw), audio_mel, audio_pcm.

B200 difference: with input_type == "mel" and `b200_gpu_frontend` (default) items carry raw PCM (`audio_pcm`) and the log-mel is computed
on the GPU inside the step; the reference computes whisper.log_mel_spectrogram in the DataLoader workers."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import whisper

IGNORE_INDEX = -100
HOP = 160                      # whisper hop length: one mel frame per 160 samples
DEFAULT_ASR_PROMPT = ("Transcribe speech to text. Output the transcription directly without redundant content. "
                      "Ensure that the output is not duplicated. ")


# ------------------------------------------------------------------------------------------------- padding primitives
def _filler(like, count: int, value):
    if isinstance(like, torch.Tensor):
        return torch.full([count] + list(like.shape[1:]), value, dtype=like.dtype)
    return np.full((count,) + like.shape[1:], value)


def fit_length(seq, length: int, value=0):
    """Right-pad (or cut) the leading axis of a list / tensor / ndarray to exactly `length`."""
    if isinstance(seq, (int, list, tuple)):
        seq = list(seq) if not isinstance(seq, int) else seq
        return seq[:length] if len(seq) >= length else seq + [value] * (length - len(seq))
    if isinstance(seq, (torch.Tensor, np.ndarray)):
        have = len(seq)
        if have >= length:
            return seq[:length]
        joiner = torch.cat if isinstance(seq, torch.Tensor) else np.concatenate
        return joiner((seq, _filler(seq, length - have, value)))
    raise Exception("Type mismatch during padding!")


def extend_by(seq, amount: int, value=0, side: str = "right"):
    """Grow the leading axis by `amount` elements on `side` (2-D tensors: the last axis, right side); negative amounts cut."""
    if isinstance(seq, (int, list, tuple)):
        return seq + [value] * amount if amount >= 0 else seq[:amount]
    if isinstance(seq, torch.Tensor):
        if seq.ndimension() == 2:
            return torch.nn.functional.pad(seq, (0, amount)) if amount >= 0 else seq[:, :amount]
        if amount < 0:
            return seq[:amount]
        fill = _filler(seq, amount, value)
        return torch.cat((fill, seq) if side == "left" else (seq, fill))
    if isinstance(seq, np.ndarray):
        return np.concatenate((seq, _filler(seq, amount, value))) if amount >= 0 else seq[:amount]
    raise Exception("Type mismatch during padding!")


# ------------------------------------------------------------------------------------------------- audio front end
def llm_audio_tokens(n_mel_frames: int) -> int:
    """Positions the audio occupies in the LLM sequence: whisper's stride-2 conv, then the k = 5 projector."""
    return ((n_mel_frames + 1) // 2) // 5


def audio_fields(wave: np.ndarray, *, input_type: str, normalize: bool, pad_or_trim: bool, gpu_frontend: bool, mel_size: int):
    """-> (fields, n_llm_tokens) with fields = {"audio", "audio_mel", "audio_pcm"} (exactly one is not None)."""
    fields = {"audio": None, "audio_mel": None, "audio_pcm": None}
    if input_type == "raw":
        raw = torch.from_numpy(wave).float()
        if normalize:
            raw = torch.nn.functional.layer_norm(raw, raw.shape)
        fields["audio"] = raw
        return fields, len(raw) // 320 // 5
    if pad_or_trim:
        wave = whisper.pad_or_trim(wave)
    if gpu_frontend:
        fields["audio_pcm"] = torch.from_numpy(np.ascontiguousarray(wave, dtype=np.float32))
        return fields, llm_audio_tokens(fields["audio_pcm"].shape[0] // HOP)
    fields["audio_mel"] = whisper.log_mel_spectrogram(wave, n_mels=mel_size).permute(1, 0)
    return fields, llm_audio_tokens(fields["audio_mel"].shape[0])


# ------------------------------------------------------------------------------------------------- token / label layout
def token_fields(tokenizer, n_audio: int, prompt: str, answer: Optional[str]) -> Dict[str, torch.Tensor]:
    """[-1 x n_audio, prompt, (answer, eos)] with labels masked outside answer + eos; answer None = inference item (no labels)."""
    audio_slots = torch.full((n_audio,), -1)
    prompt_ids = tokenizer.encode(prompt)
    if answer is None:
        ids = torch.cat((audio_slots, torch.tensor(prompt_ids, dtype=torch.int64)))
        return {"input_ids": ids, "attention_mask": ids.ge(-1), "prompt_length": len(prompt_ids)}
    text_ids = tokenizer.encode(prompt + answer) + [tokenizer.eos_token_id]
    ids = torch.cat((audio_slots, torch.tensor(text_ids, dtype=torch.int64)))
    labels = ids.clone()
    labels[: n_audio + len(prompt_ids)] = -1
    attend = ids.ge(-1)
    supervised = labels.ge(0)
    ids[~attend] = 0
    labels[~supervised] = IGNORE_INDEX
    return {"input_ids": ids, "labels": labels, "attention_mask": attend, "prompt_length": len(prompt_ids)}


# ------------------------------------------------------------------------------------------------- collation of the audio side
def _valid_mask(lengths: Sequence[int], width: int) -> torch.Tensor:
    return (torch.arange(width)[None, :] < torch.tensor(list(lengths))[:, None]).float()


def collate_audio(samples: List[dict], input_type: str, *, even_frames: bool = False) -> Dict[str, Optional[torch.Tensor]]:
    """Stack the audio payload of a batch: `audio` + `audio_mask` (raw), `audio_pcm` (+ lengths) or `audio_mel`, and the
    post-conv frame mask `audio_mel_post_mask`.  even_frames: round the PCM length up to whole, even mel frames (dynamic batches)."""
    out = dict(audio=None, audio_mask=None, audio_mel=None, audio_pcm=None, audio_pcm_lengths=None, audio_mel_post_mask=None)
    if input_type == "raw":
        lens = [s["audio"].shape[0] for s in samples]
        out["audio"] = torch.stack([fit_length(s["audio"], max(lens), 0) for s in samples])
        out["audio_mask"] = _valid_mask(lens, max(lens))
    elif samples[0].get("audio_pcm") is not None:
        lens = [s["audio_pcm"].shape[0] for s in samples]
        width = max(lens)
        if even_frames:
            width = (width + 2 * HOP - 1) // (2 * HOP) * (2 * HOP)
        out["audio_pcm"] = torch.stack([fit_length(s["audio_pcm"], width, 0) for s in samples])
        out["audio_pcm_lengths"] = torch.tensor(lens, dtype=torch.int32)
        out["audio_mel_post_mask"] = _valid_mask([(n // HOP + 1) // 2 for n in lens], (width // HOP + 1) // 2)
    else:
        lens = [s["audio_mel"].shape[0] for s in samples]
        out["audio_mel"] = torch.stack([fit_length(s["audio_mel"], max(lens), 0) for s in samples])
        out["audio_mel_post_mask"] = _valid_mask([(n + 1) // 2 for n in lens], (max(lens) + 1) // 2)
    return out


def span_mask(template: torch.Tensor, starts: Sequence[int], lengths: Sequence[int]) -> torch.Tensor:
    """Boolean [B, S] mask like `template` that is set on [start, start + length) of every row (the audio span)."""
    pos = torch.arange(template.shape[1])[None, :]
    lo = torch.tensor(list(starts))[:, None]
    return ((pos >= lo) & (pos < lo + torch.tensor(list(lengths))[:, None])).to(template.dtype)
