"""jsonl audio-caption dataset for non-speech encoders (reference: src/slam_llm/datasets/audio_dataset.py:19-231, the aac_audiocaps / EAT
recipes): items [audio(-1) x L, prompt, answer, eos] with a kaldi-fbank `audio_mel` for the frozen EAT encoder, RIGHT-padding collator with
`audio_mel_mask`.  Token / label mechanics are shared with the speech datasets (`_b200_common`).  `model_name == "eat"` is supported
(EAT_preprocess restated from models/EAT/EAT.py:5-32); BEATs preprocessing lives in the reference's vendored BEATs package."""
import json
import random

import torch
import torchaudio

from slam_llm.datasets import _b200_common as common       # absolute: recipes load this file by PATH


def eat_preprocess(source: torch.Tensor, norm_mean: float = -4.268, norm_std: float = 4.569, target_length: int = 1024, fixed_length: bool = False,
                   random_crop: bool = False) -> torch.Tensor:
    """waveform [n] @16 kHz -> normalised 128-bin kaldi fbank [frames, 128]; frames padded to a multiple of 16 (the EAT patch size) or cut /
    padded to `target_length` when `fixed_length`."""
    source = (source - source.mean()).unsqueeze(0)
    fb = torchaudio.compliance.kaldi.fbank(source, htk_compat=True, sample_frequency=16000, use_energy=False, window_type="hanning", num_mel_bins=128,
                                           dither=0.0, frame_shift=10)
    n = fb.shape[0]
    if not fixed_length:
        target_length = n if n % 16 == 0 else n + (16 - n % 16)
    if target_length > n:
        fb = torch.nn.functional.pad(fb, (0, 0, 0, target_length - n))
    elif target_length < n:
        start = random.randint(0, n - target_length) if random_crop else 0
        fb = fb[start: start + target_length]
    return (fb - norm_mean) / (norm_std * 2)


def _load_wave(path: str):
    """torchaudio.load (as the reference calls it); torchaudio >= 2.9 needs torchcodec for that - without it 16-bit PCM WAV files are read
    through scipy with the same return convention ([channels, n] float32 in [-1, 1), rate)."""
    try:
        return torchaudio.load(path)
    except ImportError:
        from scipy.io import wavfile
        rate, data = wavfile.read(path)
        if data.dtype.kind != "i":
            return torch.from_numpy(data.astype("float32")).reshape(1, -1) if data.ndim == 1 else torch.from_numpy(data.astype("float32")).t(), rate
        scale = float(2 ** (8 * data.dtype.itemsize - 1))
        wave = torch.from_numpy(data.astype("float32") / scale)
        return (wave.reshape(1, -1) if wave.ndim == 1 else wave.t().contiguous()), rate


class AudioDatasetJsonl(torch.utils.data.Dataset):
    IGNORE_INDEX = common.IGNORE_INDEX
    prompt_template = "USER: {}\n ASSISTANT:"
    answer_template = "{}"

    def __init__(self, dataset_config, tokenizer=None, split="train"):
        super().__init__()
        cfg = dataset_config
        self.dataset_config, self.tokenizer, self.split = cfg, tokenizer, split
        self.fix_length_audio = cfg.fix_length_audio
        self.inference_mode = cfg.get("inference_mode", False)
        self.input_type = cfg.get("input_type", None)
        self.model_name = cfg.get("model_name", "beats")
        if self.model_name != "eat":
            raise NotImplementedError(f"dataset_config.model_name={self.model_name!r}: this mirror restates the EAT front end only")
        with open(cfg.train_data_path if split == "train" else cfg.val_data_path, encoding="utf-8") as fin:
            self.data_list = [json.loads(line.strip()) for line in fin]

    def get_source_len(self, data_dict):
        return data_dict["source_len"]

    def get_target_len(self, data_dict):
        return data_dict.get("target_len", 0)

    def __len__(self):
        return len(self.data_list)

    def __getitem__(self, index):
        record = self.data_list[index]
        cfg = self.dataset_config
        try:
            wave, rate = _load_wave(record.get("source"))
            if wave.shape[1] == 0:
                raise ValueError("Empty audio file")
            wave = torchaudio.transforms.Resample(orig_freq=rate, new_freq=16000)(wave)
        except (FileNotFoundError, ValueError, RuntimeError):
            wave = torch.zeros(1, 16000)
        mel = eat_preprocess(wave[0], norm_mean=cfg.fbank_mean, norm_std=cfg.fbank_std, target_length=cfg.target_length, fixed_length=cfg.fixed_length,
                             random_crop=cfg.random_crop)
        n_audio = (mel.shape[0] // 2 + 1) // cfg.encoder_projector_ds_rate          # EAT: 2x time down-sampling + CLS, then the k-frame projector
        if self.fix_length_audio > 0:
            n_audio = self.fix_length_audio
        prompt = self.prompt_template.format(cfg.prompt + " ")
        target = record.get("target", None)
        item = common.token_fields(self.tokenizer, n_audio, prompt, None if self.inference_mode else self.answer_template.format(target))
        item.pop("prompt_length")
        item.update(audio=None, audio_mel=mel, audio_length=n_audio, target=target)
        if self.inference_mode:
            item.update(key=record.get("key", None))
        return item

    def pad(self, sequence, max_length, padding_idx=0):
        return common.fit_length(sequence, max_length, padding_idx)

    def collator(self, samples):
        assert samples is not None
        width = max(s["input_ids"].shape[0] for s in samples)
        frames = max(s["audio_mel"].shape[0] for s in samples)
        batch = {"input_ids": torch.stack([self.pad(s["input_ids"], width, self.tokenizer.pad_token_id) for s in samples]),
                 "attention_mask": torch.stack([self.pad(s["attention_mask"], width, False) for s in samples])}
        audio_mel = torch.stack([self.pad(s["audio_mel"], frames, 0) for s in samples])
        batch["modality_mask"] = common.span_mask(batch["attention_mask"], [0] * len(samples), [s["audio_length"] for s in samples])
        if self.inference_mode:
            batch.update(audio_mel=audio_mel if self.input_type == "mel" else None, keys=[s["key"] for s in samples], targets=[s["target"] for s in samples])
            return batch
        batch["labels"] = torch.stack([self.pad(s["labels"], width, self.IGNORE_INDEX) for s in samples])
        batch["audio_mel"] = audio_mel
        batch["audio_mel_mask"] = (torch.arange(frames)[None, :] < torch.tensor([s["audio_mel"].shape[0] for s in samples])[:, None]).float()
        return batch


def get_audio_dataset(dataset_config, tokenizer, split):
    return AudioDatasetJsonl(dataset_config, tokenizer, split)
