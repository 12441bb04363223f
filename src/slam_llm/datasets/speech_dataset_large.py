"""Dynamic-frame multitask dataset (reference: src/slam_llm/datasets/speech_dataset_large.py:23-275) - the batching rule
of BASELINE config 3 (`batching_strategy=dynamic`, `train_max_frame_length`).

Same contract as the reference: an IterableDataset sharded by `line_idx % (num_workers * world) == rank * num_workers + wid`,
items laid out [audio(-1)*L, prompt, answer, eos] (labels -100 outside the answer), a RIGHT-padding collator, and
`MultiTaskDynamicBatchDataset` which flushes its buffer when `(len(buffer) + 1) * max_len(input_ids) > max_frame_length`.
Item / batch mechanics are shared with the jsonl dataset (`_b200_common`).

B200 difference (default `dataset_config.b200_gpu_frontend: true`): items carry raw PCM; the collator emits `audio_pcm`
[B, n_max] + `audio_pcm_lengths` and the step computes each utterance's log-mel on the GPU on ITS OWN length
(slam_logmel with lengths: reflect padding at the utterance end, per-utterance max, mel frames beyond the utterance zeroed -
exactly what padding the CPU mel does in the reference collator).  Audio is read from 16 kHz wav paths, or from kaldi
`ark:offset` specifiers when `kaldiio` is installed (file I/O is outside the hot path)."""
import collections
import json
import os
import random
from functools import partial

import numpy as np
import torch
import torch.distributed as dist
import whisper
from torch.utils.data import IterableDataset

from slam_llm.datasets import _b200_common as common       # absolute: recipes load this file by PATH (dataset_config.file), outside the package

_SPLIT_KEYS = {"train": "train_scp_file_path", "val": "dev_scp_file_path", "test": "test_scp_file_path"}


def _read_wave(path: str) -> np.ndarray:
    if ".ark" not in path:
        return whisper.load_audio(path)
    try:
        import kaldiio
    except ImportError as e:
        raise ImportError("reading kaldi ark entries needs `kaldiio` (not installed in this image); use wav paths") from e
    return kaldiio.load_mat(path)[1].astype(np.float32) / 32768


def _shard_of_this_worker():
    """(index of this DataLoader worker among all workers of all ranks, their total number)."""
    info = torch.utils.data.get_worker_info()
    workers, wid = (1, 0) if info is None else (info.num_workers, info.id)
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank() * workers + wid, workers * dist.get_world_size()
    return wid, workers


class MultiTaskDataset(IterableDataset):
    IGNORE_INDEX = common.IGNORE_INDEX
    answer_template = "{}"

    def __init__(self, dataset_config, tokenizer=None, split="train"):
        super().__init__()
        cfg = dataset_config
        if split not in _SPLIT_KEYS:
            raise ValueError("split must be train val test")
        self.dataset_config, self.tokenizer, self.split = cfg, tokenizer, split
        self.data_path = cfg[_SPLIT_KEYS[split]]
        prompts = collections.defaultdict(list)
        with open(cfg.multitask_prompt_path) as f_prompt:
            for line in f_prompt:
                entry = json.loads(line.strip())
                prompts[entry["task"]].append(entry["prompt"])
        self.multitask_prompt_list = dict(prompts)
        self.append_info_tasks = cfg.append_info_tasks
        self.llm_name = cfg.get("llm_name", None)
        self.prompt_template1 = cfg.get("prompt_style", "{}")
        self.pad_or_trim = cfg.get("pad_or_trim", False)
        self.mel_size = cfg.get("mel_size", 80)
        self.fix_length_audio = cfg.get("fix_length_audio", -1)
        self.inference_mode = cfg.get("inference_mode", False)
        self.normalize = cfg.get("normalize", False)
        self.input_type = cfg.get("input_type", None)
        self.max_audio_length = cfg.get("max_audio_length", 30)
        self.audio_sample_rate = cfg.get("audio_sample_rate", 16000)
        self.gpu_frontend = cfg.get("b200_gpu_frontend", True)
        assert self.input_type in ["raw", "mel"], "input_type must be one of [raw, mel]"

    def _prompt_for(self, record) -> str:
        task = record["task"]
        text = self.prompt_template1.format(random.choice(self.multitask_prompt_list[task]))
        return text.format(record[task]) if task in self.append_info_tasks else text

    def __iter__(self):
        mine, shards = _shard_of_this_worker()
        position = 0                                   # counts only utterances that pass the length filter (as the reference does)
        with open(os.path.join(self.data_path, "multitask.jsonl")) as f_task:
            for line in f_task:
                if position % shards != mine:
                    position += 1
                    continue
                record = json.loads(line.strip())
                wave = _read_wave(record["path"])
                if len(wave) / self.audio_sample_rate > self.max_audio_length:
                    continue
                position += 1
                audio, n_audio = common.audio_fields(wave, input_type=self.input_type, normalize=self.normalize, pad_or_trim=self.pad_or_trim,
                                                     gpu_frontend=self.gpu_frontend, mel_size=self.mel_size)
                if self.fix_length_audio > 0:
                    n_audio = self.fix_length_audio
                answer = None if self.inference_mode else self.answer_template.format(record["target"])
                item = common.token_fields(self.tokenizer, n_audio, self._prompt_for(record), answer)
                item.pop("prompt_length")
                item.update(audio, audio_length=n_audio)
                if self.inference_mode:
                    item.update(key=record["key"], target=record["target"])
                yield item

    def pad(self, sequence, max_length, padding_idx=0):
        return common.fit_length(sequence, max_length, padding_idx)

    def collator(self, samples):
        assert samples is not None
        width = max(s["input_ids"].shape[0] for s in samples)

        def right_padded(field, fill):
            return torch.stack([self.pad(s[field], width, fill) for s in samples])

        batch = {"input_ids": right_padded("input_ids", self.tokenizer.pad_token_id), "attention_mask": right_padded("attention_mask", False)}
        # PCM right-padded to the longest utterance: the step's log-mel then has max(n_i // 160) frames, exactly the reference's padded mel width
        batch.update(common.collate_audio(samples, self.input_type))
        batch["modality_mask"] = common.span_mask(batch["attention_mask"], [0] * len(samples), [s["audio_length"] for s in samples])
        if self.inference_mode:
            batch["keys"] = [s["key"] for s in samples]
            batch["targets"] = [s["target"] for s in samples]
        else:
            batch["labels"] = right_padded("labels", self.IGNORE_INDEX)
        return batch


class MultiTaskDynamicBatchDataset(IterableDataset):
    """Groups the items of `dataset` into variable-size batches with the reference's window rule."""

    def __init__(self, dataset: IterableDataset, window_class) -> None:
        super().__init__()
        assert window_class is not None
        self.dp, self.window_class = dataset, window_class
        self.collator = dataset.collator
        self._buffer = []

    def __iter__(self):
        pending = self._buffer = []
        for item in self.dp:
            if self.window_class(item, pending) and pending:
                yield pending
                pending = self._buffer = []
            pending.append(item)
        if pending:
            yield pending
        self._buffer = []


def window_class(elem, buffer, max_frame_length):
    """True when `elem` must open a new batch: adding it would push (batch size) x (longest sequence) over the frame budget
    (speech_dataset_large.py:259-263); an empty buffer always "flushes" (nothing is emitted for it)."""
    if not buffer:
        return True
    longest = max(len(elem["input_ids"]), *(len(x["input_ids"]) for x in buffer))
    return (len(buffer) + 1) * longest > max_frame_length


def get_speech_dataset(dataset_config, tokenizer, split):
    frames = dataset_config.train_max_frame_length if split == "train" else dataset_config.eval_max_frame_length
    return MultiTaskDynamicBatchDataset(MultiTaskDataset(dataset_config, tokenizer, split), partial(window_class, max_frame_length=frames))
