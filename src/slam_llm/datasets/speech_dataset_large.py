"""Dynamic-frame multitask dataset (reference: src/slam_llm/datasets/speech_dataset_large.py:23-275) — the batching rule
of BASELINE config 3 (`batching_strategy=dynamic`, `train_max_frame_length`).

Same contract as the reference: an IterableDataset sharded by `line_idx % (num_workers * world) == rank * num_workers + wid`,
items laid out [audio(-1)*L, prompt, answer, eos] (labels -100 outside the answer), a RIGHT-padding collator, and
`MultiTaskDynamicBatchDataset` which flushes its buffer when `(len(buffer) + 1) * max_len(input_ids) > max_frame_length`.

B200 difference (default `dataset_config.b200_gpu_frontend: true`): items carry raw PCM; the collator emits `audio_pcm`
[B, n_max] + `audio_pcm_lengths` and the step computes each utterance's log-mel on the GPU on ITS OWN length
(slam_logmel with lengths: reflect padding at the utterance end, per-utterance max, mel frames beyond the utterance zeroed —
exactly what padding the CPU mel does in the reference collator).  Audio is read from 16 kHz wav paths, or from kaldi
`ark:offset` specifiers when `kaldiio` is installed (file I/O is outside the hot path)."""
import copy
import json
import os
import random
from functools import partial

import numpy as np
import torch
import torch.distributed as dist
import whisper
from torch.utils.data import IterableDataset


def _load_audio(path: str) -> np.ndarray:
    if ".ark" in path:
        try:
            import kaldiio
        except ImportError as e:
            raise ImportError("reading kaldi ark entries needs `kaldiio` (not installed in this image); use wav paths") from e
        return kaldiio.load_mat(path)[1].astype(np.float32) / 32768
    return whisper.load_audio(path)


class MultiTaskDataset(IterableDataset):
    def __init__(self, dataset_config, tokenizer=None, split="train"):
        super().__init__()
        self.multitask_prompt_list = {}
        self.append_info_tasks = dataset_config.append_info_tasks
        with open(dataset_config.multitask_prompt_path) as f_prompt:
            for line in f_prompt:
                item = json.loads(line.strip())
                self.multitask_prompt_list.setdefault(item["task"], []).append(item["prompt"])
        paths = {"train": "train_scp_file_path", "val": "dev_scp_file_path", "test": "test_scp_file_path"}
        if split not in paths:
            raise ValueError("split must be train val test")
        self.data_path = dataset_config[paths[split]]
        self.llm_name = dataset_config.get("llm_name", None)
        self.prompt_template1 = dataset_config.get("prompt_style", "{}")
        self.answer_template = "{}"
        self.dataset_config = dataset_config
        self.tokenizer = tokenizer
        self.split = split
        self.pad_or_trim = dataset_config.get("pad_or_trim", False)
        self.IGNORE_INDEX = -100
        self.mel_size = dataset_config.get("mel_size", 80)
        self.fix_length_audio = dataset_config.get("fix_length_audio", -1)
        self.inference_mode = dataset_config.get("inference_mode", False)
        self.normalize = dataset_config.get("normalize", False)
        self.input_type = dataset_config.get("input_type", None)
        self.max_audio_length = dataset_config.get("max_audio_length", 30)
        self.audio_sample_rate = dataset_config.get("audio_sample_rate", 16000)
        self.gpu_frontend = dataset_config.get("b200_gpu_frontend", True)
        assert self.input_type in ["raw", "mel"], "input_type must be one of [raw, mel]"

    def __iter__(self):
        multitask_task_path = os.path.join(self.data_path, "multitask.jsonl")
        worker_info = torch.utils.data.get_worker_info()
        num_workers, worker_id = (1, 0) if worker_info is None else (worker_info.num_workers, worker_info.id)
        if dist.is_available() and dist.is_initialized():
            world_size, rank = dist.get_world_size(), dist.get_rank()
        else:
            world_size, rank = 1, 0
        total_num_workers = num_workers * world_size
        worker_rank = rank * num_workers + worker_id
        data_index = 0
        with open(multitask_task_path) as f_task:
            for line in f_task:
                if (data_index % total_num_workers) == worker_rank:
                    item = json.loads(line.strip())
                    audio_raw = _load_audio(item["path"])
                    if len(audio_raw) / self.audio_sample_rate > self.max_audio_length:
                        continue                                   # (reference: skipped items do not advance data_index)
                    key, target = item["key"], item["target"]
                    audio_mel = audio_pcm = None
                    if self.input_type == "raw":
                        audio_raw = torch.from_numpy(audio_raw).float()
                        if self.normalize:
                            audio_raw = torch.nn.functional.layer_norm(audio_raw, audio_raw.shape)
                        audio_length = len(audio_raw) // 320 // 5
                    else:
                        if self.pad_or_trim:
                            audio_raw = whisper.pad_or_trim(audio_raw)
                        if self.gpu_frontend:
                            audio_pcm = torch.from_numpy(np.ascontiguousarray(audio_raw, dtype=np.float32))
                            n_frames = audio_pcm.shape[0] // 160
                        else:
                            audio_mel = whisper.log_mel_spectrogram(audio_raw, n_mels=self.mel_size).permute(1, 0)
                            n_frames = audio_mel.shape[0]
                        audio_length = ((n_frames + 1) // 2) // 5
                    if self.fix_length_audio > 0:
                        audio_length = self.fix_length_audio
                    audio_pseudo = torch.full((audio_length,), -1)

                    prompt = random.choice(self.multitask_prompt_list[item["task"]])
                    prompt = self.prompt_template1.format(prompt)
                    if item["task"] in self.append_info_tasks:
                        prompt = prompt.format(item[item["task"]])
                    prompt_ids = self.tokenizer.encode(prompt)
                    prompt_length = len(prompt_ids)
                    common = {"audio": audio_raw if self.input_type == "raw" else None, "audio_mel": audio_mel, "audio_pcm": audio_pcm,
                              "audio_length": audio_length}
                    if self.inference_mode:
                        example_ids = torch.cat((audio_pseudo, torch.tensor(prompt_ids, dtype=torch.int64)))
                        yield {"input_ids": example_ids, "attention_mask": example_ids.ge(-1), "key": key, "target": target, **common}
                    else:
                        example_ids = self.tokenizer.encode(prompt + self.answer_template.format(target))
                        example_ids.append(self.tokenizer.eos_token_id)
                        example_ids = torch.cat((audio_pseudo, torch.tensor(example_ids, dtype=torch.int64)))
                        labels_ids = copy.deepcopy(example_ids)
                        labels_ids[: audio_length + prompt_length] = -1
                        example_mask = example_ids.ge(-1)
                        label_mask = labels_ids.ge(0)
                        example_ids[~example_mask] = 0
                        labels_ids[~label_mask] = self.IGNORE_INDEX
                        yield {"input_ids": example_ids, "labels": labels_ids, "attention_mask": example_mask, **common}
                data_index += 1

    def pad(self, sequence, max_length, padding_idx=0):
        if isinstance(sequence, (int, list, tuple)):
            return sequence + [padding_idx] * (max_length - len(sequence)) if len(sequence) < max_length else sequence[:max_length]
        if isinstance(sequence, torch.Tensor):
            if len(sequence) < max_length:
                return torch.cat((sequence, torch.full([max_length - len(sequence)] + list(sequence.size())[1:], padding_idx, dtype=sequence.dtype)))
            return sequence[:max_length]
        if isinstance(sequence, np.ndarray):
            if len(sequence) < max_length:
                return np.concatenate((sequence, np.full((max_length - len(sequence),) + sequence.shape[1:], padding_idx)))
            return sequence[:max_length]
        raise Exception("Type mismatch during padding!")

    def collator(self, samples):
        assert samples is not None
        n_max = max(s["input_ids"].shape[0] for s in samples)
        input_ids = torch.stack([self.pad(s["input_ids"], n_max, self.tokenizer.pad_token_id) for s in samples])
        attention_mask = torch.stack([self.pad(s["attention_mask"], n_max, False) for s in samples])
        audio_raw = audio_mask = audio_mel = audio_mel_post_mask = audio_pcm = audio_pcm_lengths = None
        if self.input_type == "raw":
            a_max = max(s["audio"].shape[0] for s in samples)
            audio_raw = torch.stack([self.pad(s["audio"], a_max, 0) for s in samples])
            audio_mask = torch.zeros(len(samples), a_max)
            for i, s in enumerate(samples):
                audio_mask[i, : s["audio"].shape[0]] = 1
        elif samples[0].get("audio_pcm") is not None:
            a_max = max(s["audio_pcm"].shape[0] for s in samples)
            a_max = (a_max + 319) // 320 * 320                     # whole mel frames, even count (conv2 stride 2)
            audio_pcm = torch.stack([self.pad(s["audio_pcm"], a_max, 0) for s in samples])
            audio_pcm_lengths = torch.tensor([s["audio_pcm"].shape[0] for s in samples], dtype=torch.int32)
            audio_mel_post_mask = torch.zeros(len(samples), (a_max // 160 + 1) // 2)
            for i, s in enumerate(samples):
                audio_mel_post_mask[i, : (s["audio_pcm"].shape[0] // 160 + 1) // 2] = 1
        else:
            m_max = max(s["audio_mel"].shape[0] for s in samples)
            audio_mel = torch.stack([self.pad(s["audio_mel"], m_max, 0) for s in samples])
            audio_mel_post_mask = torch.zeros(len(samples), (m_max + 1) // 2)
            for i, s in enumerate(samples):
                audio_mel_post_mask[i, : (s["audio_mel"].shape[0] + 1) // 2] = 1
        modality_mask = torch.zeros_like(attention_mask)
        for i, s in enumerate(samples):
            modality_mask[i, : s["audio_length"]] = 1
        out = {"input_ids": input_ids, "attention_mask": attention_mask, "audio": audio_raw, "audio_mask": audio_mask, "audio_mel": audio_mel,
               "audio_pcm": audio_pcm, "audio_pcm_lengths": audio_pcm_lengths, "audio_mel_post_mask": audio_mel_post_mask,
               "modality_mask": modality_mask}
        if self.inference_mode:
            out["keys"] = [s["key"] for s in samples]
            out["targets"] = [s["target"] for s in samples]
            return out
        out["labels"] = torch.stack([self.pad(s["labels"], n_max, self.IGNORE_INDEX) for s in samples])
        return out


class MultiTaskDynamicBatchDataset(IterableDataset):
    def __init__(self, dataset: IterableDataset, window_class) -> None:
        super().__init__()
        self.dp = dataset
        assert window_class is not None
        self.window_class = window_class
        self.collator = self.dp.collator
        self._buffer = []

    def __iter__(self):
        for elem in self.dp:
            if not self.window_class(elem, self._buffer):
                self._buffer.append(elem)
            else:
                if len(self._buffer) > 0:
                    yield self._buffer
                self._buffer = [elem]
        if len(self._buffer) > 0:
            yield self._buffer
        self._buffer = []


def window_class(elem, buffer, max_frame_length):
    """Flush when adding `elem` would push (batch size) x (longest sequence) over the frame budget (:259-263)."""
    if len(buffer) == 0:
        return True
    max_frame = max(len(elem["input_ids"]), max(len(x["input_ids"]) for x in buffer))
    return (len(buffer) + 1) * max_frame > max_frame_length


def get_speech_dataset(dataset_config, tokenizer, split):
    dataset = MultiTaskDataset(dataset_config, tokenizer, split)
    frames = dataset_config.train_max_frame_length if split == "train" else dataset_config.eval_max_frame_length
    return MultiTaskDynamicBatchDataset(dataset, partial(window_class, max_frame_length=frames))
