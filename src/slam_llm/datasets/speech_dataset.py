"""jsonl speech dataset + collator: the batch CONTRACT of the training step
(reference: src/slam_llm/datasets/speech_dataset.py:17-298).

Token layout [audio(-1)*L, prompt, answer, eos], labels -100 outside answer+eos, prompt side LEFT-padded and answer side
RIGHT-padded by the collator, `modality_mask` marking the audio span — all as in the reference.

B200 difference: with input_type == "mel" the reference computes whisper.log_mel_spectrogram on the CPU inside the
DataLoader workers; here (default, `dataset_config.b200_gpu_frontend: true`) the item carries the padded raw waveform
(`audio_pcm`, f32 [480000]) and the log-mel runs on the GPU inside the step (slam_logmel) — the H2D copy is 1.9 MB of PCM
per utterance instead of 1.0-1.5 MB of mel, and the CPU front end stops being the 8-GPU bottleneck.
Set b200_gpu_frontend=false to get the reference's CPU `audio_mel` items.
"""
import copy
import json

import numpy as np
import torch
import whisper


class SpeechDatasetJsonl(torch.utils.data.Dataset):
    def __init__(self, dataset_config, tokenizer=None, split="train"):
        super().__init__()
        self.dataset_config = dataset_config
        self.tokenizer = tokenizer
        self.IGNORE_INDEX = -100
        self.prompt = dataset_config.get("prompt", None)
        self.mel_size = dataset_config.get("mel_size", 80)
        self.prompt_template = "USER: {}\n ASSISTANT:"
        self.answer_template = "{}"
        self.fix_length_audio = dataset_config.get("fix_length_audio", -1)
        self.inference_mode = dataset_config.get("inference_mode", False)
        self.normalize = dataset_config.get("normalize", False)
        self.input_type = dataset_config.get("input_type", None)
        self.gpu_frontend = dataset_config.get("b200_gpu_frontend", True)
        assert self.input_type in ["raw", "mel"], "input_type must be one of [raw, mel]"
        path = dataset_config.train_data_path if split == "train" else dataset_config.val_data_path
        self.data_list = []
        with open(path, encoding="utf-8") as fin:
            for line in fin:
                self.data_list.append(json.loads(line.strip()))

    def get_source_len(self, data_dict):
        return data_dict["source_len"]

    def get_target_len(self, data_dict):
        return data_dict["target_len"] if "target_len" in data_dict else 0

    def __len__(self):
        return len(self.data_list)

    def __getitem__(self, index):
        data_dict = self.data_list[index]
        audio_path = data_dict.get("source")
        target = data_dict.get("target", None)
        key = data_dict.get("key", None)

        audio_raw = whisper.load_audio(audio_path)
        audio_mel = audio_pcm = None
        if self.input_type == "raw":
            audio_raw = torch.from_numpy(audio_raw)
            if self.normalize:
                audio_raw = torch.nn.functional.layer_norm(audio_raw, audio_raw.shape)
            audio_length = len(audio_raw) // 320 // 5
        else:
            audio_raw = whisper.pad_or_trim(audio_raw)
            if self.gpu_frontend:
                audio_pcm = torch.from_numpy(np.ascontiguousarray(audio_raw, dtype=np.float32))
                n_frames = audio_pcm.shape[0] // 160
            else:
                audio_mel = whisper.log_mel_spectrogram(audio_raw, n_mels=self.mel_size).permute(1, 0)
                n_frames = audio_mel.shape[0]
            audio_length = ((n_frames + 1) // 2) // 5            # whisper 2x conv downsample, then 5x projector
        if self.fix_length_audio > 0:
            audio_length = self.fix_length_audio
        audio_pseudo = torch.full((audio_length,), -1)

        prompt = self.prompt
        if prompt is None:
            prompt = ("Transcribe speech to text. Output the transcription directly without redundant content. "
                      "Ensure that the output is not duplicated. ")
        prompt = self.prompt_template.format(prompt)
        prompt_ids = self.tokenizer.encode(prompt)
        prompt_length = len(prompt_ids)

        if self.inference_mode:
            prompt_ids = torch.tensor(prompt_ids, dtype=torch.int64)
            example_ids = torch.cat((audio_pseudo, prompt_ids))
            return {"input_ids": example_ids, "attention_mask": example_ids.ge(-1), "audio": audio_raw if self.input_type == "raw" else None,
                    "audio_mel": audio_mel, "audio_pcm": audio_pcm, "audio_length": audio_length, "key": key, "target": target,
                    "prompt_length": prompt_length}

        example = prompt + self.answer_template.format(target)
        example_ids = self.tokenizer.encode(example)
        example_ids.append(self.tokenizer.eos_token_id)
        example_ids = torch.cat((audio_pseudo, torch.tensor(example_ids, dtype=torch.int64)))
        labels_ids = copy.deepcopy(example_ids)
        labels_ids[: audio_length + prompt_length] = -1
        example_mask = example_ids.ge(-1)
        label_mask = labels_ids.ge(0)
        example_ids[~example_mask] = 0
        labels_ids[~label_mask] = self.IGNORE_INDEX
        return {"input_ids": example_ids, "labels": labels_ids, "attention_mask": example_mask,
                "audio": audio_raw if self.input_type == "raw" else None, "audio_mel": audio_mel, "audio_pcm": audio_pcm,
                "audio_length": audio_length, "prompt_length": prompt_length}

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

    @classmethod
    def padding(cls, sequence, padding_length, padding_idx=0, padding_side="right"):
        if isinstance(sequence, (int, list, tuple)):
            return sequence + [padding_idx] * padding_length if padding_length >= 0 else sequence[:padding_length]
        if isinstance(sequence, torch.Tensor):
            if sequence.ndimension() == 2:
                return torch.nn.functional.pad(sequence, (0, padding_length)) if padding_length >= 0 else sequence[:, :padding_length]
            if padding_length >= 0:
                filler = torch.full([padding_length] + list(sequence.size())[1:], padding_idx, dtype=sequence.dtype)
                return torch.cat((filler, sequence)) if padding_side == "left" else torch.cat((sequence, filler))
            return sequence[:padding_length]
        if isinstance(sequence, np.ndarray):
            if padding_length >= 0:
                return np.concatenate((sequence, np.full((padding_length,) + sequence.shape[1:], padding_idx)))
            return sequence[:padding_length]
        raise Exception("Type mismatch during padding!")

    def collator(self, samples):
        assert samples is not None
        prompt_lens = [s["audio_length"] + s["prompt_length"] for s in samples]
        answer_lens = [len(s["input_ids"]) - s["audio_length"] - s["prompt_length"] for s in samples]
        pmax, amax = max(prompt_lens), max(answer_lens)
        pad_id = self.tokenizer.pad_token_id

        def both(key, fill):
            return torch.stack([self.padding(self.padding(s[key], pmax - prompt_lens[i], fill, padding_side="left"), amax - answer_lens[i], fill)
                                for i, s in enumerate(samples)])

        input_ids = both("input_ids", pad_id)
        attention_mask = both("attention_mask", False)
        audio_raw = audio_mask = audio_mel = audio_mel_post_mask = audio_pcm = None
        if self.input_type == "raw":
            n = max(s["audio"].shape[0] for s in samples)
            audio_raw = torch.stack([self.pad(s["audio"], n, 0) for s in samples])
            audio_mask = torch.zeros(len(samples), n)
            for i, s in enumerate(samples):
                audio_mask[i, : s["audio"].shape[0]] = 1
        elif samples[0].get("audio_pcm") is not None:
            n = max(s["audio_pcm"].shape[0] for s in samples)
            audio_pcm = torch.stack([self.pad(s["audio_pcm"], n, 0) for s in samples])
            t = n // 160
            audio_mel_post_mask = torch.zeros(len(samples), (t + 1) // 2)
            for i, s in enumerate(samples):
                audio_mel_post_mask[i, : (s["audio_pcm"].shape[0] // 160 + 1) // 2] = 1
        else:
            n = max(s["audio_mel"].shape[0] for s in samples)
            audio_mel = torch.stack([self.pad(s["audio_mel"], n, 0) for s in samples])
            audio_mel_post_mask = torch.zeros(len(samples), (n + 1) // 2)
            for i, s in enumerate(samples):
                audio_mel_post_mask[i, : (s["audio_mel"].shape[0] + 1) // 2] = 1

        modality_mask = torch.zeros_like(attention_mask)
        for i, s in enumerate(samples):
            left = pmax - prompt_lens[i]
            modality_mask[i, left: left + s["audio_length"]] = True

        out = {"input_ids": input_ids, "attention_mask": attention_mask, "audio": audio_raw, "audio_mask": audio_mask, "audio_mel": audio_mel,
               "audio_pcm": audio_pcm, "audio_mel_post_mask": audio_mel_post_mask, "modality_mask": modality_mask}
        if self.inference_mode:
            out["keys"] = [s["key"] for s in samples]
            out["targets"] = [s["target"] for s in samples]
            return out
        out["labels"] = both("labels", self.IGNORE_INDEX)
        return out


def get_speech_dataset(dataset_config, tokenizer, split):
    return SpeechDatasetJsonl(dataset_config, tokenizer, split)
