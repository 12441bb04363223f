"""jsonl speech dataset + collator: the batch CONTRACT of the training step
(reference: src/slam_llm/datasets/speech_dataset.py:17-298).

Token layout [audio(-1)*L, prompt, answer, eos], labels -100 outside answer+eos, prompt side LEFT-padded and answer side
RIGHT-padded by the collator, `modality_mask` marking the audio span - all as in the reference; the mechanics live in
`_b200_common` and are shared with the dynamic-frame dataset.

B200 difference: with input_type == "mel" the reference computes whisper.log_mel_spectrogram on the CPU inside the
DataLoader workers; here (default, `dataset_config.b200_gpu_frontend: true`) the item carries the padded raw waveform
(`audio_pcm`, f32 [480000]) and the log-mel runs on the GPU inside the step (slam_logmel) - the H2D copy is 1.9 MB of PCM
per utterance instead of 1.0-1.5 MB of mel, and the CPU front end stops being the 8-GPU bottleneck.
Set b200_gpu_frontend=false to get the reference's CPU `audio_mel` items.
"""
import json

import torch
import whisper

from slam_llm.datasets import _b200_common as common       # absolute: recipes load this file by PATH (dataset_config.file), outside the package


class SpeechDatasetJsonl(torch.utils.data.Dataset):
    IGNORE_INDEX = common.IGNORE_INDEX
    prompt_template = "USER: {}\n ASSISTANT:"
    answer_template = "{}"

    def __init__(self, dataset_config, tokenizer=None, split="train"):
        super().__init__()
        cfg = dataset_config
        self.dataset_config, self.tokenizer = cfg, tokenizer
        self.prompt = cfg.get("prompt", None)
        self.mel_size = cfg.get("mel_size", 80)
        self.fix_length_audio = cfg.get("fix_length_audio", -1)
        self.inference_mode = cfg.get("inference_mode", False)
        self.normalize = cfg.get("normalize", False)
        self.input_type = cfg.get("input_type", None)
        self.gpu_frontend = cfg.get("b200_gpu_frontend", True)
        assert self.input_type in ["raw", "mel"], "input_type must be one of [raw, mel]"
        with open(cfg.train_data_path if split == "train" else cfg.val_data_path, encoding="utf-8") as fin:
            self.data_list = [json.loads(line.strip()) for line in fin]

    # ---- length hints used by length-grouped samplers
    def get_source_len(self, data_dict):
        return data_dict["source_len"]

    def get_target_len(self, data_dict):
        return data_dict.get("target_len", 0)

    def __len__(self):
        return len(self.data_list)

    def __getitem__(self, index):
        record = self.data_list[index]
        audio, n_audio = common.audio_fields(whisper.load_audio(record.get("source")), input_type=self.input_type, normalize=self.normalize,
                                             pad_or_trim=True, gpu_frontend=self.gpu_frontend, mel_size=self.mel_size)
        if self.fix_length_audio > 0:
            n_audio = self.fix_length_audio
        prompt = self.prompt_template.format(common.DEFAULT_ASR_PROMPT if self.prompt is None else self.prompt)
        target = record.get("target", None)
        item = common.token_fields(self.tokenizer, n_audio, prompt, None if self.inference_mode else self.answer_template.format(target))
        item.update(audio, audio_length=n_audio)
        if self.inference_mode:
            item.update(key=record.get("key", None), target=target)
        return item

    # ---- the reference's padding helpers (recipes call them on the dataset object)
    def pad(self, sequence, max_length, padding_idx=0):
        return common.fit_length(sequence, max_length, padding_idx)

    @classmethod
    def padding(cls, sequence, padding_length, padding_idx=0, padding_side="right"):
        return common.extend_by(sequence, padding_length, padding_idx, padding_side)

    def collator(self, samples):
        assert samples is not None
        head = [s["audio_length"] + s["prompt_length"] for s in samples]            # left-padded part: audio + prompt
        tail = [len(s["input_ids"]) - h for s, h in zip(samples, head)]              # right-padded part: answer + eos
        head_max, tail_max = max(head), max(tail)

        def aligned(field, fill):
            rows = (self.padding(self.padding(s[field], head_max - h, fill, padding_side="left"), tail_max - t, fill)
                    for s, h, t in zip(samples, head, tail))
            return torch.stack(list(rows))

        batch = {"input_ids": aligned("input_ids", self.tokenizer.pad_token_id), "attention_mask": aligned("attention_mask", False)}
        audio = common.collate_audio(samples, self.input_type)
        audio.pop("audio_pcm_lengths")                                               # fixed 30 s items: no per-utterance lengths
        batch.update(audio)
        batch["modality_mask"] = common.span_mask(batch["attention_mask"], [head_max - h for h in head], [s["audio_length"] for s in samples])
        if self.inference_mode:
            batch["keys"] = [s["key"] for s in samples]
            batch["targets"] = [s["target"] for s in samples]
        else:
            batch["labels"] = aligned("labels", self.IGNORE_INDEX)
        return batch


def get_speech_dataset(dataset_config, tokenizer, split):
    return SpeechDatasetJsonl(dataset_config, tokenizer, split)
