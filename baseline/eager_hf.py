"""B-EAGER: the reference's eager HuggingFace/PEFT training step on ONE B200 — the denominator of the
north-star "≥6x over the reference's eager HF path".  Baseline code: nothing here is used by the product path.

What the reference executes per step, restated with the modules it would instantiate:
  * models/encoder.py:13-30      conv1/conv2 + GELU, + positional_embedding[:T'], N pre-LN blocks, ln_post — HF `WhisperEncoder` sub-modules
                                 driven variable-length (HF's own forward rejects T != 3000), SDPA attention like current openai-whisper
  * models/projector.py:5-27     EncoderProjectorConcat (fp32 trainable)
  * models/slam_model.py:370-392 embedding + modality merge (python loop + .tolist() sync, as written there)
  * models/slam_model.py:400     HF `LlamaForCausalLM(inputs_embeds, attention_mask, labels)` — eager attention (transformers 4.35.2 default),
                                 fp32 master weights, peft-0.6 LoRA Linear on the target projections
  * models/slam_model.py:402-405 argmax accuracy (utils/metric.py:3-20)
  * utils/train_utils.py:70-76,112-149   torch.autocast(fp16) forward, GradScaler.scale(loss).backward(), scaler.step, scaler.update,
                                 zero_grad, and the per-step tqdm description that formats the loss (one host sync per step)
  * pipeline/finetune.py:247-251 torch.optim.AdamW over model.parameters()
"""
from __future__ import annotations

import math
import time
from typing import Dict

import torch
import torch.nn as nn
import torch.nn.functional as F


class LoraLinear(nn.Module):
    """peft 0.6.0 lora.Linear.forward: F.linear(x, W) + lora_B(lora_A(dropout(x))) * alpha / r."""

    def __init__(self, base: nn.Linear, r: int, alpha: int, dropout: float, b_std: float):
        super().__init__()
        self.weight, self.bias = base.weight, base.bias
        self.weight.requires_grad = False
        self.lora_A = nn.Linear(base.in_features, r, bias=False, device=base.weight.device)
        self.lora_B = nn.Linear(r, base.out_features, bias=False, device=base.weight.device)
        nn.init.kaiming_uniform_(self.lora_A.weight, a=math.sqrt(5))
        nn.init.normal_(self.lora_B.weight, std=b_std) if b_std > 0 else nn.init.zeros_(self.lora_B.weight)
        self.scaling = alpha / r
        self.dropout = nn.Dropout(dropout) if dropout > 0 else nn.Identity()

    def forward(self, x):
        result = F.linear(x, self.weight, self.bias)
        return result + self.lora_B(self.lora_A(self.dropout(x.to(self.lora_A.weight.dtype)))) * self.scaling


class EagerSlam(nn.Module):
    def __init__(self, enc_cfg, llm_cfg, lora_cfg, k: int = 5, hidden: int = 2048, device="cuda", lora_b_std: float = 0.02, seed: int = 42):
        super().__init__()
        from transformers import LlamaConfig, LlamaForCausalLM, WhisperConfig
        from transformers.models.whisper.modeling_whisper import WhisperEncoder
        torch.manual_seed(seed)
        wc = WhisperConfig(num_mel_bins=enc_cfg.n_mels, d_model=enc_cfg.d, encoder_layers=enc_cfg.layers, encoder_attention_heads=enc_cfg.heads,
                           encoder_ffn_dim=4 * enc_cfg.d, max_source_positions=enc_cfg.n_ctx, dropout=0.0, attention_dropout=0.0, activation_dropout=0.0)
        wc._attn_implementation = "sdpa"
        lc = LlamaConfig(vocab_size=llm_cfg.vocab, hidden_size=llm_cfg.d, intermediate_size=llm_cfg.ffn, num_hidden_layers=llm_cfg.layers,
                         num_attention_heads=llm_cfg.heads, num_key_value_heads=llm_cfg.kv_heads, rms_norm_eps=llm_cfg.eps, rope_theta=llm_cfg.rope_theta,
                         max_position_embeddings=8192, attention_bias=False, mlp_bias=False, tie_word_embeddings=False, use_cache=False,
                         attn_implementation="eager")
        with torch.device(device):
            self.encoder = WhisperEncoder(wc).eval()
            self.llm = LlamaForCausalLM(lc)
        for p in self.encoder.parameters():
            p.requires_grad = False
        for p in self.llm.parameters():                               # freeze_llm (slam_model.py:205-208)
            p.requires_grad = False
        self.llm.eval()
        if lora_cfg is not None:
            for layer in self.llm.model.layers:
                for name in lora_cfg.targets:
                    parent = layer.self_attn if name in ("q_proj", "k_proj", "v_proj", "o_proj") else layer.mlp
                    setattr(parent, name, LoraLinear(getattr(parent, name), lora_cfg.r, lora_cfg.alpha, getattr(lora_cfg, "dropout", 0.0), lora_b_std))
        self.k = k
        self.linear1 = nn.Linear(enc_cfg.d * k, hidden, device=device)
        self.relu = nn.ReLU()
        self.linear2 = nn.Linear(hidden, llm_cfg.d, device=device)

    def extract_variable_length_features(self, x):                     # models/encoder.py:13-30 on the HF sub-modules
        e = self.encoder
        x = F.gelu(e.conv1(x))
        x = F.gelu(e.conv2(x))
        x = x.permute(0, 2, 1)
        x = (x + e.embed_positions.weight[: x.shape[1]]).to(x.dtype)
        for layer in e.layers:
            out = layer(x, None)
            x = out[0] if isinstance(out, tuple) else out
        return e.layer_norm(x)

    def projector(self, x):                                            # models/projector.py:15-27
        b, t, d = x.size()
        drop = t % self.k
        if drop > 0:
            x = x[:, :-drop, :]
        x = x.contiguous().view(b, x.size(1) // self.k, d * self.k)
        return self.linear2(self.relu(self.linear1(x)))

    def forward(self, input_ids, attention_mask, labels, modality_mask, audio_mel):
        with torch.no_grad():
            self.encoder.eval()
            encoder_outs = self.extract_variable_length_features(audio_mel.permute(0, 2, 1))
        encoder_outs = self.projector(encoder_outs)
        input_ids = input_ids.clone()
        input_ids[input_ids == -1] = 0
        inputs_embeds = self.llm.model.embed_tokens(input_ids)
        start = (modality_mask == True).float().argmax(dim=1)          # noqa: E712  (slam_model.py:382-392)
        lengths = torch.clamp(modality_mask.sum(dim=1), max=encoder_outs.shape[1]).tolist()
        pad = torch.zeros_like(inputs_embeds)
        for i in range(encoder_outs.shape[0]):
            pad[i, start[i]:start[i] + lengths[i]] = encoder_outs[i][:lengths[i]]
        inputs_embeds = pad + inputs_embeds * (~modality_mask[:, :, None])
        out = self.llm(inputs_embeds=inputs_embeds, attention_mask=attention_mask, labels=labels)
        with torch.no_grad():
            preds = torch.argmax(out.logits, -1)
            mask = labels[:, 1:] != -100
            acc = (preds[:, :-1].masked_select(mask) == labels[:, 1:].masked_select(mask)).sum().float() / mask.sum().float()
        return out, acc


def run(enc_cfg, llm_cfg, lora_cfg, host_batch: Dict[str, torch.Tensor], audio_mel: torch.Tensor, steps: int, warmup: int, device, lr: float = 1e-4):
    """Returns ms per step (CUDA-event timed, batch resident on the device incl. the CPU-side log-mel the reference's DataLoader produces)."""
    model = EagerSlam(enc_cfg, llm_cfg, lora_cfg, device=device)
    model.train()                                                      # train_utils.py:92 (LoRA dropout active when configured; 0 here)
    optimizer = torch.optim.AdamW(model.parameters(), lr=lr, weight_decay=0.0)
    scaler = torch.amp.GradScaler("cuda")
    batch = dict(input_ids=host_batch["input_ids"].to(device), attention_mask=host_batch["attention_mask"].to(device),
                 labels=host_batch["labels"].to(device), modality_mask=host_batch["modality_mask"].to(device), audio_mel=audio_mel.to(device))

    def step():
        with torch.autocast("cuda", dtype=torch.float16):
            outputs, acc = model(**batch)
        loss = outputs.loss
        scaler.scale(loss).backward()
        scaler.step(optimizer)
        scaler.update()
        optimizer.zero_grad()
        return f"(loss: {loss.detach().float()}, acc: {acc})"          # the tqdm description of train_utils.py:171 (host sync)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        desc = step()
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3 / steps
    ms = e0.elapsed_time(e1) / steps
    peak_gb = torch.cuda.max_memory_allocated() / 2 ** 30
    del model, optimizer
    torch.cuda.empty_cache()
    return dict(ms_per_step=ms, wall_ms_per_step=wall, last=desc, peak_mem_gb=round(peak_gb, 1))
