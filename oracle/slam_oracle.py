"""CPU oracle for the SLAM-LLM training-step hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product path (slam_llm_b200/, src/slam_llm/) never does.

It restates, in plain PyTorch fp32/fp64 on the CPU, the arithmetic the reference executes per step:

  * log-mel              whisper.log_mel_spectrogram (called at src/slam_llm/datasets/speech_dataset.py:101-103)
  * Whisper encoder      src/slam_llm/models/encoder.py:13-30 driving whisper.model.AudioEncoder modules
  * projector            src/slam_llm/models/projector.py:5-27 (concat-linear) and :29-49 (cov1d-linear)
  * embedding merge      src/slam_llm/models/slam_model.py:370-392
  * Llama decoder + LoRA transformers v4.35.2 modeling_llama.py (LlamaRMSNorm, rotary, repeat_kv, eager attention,
                         LlamaMLP, LlamaForCausalLM loss) + peft v0.6.0 lora.Linear.forward, called at
                         src/slam_llm/models/slam_model.py:400 (model built at :181-186, peft wrap :214-218)
  * accuracy             src/slam_llm/utils/metric.py:3-20
  * optimizer / schedule src/slam_llm/pipeline/finetune.py:247-260 (torch.optim.AdamW + LambdaLR)

openai-whisper, transformers==4.35.2 and peft==0.6.0 are third-party dependencies that are NOT vendored
under /root/reference; their published algorithms are restated here from formula.
PINNING (round 2): the reference's own tests hold no golden vector for this path, so the oracle is pinned to
OUTPUTS OF THE REFERENCE ITSELF RUN IN THE BUILD CONTAINER: tests/golden/make_ref_golden.py imports /root/reference/src/slam_llm unmodified
(through tests/ref_glue.py's stand-ins for the absent peft / openai-whisper / kaldiio packages), runs its datasets + collators,
setup_encoder / setup_llm / setup_encoder_projector, slam_model.forward (and examples/s2s/model/slam_model_s2s.py), backward and
torch.optim.AdamW on the CPU and commits the results as tests/golden/ref_*.pt; tests/test_ref_pinning.py requires this oracle to
reproduce them at fp32 level (loss 1e-5, activations 1e-4, gradients 1e-4), including Whisper-large-v3 / Llama-3-8B widths at depth 1.
The third-party arithmetic is additionally pinned to the installed transformers 5.5.0 modules (WhisperFeatureExtractor, WhisperEncoder
layers, LlamaForCausalLM / Qwen2ForCausalLM eager attention + loss) in tests/test_oracle_pinning.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

# ------------------------------------------------------------------------------------------------------------------
# configuration
# ------------------------------------------------------------------------------------------------------------------


@dataclass
class EncoderCfg:
    n_mels: int = 80
    n_ctx: int = 1500
    d: int = 384
    heads: int = 6
    layers: int = 4


@dataclass
class LlmCfg:
    vocab: int = 32000
    d: int = 2048
    layers: int = 22
    heads: int = 32
    kv_heads: int = 4
    ffn: int = 5632
    rope_theta: float = 10000.0
    eps: float = 1e-5
    qkv_bias: bool = False         # Qwen2: biases on q/k/v projections
    tie_embeddings: bool = False   # Qwen2-0.5B: lm_head shares the embedding table

    @property
    def dh(self) -> int:
        return self.d // self.heads


@dataclass
class LoraCfg:
    r: int = 8
    alpha: int = 32
    targets: Tuple[str, ...] = ("q_proj", "v_proj")
    dropout: float = 0.0  # parity runs use 0

    @property
    def scaling(self) -> float:
        return self.alpha / self.r


@dataclass
class ProjCfg:
    kind: str = "linear"  # "linear" (EncoderProjectorConcat) | "cov1d-linear" (EncoderProjectorCov1d)
    k: int = 5
    hidden: int = 2048


WHISPER = {
    "tiny": EncoderCfg(80, 1500, 384, 6, 4),
    "base": EncoderCfg(80, 1500, 512, 8, 6),
    "small": EncoderCfg(80, 1500, 768, 12, 12),
    "medium": EncoderCfg(80, 1500, 1024, 16, 24),
    "large-v3": EncoderCfg(128, 1500, 1280, 20, 32),
}
LLM = {
    "tinyllama-1.1b": LlmCfg(32000, 2048, 22, 32, 4, 5632, 10000.0, 1e-5),
    "llama-3-8b": LlmCfg(128256, 4096, 32, 32, 8, 14336, 500000.0, 1e-5),
    "vicuna-7b": LlmCfg(32000, 4096, 32, 32, 32, 11008, 10000.0, 1e-5),
}

LLM_LINEARS = ("q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj")


def linear_shape(cfg: LlmCfg, name: str) -> Tuple[int, int]:
    """(out_features, in_features) of a decoder linear."""
    kv = cfg.kv_heads * cfg.dh
    return {
        "q_proj": (cfg.d, cfg.d), "k_proj": (kv, cfg.d), "v_proj": (kv, cfg.d), "o_proj": (cfg.d, cfg.d),
        "gate_proj": (cfg.ffn, cfg.d), "up_proj": (cfg.ffn, cfg.d), "down_proj": (cfg.d, cfg.ffn),
    }[name]


# ------------------------------------------------------------------------------------------------------------------
# synthetic weights (random init of the reference architectures; HF-style N(0, 0.02))
# ------------------------------------------------------------------------------------------------------------------


def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> torch.Tensor:
    inc = math.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2, dtype=torch.float32))
    t = torch.arange(length, dtype=torch.float32)[:, None] * inv[None, :]
    return torch.cat([torch.sin(t), torch.cos(t)], dim=1)


def init_encoder(cfg: EncoderCfg, seed: int = 42, std: float = 0.02) -> Dict[str, torch.Tensor]:
    """openai-whisper AudioEncoder state-dict names (what whisper.load_model(...).encoder holds)."""
    g = torch.Generator().manual_seed(seed)

    def n(*s):
        return torch.randn(*s, generator=g) * std

    w: Dict[str, torch.Tensor] = {}
    w["conv1.weight"], w["conv1.bias"] = n(cfg.d, cfg.n_mels, 3) * 3, n(cfg.d)
    w["conv2.weight"], w["conv2.bias"] = n(cfg.d, cfg.d, 3), n(cfg.d)
    w["positional_embedding"] = sinusoids(cfg.n_ctx, cfg.d)
    for i in range(cfg.layers):
        p = f"blocks.{i}."
        w[p + "attn.query.weight"], w[p + "attn.query.bias"] = n(cfg.d, cfg.d), n(cfg.d)
        w[p + "attn.key.weight"] = n(cfg.d, cfg.d)
        w[p + "attn.value.weight"], w[p + "attn.value.bias"] = n(cfg.d, cfg.d), n(cfg.d)
        w[p + "attn.out.weight"], w[p + "attn.out.bias"] = n(cfg.d, cfg.d), n(cfg.d)
        w[p + "attn_ln.weight"], w[p + "attn_ln.bias"] = 1.0 + n(cfg.d), n(cfg.d)
        w[p + "mlp.0.weight"], w[p + "mlp.0.bias"] = n(4 * cfg.d, cfg.d), n(4 * cfg.d)
        w[p + "mlp.2.weight"], w[p + "mlp.2.bias"] = n(cfg.d, 4 * cfg.d), n(cfg.d)
        w[p + "mlp_ln.weight"], w[p + "mlp_ln.bias"] = 1.0 + n(cfg.d), n(cfg.d)
    w["ln_post.weight"], w["ln_post.bias"] = 1.0 + n(cfg.d), n(cfg.d)
    return w


def init_llm(cfg: LlmCfg, seed: int = 43, std: float = 0.02) -> Dict[str, torch.Tensor]:
    """HF LlamaForCausalLM state-dict names."""
    g = torch.Generator().manual_seed(seed)

    def n(*s):
        return torch.randn(*s, generator=g) * std

    w: Dict[str, torch.Tensor] = {"model.embed_tokens.weight": n(cfg.vocab, cfg.d)}
    for i in range(cfg.layers):
        p = f"model.layers.{i}."
        for name in ("q_proj", "k_proj", "v_proj", "o_proj"):
            w[p + f"self_attn.{name}.weight"] = n(*linear_shape(cfg, name))
            if cfg.qkv_bias and name != "o_proj":
                w[p + f"self_attn.{name}.bias"] = n(linear_shape(cfg, name)[0]) * 5
        for name in ("gate_proj", "up_proj", "down_proj"):
            w[p + f"mlp.{name}.weight"] = n(*linear_shape(cfg, name))
        w[p + "input_layernorm.weight"] = 1.0 + n(cfg.d)
        w[p + "post_attention_layernorm.weight"] = 1.0 + n(cfg.d)
    w["model.norm.weight"] = 1.0 + n(cfg.d)
    if not cfg.tie_embeddings:
        w["lm_head.weight"] = n(cfg.vocab, cfg.d)
    return w


def init_lora(cfg: LlmCfg, lora: LoraCfg, seed: int = 44, b_std: float = 0.02) -> Dict[str, torch.Tensor]:
    """peft 0.6 names (without the 'base_model.model.' prefix).  A: kaiming_uniform(a=sqrt(5));
    B: N(0, b_std) instead of peft's zeros so that dA != 0 in parity tests (SURVEY.md §7 hard parts)."""
    g = torch.Generator().manual_seed(seed)
    w: Dict[str, torch.Tensor] = {}
    for i in range(cfg.layers):
        for name in lora.targets:
            out_f, in_f = linear_shape(cfg, name)
            mod = "self_attn" if name in ("q_proj", "k_proj", "v_proj", "o_proj") else "mlp"
            p = f"model.layers.{i}.{mod}.{name}."
            bound = 1.0 / math.sqrt(in_f)  # kaiming_uniform_(a=sqrt(5)) on [r, in]: bound = sqrt(6/((1+5)*in)) = 1/sqrt(in)
            w[p + "lora_A.default.weight"] = (torch.rand(lora.r, in_f, generator=g) * 2 - 1) * bound
            w[p + "lora_B.default.weight"] = torch.randn(out_f, lora.r, generator=g) * b_std
    return w


def init_projector(enc: EncoderCfg, llm: LlmCfg, proj: ProjCfg, seed: int = 45) -> Dict[str, torch.Tensor]:
    """nn.Linear / nn.Conv1d default init (kaiming_uniform(a=sqrt(5)) weights, uniform bias)."""
    g = torch.Generator().manual_seed(seed)

    def lin(out_f, in_f, fan_in=None):
        fan_in = in_f if fan_in is None else fan_in
        b = 1.0 / math.sqrt(fan_in)
        return (torch.rand(out_f, in_f, generator=g) * 2 - 1) * b, (torch.rand(out_f, generator=g) * 2 - 1) * b

    w: Dict[str, torch.Tensor] = {}
    if proj.kind == "linear":
        w["linear1.weight"], w["linear1.bias"] = lin(proj.hidden, enc.d * proj.k)
        w["linear2.weight"], w["linear2.bias"] = lin(llm.d, proj.hidden)
    elif proj.kind == "cov1d-linear":
        b = 1.0 / math.sqrt(enc.d * proj.k)
        w["conv1d.weight"] = (torch.rand(enc.d, enc.d, proj.k, generator=g) * 2 - 1) * b
        w["conv1d.bias"] = (torch.rand(enc.d, generator=g) * 2 - 1) * b
        w["linear1.weight"], w["linear1.bias"] = lin(proj.hidden, enc.d)
        w["linear2.weight"], w["linear2.bias"] = lin(llm.d, proj.hidden)
    else:
        raise ValueError(proj.kind)
    return w


# ------------------------------------------------------------------------------------------------------------------
# a1  log-mel (openai-whisper audio.py: pad_or_trim + log_mel_spectrogram)
# ------------------------------------------------------------------------------------------------------------------


def hz_to_mel(f: torch.Tensor) -> torch.Tensor:
    f_sp = 200.0 / 3
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, math.log(6.4) / 27.0
    return torch.where(f >= min_log_hz, min_log_mel + torch.log(f.clamp_min(1e-10) / min_log_hz) / logstep, f / f_sp)


def mel_to_hz(m: torch.Tensor) -> torch.Tensor:
    f_sp = 200.0 / 3
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, math.log(6.4) / 27.0
    return torch.where(m >= min_log_mel, min_log_hz * torch.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filters(n_mels: int, sr: int = 16000, n_fft: int = 400) -> torch.Tensor:
    """librosa.filters.mel(sr=16000, n_fft=400, n_mels) (slaney scale + slaney norm) -> f32 [n_mels, 201]."""
    fftfreqs = torch.linspace(0, sr / 2, n_fft // 2 + 1, dtype=torch.float64)
    mel_pts = torch.linspace(hz_to_mel(torch.tensor(0.0, dtype=torch.float64)).item(),
                             hz_to_mel(torch.tensor(sr / 2, dtype=torch.float64)).item(), n_mels + 2, dtype=torch.float64)
    mel_f = mel_to_hz(mel_pts)
    fdiff = mel_f[1:] - mel_f[:-1]
    ramps = mel_f[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    weights = torch.clamp(torch.minimum(lower, upper), min=0.0)
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    return (weights * enorm[:, None]).float()


def pad_or_trim(x: torch.Tensor, length: int = 480000) -> torch.Tensor:
    n = x.shape[-1]
    if n > length:
        return x[..., :length]
    if n < length:
        return F.pad(x, (0, length - n))
    return x


def log_mel_spectrogram(wav: torch.Tensor, n_mels: int = 80, dtype=torch.float32) -> torch.Tensor:
    """One utterance [n] -> [n_mels, n//160] exactly as whisper.log_mel_spectrogram."""
    wav = wav.to(dtype)
    window = torch.hann_window(400, dtype=dtype)
    stft = torch.stft(wav, 400, 160, window=window, return_complex=True)
    magnitudes = stft[..., :-1].abs() ** 2
    mel_spec = mel_filters(n_mels).to(dtype) @ magnitudes
    log_spec = torch.clamp(mel_spec, min=1e-10).log10()
    log_spec = torch.maximum(log_spec, log_spec.max() - 8.0)
    return (log_spec + 4.0) / 4.0


def batch_log_mel(wavs: torch.Tensor, n_mels: int, lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[B, n] -> [B, T, n_mels] (the dataset's .permute(1, 0) + collator stacking, speech_dataset.py:103,246-249).
    lengths (dynamic-frame recipe, speech_dataset_large.py:102-104,196-199): every utterance's log-mel is computed on ITS OWN samples
    (own reflect padding, own max), then the collator right-pads the mel with zeros to the longest utterance of the batch."""
    if lengths is None:
        return torch.stack([log_mel_spectrogram(w, n_mels).permute(1, 0) for w in wavs])
    mels = [log_mel_spectrogram(w[: int(n)], n_mels).permute(1, 0) for w, n in zip(wavs, lengths)]
    t_max = max(m.shape[0] for m in mels)
    return torch.stack([F.pad(m, (0, 0, 0, t_max - m.shape[0])) for m in mels])


# ------------------------------------------------------------------------------------------------------------------
# a2  Whisper encoder, variable length (src/slam_llm/models/encoder.py:13-30)
# ------------------------------------------------------------------------------------------------------------------


def whisper_encoder(w: Dict[str, torch.Tensor], cfg: EncoderCfg, mel: torch.Tensor, return_layers: bool = False):
    """mel [B, T, n_mels] (time-major, as collated) -> [B, ceil(T/2), d]."""
    x = mel.permute(0, 2, 1)                                           # slam_model.py:321 passes [B, n_mels, T]
    x = F.gelu(F.conv1d(x, w["conv1.weight"], w["conv1.bias"], padding=1))
    x = F.gelu(F.conv1d(x, w["conv2.weight"], w["conv2.bias"], stride=2, padding=1))
    x = x.permute(0, 2, 1)
    x = x + w["positional_embedding"][: x.shape[1]]                    # encoder.py:24 (variable length)
    dh = cfg.d // cfg.heads
    B, T, _ = x.shape
    outs = []
    for i in range(cfg.layers):
        p = f"blocks.{i}."
        h = F.layer_norm(x, (cfg.d,), w[p + "attn_ln.weight"], w[p + "attn_ln.bias"], 1e-5)
        q = F.linear(h, w[p + "attn.query.weight"], w[p + "attn.query.bias"])
        k = F.linear(h, w[p + "attn.key.weight"])
        v = F.linear(h, w[p + "attn.value.weight"], w[p + "attn.value.bias"])
        q = q.view(B, T, cfg.heads, dh).permute(0, 2, 1, 3) * dh ** -0.25
        k = k.view(B, T, cfg.heads, dh).permute(0, 2, 3, 1) * dh ** -0.25
        v = v.view(B, T, cfg.heads, dh).permute(0, 2, 1, 3)
        a = torch.softmax((q @ k).float(), dim=-1).to(q.dtype)         # no mask
        h = (a @ v).permute(0, 2, 1, 3).reshape(B, T, cfg.d)
        x = x + F.linear(h, w[p + "attn.out.weight"], w[p + "attn.out.bias"])
        h = F.layer_norm(x, (cfg.d,), w[p + "mlp_ln.weight"], w[p + "mlp_ln.bias"], 1e-5)
        h = F.gelu(F.linear(h, w[p + "mlp.0.weight"], w[p + "mlp.0.bias"]))
        x = x + F.linear(h, w[p + "mlp.2.weight"], w[p + "mlp.2.bias"])
        if return_layers:
            outs.append(x)
    x = F.layer_norm(x, (cfg.d,), w["ln_post.weight"], w["ln_post.bias"], 1e-5)
    return (x, outs) if return_layers else x


# ------------------------------------------------------------------------------------------------------------------
# a3  projectors (src/slam_llm/models/projector.py)
# ------------------------------------------------------------------------------------------------------------------


def projector(w: Dict[str, torch.Tensor], proj: ProjCfg, x: torch.Tensor) -> torch.Tensor:
    if proj.kind == "linear":                                           # projector.py:15-27
        B, T, d = x.shape
        discard = T % proj.k
        if discard > 0:
            x = x[:, :-discard, :]
        x = x.contiguous().view(B, T // proj.k, d * proj.k)
        x = F.relu(F.linear(x, w["linear1.weight"], w["linear1.bias"]))
        return F.linear(x, w["linear2.weight"], w["linear2.bias"])
    # cov1d-linear, projector.py:41-49
    x = x.transpose(1, 2)
    x = F.conv1d(x, w["conv1d.weight"], w["conv1d.bias"], stride=proj.k)
    x = F.relu(x).transpose(1, 2).contiguous()
    x = F.relu(F.linear(x, w["linear1.weight"], w["linear1.bias"]))
    return F.linear(x, w["linear2.weight"], w["linear2.bias"])


# ------------------------------------------------------------------------------------------------------------------
# a4  embedding gather + modality merge (src/slam_llm/models/slam_model.py:370-392)
# ------------------------------------------------------------------------------------------------------------------


def merge(embed: torch.Tensor, input_ids: torch.Tensor, modality_mask: torch.Tensor, encoder_outs: torch.Tensor) -> torch.Tensor:
    ids = input_ids.clone()
    ids[ids == -1] = 0
    inputs_embeds = F.embedding(ids, embed)
    start = (modality_mask == True).float().argmax(dim=1)               # noqa: E712
    lengths = torch.clamp(modality_mask.sum(dim=1), max=encoder_outs.shape[1]).tolist()
    pad = torch.zeros_like(inputs_embeds)
    for i in range(encoder_outs.shape[0]):
        pad[i, start[i]:start[i] + lengths[i]] = encoder_outs[i][:lengths[i]]
    return pad + inputs_embeds * (~modality_mask[:, :, None])


# ------------------------------------------------------------------------------------------------------------------
# a5/a6  Llama decoder with LoRA (transformers v4.35.2 semantics + peft v0.6.0 lora.Linear)
# ------------------------------------------------------------------------------------------------------------------


def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    v = x.float()
    v = v * torch.rsqrt(v.pow(2).mean(-1, keepdim=True) + eps)
    return weight * v.to(x.dtype)


def rope_tables(seq: int, dh: int, theta: float, dtype=torch.float32) -> Tuple[torch.Tensor, torch.Tensor]:
    inv_freq = 1.0 / (theta ** (torch.arange(0, dh, 2, dtype=torch.float32) / dh))
    freqs = torch.outer(torch.arange(seq, dtype=torch.float32), inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def lora_linear(x: torch.Tensor, weight: torch.Tensor, lw: Dict[str, torch.Tensor], prefix: str, lora: Optional[LoraCfg],
                masks: Optional[Dict[str, torch.Tensor]] = None, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """peft 0.6 lora.Linear.forward: F.linear(x, W) + B(A(dropout(x))) * alpha/r.  Dropout is deterministic here: `masks`
    (prefix -> keep/(1-p) tensor shaped like x) stands in for nn.Dropout's random mask; None = dropout 0."""
    y = F.linear(x, weight, bias)
    a_key = prefix + "lora_A.default.weight"
    if lora is not None and a_key in lw:
        xd = x if masks is None or prefix not in masks else x * masks[prefix].view_as(x)
        y = y + F.linear(F.linear(xd, lw[a_key]), lw[prefix + "lora_B.default.weight"]) * lora.scaling
    return y


def llama_forward(w: Dict[str, torch.Tensor], lw: Dict[str, torch.Tensor], cfg: LlmCfg, lora: Optional[LoraCfg],
                  inputs_embeds: torch.Tensor, attention_mask: torch.Tensor, return_hidden: bool = False,
                  lora_masks: Optional[Dict[str, torch.Tensor]] = None):
    """inputs_embeds [B,S,D], attention_mask bool/int [B,S] (1 = real token) -> logits f32 [B,S,V]."""
    B, S, D = inputs_embeds.shape
    H, Hkv, dh = cfg.heads, cfg.kv_heads, cfg.dh
    cos, sin = rope_tables(S, dh, cfg.rope_theta, inputs_embeds.dtype)   # position_ids = arange(S), padding-agnostic
    neg = torch.finfo(inputs_embeds.dtype).min
    causal = torch.ones(S, S, dtype=torch.bool).tril()
    allowed = causal[None, None] & attention_mask.bool()[:, None, None, :]
    add_mask = torch.zeros(B, 1, S, S, dtype=inputs_embeds.dtype).masked_fill(~allowed, neg)
    x = inputs_embeds
    hiddens = []
    for i in range(cfg.layers):
        p = f"model.layers.{i}."
        h = rms_norm(x, w[p + "input_layernorm.weight"], cfg.eps)
        q = lora_linear(h, w[p + "self_attn.q_proj.weight"], lw, p + "self_attn.q_proj.", lora, lora_masks, w.get(p + "self_attn.q_proj.bias"))
        k = lora_linear(h, w[p + "self_attn.k_proj.weight"], lw, p + "self_attn.k_proj.", lora, lora_masks, w.get(p + "self_attn.k_proj.bias"))
        v = lora_linear(h, w[p + "self_attn.v_proj.weight"], lw, p + "self_attn.v_proj.", lora, lora_masks, w.get(p + "self_attn.v_proj.bias"))
        q = q.view(B, S, H, dh).transpose(1, 2)
        k = k.view(B, S, Hkv, dh).transpose(1, 2)
        v = v.view(B, S, Hkv, dh).transpose(1, 2)
        q = q * cos[None, None] + rotate_half(q) * sin[None, None]
        k = k * cos[None, None] + rotate_half(k) * sin[None, None]
        k = k.repeat_interleave(H // Hkv, dim=1)                         # repeat_kv
        v = v.repeat_interleave(H // Hkv, dim=1)
        att = q @ k.transpose(2, 3) / math.sqrt(dh) + add_mask
        att = torch.softmax(att, dim=-1, dtype=torch.float32).to(q.dtype)
        o = (att @ v).transpose(1, 2).reshape(B, S, D)
        x = x + lora_linear(o, w[p + "self_attn.o_proj.weight"], lw, p + "self_attn.o_proj.", lora, lora_masks)
        h = rms_norm(x, w[p + "post_attention_layernorm.weight"], cfg.eps)
        g = lora_linear(h, w[p + "mlp.gate_proj.weight"], lw, p + "mlp.gate_proj.", lora, lora_masks)
        u = lora_linear(h, w[p + "mlp.up_proj.weight"], lw, p + "mlp.up_proj.", lora, lora_masks)
        x = x + lora_linear(F.silu(g) * u, w[p + "mlp.down_proj.weight"], lw, p + "mlp.down_proj.", lora, lora_masks)
        if return_hidden:
            hiddens.append(x)
    x = rms_norm(x, w["model.norm.weight"], cfg.eps)
    logits = F.linear(x, w.get("lm_head.weight", w["model.embed_tokens.weight"])).float()      # tied embeddings: lm_head IS the embedding table
    return (logits, hiddens) if return_hidden else logits


# ------------------------------------------------------------------------------------------------------------------
# a7  loss + accuracy
# ------------------------------------------------------------------------------------------------------------------


def causal_lm_loss(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """HF LlamaForCausalLM loss block: shift, CrossEntropyLoss() mean over labels != -100."""
    shift_logits = logits[..., :-1, :].contiguous()
    shift_labels = labels[..., 1:].contiguous()
    return F.cross_entropy(shift_logits.view(-1, shift_logits.shape[-1]), shift_labels.view(-1), ignore_index=-100)


def compute_accuracy(pad_outputs: torch.Tensor, pad_targets: torch.Tensor, ignore_label: int) -> torch.Tensor:
    """src/slam_llm/utils/metric.py:3-20."""
    mask = pad_targets != ignore_label
    numerator = torch.sum(pad_outputs.masked_select(mask) == pad_targets.masked_select(mask))
    denominator = torch.sum(mask)
    return numerator.float() / denominator.float()


# ------------------------------------------------------------------------------------------------------------------
# the step
# ------------------------------------------------------------------------------------------------------------------


@dataclass
class OracleModel:
    enc_cfg: EncoderCfg
    llm_cfg: LlmCfg
    lora_cfg: Optional[LoraCfg]
    proj_cfg: ProjCfg
    enc_w: Dict[str, torch.Tensor]
    llm_w: Dict[str, torch.Tensor]
    lora_w: Dict[str, torch.Tensor]
    proj_w: Dict[str, torch.Tensor]
    adam_state: dict = field(default_factory=dict)
    train_llm: bool = False        # freeze_llm=false (full fine-tune, examples/s2s): every decoder parameter is trainable

    @classmethod
    def build(cls, enc_cfg, llm_cfg, lora_cfg, proj_cfg, seed: int = 42):
        return cls(enc_cfg, llm_cfg, lora_cfg, proj_cfg, init_encoder(enc_cfg, seed), init_llm(llm_cfg, seed + 1),
                   init_lora(llm_cfg, lora_cfg, seed + 2) if lora_cfg is not None else {}, init_projector(enc_cfg, llm_cfg, proj_cfg, seed + 3))

    def to(self, dtype):
        for d in (self.enc_w, self.llm_w, self.lora_w, self.proj_w):
            for k in d:
                d[k] = d[k].to(dtype)
        return self

    def trainable(self) -> Dict[str, torch.Tensor]:
        """Reference checkpoint key names (SURVEY.md §5 checkpoint row)."""
        out = {f"encoder_projector.{k}": v for k, v in self.proj_w.items()}
        out.update({f"llm.base_model.model.{k}": v for k, v in self.lora_w.items()})
        if self.train_llm:
            out.update({f"llm.{k}": v for k, v in self.llm_w.items()})                   # HF names under the `llm.` attribute (no peft wrapper)
        return out

    def forward(self, batch: Dict[str, torch.Tensor], return_all: bool = False, lora_masks=None):
        """batch keys as produced by SpeechDatasetJsonl.collator (speech_dataset.py:216-291); audio either as
        `audio_mel` [B,T,n_mels] or raw `audio_pcm` [B,n] (log-mel computed here)."""
        dtype = self.llm_w["model.embed_tokens.weight"].dtype
        if "encoder_out" in batch:                                       # a non-Whisper (foreign, frozen) encoder produced the features
            mel, enc = None, batch["encoder_out"].to(dtype)
        else:
            mel = batch.get("audio_mel")
            if mel is None:
                mel = batch_log_mel(batch["audio_pcm"], self.enc_cfg.n_mels, batch.get("audio_pcm_lengths"))
            mel = mel.to(dtype)
            with torch.no_grad():                                        # encoder frozen (slam_model.py:110-113)
                enc = whisper_encoder(self.enc_w, self.enc_cfg, mel)
        aud = projector(self.proj_w, self.proj_cfg, enc)
        x = merge(self.llm_w["model.embed_tokens.weight"], batch["input_ids"], batch["modality_mask"].bool(), aud)
        logits = llama_forward(self.llm_w, self.lora_w, self.llm_cfg, self.lora_cfg, x, batch["attention_mask"], lora_masks=lora_masks)
        labels = batch["labels"]
        loss = causal_lm_loss(logits, labels)
        preds = torch.argmax(logits, -1)
        acc = compute_accuracy(preds[:, :-1], labels[:, 1:], ignore_label=-100)
        if return_all:
            return {"loss": loss, "acc": acc, "logits": logits, "encoder_out": enc, "audio_tokens": aud, "inputs_embeds": x, "mel": mel}
        return loss, acc

    def step(self, batch, lr: float = 1e-4, weight_decay: float = 0.0, do_update: bool = True, lora_masks=None):
        """One optimizer step: forward, backward (grads only for projector + LoRA), AdamW.  Returns dict with
        loss, acc and gradients keyed by the reference checkpoint names."""
        params = self.trainable()
        for p in params.values():
            p.requires_grad_(True)
            p.grad = None
        out = self.forward(batch, return_all=True, lora_masks=lora_masks)
        out["loss"].backward()
        grads = {k: p.grad.detach().clone() for k, p in params.items()}
        if do_update:
            if "opt" not in self.adam_state:
                self.adam_state["opt"] = torch.optim.AdamW(list(params.values()), lr=lr, weight_decay=weight_decay)
            opt = self.adam_state["opt"]
            for gq in opt.param_groups:
                gq["lr"] = lr
            opt.step()
        for p in params.values():
            p.requires_grad_(False)
        return {"loss": out["loss"].detach(), "acc": out["acc"], "grads": grads, "logits": out["logits"].detach(),
                "encoder_out": out["encoder_out"], "audio_tokens": out["audio_tokens"].detach(), "inputs_embeds": out["inputs_embeds"].detach()}


# ------------------------------------------------------------------------------------------------------------------
# f3  SLAM-Omni (s2s) head: multi-layer token input, group CE  (examples/s2s/model/slam_model_s2s.py:160-306)
# ------------------------------------------------------------------------------------------------------------------


def s2s_forward(om: "OracleModel", batch: Dict[str, torch.Tensor], code_layer: int, text_vocab: int, audio_vocab: int):
    """input_ids / labels are [B, code_layer + 1, S]: `code_layer` audio-codebook layers + one text layer (slam_model_s2s.py:220-268).
    Embeddings of all layers are averaged (the audio features replace the masked span in the audio layers only), the decoder's logits over the
    expanded vocabulary are split into a text slice and code_layer audio slices, and the loss is the mean of the code_layer + 1 shifted CEs
    (compute_parallel_loss, :285-306).  Returns (loss, text_acc, layer_losses, logits)."""
    mel = batch.get("audio_mel")
    if mel is None:
        mel = batch_log_mel(batch["audio_pcm"], om.enc_cfg.n_mels, batch.get("audio_pcm_lengths"))
    with torch.no_grad():
        enc = whisper_encoder(om.enc_w, om.enc_cfg, mel)
    aud = projector(om.proj_w, om.proj_cfg, enc)
    ids = batch["input_ids"].clone()
    ids[ids == -1] = 0
    emb = F.embedding(ids, om.llm_w["model.embed_tokens.weight"])                       # [B, L+1, S, D]
    mm = batch["modality_mask"].bool().unsqueeze(1).repeat(1, code_layer, 1)             # [B, L, S]
    start = (mm == True).float().argmax(dim=2)                                          # noqa: E712
    lengths = torch.clamp(mm.sum(dim=2), max=aud.shape[1]).tolist()
    pad = torch.zeros_like(emb)
    for i in range(aud.shape[0]):
        for j in range(code_layer):
            s0, n = start[i, j].item(), lengths[i][j]
            pad[i, j, s0:s0 + n] = aud[i, :n]
    emb = torch.cat([pad[:, :code_layer] + emb[:, :code_layer] * (~mm[:, :, :, None]), emb[:, code_layer:]], dim=1)
    x = emb.mean(dim=1)
    logits = llama_forward(om.llm_w, om.lora_w, om.llm_cfg, om.lora_cfg, x, batch["attention_mask"])
    labels = batch["labels"]
    text_labels, audio_labels = labels[:, code_layer], labels[:, :code_layer]
    xt = logits[..., :text_vocab]
    layer_loss = [None] * (code_layer + 1)
    layer_loss[code_layer] = F.cross_entropy(xt[:, :-1].reshape(-1, text_vocab), text_labels[:, 1:].reshape(-1), ignore_index=-100)
    total = layer_loss[code_layer]
    for i in range(code_layer):
        xa = logits[..., text_vocab + audio_vocab * i: text_vocab + audio_vocab * (i + 1)]
        layer_loss[i] = F.cross_entropy(xa[:, :-1].reshape(-1, audio_vocab), audio_labels[:, i, 1:].reshape(-1), ignore_index=-100)
        total = total + layer_loss[i]
    loss = total / (code_layer + 1)
    text_acc = compute_accuracy(torch.argmax(xt, -1)[:, :-1], text_labels[:, 1:], ignore_label=-100)
    return loss, text_acc, layer_loss, logits


def s2s_step(om: "OracleModel", batch, code_layer: int, text_vocab: int, audio_vocab: int):
    params = om.trainable()
    for p in params.values():
        p.requires_grad_(True)
        p.grad = None
    loss, acc, layer_loss, logits = s2s_forward(om, batch, code_layer, text_vocab, audio_vocab)
    loss.backward()
    grads = {k: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p)) for k, p in params.items()}
    for p in params.values():
        p.requires_grad_(False)
    return {"loss": loss.detach(), "acc": acc, "layer_loss": [l.detach() for l in layer_loss], "grads": grads, "logits": logits.detach()}


def s2s_synthetic_batch(B: int, n_samples: int, code_layer: int, text_vocab: int, audio_vocab: int, k: int = 5, prompt_len: int = 5, answer_len: int = 8,
                        seed: int = 42) -> Dict[str, torch.Tensor]:
    """[audio(-1)*Ta, prompt, answer] in code_layer + 1 parallel token layers: layer i < code_layer draws from its own audio-code slice of the
    expanded vocabulary, the last layer from the text vocabulary; labels -100 outside the answer span (all layers)."""
    g = torch.Generator().manual_seed(seed)
    wav = torch.randn(B, n_samples, generator=g) * 0.1
    ta = ((n_samples // 160 + 1) // 2) // k
    S = ta + prompt_len + answer_len
    ids = torch.zeros(B, code_layer + 1, S, dtype=torch.int64)
    labels = torch.zeros(B, code_layer + 1, S, dtype=torch.int64)
    for i in range(code_layer):
        codes = torch.randint(0, audio_vocab, (B, S), generator=g)
        ids[:, i] = text_vocab + audio_vocab * i + codes              # inputs index the expanded table (layer shift) ...
        labels[:, i] = codes                                          # ... targets index the layer's own slice of the logits
    ids[:, code_layer] = labels[:, code_layer] = torch.randint(0, text_vocab, (B, S), generator=g)
    labels[:, :, : ta + prompt_len] = -100
    ids[:, :, :ta] = -1
    att = torch.ones(B, S, dtype=torch.bool)
    mod = torch.zeros(B, S, dtype=torch.bool)
    mod[:, :ta] = True
    return {"input_ids": ids, "labels": labels, "attention_mask": att, "modality_mask": mod, "audio_pcm": wav}


def lr_lambda(step: int, warmup: int, total: int) -> float:
    """src/slam_llm/pipeline/finetune.py:253-260."""
    if step < warmup:
        return min(step / warmup, 1.0)
    return max(0.0, 1 - (step - warmup) / (total - warmup))


# ------------------------------------------------------------------------------------------------------------------
# synthetic batch in the collator's contract (BASELINE.md §2 / SURVEY.md §8d)
# ------------------------------------------------------------------------------------------------------------------


def synthetic_batch(B: int, n_samples: int, vocab: int, k: int = 5, prompt_len: int = 24, answer_len: int = 76, seed: int = 42,
                    left_pad: Optional[Sequence[int]] = None, with_mel: bool = False, n_mels: int = 80) -> Dict[str, torch.Tensor]:
    """Token layout [audio(-1)*L, prompt, answer, eos] with labels -100 outside answer+eos (speech_dataset.py:109-161);
    optional per-sample left padding (collator left-pads, speech_dataset.py:224-236)."""
    g = torch.Generator().manual_seed(seed)
    wav = torch.randn(B, n_samples, generator=g) * 0.1
    n_frames = n_samples // 160
    audio_len = ((n_frames + 1) // 2) // k                              # speech_dataset.py:104-105
    left_pad = [0] * B if left_pad is None else list(left_pad)
    S = max(left_pad) + audio_len + prompt_len + answer_len + 1
    ids = torch.zeros(B, S, dtype=torch.int64)
    labels = torch.full((B, S), -100, dtype=torch.int64)
    att = torch.zeros(B, S, dtype=torch.bool)
    mod = torch.zeros(B, S, dtype=torch.bool)
    for b in range(B):
        off = left_pad[b]
        n_tok = audio_len + prompt_len + answer_len + 1
        ids[b, off:off + audio_len] = -1
        ids[b, off + audio_len: off + audio_len + prompt_len] = torch.randint(0, vocab, (prompt_len,), generator=g)
        ans = torch.randint(0, vocab, (answer_len + 1,), generator=g)
        ids[b, off + audio_len + prompt_len: off + n_tok] = ans
        labels[b, off + audio_len + prompt_len: off + n_tok] = ans
        att[b, off:off + n_tok] = True
        mod[b, off:off + audio_len] = True
    batch = {"input_ids": ids, "labels": labels, "attention_mask": att, "modality_mask": mod, "audio_pcm": wav}
    if with_mel:
        batch["audio_mel"] = batch_log_mel(wav, n_mels)
    return batch
