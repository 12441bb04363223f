"""Architecture descriptors of the hot path (dims only; mirrors what the reference reads from
whisper `dims`, HF `config.json` and train_config.peft_config — SURVEY.md §8a / Appendix A)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple


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
    qkv_bias: bool = False         # Qwen2: biases on the q/k/v projections
    tie_embeddings: bool = False   # lm_head shares the embedding table (Qwen2-0.5B)

    @property
    def dh(self) -> int:
        return self.d // self.heads

    @property
    def dq(self) -> int:
        return self.heads * self.dh

    @property
    def dkv(self) -> int:
        return self.kv_heads * self.dh


@dataclass
class LoraCfg:
    r: int = 8
    alpha: int = 32
    targets: Tuple[str, ...] = ("q_proj", "v_proj")
    dropout: float = 0.0

    @property
    def scaling(self) -> float:
        return self.alpha / self.r


@dataclass
class ProjCfg:
    kind: str = "linear"
    k: int = 5
    hidden: int = 2048


WHISPER = {
    "tiny": EncoderCfg(80, 1500, 384, 6, 4),
    "base": EncoderCfg(80, 1500, 512, 8, 6),
    "small": EncoderCfg(80, 1500, 768, 12, 12),
    "medium": EncoderCfg(80, 1500, 1024, 16, 24),
    "large": EncoderCfg(80, 1500, 1280, 20, 32),
    "large-v2": EncoderCfg(80, 1500, 1280, 20, 32),
    "large-v3": EncoderCfg(128, 1500, 1280, 20, 32),
}
LLM = {
    "tinyllama-1.1b": LlmCfg(32000, 2048, 22, 32, 4, 5632, 10000.0, 1e-5),
    "llama-3-8b": LlmCfg(128256, 4096, 32, 32, 8, 14336, 500000.0, 1e-5),
    "vicuna-7b": LlmCfg(32000, 4096, 32, 32, 32, 11008, 10000.0, 1e-5),
    "qwen2-0.5b": LlmCfg(151936, 896, 24, 14, 2, 4864, 1000000.0, 1e-6, True, True),
}

ATTN_LINEARS = ("q_proj", "k_proj", "v_proj", "o_proj")
MLP_LINEARS = ("gate_proj", "up_proj", "down_proj")


def linear_shape(cfg: LlmCfg, name: str) -> Tuple[int, int]:
    """(out_features, in_features)."""
    return {
        "q_proj": (cfg.dq, cfg.d), "k_proj": (cfg.dkv, cfg.d), "v_proj": (cfg.dkv, cfg.d), "o_proj": (cfg.d, cfg.dq),
        "gate_proj": (cfg.ffn, cfg.d), "up_proj": (cfg.ffn, cfg.d), "down_proj": (cfg.d, cfg.ffn),
    }[name]


def step_flops(enc: EncoderCfg, llm: LlmCfg, proj: ProjCfg, lora: Optional[LoraCfg], B: int, T: int, S: int, n_label_rows: Optional[int] = None) -> dict:
    """Algorithmic FLOPs of one training step (SURVEY.md §8d formulas; matmul = 2MNK, attention counted full S^2).
    n_label_rows: rows of the lm_head actually needed (labels != -100); None = all B*S rows."""
    Tp = (T + 1) // 2
    d, L_e = enc.d, enc.layers
    E = 6 * enc.n_mels * d * T + 6 * d * d * Tp + L_e * (8 * Tp * d * d + 4 * Tp * Tp * d + 16 * Tp * d * d)
    Ta = Tp // proj.k
    P_f = 2 * Ta * (proj.k * d * proj.hidden + proj.hidden * llm.d)
    P_b = 2 * Ta * (proj.k * d * proj.hidden) + 4 * Ta * proj.hidden * llm.d
    D, Dkv, F, V, L = llm.d, llm.dkv, llm.ffn, llm.vocab, llm.layers
    layer_lin = 4 * S * D * D + 4 * S * D * Dkv + 6 * S * D * F
    layer_att = 4 * S * S * D
    head_rows = S if n_label_rows is None else n_label_rows / B
    lm = 2 * head_rows * D * V
    llm_f = L * (layer_lin + layer_att) + lm
    llm_b = L * (layer_lin + 2 * layer_att) + lm
    lo = 0
    if lora is not None:
        from .config import linear_shape as _ls
        tot = sum(sum(_ls(llm, t)) for t in lora.targets)
        lo = 2 * S * lora.r * tot * L * 3
    per_utt = E + P_f + P_b + llm_f + llm_b + lo
    return {"encoder_fwd": B * E, "projector": B * (P_f + P_b), "llm_fwd": B * llm_f, "llm_bwd": B * llm_b, "lora": B * lo, "total": B * per_utt}
