"""Whisper encoder wrapper (reference: src/slam_llm/models/encoder.py:8-45): an object exposing
`extract_variable_length_features(x[B, n_mels, T]) -> [B, ceil(T/2), d]`, frozen, running on slam_llm_b200 kernels
(im2col + tcgen05 GEMM conv stem, LayerNorm, fused-QKV GEMM, flash attention, GELU/residual GEMM epilogues)."""
import logging
import os

import torch
import torch.nn as nn

from slam_llm_b200.config import WHISPER, EncoderCfg
from slam_llm_b200.engine import WhisperEncoderB200

logger = logging.getLogger(__name__)


class WhisperEncoderModule(nn.Module):
    def __init__(self, engine_encoder: WhisperEncoderB200):
        super().__init__()
        self.b200 = engine_encoder
        cfg = engine_encoder.cfg
        d = cfg.d
        self.num_frozen_params = 3 * cfg.n_mels * d + 3 * d * d + 2 * d + cfg.layers * (12 * d * d + 13 * d) + 2 * d

    def extract_variable_length_features(self, x: torch.Tensor) -> torch.Tensor:
        """x [B, n_mels, T] (any float dtype, any device) -> bf16 [B, ceil(T/2), d] on the GPU."""
        mel = x.permute(0, 2, 1).to(self.b200.device, torch.float32).contiguous()
        return self.b200.forward(mel)

    def forward(self, x):
        return self.extract_variable_length_features(x)


def _dims_from_name(name: str) -> EncoderCfg:
    key = os.path.basename(str(name)).replace(".pt", "").replace("whisper-", "")
    key = key.split(".")[0] if key.split(".")[0] in WHISPER else key
    if key not in WHISPER:
        raise ValueError(f"unknown whisper model {name!r}: expected a path to an openai-whisper .pt or one of {sorted(WHISPER)}")
    return WHISPER[key]


class WhisperWrappedEncoder:
    @classmethod
    def load(cls, model_config):
        if model_config.get("whisper_decode", False):                         # reference quirk Q1: use .get
            raise NotImplementedError("whisper_decode (full Whisper enc-dec) is outside the B200 hot path")
        if model_config.get("encoder_path_hf", None) is not None:
            raise NotImplementedError("encoder_path_hf: load an openai-whisper .pt via encoder_path instead")
        path = model_config.encoder_path
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else "cuda:0"
        if path is not None and os.path.isfile(str(path)):
            ckpt = torch.load(path, map_location="cpu")
            dims = ckpt["dims"]
            cfg = EncoderCfg(dims["n_mels"], dims["n_audio_ctx"], dims["n_audio_state"], dims["n_audio_head"], dims["n_audio_layer"])
            weights = {k[len("encoder."):]: v for k, v in ckpt["model_state_dict"].items() if k.startswith("encoder.")}
            eng = WhisperEncoderB200(cfg, weights, device)
        else:
            cfg = _dims_from_name(path)
            logger.warning(f"encoder_path={path!r} is not a file: RANDOM-INIT Whisper encoder with dims {cfg} (offline / benchmark mode)")
            eng = WhisperEncoderB200(cfg, None, device)
        return WhisperEncoderModule(eng)
