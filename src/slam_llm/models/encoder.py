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
    def __init__(self, engine_encoder: WhisperEncoderB200, hf_style: bool = False):
        super().__init__()
        self.b200 = engine_encoder
        self.hf_style = hf_style      # loaded via encoder_path_hf: recipes call `self.encoder(mel).last_hidden_state` (st_covost2/model/slam_model_st.py:304-305)
        cfg = engine_encoder.cfg
        d = cfg.d
        self.num_frozen_params = 3 * cfg.n_mels * d + 3 * d * d + 2 * d + cfg.layers * (12 * d * d + 13 * d) + 2 * d

    def extract_variable_length_features(self, x: torch.Tensor) -> torch.Tensor:
        """x [B, n_mels, T] (any float dtype, any device) -> bf16 [B, ceil(T/2), d] on the GPU."""
        mel = x.permute(0, 2, 1).to(self.b200.device, torch.float32).contiguous()
        return self.b200.forward(mel)

    def forward(self, x):
        out = self.extract_variable_length_features(x)
        if self.hf_style:
            import types
            return types.SimpleNamespace(last_hidden_state=out)
        return out


def _dims_from_name(name: str) -> EncoderCfg:
    key = os.path.basename(str(name)).replace(".pt", "").replace("whisper-", "")
    key = key.split(".")[0] if key.split(".")[0] in WHISPER else key
    if key not in WHISPER:
        raise ValueError(f"unknown whisper model {name!r}: expected a path to an openai-whisper .pt or one of {sorted(WHISPER)}")
    return WHISPER[key]


_HF_TO_OPENAI = {"self_attn.q_proj": "attn.query", "self_attn.k_proj": "attn.key", "self_attn.v_proj": "attn.value", "self_attn.out_proj": "attn.out",
                 "self_attn_layer_norm": "attn_ln", "fc1": "mlp.0", "fc2": "mlp.2", "final_layer_norm": "mlp_ln"}


def hf_whisper_encoder_weights(sd):
    """HF Whisper state dict (`model.encoder.*` / `encoder.*` / bare) -> openai-whisper AudioEncoder names
    (`layers.i.self_attn.q_proj` -> `blocks.i.attn.query`, `embed_positions.weight` -> `positional_embedding`, `layer_norm` -> `ln_post`)."""
    out = {}
    for k, v in sd.items():
        if "encoder." in k:
            k = k[k.index("encoder.") + len("encoder."):]
        elif k.startswith(("model.decoder.", "decoder.", "proj_out.")):
            continue
        if k.startswith("layers."):
            _, i, rest = k.split(".", 2)
            mod, suffix = rest.rsplit(".", 1)
            if mod not in _HF_TO_OPENAI:
                continue
            out[f"blocks.{i}.{_HF_TO_OPENAI[mod]}.{suffix}"] = v
        elif k == "embed_positions.weight":
            out["positional_embedding"] = v
        elif k.startswith("layer_norm."):
            out["ln_post." + k.split(".", 1)[1]] = v
        elif k.startswith(("conv1.", "conv2.")):
            out[k] = v
    return out


def _load_hf_whisper(path: str):
    import json
    with open(os.path.join(path, "config.json")) as f:
        c = json.load(f)
    cfg = EncoderCfg(c["num_mel_bins"], c["max_source_positions"], c["d_model"], c["encoder_attention_heads"], c["encoder_layers"])
    names = sorted(os.listdir(path))
    sd = {}
    st = [f for f in names if f.endswith(".safetensors")]
    if st:
        from safetensors.torch import load_file
        for f in st:
            sd.update(load_file(os.path.join(path, f)))
    else:
        for f in (n for n in names if n.startswith("pytorch_model") and n.endswith(".bin")):
            sd.update(torch.load(os.path.join(path, f), map_location="cpu", weights_only=True))
    if not sd:
        raise FileNotFoundError(f"no *.safetensors / pytorch_model*.bin under encoder_path_hf={path!r}")
    return cfg, hf_whisper_encoder_weights(sd)


class WhisperWrappedEncoder:
    @classmethod
    def load(cls, model_config):
        if model_config.get("whisper_decode", False):                         # reference quirk Q1: use .get
            raise NotImplementedError("whisper_decode (full Whisper enc-dec) is outside the B200 hot path")
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else "cuda:0"
        hf_path = model_config.get("encoder_path_hf", None)
        if hf_path is not None:                                               # encoder.py:38-40: WhisperModel.from_pretrained(...).encoder
            cfg, weights = _load_hf_whisper(str(hf_path))
            return WhisperEncoderModule(WhisperEncoderB200(cfg, weights, device), hf_style=True)
        path = model_config.encoder_path
        if path is not None and os.path.isfile(str(path)):                    # whisper.load_model(name=<path to .pt>)
            ckpt = torch.load(path, map_location="cpu", weights_only=False)
            dims = ckpt["dims"]
            cfg = EncoderCfg(dims["n_mels"], dims["n_audio_ctx"], dims["n_audio_state"], dims["n_audio_head"], dims["n_audio_layer"])
            weights = {k[len("encoder."):]: v for k, v in ckpt["model_state_dict"].items() if k.startswith("encoder.")}
            eng = WhisperEncoderB200(cfg, weights, device)
        else:
            from slam_llm.models.slam_model import random_init_allowed
            if not random_init_allowed(model_config):
                raise FileNotFoundError(f"encoder_path={path!r} is not a file (openai-whisper names are downloaded by whisper.load_model in the reference; "
                                        "offline, pass the .pt path - or set model_config.b200_random_init=true / SLAM_B200_RANDOM_INIT=1 to benchmark "
                                        "with random frozen weights)")
            cfg = _dims_from_name(path)
            logger.warning(f"encoder_path={path!r} is not a file: RANDOM-INIT Whisper encoder with dims {cfg} (b200_random_init)")
            eng = WhisperEncoderB200(cfg, None, device)
        return WhisperEncoderModule(eng)


# ---------------------------------------------------------------------------------------------------------------------
# Non-Whisper ("foreign") modality encoders: frozen torch modules that the user's environment provides (fairseq EAT, BEATs, WavLM, ...).
# They run as they are (torch, no B200 kernels: their networks live in un-vendored third-party code); everything that trains -
# projector, merge, decoder, loss, optimizer - runs on the B200 step, which is entered with their output (SlamStepB200.forward_rest).
# ---------------------------------------------------------------------------------------------------------------------
_FOREIGN = {}


def register_encoder(name: str, loader, call) -> None:
    """loader(model_config) -> nn.Module;  call(module, batch_kwargs) -> features [B, T, encoder_dim] (what slam_model.forward computes for
    this `model_config.encoder_name`, models/slam_model.py:319-353)."""
    _FOREIGN[name] = (loader, call)


def foreign_encoder(name: str):
    return _FOREIGN.get(name)


class EATEncoder:
    """models/encoder.py:65-78: the EAT network is fairseq user code (`model_config.encoder_fairseq_dir`), loaded exactly like the reference."""

    @classmethod
    def load(cls, model_config):
        try:
            import fairseq
        except ImportError as e:
            raise ImportError("encoder_name=eat needs `fairseq` and the EAT user directory (model_config.encoder_fairseq_dir), as in the reference; "
                              "neither ships with this package") from e
        from dataclasses import dataclass

        @dataclass
        class UserDirModule:
            user_dir: str
        fairseq.utils.import_user_module(UserDirModule(model_config.encoder_fairseq_dir))
        models, _, _ = fairseq.checkpoint_utils.load_model_ensemble_and_task([model_config.encoder_path])
        return models[0]


register_encoder("eat", EATEncoder.load,
                 lambda enc, kw: enc.model.extract_features(kw["audio_mel"].unsqueeze(dim=1), padding_mask=None, mask=False, remove_extra_tokens=False)["x"])
