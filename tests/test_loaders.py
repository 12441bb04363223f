"""Host-side checkpoint loading and its failure modes — CPU tests (no kernels run)."""
import json
import os

import pytest
import torch

from oracle import slam_oracle as so


def _cfg_dir(tmp_path, **over):
    c = dict(model_type="llama", vocab_size=512, hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
             num_key_value_heads=2, rms_norm_eps=1e-5, rope_theta=10000.0)
    c.update(over)
    d = tmp_path / "llm"
    d.mkdir(exist_ok=True)
    (d / "config.json").write_text(json.dumps(c))
    return str(d)


def test_llm_config_reader_rejects_what_the_kernels_do_not_implement(tmp_path):
    from slam_llm.models.slam_model import _load_llm_cfg
    cfg = _load_llm_cfg(_cfg_dir(tmp_path))
    assert (cfg.vocab, cfg.d, cfg.layers, cfg.heads, cfg.kv_heads, cfg.ffn) == (512, 256, 2, 4, 2, 512)
    for bad in (dict(rope_scaling={"rope_type": "llama3", "factor": 8.0}), dict(head_dim=32), dict(attention_bias=True), dict(mlp_bias=True),
                dict(hidden_act="gelu"), dict(model_type="mistral", sliding_window=4096), dict(model_type="gemma"),
                dict(model_type="qwen2", use_sliding_window=True)):
        with pytest.raises(NotImplementedError):
            _load_llm_cfg(_cfg_dir(tmp_path, **bad))
    assert _load_llm_cfg(_cfg_dir(tmp_path, head_dim=64, rope_scaling=None)).d == 256
    q = _load_llm_cfg(_cfg_dir(tmp_path, model_type="qwen2", tie_word_embeddings=True, rms_norm_eps=1e-6, rope_theta=1e6))
    assert q.qkv_bias and q.tie_embeddings and q.eps == 1e-6 and not cfg.qkv_bias and not cfg.tie_embeddings


def test_llm_weight_reader_safetensors_bin_and_tied_head(tmp_path):
    from safetensors.torch import save_file
    from slam_llm.models.slam_model import _load_llm_weights
    w = so.init_llm(so.LlmCfg(64, 32, 1, 2, 1, 48, 1e4, 1e-5), seed=3)
    d = tmp_path / "st"
    d.mkdir()
    half = len(w) // 2
    items = list(w.items())
    save_file({k: v.contiguous() for k, v in items[:half]}, str(d / "model-00001-of-00002.safetensors"))
    save_file({k: v.contiguous() for k, v in items[half:]}, str(d / "model-00002-of-00002.safetensors"))
    got = _load_llm_weights(str(d))
    assert set(got) == set(w) and all(torch.equal(got[k], w[k]) for k in w)
    b = tmp_path / "bin"
    b.mkdir()
    tied = {k: v for k, v in w.items() if k != "lm_head.weight"}
    torch.save(tied, str(b / "pytorch_model.bin"))
    got = _load_llm_weights(str(b))
    assert "lm_head.weight" not in got and torch.equal(got["model.embed_tokens.weight"], w["model.embed_tokens.weight"])   # tied: the engine shares the table
    e = tmp_path / "empty"
    e.mkdir()
    assert _load_llm_weights(str(e)) is None


def test_missing_frozen_weights_raise_unless_random_init_is_requested(tmp_path, monkeypatch):
    from omegaconf import OmegaConf
    from slam_llm.models.encoder import WhisperWrappedEncoder
    from slam_llm.models.slam_model import LlamaB200ForCausalLM, _load_llm_cfg, random_init_allowed
    monkeypatch.delenv("SLAM_B200_RANDOM_INIT", raising=False)
    mc = OmegaConf.create(dict(encoder_path="tiny", encoder_name="whisper"))
    with pytest.raises(FileNotFoundError):
        WhisperWrappedEncoder.load(mc)
    assert not random_init_allowed(mc) and random_init_allowed(OmegaConf.create(dict(b200_random_init=True)))
    d = _cfg_dir(tmp_path)
    m = LlamaB200ForCausalLM(_load_llm_cfg(d), d, None, False)
    with pytest.raises(FileNotFoundError):
        m.bind(None, "cuda:0")
    monkeypatch.setenv("SLAM_B200_RANDOM_INIT", "1")
    assert random_init_allowed(mc)


def test_hf_whisper_names_map_back_to_openai_names():
    from slam_llm.models.encoder import hf_whisper_encoder_weights
    cfg = so.EncoderCfg(80, 1500, 64, 2, 2)
    w = so.init_encoder(cfg, seed=5)
    names = {"attn.query": "self_attn.q_proj", "attn.key": "self_attn.k_proj", "attn.value": "self_attn.v_proj", "attn.out": "self_attn.out_proj",
             "attn_ln": "self_attn_layer_norm", "mlp.0": "fc1", "mlp.2": "fc2", "mlp_ln": "final_layer_norm"}
    hf = {"model.encoder.conv1.weight": w["conv1.weight"], "model.encoder.conv1.bias": w["conv1.bias"], "model.encoder.conv2.weight": w["conv2.weight"],
          "model.encoder.conv2.bias": w["conv2.bias"], "model.encoder.embed_positions.weight": w["positional_embedding"],
          "model.encoder.layer_norm.weight": w["ln_post.weight"], "model.encoder.layer_norm.bias": w["ln_post.bias"],
          "model.decoder.embed_tokens.weight": torch.zeros(4, 4), "proj_out.weight": torch.zeros(4, 4)}
    for i in range(cfg.layers):
        for o, h in names.items():
            for sfx in ("weight", "bias"):
                if f"blocks.{i}.{o}.{sfx}" in w:
                    hf[f"model.encoder.layers.{i}.{h}.{sfx}"] = w[f"blocks.{i}.{o}.{sfx}"]
    back = hf_whisper_encoder_weights(hf)
    assert set(back) == set(w) and all(torch.equal(back[k], w[k]) for k in w)


def test_peft_directory_reader(tmp_path):
    from slam_llm.models.slam_model import _peft_dir_config, _peft_dir_state
    d = tmp_path / "peft"
    d.mkdir()
    (d / "adapter_config.json").write_text(json.dumps(dict(peft_type="LORA", r=4, lora_alpha=16, target_modules=["q_proj", "v_proj"], lora_dropout=0.05,
                                                           bias="none", task_type="CAUSAL_LM")))
    a = torch.randn(4, 32)
    torch.save({"base_model.model.model.layers.0.self_attn.q_proj.lora_A.weight": a}, str(d / "adapter_model.bin"))
    c = _peft_dir_config(str(d))
    assert (c.r, c.alpha, c.targets, c.dropout) == (4, 16, ("q_proj", "v_proj"), 0.05)
    sd = _peft_dir_state(str(d))
    assert list(sd) == ["llm.base_model.model.model.layers.0.self_attn.q_proj.lora_A.default.weight"] and torch.equal(list(sd.values())[0], a)
    (d / "adapter_config.json").write_text(json.dumps(dict(peft_type="PREFIX_TUNING")))
    with pytest.raises(NotImplementedError):
        _peft_dir_config(str(d))
