"""Pin the CPU oracle (oracle/slam_oracle.py) against the third-party arithmetic the reference delegates to,
using the implementations installed in this image (transformers 5.5.0): WhisperFeatureExtractor, WhisperEncoder
sub-modules driven variable-length exactly like src/slam_llm/models/encoder.py:13-30, and LlamaForCausalLM
(eager attention, inputs_embeds + attention_mask + labels) with the peft-0.6 LoRA update merged into the weights
(W + (alpha/r) B A is algebraically what lora.Linear.forward computes).  The reference's own tests hold no golden
vectors for this path, so these are the anchors."""
import math

import numpy as np
import pytest
import torch

from oracle import slam_oracle as so

torch.manual_seed(0)


def test_mel_filterbank_matches_transformers_and_frontend():
    from transformers.audio_utils import mel_filter_bank
    from slam_llm_b200.frontend import mel_filterbank
    for n in (80, 128):
        ref = torch.from_numpy(mel_filter_bank(201, n, 0.0, 8000.0, 16000, norm="slaney", mel_scale="slaney")).float().t()
        assert (so.mel_filters(n) - ref).abs().max().item() < 3e-7
        assert (mel_filterbank(n) - ref).abs().max().item() < 3e-7


@pytest.mark.parametrize("n_mels", [80, 128])
def test_logmel_matches_whisper_feature_extractor(n_mels):
    from transformers import WhisperFeatureExtractor
    fe = WhisperFeatureExtractor(feature_size=n_mels)
    g = torch.Generator().manual_seed(1)
    wav = torch.randn(480000, generator=g) * 0.1
    wav[300000:] = 0.0
    ref = fe(wav.numpy(), sampling_rate=16000, return_tensors="pt").input_features[0]  # [n_mels, 3000]
    got = so.log_mel_spectrogram(so.pad_or_trim(wav), n_mels)
    assert got.shape == ref.shape == (n_mels, 3000)
    assert (got - ref).abs().max().item() < 5e-5


def _hf_encoder(cfg: so.EncoderCfg, w):
    from transformers import WhisperConfig
    from transformers.models.whisper.modeling_whisper import WhisperEncoder
    hc = WhisperConfig(num_mel_bins=cfg.n_mels, d_model=cfg.d, encoder_layers=cfg.layers, encoder_attention_heads=cfg.heads,
                       encoder_ffn_dim=4 * cfg.d, max_source_positions=cfg.n_ctx, dropout=0.0, attention_dropout=0.0, activation_dropout=0.0)
    hc._attn_implementation = "eager"
    enc = WhisperEncoder(hc).eval()
    sd = {"conv1.weight": w["conv1.weight"], "conv1.bias": w["conv1.bias"], "conv2.weight": w["conv2.weight"], "conv2.bias": w["conv2.bias"],
          "embed_positions.weight": w["positional_embedding"], "layer_norm.weight": w["ln_post.weight"], "layer_norm.bias": w["ln_post.bias"]}
    names = {"attn.query": "self_attn.q_proj", "attn.key": "self_attn.k_proj", "attn.value": "self_attn.v_proj", "attn.out": "self_attn.out_proj",
             "attn_ln": "self_attn_layer_norm", "mlp.0": "fc1", "mlp.2": "fc2", "mlp_ln": "final_layer_norm"}
    for i in range(cfg.layers):
        for o, h in names.items():
            for suffix in ("weight", "bias"):
                k = f"blocks.{i}.{o}.{suffix}"
                if k in w:
                    sd[f"layers.{i}.{h}.{suffix}"] = w[k]
    missing, unexpected = enc.load_state_dict(sd, strict=False)
    assert not unexpected and all("k_proj.bias" in m for m in missing), (missing, unexpected)
    return enc


@pytest.mark.parametrize("T", [3000, 500, 333])
def test_encoder_matches_hf_whisper_modules_variable_length(T):
    cfg = so.EncoderCfg(80, 1500, 64, 4, 2)
    w = so.init_encoder(cfg, seed=3)
    enc = _hf_encoder(cfg, w)
    mel = torch.randn(2, T, cfg.n_mels, generator=torch.Generator().manual_seed(4))
    got = so.whisper_encoder(w, cfg, mel)
    with torch.no_grad():
        # drive the HF sub-modules the way the reference's extract_variable_length_features drives whisper's
        x = torch.nn.functional.gelu(enc.conv1(mel.permute(0, 2, 1)))
        x = torch.nn.functional.gelu(enc.conv2(x)).permute(0, 2, 1)
        x = x + enc.embed_positions.weight[: x.shape[1]]
        for layer in enc.layers:
            out = layer(x, None)
            x = out[0] if isinstance(out, tuple) else out
        ref = enc.layer_norm(x)
    assert got.shape == ref.shape == (2, (T + 1) // 2, cfg.d)
    assert (got - ref).abs().max().item() < 2e-5
    if T == 3000:
        with torch.no_grad():
            full = enc(mel.permute(0, 2, 1)).last_hidden_state
        assert (got - full).abs().max().item() < 2e-5


def _hf_llama(cfg: so.LlmCfg, w, lw, lora):
    from transformers import LlamaConfig, LlamaForCausalLM
    hc = LlamaConfig(vocab_size=cfg.vocab, hidden_size=cfg.d, intermediate_size=cfg.ffn, num_hidden_layers=cfg.layers,
                     num_attention_heads=cfg.heads, num_key_value_heads=cfg.kv_heads, rms_norm_eps=cfg.eps, rope_theta=cfg.rope_theta,
                     max_position_embeddings=4096, attention_bias=False, tie_word_embeddings=False)
    hc._attn_implementation = "eager"
    m = LlamaForCausalLM(hc).eval()
    sd = {k: v.clone() for k, v in w.items()}
    for k, a in lw.items():
        if "lora_A" in k:
            base = k.replace("lora_A.default.weight", "")
            b = lw[base + "lora_B.default.weight"]
            sd[base + "weight"] = sd[base + "weight"] + lora.scaling * b @ a
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all("rotary" in x or "inv_freq" in x for x in missing), (missing, unexpected)
    return m


@pytest.mark.parametrize("kv_heads,theta", [(2, 10000.0), (4, 500000.0)])
def test_llama_lora_loss_and_logits_match_hf(kv_heads, theta):
    cfg = so.LlmCfg(vocab=300, d=64, layers=2, heads=4, kv_heads=kv_heads, ffn=176, rope_theta=theta, eps=1e-5)
    lora = so.LoraCfg(r=4, alpha=16, targets=("q_proj", "v_proj", "down_proj"))
    w, lw = so.init_llm(cfg, seed=5, std=0.08), so.init_lora(cfg, lora, seed=6, b_std=0.1)
    hf = _hf_llama(cfg, w, lw, lora)
    g = torch.Generator().manual_seed(7)
    B, S = 3, 21
    x = torch.randn(B, S, cfg.d, generator=g)
    att = torch.ones(B, S, dtype=torch.bool)
    att[0, :4] = False          # left padding
    att[1, 17:] = False         # right padding
    labels = torch.randint(0, cfg.vocab, (B, S), generator=g)
    labels[~att] = -100
    labels[:, :6] = -100
    logits = so.llama_forward(w, lw, cfg, lora, x, att)
    loss = so.causal_lm_loss(logits, labels)
    with torch.no_grad():
        ref = hf(inputs_embeds=x, attention_mask=att.long(), labels=labels)
    valid = att[:, :, None].expand_as(logits)
    assert (logits[valid] - ref.logits[valid]).abs().max().item() < 2e-4
    assert abs(loss.item() - ref.loss.item()) < 1e-5


def test_accuracy_and_merge_follow_reference_semantics():
    preds = torch.tensor([[1, 2, 3, 4], [5, 6, 7, 8]])
    tgts = torch.tensor([[1, 0, -100, 4], [-100, 6, 7, 0]])
    assert so.compute_accuracy(preds, tgts, -100).item() == pytest.approx(4 / 6)
    embed = torch.arange(40, dtype=torch.float32).view(10, 4)
    ids = torch.tensor([[-1, -1, -1, 3, 4], [7, -1, -1, 2, 0]])
    mask = torch.tensor([[1, 1, 1, 0, 0], [0, 1, 1, 0, 0]], dtype=torch.bool)
    aud = torch.full((2, 2, 4), 100.0)
    aud[1] = 200.0
    x = so.merge(embed, ids, mask, aud)
    # sample 0: three masked slots but only two audio frames (clamp) -> third masked row is zero
    assert torch.equal(x[0, 0], aud[0, 0]) and torch.equal(x[0, 1], aud[0, 1]) and x[0, 2].abs().sum() == 0
    assert torch.equal(x[0, 3], embed[3]) and torch.equal(x[1, 0], embed[7]) and torch.equal(x[1, 1], aud[1, 0])


def test_adamw_schedule_restatement():
    assert so.lr_lambda(0, 10, 100) == 0.0 and so.lr_lambda(5, 10, 100) == 0.5
    assert so.lr_lambda(10, 10, 100) == 1.0 and so.lr_lambda(100, 10, 100) == 0.0
    assert so.lr_lambda(55, 10, 100) == pytest.approx(0.5)


def test_lora_branch_matches_peft_formula_and_grads():
    # peft 0.6: result = F.linear(x, W) + lora_B(lora_A(x)) * scaling ; grads flow only to A and B
    g = torch.Generator().manual_seed(8)
    x = torch.randn(5, 16, generator=g)
    W = torch.randn(12, 16, generator=g)
    A = torch.randn(4, 16, generator=g).requires_grad_(True)
    Bm = torch.randn(12, 4, generator=g).requires_grad_(True)
    lora = so.LoraCfg(r=4, alpha=8, targets=("q_proj",))
    y = so.lora_linear(x, W, {"p.lora_A.default.weight": A, "p.lora_B.default.weight": Bm}, "p.", lora)
    dy = torch.randn(5, 12, generator=g)
    y.backward(dy)
    s = 2.0
    assert torch.allclose(y.detach(), x @ W.t() + s * (x @ A.t()) @ Bm.t(), atol=1e-5)
    assert torch.allclose(Bm.grad, s * dy.t() @ (x @ A.t()).detach(), atol=1e-5)
    assert torch.allclose(A.grad, s * (dy @ Bm.detach()).t() @ x, atol=1e-5)


def test_qwen2_style_decoder_matches_hf_qwen2():
    """q/k/v biases + tied embeddings (the s2s recipes' Qwen2-0.5B): the oracle's decoder vs HF Qwen2ForCausalLM (eager)."""
    from transformers import Qwen2Config, Qwen2ForCausalLM
    cfg = so.LlmCfg(300, 128, 2, 4, 2, 256, 1e6, 1e-6, True, True)
    w = so.init_llm(cfg, seed=3)
    assert "lm_head.weight" not in w and "model.layers.0.self_attn.q_proj.bias" in w and "model.layers.0.self_attn.o_proj.bias" not in w
    hc = Qwen2Config(vocab_size=cfg.vocab, hidden_size=cfg.d, intermediate_size=cfg.ffn, num_hidden_layers=cfg.layers, num_attention_heads=cfg.heads,
                     num_key_value_heads=cfg.kv_heads, rms_norm_eps=cfg.eps, rope_theta=cfg.rope_theta, tie_word_embeddings=True,
                     max_position_embeddings=512, use_sliding_window=False, attn_implementation="eager")
    m = Qwen2ForCausalLM(hc).eval()
    missing, unexpected = m.load_state_dict(w, strict=False)
    assert not unexpected and set(missing) <= {"lm_head.weight"}
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 17, cfg.d, generator=g)
    att = torch.ones(2, 17, dtype=torch.bool)
    att[0, :3] = False
    labels = torch.randint(0, cfg.vocab, (2, 17), generator=g)
    labels[:, :5] = -100
    out = m(inputs_embeds=x, attention_mask=att, labels=labels)
    lg = so.llama_forward(w, {}, cfg, None, x, att)
    sel = att[:, :, None].expand_as(lg)
    assert (lg[sel] - out.logits[sel]).abs().max().item() < 2e-4
    assert abs(so.causal_lm_loss(lg, labels).item() - out.loss.item()) < 1e-5
