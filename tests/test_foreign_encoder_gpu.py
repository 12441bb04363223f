"""Alt-encoder recipes (aac_audiocaps / EAT): a FROZEN torch encoder from the user's environment produces the audio features;
projector, merge, decoder (+LoRA), loss, backward and optimizer run on the B200 step, entered at the projector.  The test registers a small
stand-in with EAT's calling convention (`encoder.model.extract_features(mel.unsqueeze(1), padding_mask=None, mask=False,
remove_extra_tokens=False)['x']`, models/slam_model.py:324-325 — the real EAT network is fairseq user code that this image does not have) and
requires loss and gradients to equal the oracle's, which is handed the same features."""
import os

import pytest
import torch
import torch.nn as nn

import ref_fixture as rf
from oracle import slam_oracle as so

pytestmark = pytest.mark.gpu


class _PatchEncoder(nn.Module):
    """EAT-shaped stand-in: 16 x 16 patches of a [B, 1, T, 128] fbank -> tokens + CLS, one linear mixing layer."""

    def __init__(self, d):
        super().__init__()
        torch.manual_seed(3)
        self.patch = nn.Conv2d(1, d, kernel_size=16, stride=16)
        self.cls = nn.Parameter(torch.randn(1, 1, d) * 0.02)
        self.mix = nn.Linear(d, d)

    def extract_features(self, x, padding_mask=None, mask=False, remove_extra_tokens=False):
        h = self.patch(x).flatten(2).transpose(1, 2)                       # [B, (T/16) * 8, d]
        h = torch.cat([self.cls.expand(h.shape[0], -1, -1), h], dim=1)
        return {"x": torch.tanh(self.mix(h))}


class _Wrapper(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.model = _PatchEncoder(d)


def test_foreign_encoder_features_enter_the_b200_step(tmp_path):
    import slam_llm  # noqa: F401
    from omegaconf import OmegaConf
    from slam_llm.models import encoder as enc_mod
    from slam_llm.models.slam_model import model_factory
    from test_loaders_gpu import _write_assets
    fix = rf.load("ref_tiny.pt")
    om = rf.oracle_model(fix)
    enc, llm, lora, proj = rf.cfgs(fix)
    d_feat = 64
    om.proj_w = so.init_projector(so.EncoderCfg(0, 0, d_feat, 1, 0), llm, proj, seed=5)
    mc, kwargs = _write_assets(str(tmp_path), fix, om, False, False)
    mc.update(encoder_name="toy-eat", encoder_dim=d_feat, encoder_path=None)
    enc_mod.register_encoder("toy-eat", lambda cfg: _Wrapper(d_feat), enc_mod.foreign_encoder("eat")[1])     # EAT's own calling convention
    tc = dict(model_name="aac", enable_fsdp=False, enable_ddp=False, quantization=False, freeze_llm=True, freeze_encoder=True, use_peft=True, seed=42,
              peft_config=dict(peft_method="lora", r=lora.r, lora_alpha=lora.alpha, target_modules=list(lora.targets), bias="none", task_type="CAUSAL_LM",
                               lora_dropout=0.0, inference_mode=False))
    model, _ = model_factory(OmegaConf.create(tc), OmegaConf.create(mc), metric="acc", **kwargs)
    model.train()
    assert model.b200.encoder is None and not any(p.requires_grad for p in model.encoder.parameters())
    # batch: fbank-like features [B, T, 128] with T a multiple of 16; audio_length = (T/16*8 + 1) // 5 tokens, right padding only
    g = torch.Generator().manual_seed(9)
    B, T = 2, 160
    mel = torch.randn(B, T, 128, generator=g)
    n_tok = T // 16 * 8 + 1
    ta = n_tok // 5
    S = ta + 5 + 7
    ids = torch.randint(3, llm.vocab, (B, S), generator=g)
    ids[:, :ta] = -1
    labels = ids.clone()
    labels[:, : ta + 5] = -100
    att = torch.ones(B, S, dtype=torch.bool)
    mod = torch.zeros(B, S, dtype=torch.bool)
    mod[:, :ta] = True
    batch = dict(input_ids=ids, labels=labels, attention_mask=att, modality_mask=mod, audio_mel=mel, audio_mel_mask=torch.ones(B, T))
    outputs, acc = model(**{k: v.cuda() for k, v in batch.items()})
    with torch.no_grad():
        feats = model.encoder.model.extract_features(mel.cuda().unsqueeze(1))["x"].float().cpu()
    assert feats.shape == (B, n_tok, d_feat)
    ref = om.step(dict(input_ids=ids.clone(), labels=labels, attention_mask=att, modality_mask=mod, encoder_out=feats.bfloat16().float()), do_update=False)
    assert abs(outputs.loss.item() - ref["loss"].item()) <= 5e-3 * abs(ref["loss"].item()), (outputs.loss.item(), ref["loss"].item())
    outputs.loss.backward()
    named = dict(model.named_parameters())
    gmax = max(g_.norm().item() for g_ in ref["grads"].values())
    checked = 0
    for k, g_ref in ref["grads"].items():
        if g_ref.norm().item() >= 1e-3 * gmax:
            assert rf.cosine(named[k].grad, g_ref) > 0.99 and rf.rel_l2(named[k].grad, g_ref) < 3e-2, (k, rf.cosine(named[k].grad, g_ref))
            checked += 1
    assert checked >= 6
    # decode entry: inputs_embeds through the foreign encoder, then generate
    model.eval()
    model.tokenizer = type("T", (), dict(bos_token_id=1, eos_token_id=2, pad_token_id=2))()
    gen = model.generate(input_ids=ids[:, : ta + 5].cuda(), attention_mask=att[:, : ta + 5].cuda(), modality_mask=mod[:, : ta + 5].cuda(), audio_mel=mel.cuda(),
                         max_new_tokens=4, num_beams=2)
    assert gen.shape[0] == B and gen.shape[1] <= 4
