"""SLAM-Omni-style recipe on the slam_llm surface: Whisper encoder + projector + FULL fine-tune of a Qwen2-architecture decoder
with an expanded vocabulary, multi-layer token inputs, group cross-entropy built BY THE RECIPE from `outputs.logits`
(tests/recipe_s2s_model.py, following examples/s2s/model/slam_model_s2s.py:160-306).  The decoder runs as one autograd node that is
differentiable w.r.t. the returned logits (train_config.b200_full_logits=true); loss and the gradient of every parameter are compared with the
oracle's restatement (oracle.s2s_step).  Tolerances: loss rel <= 5e-3; gradients cosine >= 0.99, rel-L2 <= 3e-2."""
import json
import os

import pytest
import torch

from oracle import slam_oracle as so
from parity_util import round_frozen

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
L, TV, AV = 3, 400, 80


def cosine(a, b):
    return torch.nn.functional.cosine_similarity(a.float().cpu().flatten(), b.float().cpu().flatten(), dim=0).item()


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _assets(tmp, om, enc, llm_ckpt_vocab):
    from safetensors.torch import save_file
    llm_dir = os.path.join(tmp, "llm")
    os.makedirs(llm_dir)
    c = om.llm_cfg
    json.dump(dict(model_type="qwen2", vocab_size=llm_ckpt_vocab, hidden_size=c.d, intermediate_size=c.ffn, num_hidden_layers=c.layers,
                   num_attention_heads=c.heads, num_key_value_heads=c.kv_heads, rms_norm_eps=c.eps, rope_theta=c.rope_theta, tie_word_embeddings=True,
                   use_sliding_window=False), open(os.path.join(llm_dir, "config.json"), "w"))
    sd = {k: v.contiguous() for k, v in om.llm_w.items()}
    save_file(sd, os.path.join(llm_dir, "model.safetensors"))
    wpt = os.path.join(tmp, "whisper.pt")
    dims = dict(n_mels=enc.n_mels, n_audio_ctx=enc.n_ctx, n_audio_state=enc.d, n_audio_head=enc.heads, n_audio_layer=enc.layers, n_vocab=8, n_text_ctx=8,
                n_text_state=enc.d, n_text_head=enc.heads, n_text_layer=0)
    torch.save({"dims": dims, "model_state_dict": {"encoder." + k: v for k, v in om.enc_w.items()}}, wpt)
    ckpt = os.path.join(tmp, "model.pt")
    torch.save({f"encoder_projector.{k}": v for k, v in om.proj_w.items()}, ckpt)
    return llm_dir, wpt, ckpt


def test_s2s_style_recipe_full_finetune_matches_oracle(tmp_path):
    import slam_llm  # noqa: F401
    from omegaconf import OmegaConf
    from slam_llm.utils.dataset_utils import load_module_from_py_file
    from slam_llm_b200.optim import FlatAdamW
    enc = so.EncoderCfg(80, 1500, 128, 2, 1)
    cfg = so.LlmCfg(TV + L * AV, 256, 2, 4, 1, 384, 1000000.0, 1e-6, True, True)
    om = round_frozen(so.OracleModel.build(enc, cfg, None, so.ProjCfg("linear", 5, 2048), seed=21))
    om.train_llm = True
    llm_dir, wpt, ckpt = _assets(str(tmp_path), om, enc, cfg.vocab)
    tc = OmegaConf.create(dict(model_name="s2s", enable_fsdp=False, enable_ddp=False, quantization=False, freeze_llm=False, freeze_encoder=True, use_peft=False,
                               seed=42, b200_full_logits=True))
    mc = OmegaConf.create(dict(llm_name="qwen2-toy", llm_path=llm_dir, llm_dim=cfg.d, encoder_name="whisper", encoder_path=wpt, encoder_dim=enc.d,
                               encoder_projector="linear", encoder_projector_ds_rate=5, group_decode=False,
                               vocab_config=dict(code_layer=L, padded_text_vocabsize=TV, padded_audio_vocabsize=AV, total_vocabsize=TV + L * AV)))
    plugin = load_module_from_py_file(os.path.join(HERE, "recipe_s2s_model.py"))
    model, _ = plugin.model_factory(tc, mc, ckpt_path=ckpt)
    model.train()
    names = {n for n, p in model.named_parameters() if p.requires_grad}
    assert "llm.model.embed_tokens.weight" in names and "llm.model.layers.1.mlp.down_proj.weight" in names and "llm.model.layers.0.self_attn.q_proj.bias" in names
    assert "llm.lm_head.weight" not in names                                        # tied
    batch = so.s2s_synthetic_batch(2, 32000, L, TV, AV, seed=5)
    ref = so.s2s_step(om, {k: v.clone() for k, v in batch.items()}, L, TV, AV)
    outputs, text_acc, _, losses = model(**{k: v.cuda() for k, v in batch.items()})
    assert abs(outputs.loss.item() - ref["loss"].item()) <= 5e-3 * abs(ref["loss"].item()), (outputs.loss.item(), ref["loss"].item())
    for got, want in zip(losses, ref["layer_loss"]):
        assert abs(got.item() - want.item()) <= 1e-2 * abs(want.item())
    outputs.loss.backward()
    named = dict(model.named_parameters())
    gmax = max(g.norm().item() for g in ref["grads"].values())
    checked = 0
    for k, g_ref in ref["grads"].items():
        if g_ref.norm().item() < 1e-3 * gmax:
            continue
        g = named[k].grad
        assert g is not None, k
        assert cosine(g, g_ref) > 0.99 and rel_l2(g, g_ref) < 3e-2, (k, cosine(g, g_ref), rel_l2(g, g_ref))
        checked += 1
    assert checked >= 15, checked
    # ... and with what the REFERENCE'S OWN slam_model_s2s produced on the CPU for this configuration (tests/golden/ref_s2s.pt)
    import ref_fixture as rf
    fix = rf.load("ref_s2s.pt")
    assert fix["cfg"]["seed"] == 21 and fix["cfg"]["batch_seed"] == 5 and tuple(fix["cfg"]["llm"]) == tuple(vars(cfg).values())
    assert abs(outputs.loss.item() - fix["loss"]) <= 5e-3 * abs(fix["loss"]), (outputs.loss.item(), fix["loss"])
    fmax = max((g["norm"] if rf.is_probe(g) else g.norm().item()) for g in fix["grads"].values())
    for k, g_ref in fix["grads"].items():
        g = named[k].grad
        if rf.is_probe(g_ref):
            if g_ref["norm"] >= 1e-3 * fmax:
                rf.check_probe(g, g_ref, norm_rel=3e-2, head_cos=0.99, what=k)
        elif g_ref.norm().item() >= 1e-3 * fmax:
            assert cosine(g, g_ref) > 0.99 and rel_l2(g, g_ref) < 3e-2, (k, cosine(g, g_ref), rel_l2(g, g_ref))
    # p.grad are arena views again; one optimizer step lowers the recipe's loss
    assert named["llm.model.embed_tokens.weight"].grad.data_ptr() == model.b200.trainable_state("grad")["llm.model.embed_tokens.weight"].data_ptr()
    opt = FlatAdamW(model, lr=1e-3)
    opt.step(); opt.zero_grad()
    out2, _, _, _ = model(**{k: v.cuda() for k, v in batch.items()})
    assert out2.loss.item() < outputs.loss.item()


def test_recipe_parameters_outside_the_arena_are_trained_too(tmp_path):
    """group_decode_adapter-style module (a torch nn.Linear on the audio logits, examples/s2s/utils/projector_utils.py): its gradient comes from
    autograd and FlatAdamW's inner AdamW steps it; a grown vocabulary (checkpoint smaller than vocab_config.total_vocabsize) gets new rows."""
    import slam_llm  # noqa: F401
    from omegaconf import OmegaConf
    from slam_llm.utils.dataset_utils import load_module_from_py_file
    from slam_llm_b200.optim import FlatAdamW
    enc = so.EncoderCfg(80, 1500, 128, 2, 1)
    small = so.LlmCfg(TV, 256, 2, 4, 1, 384, 1000000.0, 1e-6, True, True)           # checkpoint: text vocabulary only
    om = round_frozen(so.OracleModel.build(enc, small, None, so.ProjCfg("linear", 5, 2048), seed=22))
    llm_dir, wpt, ckpt = _assets(str(tmp_path), om, enc, TV)
    tc = OmegaConf.create(dict(model_name="s2s", enable_fsdp=False, enable_ddp=False, quantization=False, freeze_llm=False, freeze_encoder=True, use_peft=False,
                               seed=42, b200_full_logits=True))
    mc = OmegaConf.create(dict(llm_name="qwen2-toy", llm_path=llm_dir, llm_dim=small.d, encoder_name="whisper", encoder_path=wpt, encoder_dim=enc.d,
                               encoder_projector="linear", encoder_projector_ds_rate=5, group_decode=True,
                               vocab_config=dict(code_layer=L, padded_text_vocabsize=TV, padded_audio_vocabsize=AV, total_vocabsize=TV + L * AV)))
    plugin = load_module_from_py_file(os.path.join(HERE, "recipe_s2s_model.py"))
    model, _ = plugin.model_factory(tc, mc, ckpt_path=ckpt)
    model.train()
    E = dict(model.named_parameters())["llm.model.embed_tokens.weight"]
    assert E.shape[0] == TV + L * AV and torch.equal(E[:TV].detach().cpu(), om.llm_w["model.embed_tokens.weight"])
    assert E[TV:].detach().abs().max().item() > 0                                   # grown rows are initialised, not zeros
    opt = FlatAdamW(model, lr=2e-3)
    assert len(opt.foreign) == 1 and opt.foreign[0] is model.group_decode_adapter.weight
    batch = {k: v.cuda() for k, v in so.s2s_synthetic_batch(2, 32000, L, TV, AV, seed=6).items()}
    w0 = model.group_decode_adapter.weight.detach().clone()
    losses = []
    for _ in range(4):
        out, _, _, _ = model(**batch)
        out.loss.backward()
        opt.step(); opt.zero_grad()
        losses.append(out.loss.item())
    assert losses[-1] < losses[0] and not torch.equal(w0, model.group_decode_adapter.weight.detach())
