"""Full fine-tune of the decoder (train_config.freeze_llm=false — the SLAM-Omni recipes, examples/s2s: SURVEY §8 f3): every decoder
parameter trains.  CUDA step vs the oracle with `train_llm=True` on identical weights and batch: loss, and the gradient of EVERY parameter
(embedding table via the merge, q/k/v/o/gate/up/down weights via tcgen05 wgrad GEMMs, q/k/v biases, RMSNorm weights, lm_head — tied or not),
then one AdamW step and a second forward (the bf16 GEMM operands are re-derived from the updated fp32 masters).
Tolerances as in test_step_parity_gpu.py: loss rel <= 5e-3; gradients cosine >= 0.99, rel-L2 <= 3e-2 (tensors with non-negligible norm)."""
import pytest
import torch

from oracle import slam_oracle as so
from parity_util import round_frozen

pytestmark = pytest.mark.gpu

CASES = {
    # Llama architecture, untied lm_head, GQA 2:1, dh = 64
    "llama": dict(enc=so.EncoderCfg(80, 1500, 128, 2, 1), llm=so.LlmCfg(512, 256, 2, 4, 2, 512, 10000.0, 1e-5), proj=so.ProjCfg("linear", 5, 128)),
    # Qwen2 architecture (the s2s recipes' LLM): q/k/v biases, tied embeddings, eps 1e-6, theta 1e6, dh = 64, GQA 4:1... at toy width
    "qwen2": dict(enc=so.EncoderCfg(80, 1500, 128, 2, 1), llm=so.LlmCfg(640, 256, 2, 4, 1, 384, 1000000.0, 1e-6, True, True), proj=so.ProjCfg("linear", 5, 128)),
}


def cosine(a, b):
    return torch.nn.functional.cosine_similarity(a.float().cpu().flatten(), b.float().cpu().flatten(), dim=0).item()


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("name", list(CASES))
def test_full_finetune_step_matches_oracle(name):
    from slam_llm_b200 import config as C
    from slam_llm_b200.engine import SlamStepB200
    c = CASES[name]
    om = round_frozen(so.OracleModel.build(c["enc"], c["llm"], None, c["proj"], seed=11))
    om.train_llm = True
    eng = SlamStepB200(C.EncoderCfg(**vars(c["enc"])), C.LlmCfg(**vars(c["llm"])), None, C.ProjCfg(**vars(c["proj"])), device="cuda:0",
                       enc_weights=om.enc_w, llm_weights=om.llm_w, proj_weights=om.proj_w, train_llm=True)
    batch = so.synthetic_batch(2, 32000, c["llm"].vocab, prompt_len=6, answer_len=9, left_pad=[0, 3], seed=13)
    ref = om.step(dict(batch), lr=1e-3, weight_decay=0.0)
    gb = {k: v.cuda() for k, v in batch.items()}
    loss, acc, _ = eng.forward(gb, train=True)
    assert abs(loss.item() - ref["loss"].item()) <= 5e-3 * abs(ref["loss"].item()), (loss.item(), ref["loss"].item())
    eng.backward()
    grads = eng.trainable_state("grad")
    assert set(grads) == set(ref["grads"]), sorted(set(grads) ^ set(ref["grads"]))[:6]
    gmax = max(g.norm().item() for g in ref["grads"].values())
    checked = []
    for k, g_ref in ref["grads"].items():
        if g_ref.norm().item() < 1e-3 * gmax:
            continue
        g = grads[k]
        assert cosine(g, g_ref) > 0.99, (k, cosine(g, g_ref))
        assert rel_l2(g, g_ref) < 3e-2, (k, rel_l2(g, g_ref))
        checked.append(k)
    kinds = ("embed_tokens", "q_proj.weight", "k_proj.weight", "v_proj.weight", "o_proj.weight", "gate_proj", "up_proj", "down_proj", "layernorm", "model.norm",
             "encoder_projector") + (("q_proj.bias", "v_proj.bias") if c["llm"].qkv_bias else ("lm_head",))
    for kind in kinds:
        assert any(kind in k for k in checked), f"no gradient of kind {kind} was compared"
    # AdamW on the fp32 masters, then a second forward through re-derived bf16 operands
    eng.optimizer_step(1e-3, 0.0)
    loss2, _, _ = eng.forward(gb, train=False)
    ref2 = om.forward(dict(batch), return_all=True)
    assert abs(loss2.item() - ref2["loss"].item()) <= 1e-2 * abs(ref2["loss"].item()), (loss2.item(), ref2["loss"].item())
    assert loss2.item() < loss.item()                                   # one step on the same batch lowers the loss


def test_frozen_qwen2_with_lora_runs_biased_projections():
    """Qwen2 as a FROZEN base with LoRA on q/v (asr recipes with a Qwen2 LLM): the q/k/v biases ride in the fused base+LoRA GEMM epilogue."""
    from slam_llm_b200 import config as C
    from slam_llm_b200.engine import SlamStepB200
    c = CASES["qwen2"]
    lora = so.LoraCfg(8, 32, ("q_proj", "v_proj"))
    om = round_frozen(so.OracleModel.build(c["enc"], c["llm"], lora, c["proj"], seed=12))
    eng = SlamStepB200(C.EncoderCfg(**vars(c["enc"])), C.LlmCfg(**vars(c["llm"])), C.LoraCfg(8, 32, ("q_proj", "v_proj")), C.ProjCfg(**vars(c["proj"])),
                       device="cuda:0", enc_weights=om.enc_w, llm_weights=om.llm_w, lora_weights=om.lora_w, proj_weights=om.proj_w)
    batch = so.synthetic_batch(2, 32000, c["llm"].vocab, prompt_len=6, answer_len=9, left_pad=[1, 0], seed=14)
    ref = om.step(dict(batch), do_update=False)
    loss, _, _ = eng.forward({k: v.cuda() for k, v in batch.items()}, train=True)
    assert abs(loss.item() - ref["loss"].item()) <= 5e-3 * abs(ref["loss"].item())
    eng.backward()
    grads = eng.trainable_state("grad")
    gmax = max(g.norm().item() for g in ref["grads"].values())
    for k, g_ref in ref["grads"].items():
        if g_ref.norm().item() >= 1e-3 * gmax:
            assert cosine(grads[k], g_ref) > 0.99 and rel_l2(grads[k], g_ref) < 3e-2, (k, cosine(grads[k], g_ref), rel_l2(grads[k], g_ref))
