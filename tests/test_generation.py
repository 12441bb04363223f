"""slam_llm_b200/generation.py (the control flow behind slam_model.generate) against HuggingFace `generate(inputs_embeds=...)` on the CPU:
same tiny LlamaForCausalLM, same knobs the reference passes (models/slam_model.py:439-454) -> identical token ids.  The next-token logits
come from the HF model itself here, so this pins the search / processor logic alone; the GPU test pins the B200 decoder under it."""
import pytest
import torch

from slam_llm_b200.generation import generate


def _tiny_lm(vocab=40, seed=0):
    from transformers import LlamaConfig, LlamaForCausalLM
    torch.manual_seed(seed)
    hc = LlamaConfig(vocab_size=vocab, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                     max_position_embeddings=256, attn_implementation="eager", tie_word_embeddings=False)
    m = LlamaForCausalLM(hc).eval()
    with torch.no_grad():
        m.lm_head.weight.mul_(6.0)            # spread the logits: top-1 margins far above fp noise
    return m


def _next_logits_fn(m, prompt_embeds, mask, nb):
    def fn(tokens, beam_src):
        rows = tokens.shape[0]
        pe = prompt_embeds.repeat_interleave(rows // prompt_embeds.shape[0], dim=0)
        am = mask.repeat_interleave(rows // mask.shape[0], dim=0)
        x = torch.cat([pe, m.model.embed_tokens(tokens)], dim=1)
        am = torch.cat([am, torch.ones(rows, tokens.shape[1], dtype=am.dtype)], dim=1)
        return m(inputs_embeds=x, attention_mask=am).logits[:, -1]
    return fn


@pytest.mark.parametrize("kw", [
    dict(num_beams=1, max_new_tokens=14, min_length=1, repetition_penalty=1.0, length_penalty=1.0),
    dict(num_beams=1, max_new_tokens=10, min_length=4, repetition_penalty=1.3, length_penalty=1.0),
    dict(num_beams=4, max_new_tokens=14, min_length=1, repetition_penalty=1.0, length_penalty=1.0),      # the reference defaults (slam_model.py:441-449)
    dict(num_beams=4, max_new_tokens=9, min_length=3, repetition_penalty=1.2, length_penalty=2.0),
    dict(num_beams=3, max_new_tokens=12, min_length=1, repetition_penalty=1.0, length_penalty=0.5),
])
def test_generation_matches_hf_generate(kw):
    eos, pad, bos = 2, 2, 1                                   # pad = eos as setup_tokenizer sets it (slam_model.py:64)
    for seed in (0, 1, 2):
        m = _tiny_lm(seed=seed)
        g = torch.Generator().manual_seed(10 + seed)
        B, S = 3, 7
        emb = torch.randn(B, S, 32, generator=g) * 0.5
        mask = torch.ones(B, S, dtype=torch.long)
        mask[0, :2] = 0                                       # left padding, as the inference collator produces
        with torch.no_grad():
            want = m.generate(inputs_embeds=emb, attention_mask=mask, do_sample=False, top_p=1.0, temperature=1.0, bos_token_id=bos, eos_token_id=eos,
                              pad_token_id=pad, **kw)
            # the installed transformers (>= 4.50) counts the eos token in a finished hypothesis' length; v4.35.2 (the reference's pin,
            # this module's default) does not - everything else must agree
            got = generate(_next_logits_fn(m, emb, mask, kw["num_beams"]), B, do_sample=False, eos_token_id=eos, pad_token_id=pad,
                           beam_length_counts_eos=True, **kw)
        width = max(want.shape[1], got.shape[1])
        pw = torch.full((B, width), pad)
        pg = torch.full((B, width), pad)
        pw[:, : want.shape[1]] = want
        pg[:, : got.shape[1]] = got
        assert torch.equal(pw, pg), (kw, seed, want.tolist(), got.tolist())


def test_sampling_is_seeded_and_respects_top_p():
    m = _tiny_lm(seed=3)
    emb = torch.randn(2, 5, 32, generator=torch.Generator().manual_seed(1)) * 0.5
    mask = torch.ones(2, 5, dtype=torch.long)
    fn = _next_logits_fn(m, emb, mask, 1)
    a = generate(fn, 2, num_beams=1, do_sample=True, top_p=0.8, temperature=0.7, max_new_tokens=8, eos_token_id=2, generator=torch.Generator().manual_seed(5))
    b = generate(fn, 2, num_beams=1, do_sample=True, top_p=0.8, temperature=0.7, max_new_tokens=8, eos_token_id=2, generator=torch.Generator().manual_seed(5))
    assert torch.equal(a, b)
    greedy = generate(fn, 2, num_beams=1, do_sample=False, max_new_tokens=8, eos_token_id=2)
    near_greedy = generate(fn, 2, num_beams=1, do_sample=True, top_p=1e-6, max_new_tokens=8, eos_token_id=2, generator=torch.Generator().manual_seed(1))
    assert torch.equal(greedy, near_greedy)                   # top_p -> 0 keeps only the arg-max token
