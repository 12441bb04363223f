"""Token generation control flow for `slam_model.generate` (reference: src/slam_llm/models/slam_model.py:409-456, which hands
`inputs_embeds` + the knobs below to HuggingFace `generate`).  Host logic only: the decoder forward that produces the next-token logits is
a callable supplied by the caller (the B200 decoder on the GPU; the fp32 oracle in the CPU tests).

Restated from the published transformers v4.35.2 algorithms (generation/utils.py greedy_search / beam_search / sample,
generation/beam_search.py BeamSearchScorer + BeamHypotheses, generation/logits_process.py), for the arguments the reference passes:
max_new_tokens, num_beams, do_sample, min_length, top_p, repetition_penalty, length_penalty, temperature, bos/eos/pad ids.
With `inputs_embeds` the generated sequence starts EMPTY (no prompt ids): processors and length penalties see generated tokens only.

  * greedy (num_beams = 1, do_sample = False): processors act on the raw logits; finished rows emit pad_token_id.
  * beam search (num_beams > 1): log_softmax first, processors on the log-probabilities, 2 x num_beams candidates per batch entry,
    hypotheses scored sum_logprobs / len ** length_penalty, early_stopping = False ("heuristic": stop when the worst kept hypothesis beats
    the best attainable score of the running beams), finalisation from the running beams, eos appended to finished hypotheses.
  * sampling (do_sample = True): temperature, top-p, torch.multinomial with the caller's generator (not bit-comparable with HF's CUDA RNG).
"""
from __future__ import annotations

from typing import Callable, List, Optional

import torch

NextLogits = Callable[[torch.Tensor, Optional[torch.Tensor]], torch.Tensor]
"""next_logits(tokens i64 [n, t] (generated so far, row-major over batch x beams), beam_src i64 [n] | None) -> f32 [n, vocab] on any device.
beam_src[i] = the row of the PREVIOUS call whose sequence row i continues (beam re-ordering; None on the first call / greedy)."""


def _repetition_penalty(scores: torch.Tensor, tokens: torch.Tensor, penalty: float) -> torch.Tensor:
    if penalty == 1.0 or tokens.shape[1] == 0:
        return scores
    picked = torch.gather(scores, 1, tokens)
    picked = torch.where(picked < 0, picked * penalty, picked / penalty)
    return scores.scatter(1, tokens, picked)


def _min_length(scores: torch.Tensor, cur_len: int, min_length: int, eos_token_id: Optional[int]) -> torch.Tensor:
    if eos_token_id is not None and min_length > 0 and cur_len < min_length:
        scores = scores.clone()
        scores[:, eos_token_id] = -float("inf")
    return scores


def _top_p(scores: torch.Tensor, top_p: float) -> torch.Tensor:
    if top_p >= 1.0:
        return scores
    sorted_logits, sorted_idx = torch.sort(scores, descending=False)
    cum = sorted_logits.softmax(dim=-1).cumsum(dim=-1)
    remove = cum <= (1 - top_p)
    remove[..., -1:] = False
    return scores.masked_fill(remove.scatter(1, sorted_idx, remove), -float("inf"))


@torch.no_grad()
def generate(next_logits: NextLogits, batch_size: int, *, max_new_tokens: int = 200, num_beams: int = 4, do_sample: bool = False, min_length: int = 1,
             top_p: float = 1.0, repetition_penalty: float = 1.0, length_penalty: float = 1.0, temperature: float = 1.0,
             eos_token_id: Optional[int] = None, pad_token_id: Optional[int] = None, generator: Optional[torch.Generator] = None,
             beam_length_counts_eos: bool = False) -> torch.Tensor:
    """-> i64 [batch, <= max_new_tokens] generated token ids (eos included, pad_token_id after it), on the CPU."""
    if pad_token_id is None:
        pad_token_id = eos_token_id if eos_token_id is not None else 0
    if num_beams > 1 and not do_sample:
        return _beam_search(next_logits, batch_size, max_new_tokens, num_beams, min_length, repetition_penalty, length_penalty, eos_token_id, pad_token_id,
                            count_eos=beam_length_counts_eos)
    if num_beams > 1:
        raise NotImplementedError("beam-search multinomial sampling (num_beams > 1 with do_sample=True) is not implemented")
    tokens = torch.zeros((batch_size, 0), dtype=torch.int64)
    unfinished = torch.ones(batch_size, dtype=torch.int64)
    for step in range(max_new_tokens):
        scores = next_logits(tokens, None).float().cpu()
        scores = _repetition_penalty(scores, tokens, repetition_penalty)
        scores = _min_length(scores, step, min_length, eos_token_id)
        if do_sample:
            if temperature != 1.0:
                scores = scores / temperature
            scores = _top_p(scores, top_p)
            nxt = torch.multinomial(scores.softmax(dim=-1), 1, generator=generator).squeeze(1)
        else:
            nxt = scores.argmax(dim=-1)
        if eos_token_id is not None:
            nxt = nxt * unfinished + pad_token_id * (1 - unfinished)
        tokens = torch.cat([tokens, nxt[:, None]], dim=1)
        if eos_token_id is not None:
            unfinished = unfinished * (nxt != eos_token_id).long()
            if int(unfinished.max()) == 0:
                break
    return tokens


class _Hypotheses:
    def __init__(self, num_beams: int, length_penalty: float):
        self.num_beams, self.length_penalty = num_beams, length_penalty
        self.beams: List = []
        self.worst = 1e9

    def add(self, hyp: torch.Tensor, sum_logprobs: float, extra_len: int = 0) -> None:
        score = sum_logprobs / ((hyp.shape[-1] + extra_len) ** self.length_penalty)
        if len(self.beams) < self.num_beams or score > self.worst:
            self.beams.append((score, hyp))
            if len(self.beams) > self.num_beams:
                order = sorted((s, i) for i, (s, _) in enumerate(self.beams))
                del self.beams[order[0][1]]
                self.worst = order[1][0]
            else:
                self.worst = min(score, self.worst)

    def is_done(self, best_sum_logprobs: float, cur_len: int) -> bool:
        if len(self.beams) < self.num_beams:
            return False
        return self.worst >= best_sum_logprobs / cur_len ** self.length_penalty      # early_stopping = False


def _beam_search(next_logits, B, max_new_tokens, nb, min_length, repetition_penalty, length_penalty, eos, pad, count_eos: bool = False) -> torch.Tensor:
    """count_eos=False: transformers v4.35.2 (the reference's pin) - a finished hypothesis is normalised by its length WITHOUT the eos token
    (BeamHypotheses.add: sum_logprobs / hyp.shape[-1] ** length_penalty).  count_eos=True: the length includes the eos token, which is what the
    rewritten beam search of transformers >= 4.50 does; the CPU test uses it to check every other rule against the installed transformers."""
    eos_len = 1 if count_eos else 0
    tokens = torch.zeros((B * nb, 0), dtype=torch.int64)
    beam_scores = torch.zeros((B, nb))
    beam_scores[:, 1:] = -1e9
    beam_scores = beam_scores.view(-1)
    hyps = [_Hypotheses(nb, length_penalty) for _ in range(B)]
    done = [False] * B
    beam_src = None
    for step in range(max_new_tokens):
        logits = next_logits(tokens, beam_src).float().cpu()
        V = logits.shape[-1]
        scores = torch.log_softmax(logits, dim=-1)
        scores = _repetition_penalty(scores, tokens, repetition_penalty)
        scores = _min_length(scores, step, min_length, eos)
        scores = (scores + beam_scores[:, None]).view(B, nb * V)
        top_scores, top_idx = torch.topk(scores, 2 * nb, dim=1, largest=True, sorted=True)
        src_beam, tok = top_idx // V, top_idx % V
        new_scores, new_tok, new_src = torch.zeros(B, nb), torch.zeros(B, nb, dtype=torch.int64), torch.zeros(B, nb, dtype=torch.int64)
        cur_len = tokens.shape[1] + 1
        for b in range(B):
            if done[b]:
                new_scores[b], new_tok[b], new_src[b] = 0.0, pad, b * nb
                continue
            k = 0
            for rank in range(2 * nb):
                t, s, row = int(tok[b, rank]), float(top_scores[b, rank]), b * nb + int(src_beam[b, rank])
                if eos is not None and t == eos:
                    if rank >= nb:
                        continue
                    hyps[b].add(tokens[row].clone(), s, eos_len)
                else:
                    new_scores[b, k], new_tok[b, k], new_src[b, k] = s, t, row
                    k += 1
                if k == nb:
                    break
            done[b] = done[b] or hyps[b].is_done(float(top_scores[b].max()), cur_len)
        beam_scores = new_scores.view(-1)
        beam_src = new_src.view(-1)
        tokens = torch.cat([tokens[beam_src], new_tok.view(-1, 1)], dim=1)
        if all(done):
            break
    for b in range(B):                                               # finalize: running beams of unfinished entries become hypotheses
        if done[b]:
            continue
        for j in range(nb):
            hyps[b].add(tokens[b * nb + j], float(beam_scores[b * nb + j]))
    best = [sorted(h.beams, key=lambda x: x[0])[-1][1] for h in hyps]
    lengths = [int(h.shape[0]) for h in best]
    width = min(max(lengths) + 1, max_new_tokens)
    out = torch.full((B, width), pad, dtype=torch.int64)
    for b, h in enumerate(best):
        out[b, : lengths[b]] = h
        if lengths[b] < width and eos is not None:
            out[b, lengths[b]] = eos
    return out
