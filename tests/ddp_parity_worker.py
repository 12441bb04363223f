"""Worker of tests/test_ddp_gpu.py: launched as `python -m torch.distributed.run --nproc-per-node 2 tests/ddp_parity_worker.py`.

Reference semantics (DDP, pipeline/finetune.py:181-184): every rank computes the loss of ITS micro-batch, gradients are averaged over
ranks, every replica applies the same AdamW update.  With equally many labelled tokens per rank the rank-mean gradient equals the gradient of
the concatenated batch on one GPU, so:
    all-reduced arena.grad / world  ==  single-GPU grad on cat(batches)      (cosine >= 0.999, rel-L2 <= 2e-2: two bf16 evaluation orders)
    parameters after train_step() are bit-identical on all ranks and match the single-GPU update direction.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

from oracle import slam_oracle as so  # noqa: E402
from parity_util import round_frozen  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    from slam_llm_b200 import config as C
    from slam_llm_b200.engine import SlamStepB200
    enc, llm = so.EncoderCfg(80, 1500, 128, 2, 2), so.LlmCfg(512, 256, 2, 4, 2, 512, 10000.0, 1e-5)
    lora, proj = so.LoraCfg(8, 32, ("q_proj", "v_proj")), so.ProjCfg("linear", 5, 128)
    om = round_frozen(so.OracleModel.build(enc, llm, lora, proj, seed=9))

    def engine():
        return SlamStepB200(C.EncoderCfg(**vars(enc)), C.LlmCfg(**vars(llm)), C.LoraCfg(8, 32, ("q_proj", "v_proj")), C.ProjCfg(**vars(proj)),
                            device=dev, enc_weights=om.enc_w, llm_weights=om.llm_w, lora_weights=om.lora_w, proj_weights=om.proj_w)

    per = 2
    full = so.synthetic_batch(per * world, 32000, llm.vocab, prompt_len=6, answer_len=9, left_pad=[0, 2, 1, 0][: per * world], seed=17)
    mine = {k: v[rank * per:(rank + 1) * per].to(dev) for k, v in full.items()}
    eng = engine()
    loss, acc, _ = eng.forward(mine, train=True)
    eng.backward()
    local = eng.arena.grad.clone()
    dist.all_reduce(eng.arena.grad)
    mean_grad = eng.arena.grad / world
    # cross-check the collective itself against a gather of the local gradients
    parts = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(parts, local)
    assert torch.allclose(mean_grad, torch.stack(parts).sum(0) / world, rtol=1e-6, atol=1e-9)

    single = engine()
    loss1, _, _ = single.forward({k: v.to(dev) for k, v in full.items()}, train=True)
    single.backward()
    g1 = single.arena.grad
    cos = torch.nn.functional.cosine_similarity(mean_grad, g1, dim=0).item()
    rel = ((mean_grad - g1).norm() / g1.norm()).item()
    losses = [torch.zeros((), device=dev) for _ in range(world)]
    dist.all_gather(losses, loss.detach())
    mean_loss = torch.stack(losses).mean().item()
    assert abs(mean_loss - loss1.item()) <= 2e-3 * abs(loss1.item()), (mean_loss, loss1.item())
    assert cos > 0.999 and rel < 2e-2, (cos, rel)

    # the public step: train_step(world_size) = forward, backward, all-reduce, AdamW(grad / world)
    eng2, ref2 = engine(), engine()
    eng2.train_step(mine, lr=1e-3, world_size=world)
    ref2.train_step({k: v.to(dev) for k, v in full.items()}, lr=1e-3, world_size=1)
    p = eng2.arena.param.clone()
    gathered = [torch.empty_like(p) for _ in range(world)]
    dist.all_gather(gathered, p)
    assert all(torch.equal(gathered[0], g) for g in gathered), "replicas diverged after one step"
    p0 = engine().arena.param
    upd, upd_ref = p - p0, ref2.arena.param - p0
    ucos = torch.nn.functional.cosine_similarity(upd, upd_ref, dim=0).item()
    assert ucos > 0.97, ucos
    # deferred mode (async all-reduce, AdamW applied behind the next step's frozen front end) is the same arithmetic: 3 steps
    eng3, eng4 = engine(), engine()
    eng4.defer_update = True
    for _ in range(3):
        eng3.train_step(mine, lr=1e-3, world_size=world)
        eng4.train_step(mine, lr=1e-3, world_size=world)
    eng4.flush_update()
    torch.cuda.synchronize()
    drift = ((eng3.arena.param - eng4.arena.param).norm() / eng3.arena.param.norm()).item()
    assert drift < 1e-4, f"deferred update diverged from the blocking step: {drift}"   # equal up to fp32 atomics reordering
    if rank == 0:
        print(f"DDP_PARITY_OK world={world} grad_cos={cos:.6f} grad_rel={rel:.2e} update_cos={ucos:.4f} mean_loss={mean_loss:.5f} single_loss={loss1.item():.5f}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
