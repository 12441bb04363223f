"""Generate tests/golden/step_tiny.pt from the CPU oracle (run here, committed; the GPU box only reads it).

    python tests/golden/make_golden.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import slam_oracle as so  # noqa: E402
from parity_util import round_frozen  # noqa: E402

CFG = dict(enc=(80, 1500, 128, 2, 2), llm=(512, 256, 2, 4, 2, 512, 10000.0, 1e-5), lora=(8, 32, ("q_proj", "v_proj")), proj=("linear", 5, 128),
           seed=1234, batch_args=(2, 32000, 512), batch_kwargs=dict(prompt_len=6, answer_len=9, left_pad=[1, 0], seed=99))


def main():
    enc, llm = so.EncoderCfg(*CFG["enc"]), so.LlmCfg(*CFG["llm"])
    lora, proj = so.LoraCfg(CFG["lora"][0], CFG["lora"][1], CFG["lora"][2]), so.ProjCfg(*CFG["proj"])
    om = round_frozen(so.OracleModel.build(enc, llm, lora, proj, seed=CFG["seed"]))
    batch = so.synthetic_batch(*CFG["batch_args"], **CFG["batch_kwargs"])
    r = om.step(batch, do_update=False)
    probe = {k: (g.norm().item(), g.flatten()[:256].clone()) for k, g in r["grads"].items()}
    out = dict(cfg=CFG, loss=r["loss"].item(), acc=r["acc"].item(), grad_probe=probe,
               encoder_out_head=r["encoder_out"][0, :4, :16].clone(), logits_last=r["logits"][:, -1, :32].clone())
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "step_tiny.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes; loss", out["loss"], "acc", out["acc"])


if __name__ == "__main__":
    main()
