"""Time the full-size step (random-init weights, synthetic batch) phase by phase with CUDA events."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slam_llm_b200 import config as C, ops
from slam_llm_b200.engine import SlamStepB200

ap = argparse.ArgumentParser()
ap.add_argument("--enc", default="base")
ap.add_argument("--llm", default="llama-3-8b")
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--r", type=int, default=16)
ap.add_argument("--steps", type=int, default=5)
args = ap.parse_args()

enc, llm = C.WHISPER[args.enc], C.LLM[args.llm]
lora, proj = C.LoraCfg(args.r, 32, ("q_proj", "v_proj")), C.ProjCfg("linear", 5, 2048)
t0 = time.time()
eng = SlamStepB200(enc, llm, lora, proj, device="cuda:0", lora_b_std=0.02)
torch.cuda.synchronize()
print(f"build {time.time() - t0:.1f}s, mem {torch.cuda.memory_allocated() / 2**30:.1f} GiB", flush=True)

B, S_audio, prompt, answer = args.batch, 300, 24, 76
S = S_audio + prompt + answer + 1
g = torch.Generator().manual_seed(0)
ids = torch.randint(0, llm.vocab, (B, S), generator=g)
ids[:, :S_audio] = -1
labels = torch.full((B, S), -100, dtype=torch.int64)
labels[:, S_audio + prompt:] = ids[:, S_audio + prompt:]
batch = dict(input_ids=ids.cuda(), labels=labels.cuda(), attention_mask=torch.ones(B, S, dtype=torch.bool).cuda(),
             modality_mask=(ids == -1).cuda(), audio_pcm=(torch.randn(B, 480000, generator=g) * 0.1).cuda())
rows, tgts = SlamStepB200.label_rows(labels)
batch["_rows"], batch["_targets"] = rows.cuda(), tgts.cuda()

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

for step in range(args.steps):
    l0 = ops.launch_count()
    e = [ev()]
    mel = eng.log_mel(batch["audio_pcm"]); e.append(ev())
    b2 = dict(batch); b2["audio_mel"] = mel
    eng.llm.pack_lora(); e.append(ev())
    enc_out = eng.encoder.forward(mel); e.append(ev())
    loss, acc, _ = eng.forward(b2, train=True); e.append(ev())   # includes a second encoder pass (subtract)
    eng.backward(); e.append(ev())
    eng.optimizer_step(1e-4); e.append(ev())
    torch.cuda.synchronize()
    names = ["logmel", "pack_lora", "encoder", "forward(incl enc+pack)", "backward", "adamw"]
    ts = {n: round(e[i].elapsed_time(e[i + 1]), 3) for i, n in enumerate(names)}
    ts["launches"] = ops.launch_count() - l0
    ts["loss"] = round(loss.item(), 4)
    print(json.dumps(ts), flush=True)

# clean whole-step timing
for _ in range(2):
    eng.train_step(batch)
torch.cuda.synchronize()
e0 = ev()
n = 5
for _ in range(n):
    eng.train_step(batch)
e1 = ev(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
fl = C.step_flops(enc, llm, proj, lora, B, 3000, S, n_label_rows=rows.numel())
print(json.dumps({"ms_per_step": round(ms, 2), "audio_s_per_s": round(B * 30 / ms * 1e3, 1), "tflops_alg": round(fl["total"] / 1e12, 2),
                  "achieved_tflops": round(fl["total"] / ms / 1e9, 1), "mem_GiB": round(torch.cuda.max_memory_allocated() / 2**30, 1)}))
