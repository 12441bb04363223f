"""Where does the recipe surface lose time against engine.train_step?  (run under gpurun, one GPU)"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from slam_llm_b200 import config as C

wl = bench.WORKLOADS["c3"]
enc, llm = C.WHISPER[wl["enc"]], C.LLM[wl["llm"]]
lora = C.LoraCfg(wl["r"], wl["alpha"], tuple(wl["targets"]))
torch.cuda.set_device(0)
model, train_config = bench.build_recipe_model(wl, enc, llm, lora, 1)
eng = model.b200
from slam_llm.utils.train_utils import _move_batch, train
from slam_llm_b200.optim import FlatAdamW
from slam_llm_b200.engine import SlamStepB200
from omegaconf import OmegaConf
hb, S = bench.make_batch(wl, llm.vocab, seed=1, pin=True)
rows, tg = SlamStepB200.label_rows(hb["labels"])
hb["_rows"], hb["_targets"] = rows.pin_memory(), tg.pin_memory()
db = {k: v.cuda() for k, v in hb.items()}
opt = FlatAdamW(model, lr=1e-4)
model.train()
N = 12


def timeit(name, fn, n=N):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    print(f"{name:60s} {(time.perf_counter() - t0) * 1e3 / n:8.2f} ms/step", flush=True)


def a():
    loss, acc = eng.train_step(db, lr=1e-4)
    return loss.item()


def b():
    out, acc = model(**db)
    out.loss.backward()
    opt.step(); opt.zero_grad()
    return out.loss.item()


def c():
    batch = _move_batch({k: v for k, v in hb.items() if not k.startswith("_")}, torch.device("cuda:0"))
    out, acc = model(**batch)
    out.loss.backward()
    opt.step(); opt.zero_grad()
    return out.loss.item()


def c2():
    batch = _move_batch({k: v for k, v in hb.items() if not k.startswith("_")}, torch.device("cuda:0"))
    out, acc = model(**batch)
    out.loss.backward()
    opt.step(); opt.zero_grad()
    return f"{out.loss.detach().float()} {acc}"


timeit("A engine.train_step(resident) + loss.item()", a)
timeit("B model(**batch) + loss.backward + FlatAdamW (resident)", b)
timeit("C B + _move_batch from pinned host (label rows on CPU)", c)
timeit("C2 C + f-string of loss AND acc (two D2H reads)", c2)
log_config = OmegaConf.create(dict(use_wandb=False, log_interval=10))


class ListLoader:
    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def __iter__(self):
        for _ in range(self.n):
            yield {k: v for k, v in hb.items() if not k.startswith("_")}


for name, loader in (("D train() over pre-collated pinned batches (no DataLoader)", ListLoader(N)),):
    train(model, ListLoader(4), None, None, opt, None, 1, train_config, log_config)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    train(model, loader, None, None, opt, None, 1, train_config, log_config)
    torch.cuda.synchronize()
    print(f"{name:60s} {(time.perf_counter() - t0) * 1e3 / N:8.2f} ms/step (whole epoch incl. MemoryTrace)", flush=True)
ds = bench._SyntheticUtterances(wl, llm.vocab, N * wl["batch"], seed=2)
for workers in (2, 4):
    dl = torch.utils.data.DataLoader(ds, batch_size=wl["batch"], num_workers=workers, pin_memory=True, collate_fn=ds.collator, drop_last=True,
                                     persistent_workers=True, prefetch_factor=4)
    train(model, dl, None, None, opt, None, 1, train_config, log_config)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    train(model, dl, None, None, opt, None, 1, train_config, log_config)
    torch.cuda.synchronize()
    print(f"E train() over DataLoader({workers} persistent workers)                  {(time.perf_counter() - t0) * 1e3 / N:8.2f} ms/step (whole epoch)", flush=True)
    t0 = time.perf_counter()
    n = 0
    for batch in dl:
        n += 1
    print(f"  DataLoader({workers}) alone: {(time.perf_counter() - t0) * 1e3 / n:8.2f} ms/batch", flush=True)
    del dl
