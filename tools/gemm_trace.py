"""Where does the time of a CTA-pair GEMM launch go INSIDE the training step?  (debug build with in-kernel time stamps)

    SLAM_NVCC_EXTRA=-DSLAM_GEMM_TRACE SLAM_B200_LIB_OUT=slam_llm_b200/libslam_b200_trace.so python -m slam_llm_b200.build
    SLAM_B200_LIB=slam_llm_b200/libslam_b200_trace.so python tools/gemm_trace.py > gpurun_out/gemm_trace.log

Runs eager C3 steps, records the stamps of every pair-kernel launch of the last step (gemm_2cta.cuh SLAM_TRACE events), and prints per launch
(times in us, globaltimer; medians over the CTAs unless stated):
    dur     first CTA entry -> last CTA exit                         gap    previous pair launch's last exit -> this launch's first entry
    pro     entry -> prologue done                                   wait   prologue done -> griddepcontrol.wait returned
    fill    wait returned -> first operands landed (TMA latency)     mma    first operands -> last MMA issued (all tiles)
    drain   last MMA issued -> accumulator complete                  epi    accumulator complete -> last tile's epilogue done
    skew    max - min over CTAs of "role finished"
and totals per phase over the step."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from slam_llm_b200 import config as Cfg, lib as L
from slam_llm_b200.engine import SlamStepB200

EV, CTAS, MAXL = 16, 160, 512
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
wl = bench.WORKLOADS["c3"]
enc, llm = Cfg.WHISPER[wl["enc"]], Cfg.LLM[wl["llm"]]
lora = Cfg.LoraCfg(wl["r"], wl["alpha"], tuple(wl["targets"]))
model, _ = bench.build_recipe_model(wl, enc, llm, lora, 1)
eng = model.b200
hb, S = bench.make_batch(wl, llm.vocab, seed=42)
rows, tgts = SlamStepB200.label_rows(hb["labels"])
hb["_rows"], hb["_targets"] = rows, tgts
batch = {k: v.to(dev) for k, v in hb.items()}
lib = L.load()
lib.slam_debug_gemm_trace.argtypes = [C.c_void_p, C.c_int]
lib.slam_debug_gemm_trace.restype = C.c_int
for _ in range(3):
    eng.train_step(batch, lr=1e-4)
torch.cuda.synchronize()
buf = torch.zeros(MAXL * CTAS * EV * 2, dtype=torch.int64, device=dev)
lib.slam_debug_gemm_trace(C.c_void_p(buf.data_ptr()), MAXL)
eng.train_step(batch, lr=1e-4)
torch.cuda.synchronize()
n = lib.slam_debug_gemm_trace(None, 0)
t = buf.cpu().numpy().reshape(MAXL, CTAS, EV, 2).astype(np.float64)
print(f"# {n} pair-kernel launches in the step", flush=True)
tot = {}
prev_exit = None
print(f"{'#':>4s} {'ctas':>4s} {'tiles':>5s} {'dur':>7s} {'gap':>6s} {'pro':>5s} {'wait':>6s} {'fill':>5s} {'fill9':>5s} {'mma':>7s} {'drain':>5s} {'epi':>5s} {'skew':>5s}")
for i in range(min(n, MAXL)):
    g = t[i, :, :, 0]
    live = g[:, 0] > 0
    if not live.any():
        continue
    g = g[live]
    lead = g[0::2]                                             # leader CTAs (even blockIdx) carry the MMA events
    us = lambda a: float(np.median(a)) / 1e3
    last_mma = np.where(lead[:, 6] > 0, lead[:, 6], lead[:, 5])
    acc_done = np.where(lead[:, 9] > 0, lead[:, 9], lead[:, 7])
    epi_done = np.where(lead[:, 10] > 0, lead[:, 10], lead[:, 8])
    row = dict(dur=(g[:, 12].max() - g[:, 0].min()) / 1e3, gap=((g[:, 0].min() - prev_exit) / 1e3 if prev_exit else 0.0), pro=us(g[:, 1] - g[:, 0]),
               wait=us(g[:, 2] - g[:, 1]), fill=us(lead[:, 4] - lead[:, 2]), fill9=us(lead[:, 14] - lead[:, 4]), mma=us(last_mma - lead[:, 4]),
               drain=us(acc_done - last_mma), epi=us(epi_done - acc_done), skew=(g[:, 11].max() - g[:, 11].min()) / 1e3)
    prev_exit = g[:, 12].max()
    for k, v in row.items():
        tot[k] = tot.get(k, 0.0) + v
    print(f"{i:4d} {len(g):4d} {int(g[:, 13].max()):5d} {row['dur']:7.1f} {row['gap']:6.1f} {row['pro']:5.1f} {row['wait']:6.1f} {row['fill']:5.1f} {row['fill9']:5.1f} "
          f"{row['mma']:7.1f} {row['drain']:5.1f} {row['epi']:5.1f} {row['skew']:5.1f}")
print("# totals over the step (ms): " + "  ".join(f"{k} {v / 1e3:.2f}" for k, v in tot.items()))
