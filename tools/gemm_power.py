"""Sustained (power-capped) throughput of GEMM configurations vs cuBLAS, measured FAIRLY (round-2 rewrite; the round-1 version ran cuBLAS
first after an idle gap for 0.7 s, so its power and rate were not steady-state):

  * the GPU is pre-heated for 3 s with back-to-back 8192^3 cuBLAS GEMMs, and never idles afterwards;
  * per shape the configurations are INTERLEAVED round-robin for ROUNDS rounds; every slot runs 0.4 s untimed (clock/power settle to
    this kernel's own steady state), then >= 0.8 s timed with CUDA events while NVML samples power and SM clock;
  * a configuration's figure is total flops / total timed seconds over its slots (3 x 0.8 s = 2.4 s of steady state).

    python tools/gemm_power.py [shape,shape,...] > gpurun_out/r02_gemm_power.log
"""
import os
import sys
import threading
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pynvml
import torch
from slam_llm_b200 import ops

pynvml.nvmlInit()
H = pynvml.nvmlDeviceGetHandleByIndex(0)
ROUNDS, SETTLE_S, TIMED_S = 3, 0.4, 0.8


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.stop = False
        self.p, self.c = [], []

    def run(self):
        while not self.stop:
            self.p.append(pynvml.nvmlDeviceGetPowerUsage(H) / 1000.0)
            self.c.append(pynvml.nvmlDeviceGetClockInfo(H, pynvml.NVML_CLOCK_SM))
            time.sleep(0.02)


def spin(fn, secs):
    n, t0 = 0, time.time()
    while time.time() - t0 < secs:
        for _ in range(50):
            fn()
        n += 50
        torch.cuda.synchronize()
    return n


def slot(fn):
    spin(fn, SETTLE_S)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s = Sampler()
    s.start()
    e0.record()
    n = spin(fn, TIMED_S)
    e1.record()
    torch.cuda.synchronize()
    s.stop = True
    s.join()
    return n, e0.elapsed_time(e1), s.p, s.c


SHAPES = [(1604, 6144, 4160, "qkv"), (1604, 4096, 4096, "o"), (1604, 28672, 4096, "gate_up"), (1604, 4096, 14336, "down"),
          (1604, 14336, 4096, "d_down"), (1604, 4096, 28672, "d_gate_up"), (1604, 4096, 6208, "d_qkv"),
          (6000, 3840, 1280, "enc_qkv"), (6000, 1280, 1280, "enc_o"), (6000, 5120, 1280, "enc_fc1"), (6000, 1280, 5120, "enc_fc2"),
          (1200, 2048, 6400, "proj1"), (308, 128256, 4096, "lm_head"), (308, 4096, 128256, "d_lm_head")]
CFGS = [("cublas", None), ("auto", 0), ("128x192", 128192), ("128x256", 128256), ("pair256", 2000256), ("pair224", 2000224)]
if len(sys.argv) > 1:
    SHAPES = [s for s in SHAPES if s[3] in sys.argv[1].split(",")]
big = torch.randn(8192, 8192, device="cuda").bfloat16()
spin(lambda: torch.matmul(big, big), 3.0)
print(f"# ROUNDS={ROUNDS} settle={SETTLE_S}s timed={TIMED_S}s per slot, configurations interleaved per shape; pre-heated 3 s", flush=True)
for M, N, K, tag in SHAPES:
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = torch.randn(N, K, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    flops = 2.0 * M * N * K
    fns = {}
    for name, bn in CFGS:
        if bn is None:
            fns[name] = lambda: torch.matmul(a, b.t(), out=out)
        else:
            fns[name] = (lambda bn=bn: ops.gemm(a, b, out=out, block_n=bn, tail_split=0))
    acc = {name: [0, 0.0, [], []] for name in fns}
    for _ in range(ROUNDS):
        for name, fn in fns.items():
            try:
                n, ms, p, c = slot(fn)
            except Exception as e:  # a tile that does not support the shape
                acc[name] = None
                continue
            if acc[name] is not None:
                acc[name][0] += n; acc[name][1] += ms; acc[name][2] += p; acc[name][3] += c
    base = None
    for name, v in acc.items():
        if v is None or v[1] == 0:
            print(f"{tag:10s} {name:8s} unsupported", flush=True)
            continue
        tf = flops * v[0] / v[1] / 1e9
        p, c = sum(v[2]) / len(v[2]), sum(v[3]) / len(v[3])
        base = tf if name == "cublas" else base
        print(f"{tag:10s} {name:8s} {tf:7.1f} TF/s  {p:6.1f} W  {c:6.0f} MHz  {tf / p:5.2f} TF/s/W  {tf / c * 1000:6.1f} TF/s/GHz  "
              f"{(tf / base if base else 0):5.3f} x cublas", flush=True)
