"""Sustained (power-capped) throughput of GEMM configurations: each config runs back to back for ~1.5 s while NVML samples
power and SM clock.  Under the B200 power cap the step's GEMM rate is set by energy per flop, not by the burst rate."""
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


def sustained(fn, flops, secs=1.5):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s = Sampler()
    s.start()
    n = 0
    t0 = time.time()
    e0.record()
    while time.time() - t0 < secs:
        for _ in range(50):
            fn()
        n += 50
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    s.stop = True
    s.join()
    ms = e0.elapsed_time(e1)
    half = len(s.p) // 2
    p = sum(s.p[half:]) / max(1, len(s.p) - half)
    c = sum(s.c[half:]) / max(1, len(s.c) - half)
    tf = flops * n / ms / 1e9
    return tf, p, c


SHAPES = [(1604, 6144, 4160, "qkv"), (1604, 4096, 4096, "o"), (1604, 28672, 4096, "gate_up"), (1604, 4096, 14336, "down"),
          (1604, 14336, 4096, "d_down"), (1604, 4096, 28672, "d_gate_up"), (1604, 4096, 6208, "d_qkv"),
          (6000, 3840, 1280, "enc_qkv"), (6000, 1280, 1280, "enc_o"), (6000, 5120, 1280, "enc_fc1"), (6000, 1280, 5120, "enc_fc2"),
          (1200, 2048, 6400, "proj1"), (308, 128256, 4096, "lm_head"), (308, 4096, 128256, "d_lm_head")]
CFGS = [("cublas", None), ("auto", 0), ("128x256", 128256), ("128x192", 128192), ("128x128", 128128), ("256x224", 256224),
        ("pair256", 2000256), ("pair224", 2000224), ("pair192", 2000192), ("pair160", 2000160), ("pair128", 2000128),
        ("128x256+ts", -128256), ("128x192+ts", -128192)]
if len(sys.argv) > 1:
    SHAPES = [s for s in SHAPES if s[3] in sys.argv[1].split(",")]
for M, N, K, tag in SHAPES:
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = torch.randn(N, K, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    flops = 2.0 * M * N * K
    for name, bn in CFGS:
        if bn is None:
            fn = lambda: torch.matmul(a, b.t(), out=out)
        elif bn < 0:
            fn = lambda: ops.gemm(a, b, out=out, block_n=-bn, tail_split=0)
        else:
            fn = lambda: ops.gemm(a, b, out=out, block_n=bn, tail_split=-1 if bn else 0)
        tf, p, c = sustained(fn, flops, 0.7)
        print(f"{tag:8s} {name:8s} {tf:7.1f} TF/s  {p:6.1f} W  {c:6.0f} MHz  {tf / p:5.2f} TF/s/W  {tf / c * 1000:6.1f} TF/s/GHz", flush=True)
    time.sleep(0.5)
