"""Experiment: is the GEMM mainloop latency-bound (throughput ~ stages in flight) or bandwidth-bound?"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slam_llm_b200 import ops

def bench(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts)//2]

for (M, N, K) in ((1600, 28672, 4096), (1600, 4096, 14336), (6000, 5120, 1280)):
    a = torch.randn(M, K, device="cuda").bfloat16(); b = torch.randn(N, K, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    row = {"M": M, "N": N, "K": K, "lib": os.environ.get("SLAM_B200_LIB", "default")}
    for t in (128256, 128192, 256256):
        ms = bench(lambda: ops.gemm(a, b, out=out, block_n=t))
        row[str(t)] = round(2.0 * M * N * K / ms / 1e9, 1)
    print(json.dumps(row), flush=True)
