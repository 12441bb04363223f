"""Epilogue cost per tile: GEMMs with K = 64 (one k-block: the mainloop is negligible) over many tiles."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slam_llm_b200 import ops


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1000.0


M, N, K = 6000, 5120, 64
a = torch.randn(M, K, device="cuda").bfloat16()
b = torch.randn(N, K, device="cuda").bfloat16()
bias = torch.randn(N, device="cuda")
res = torch.randn(M, N, device="cuda").bfloat16()
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for tile in [int(t) for t in os.environ.get("PROBE_TILES", "128256,128192,2000256,2000224").split(",")]:
    bn = tile % 1000
    rows = 256 if tile >= 2000000 else 128
    units = 74 if tile >= 2000000 else 148
    tiles = -(-M // rows) * -(-N // bn)
    waves = -(-tiles // units)
    for name, kw in (("plain", {}), ("bias+gelu", {"bias": bias, "act": 1}), ("residual", {"residual": res}), ("f32", {"out_f32": True})):
        o = torch.empty(M, N, device="cuda", dtype=torch.float32) if name == "f32" else out
        t = timeit(lambda: ops.gemm(a, b, out=o, block_n=tile, tail_split=-1, **kw))
        print(f"tile {tile:8d} {name:10s} {t:8.1f} us total, {t / waves:6.2f} us per tile-wave ({waves} waves)", flush=True)
