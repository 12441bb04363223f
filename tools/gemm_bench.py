"""Micro-benchmark of the tcgen05 GEMM against cuBLAS (torch.matmul) on the step's shapes."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slam_llm_b200 import ops

SHAPES = [  # (M, N, K, K2, tag)
    (1604, 6144, 4096, 64, "qkv+lora"), (1604, 4096, 4096, 0, "o"), (1604, 28672, 4096, 0, "gate_up"),
    (1604, 4096, 14336, 0, "down"), (1604, 14336, 4096, 0, "d_down"), (1604, 4096, 28672, 0, "d_gate_up"),
    (6000, 3840, 1280, 0, "enc_qkv"), (6000, 5120, 1280, 0, "enc_fc1"), (6000, 1280, 5120, 0, "enc_fc2"),
    (308, 128256, 4096, 0, "lm_head"), (308, 4096, 128256, 0, "d_lm_head"), (1604, 64, 4096, 0, "lora_T"),
]

def bench(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]

res = []
for M, N, K, K2, tag in SHAPES:
    a = torch.randn(M, K, device="cuda").bfloat16(); b = torch.randn(N, K, device="cuda").bfloat16()
    a2 = torch.randn(M, K2, device="cuda").bfloat16() if K2 else None
    b2 = torch.randn(N, K2, device="cuda").bfloat16() if K2 else None
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    row = {"tag": tag, "M": M, "N": N, "K": K, "K2": K2}
    flops = 2.0 * M * N * (K + K2)
    for bn in (0, 128256, 128192, 256256, 256224, 2000256) if N >= 256 else (0,):
        t = bench(lambda: ops.gemm(a, b, a2=a2, b2=b2, out=out, block_n=bn, tail_split=-1))
        row[f"slam_bn{bn}_ms"] = round(t, 4); row[f"slam_bn{bn}_tflops"] = round(flops / t / 1e9, 1)
    for bn, sk in ((0, 0), (128256, 0), (128256, 2), (128256, 8), (128192, 0)) if N >= 256 else ((0, 0),):
        t = bench(lambda: ops.gemm(a, b, a2=a2, b2=b2, out=out, block_n=bn, tail_split=sk))
        row[f"slam_sk{sk}_bn{bn}_ms"] = round(t, 4); row[f"slam_sk{sk}_bn{bn}_tflops"] = round(flops / t / 1e9, 1)
    t = bench(lambda: torch.matmul(a, b.t(), out=out))
    row["cublas_ms"] = round(t, 4); row["cublas_tflops"] = round(2.0 * M * N * K / t / 1e9, 1)
    res.append(row)
    print(json.dumps(row), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/gemm_bench.json", "w"), indent=1)
