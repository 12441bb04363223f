"""LoRA thin products (T = x A_cat^T: M=1604, N=64, K=4096 and U = dY (sB): K=6144/1024...): time per launch for split_k variants (CUDA events, 200 reps)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slam_llm_b200 import ops

def t(fn, reps=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

for M, N, K, tag in [(1604, 64, 4096, "T=x*A^T (qkv fwd)"), (1604, 64, 6144, "U=dqkv*(sB) (qkv bwd)"), (1604, 64, 14336, "K=14336")]:
    a = torch.randn(M, K, device="cuda").bfloat16(); b = torch.randn(N, K, device="cuda").bfloat16()
    for sk in (1, 2, 4, 8):
        if sk == 1:
            out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
            us = t(lambda: ops.gemm(a, b, out=out, tail_split=-1))
            us_ts = t(lambda: ops.gemm(a, b, out=out))
            print(f"{tag:26s} split_k=1 bf16 out: {us:6.1f} us  (auto tail split: {us_ts:6.1f} us)")
        else:
            acc = torch.zeros(M, N, device="cuda", dtype=torch.float32)
            o2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
            us = t(lambda: ops.cast_bf16(ops.gemm(a, b, out=acc, out_f32=True, split_k=sk), out=o2))
            print(f"{tag:26s} split_k={sk} f32 atomics + cast: {us:6.1f} us")
