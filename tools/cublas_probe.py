"""Side-by-side launches of cuBLAS (torch.matmul) and slam_gemm_bf16 on a few step shapes, for an ncu capture."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slam_llm_b200 import ops

TILES = [int(t) for t in os.environ.get("PROBE_TILES", "0").split(",")]
for (M, N, K) in [(1604, 4096, 4096), (1604, 6144, 4096), (1604, 4096, 14336), (6000, 5120, 1280), (1604, 28672, 4096)]:
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = torch.randn(N, K, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    if os.environ.get("PROBE_CUBLAS", "1") == "1":
        torch.matmul(a, b.t(), out=out)
    for t in TILES:
        ops.gemm(a, b, out=out, block_n=t, tail_split=-1)
    torch.cuda.synchronize()
