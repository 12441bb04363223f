"""Side-by-side launches of cuBLAS (torch.matmul) and slam_gemm_bf16 on a few step shapes, for an ncu capture."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slam_llm_b200 import ops

for (M, N, K) in [(1604, 4096, 4096), (1604, 6144, 4096), (1604, 4096, 14336), (6000, 5120, 1280), (1604, 28672, 4096)]:
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = torch.randn(N, K, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        torch.matmul(a, b.t(), out=out)
        ops.gemm(a, b, out=out)
    torch.cuda.synchronize()
