"""What does cuBLAS launch for the decoder GEMM shapes?  One profiled call per (shape, implementation) between cudaProfilerStart/Stop:

    ncu --set full --clock-control none --profile-from-start off -o gpurun_out/cublas_peek python tools/cublas_peek.py [shape,...]

The kernel NAMES (tile / cluster / stage counts are encoded in cuBLAS's names), launch geometry, L2 -> SM sectors and instruction counts of both
implementations are then read from the report (tools/cublas_peek_read.py)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slam_llm_b200 import ops

SHAPES = [(1604, 6144, 4160, "qkv"), (1604, 4096, 4096, "o"), (1604, 28672, 4096, "gate_up"), (1604, 4096, 14336, "down"),
          (1604, 14336, 4096, "d_down"), (1604, 4096, 28672, "d_gate_up"), (1604, 4096, 6208, "d_qkv"),
          (6000, 3840, 1280, "enc_qkv"), (6000, 5120, 1280, "enc_fc1"), (6000, 1280, 5120, "enc_fc2")]
if len(sys.argv) > 1:
    SHAPES = [s for s in SHAPES if s[3] in sys.argv[1].split(",")]
from slam_llm_b200.engine import _swap_ab

for M, N, K, tag in SHAPES:
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = torch.randn(N, K, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    swap = tag in ("qkv", "o", "down", "d_gate_up", "d_qkv") and _swap_ab(M, N)      # the fused-SwiGLU GEMMs (gate_up, d_down) are never swapped
    fns = [lambda: torch.matmul(a, b.t(), out=out), (lambda: ops.gemm(b, a, out=out, transpose_out=True)) if swap else (lambda: ops.gemm(a, b, out=out))]
    for fn in fns:
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        fn()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    print(tag, "swap" if swap else "plain", flush=True)
