"""Swap-AB CTA-pair tiles (weight = 256-row M operand, tokens = N, transposed epilogue) vs the current `auto` tile and cuBLAS on the decoder
shapes with M = 1604 tokens.  Interleaved, pre-heated, steady-state (same protocol as tools/gemm_power.py)."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pynvml, torch
from slam_llm_b200 import ops
pynvml.nvmlInit(); H = pynvml.nvmlDeviceGetHandleByIndex(0)

def spin(fn, secs):
    n, t0 = 0, time.time()
    while time.time() - t0 < secs:
        for _ in range(50): fn()
        n += 50; torch.cuda.synchronize()
    return n

def slot(fn):
    spin(fn, 0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); n = spin(fn, 0.7); e1.record(); torch.cuda.synchronize()
    return n, e0.elapsed_time(e1)

SHAPES = [(1604, 6144, 4160, "qkv"), (1604, 4096, 4096, "o"), (1604, 4096, 14336, "down"), (1604, 4096, 28672, "d_gate_up"), (1604, 4096, 6208, "d_qkv")]
if len(sys.argv) > 1: SHAPES = [s for s in SHAPES if s[3] in sys.argv[1].split(",")]
big = torch.randn(8192, 8192, device="cuda").bfloat16(); spin(lambda: torch.matmul(big, big), 3.0)
for M, N, K, tag in SHAPES:
    x = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    fns = {"cublas": lambda: torch.matmul(x, w.t(), out=out), "auto": lambda: ops.gemm(x, w, out=out)}
    for bn in (2000160, 2000192, 2000224, 2000256, 0):
        fns[f"swap{bn % 1000 if bn else 'auto'}"] = (lambda bn=bn: ops.gemm(w, x, out=out, transpose_out=True, block_n=bn))
    acc = {k: [0, 0.0] for k in fns}
    for _ in range(2):
        for k, fn in fns.items():
            n, ms = slot(fn); acc[k][0] += n; acc[k][1] += ms
    base = None
    for k, (n, ms) in acc.items():
        tf = 2.0 * M * N * K * n / ms / 1e9
        base = tf if k == "cublas" else base
        print(f"{tag:10s} {k:9s} {tf:7.1f} TF/s  {tf / base:5.3f} x cublas", flush=True)
