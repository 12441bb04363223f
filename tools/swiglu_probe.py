"""Fused SwiGLU GEMM epilogues against GEMM + element-wise kernel (decoder shapes), sustained timing."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slam_llm_b200 import ops


def timeit(fn, iters=30):
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


M, D, F = 1604, 4096, 14336
x = torch.randn(M, D, device="cuda").bfloat16()
wgu = (torch.randn(2 * F, D, device="cuda") * 0.02).bfloat16()
wdT = (torch.randn(F, D, device="cuda") * 0.02).bfloat16()
dy = torch.randn(M, D, device="cuda").bfloat16()
gu = torch.empty(M, 2 * F, device="cuda", dtype=torch.bfloat16)
h = torch.empty(M, F, device="cuda", dtype=torch.bfloat16)
dh = torch.empty(M, F, device="cuda", dtype=torch.bfloat16)
dgu = torch.empty(M, 2 * F, device="cuda", dtype=torch.bfloat16)
for tile in (0, 128256, 2000256):
    t_plain = timeit(lambda: ops.gemm(x, wgu, out=gu, block_n=tile))
    t_sw = timeit(lambda: ops.swiglu_fwd(gu, out=h, block=64))
    t_fused = timeit(lambda: ops.gemm(x, wgu, out=gu, act=3, aux=h, block_n=tile))
    print(f"fwd tile {tile}: gemm {t_plain:.1f} + swiglu {t_sw:.1f} = {t_plain + t_sw:.1f} us ; fused {t_fused:.1f} us", flush=True)
for tile in (0, 128256, 128192, 2000256, 2000224):
    t_plain = timeit(lambda: ops.gemm(dy, wdT, out=dh, block_n=tile))
    t_sw = timeit(lambda: ops.swiglu_bwd(gu, dh, out=dgu, block=64))
    t_fused = timeit(lambda: ops.gemm(dy, wdT, out=dgu, act=4, aux=gu, block_n=tile))
    print(f"bwd tile {tile}: gemm {t_plain:.1f} + swiglu {t_sw:.1f} = {t_plain + t_sw:.1f} us ; fused {t_fused:.1f} us", flush=True)
