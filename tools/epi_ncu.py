"""Two epilogue-bound encoder GEMMs (K = 1280) for an `ncu --set full --import-source on` capture: where do the epilogue warps stall?
    ncu --set full --import-source on --clock-control none --profile-from-start off -o gpurun_out/epi python tools/epi_ncu.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slam_llm_b200 import ops

M, d = 6000, 1280
x = torch.randn(M, d, device="cuda").bfloat16()
res = torch.randn(M, d, device="cuda").bfloat16()
wo = (torch.randn(d, d, device="cuda") * 0.02).bfloat16()
w1 = (torch.randn(4 * d, d, device="cuda") * 0.02).bfloat16()
bo = torch.randn(d, device="cuda")
b1 = torch.randn(4 * d, device="cuda")
fns = [lambda: ops.gemm(x, wo, bias=bo, residual=res, static_w=True), lambda: ops.gemm(x, w1, bias=b1, act=1, static_w=True)]
for fn in fns:
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
