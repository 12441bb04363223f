"""Correctness of the tail-split schedule and of the CTA-pair tiles against the plain single-CTA kernel (run on a B200)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slam_llm_b200 import ops

torch.manual_seed(0)
ok = True
SHAPES = [(256, 256, 64, 0), (256, 512, 256, 0), (300, 776, 1280, 0), (1604, 6144, 4096, 64), (1604, 4096, 4096, 0), (1604, 4096, 14336, 0),
          (6000, 1280, 5120, 0), (6000, 3840, 1280, 0), (77, 264, 200, 64), (308, 4096, 32064, 0), (1604, 28672, 4096, 0), (4000, 9000, 512, 0)]
for (M, N, K, K2) in SHAPES:
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = torch.randn(N, K, device="cuda").bfloat16()
    a2 = torch.randn(M, K2, device="cuda").bfloat16() if K2 else None
    b2 = torch.randn(N, K2, device="cuda").bfloat16() if K2 else None
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda").bfloat16()
    ref = ops.gemm(a, b, a2=a2, b2=b2, bias=bias, residual=res, act=1, block_n=128256, tail_split=-1)
    ref32 = torch.nn.functional.gelu(a.float() @ b.float().t() + (a2.float() @ b2.float().t() if K2 else 0) + bias) + res.float()
    base = ((ref.float() - ref32).norm() / ref32.norm()).item()
    for bn in (128256, 128192, 128128, 0):
        outs = [ops.gemm(a, b, a2=a2, b2=b2, bias=bias, residual=res, act=1, block_n=bn, tail_split=0) for _ in range(3)]
        torch.cuda.synchronize()
        rel = ((outs[0].float() - ref32).norm() / ref32.norm()).item()
        det = torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])
        print(f"M={M} N={N} K={K}+{K2} tile={bn} tail-split: rel_l2 vs fp32 {rel:.3e} (plain {base:.3e}) repeatable={det}", flush=True)
        ok = ok and rel < 1.5 * base + 1e-4 and det
    for bn in (2000256, 2000192):
        out = ops.gemm(a, b, a2=a2, b2=b2, bias=bias, residual=res, act=1, block_n=bn, tail_split=-1)
        same = torch.equal(out, ref)
        ok = ok and same
        if not same:
            print(f"  pair tile {bn}: NOT bit-identical", flush=True)
    out32 = ops.gemm(a, b, out_f32=True, tail_split=0)
    r32 = a.float() @ b.float().t()
    rel = ((out32 - r32).norm() / r32.norm()).item()
    ok = ok and rel < 1e-3
    print(f"  f32-out tail-split rel_l2={rel:.3e}", flush=True)
print("GEMM_CHECK", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
