"""C-ABI checks that need no GPU: the shared library loads, exports every symbol include/slam_b200.h declares,
and the ctypes struct mirrors agree with the C compiler's layout of the header structs."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from slam_llm_b200 import build, lib as L
    build.build()
    return L


def test_library_exports_every_header_symbol(lib):
    handle = lib.load()
    syms = lib.header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(handle, s), f"{s} declared in include/slam_b200.h but not exported"
    assert set(syms) == set(lib._SIGS), (set(syms) ^ set(lib._SIGS))
    assert handle.slam_abi_version() == 6
    assert handle.slam_launch_count() == 0          # nothing launched: no compute without a GPU


def test_struct_layout_matches_c_compiler(lib, tmp_path):
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "slam_b200.h"\nint main(){\n'
                   'printf("%zu %zu %zu %zu %zu\\n", sizeof(slam_gemm_args), offsetof(slam_gemm_args, out), offsetof(slam_gemm_args, alpha), '
                   'offsetof(slam_gemm_args, split_k), offsetof(slam_gemm_args, k2));\n'
                   'printf("%zu %zu %zu %zu\\n", sizeof(slam_attn_args), offsetof(slam_attn_args, scale), offsetof(slam_attn_args, dout), '
                   'offsetof(slam_attn_args, dkv_part));\nreturn 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).split()
    g, a = lib.GemmArgs, lib.AttnArgs
    assert [int(x) for x in out[:5]] == [ctypes.sizeof(g), g.out.offset, g.alpha.offset, g.split_k.offset, g.k2.offset]
    assert [int(x) for x in out[5:]] == [ctypes.sizeof(a), a.scale.offset, a.dout.offset, a.dkv_part.offset]


def test_header_is_plain_c_and_cites_reference(lib):
    text = open(os.path.join(ROOT, "include", "slam_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)                  # strip comments: signatures only
    assert "torch" not in code.lower() and "at::" not in code and "Tensor" not in code   # no torch types in the signatures
    for cite in ("models/slam_model.py", "models/encoder.py", "models/projector.py", "datasets/speech_dataset.py", "utils/metric.py",
                 "pipeline/finetune.py"):
        assert cite in text, f"header should cite the reference site {cite}"
    subprocess.check_call(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "slam_b200.h")])


def test_product_path_fails_loudly_without_gpu_or_library(lib, monkeypatch):
    import torch
    from slam_llm_b200 import ops
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA"):
            ops.cast_bf16(torch.zeros(8))
        from slam_llm_b200.engine import SlamStepB200
        from slam_llm_b200 import config as C
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            SlamStepB200(C.WHISPER["tiny"], C.LLM["tinyllama-1.1b"], None, C.ProjCfg(), device="cuda:0")
    monkeypatch.setenv("SLAM_B200_LIB", "/nonexistent/libslam_b200.so")
    monkeypatch.setattr(lib, "_lib", None)
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        lib.load()


def test_product_never_imports_oracle():
    bad = []
    for base in ("slam_llm_b200", os.path.join("src", "slam_llm")):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith(".py"):
                    t = open(os.path.join(dp, f)).read()
                    if re.search(r"^\s*(from|import)\s+oracle\b", t, flags=re.M):
                        bad.append(os.path.join(dp, f))
    assert not bad, f"product code imports the oracle: {bad}"
