"""Build libslam_b200.so (all sm_100a kernels + the C ABI) in-tree with nvcc.

Usage: python -m slam_llm_b200.build [--force] [--no-watchdog]
The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = PKG / "csrc" / "_obj"
LIB = PKG / "libslam_b200.so"
INCLUDE = PKG.parent / "include"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-I", str(INCLUDE),
]


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _digest(src: Path, flags) -> str:
    h = hashlib.sha256()
    h.update(" ".join(flags).encode())
    h.update(src.read_bytes())
    for hdr in sorted(CSRC.glob("*.cuh")) + sorted(INCLUDE.glob("*.h")):
        h.update(hdr.read_bytes())
    return h.hexdigest()


def build(force: bool = False, watchdog: bool = True, verbose: bool = True) -> Path:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = list(NVCC_FLAGS) + [f"-DSLAM_WATCHDOG={1 if watchdog else 0}"] + os.environ.get("SLAM_NVCC_EXTRA", "").split()
    lib = Path(os.environ["SLAM_B200_LIB_OUT"]) if os.environ.get("SLAM_B200_LIB_OUT") else LIB     # experiment builds: another output file
    OBJ.mkdir(exist_ok=True)
    objs, jobs = [], []
    for src in _sources():
        obj = OBJ / (src.stem + ".o")
        stamp = OBJ / (src.stem + ".sha")
        dig = _digest(src, flags)
        objs.append(obj)
        if force or not obj.exists() or not stamp.exists() or stamp.read_text() != dig:
            jobs.append((src, obj, stamp, dig))

    def compile_one(job):
        src, obj, stamp, dig = job
        cmd = [nvcc] + flags + ["-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        stamp.write_text(dig)
        return src.name

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for name in ex.map(compile_one, jobs):
                if verbose:
                    print(f"[slam_b200.build] compiled {name}", file=sys.stderr)
    if jobs or not lib.exists():
        cmd = [nvcc, "-shared", "-o", str(lib)] + [str(o) for o in objs] + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"[slam_b200.build] linked {lib}", file=sys.stderr)
    return lib


if __name__ == "__main__":
    build(force="--force" in sys.argv, watchdog="--no-watchdog" not in sys.argv)
