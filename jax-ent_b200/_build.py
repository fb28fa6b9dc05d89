"""In-tree nvcc build of libjaxent_b200.so for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libjaxent_b200.so")
SOURCES = ["bv_pass.cu", "bv_tail.cu", "api.cu", "bg_gemm.cu", "bg_step.cu"]
HEADERS = [os.path.join(CSRC, "jx_internal.cuh"), os.path.join(CSRC, "bv_tail_body.cuh"), os.path.join(ROOT, "include", "jaxent_b200.h")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libjaxent_b200.so cannot be built (there is no CPU fallback)")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [
        nvcc_path(),
        "-shared",
        "-Xcompiler",
        "-fPIC",
        "-gencode",
        "arch=compute_100a,code=sm_100a",
        "-lineinfo",
        "-O3",
        "-std=c++17",
        "-I",
        os.path.join(ROOT, "include"),
        "-o",
        LIB,
    ] + [os.path.join(CSRC, s) for s in SOURCES]
    extra = os.environ.get("JAXENT_NVCC_FLAGS")
    if extra:
        cmd[1:1] = extra.split()
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed ({res.returncode}):\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr, file=sys.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
