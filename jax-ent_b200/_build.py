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
SOURCES = ["bv_pass.cu", "bv_tail.cu", "api.cu", "bg_gemm.cu", "bg_step.cu", "bv_contacts.cu"]
HEADERS = [os.path.join(CSRC, "jx_internal.cuh"), os.path.join(CSRC, "bv_tail_body.cuh"), os.path.join(CSRC, "bg_mma.cuh"), os.path.join(ROOT, "include", "jaxent_b200.h")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libjaxent_b200.so cannot be built (there is no CPU fallback)")


HASH = LIB + ".srchash"


def _flags() -> list:
    flags = ["-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17"]
    extra = os.environ.get("JAXENT_NVCC_FLAGS")
    return (extra.split() if extra else []) + flags


def source_hash() -> str:
    """content hash of every source / header and of the compile flags: decides whether the in-tree .so is current
    (file times are not preserved when the tree is copied to another box)"""
    import hashlib

    h = hashlib.sha256()
    for path in [os.path.join(CSRC, s) for s in SOURCES] + HEADERS:
        with open(path, "rb") as f:
            h.update(f.read())
    h.update(" ".join(_flags()).encode())
    return h.hexdigest()


def needs_build() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(HASH):
        return True
    try:
        with open(HASH) as f:
            return f.read().strip() != source_hash()
    except OSError:
        return True


def build(force: bool = False, verbose: bool = False) -> str:
    """Compiles into a temporary file and renames it into place under an exclusive lock, so concurrent processes
    (one per GPU under torchrun) never load a half-written library."""
    import fcntl

    if not force and not needs_build():
        return LIB
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():  # another process built it while we waited
                return LIB
            tmp = f"{LIB}.tmp.{os.getpid()}"
            cmd = [nvcc_path()] + _flags() + ["-I", os.path.join(ROOT, "include"), "-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), file=sys.stderr)
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError(f"nvcc failed ({res.returncode}):\n{res.stdout}\n{res.stderr}")
            if verbose:
                print(res.stderr, file=sys.stderr)
            os.replace(tmp, LIB)
            with open(HASH, "w") as f:
                f.write(source_hash())
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


FFI_TEST_LIB = os.path.join(ROOT, "tests", "ffi_stub", "libjaxent_ffi_test.so")


def build_ffi_test(force: bool = False) -> str:
    """csrc/xla_ffi.cc (the jax.ffi handlers) compiled against the stand-in header of tests/ffi_stub, plus the test driver,
    linked against the in-tree library.  Test infrastructure: type-checks the handlers and lets the tests call them."""
    stub = os.path.join(ROOT, "tests", "ffi_stub")
    srcs = [os.path.join(CSRC, "xla_ffi.cc"), os.path.join(stub, "driver.cc")]
    deps = srcs + [os.path.join(stub, "xla", "ffi", "api", "ffi.h"), os.path.join(ROOT, "include", "jaxent_b200.h")]
    if not force and os.path.exists(FFI_TEST_LIB) and all(os.path.getmtime(FFI_TEST_LIB) >= os.path.getmtime(d) for d in deps):
        return FFI_TEST_LIB
    build()
    tmp = f"{FFI_TEST_LIB}.tmp.{os.getpid()}"
    cmd = [nvcc_path(), "-shared", "-Xcompiler", "-fPIC", "-std=c++17", "-O2", "-x", "cu", "-I", os.path.join(ROOT, "include"), "-I", stub,
           "-o", tmp] + srcs + ["-L", HERE, "-ljaxent_b200", "-Xlinker", "-rpath=$ORIGIN/../../jax-ent_b200"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on the FFI shim ({res.returncode}):\n{res.stdout}\n{res.stderr}")
    os.replace(tmp, FFI_TEST_LIB)
    return FFI_TEST_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
