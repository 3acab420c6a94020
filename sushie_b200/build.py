"""Build ``sushie_b200/_lib/libsushie_b200.so`` in-tree with nvcc for sm_100a.

    python -m sushie_b200.build [--force]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = [os.path.join(HERE, "csrc", "capi.cu")]
DEPS = SRC + [os.path.join(HERE, "csrc", "ibss_kernel.cuh"), os.path.join(ROOT, "include", "sushie_b200.h")]
OUT = os.path.join(HERE, "_lib", "libsushie_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-cudart", "static",
]


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build the sushie_b200 CUDA engine")
    return nvcc


STAMP = OUT + ".srchash"


def _source_hash() -> str:
    """Hash of everything the library is built from (sources, header, flags, development switches)."""
    import hashlib

    h = hashlib.sha256()
    for path in DEPS:
        with open(path, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(os.environ.get("SUSHIE_B200_ONLY_K", "").encode())
    return h.hexdigest()


def up_to_date() -> bool:
    """True when the library was built from exactly these sources.  Content hash, not mtimes: a snapshot
    copied to another machine (the GPU box) must not trigger a multi-minute rebuild."""
    if not os.path.exists(OUT):
        return False
    try:
        with open(STAMP) as fh:
            return fh.read().strip() == _source_hash()
    except OSError:
        t = os.path.getmtime(OUT)
        return all(os.path.getmtime(d) <= t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    dev = os.environ.get("SUSHIE_B200_ONLY_K")  # development: compile one ancestry count only
    extra = [f"-DSB_ONLY_K={int(dev)}"] if dev else []
    cmd = [find_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SRC
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if verbose:
        sys.stderr.write(proc.stderr)
    with open(STAMP, "w") as fh:
        fh.write(_source_hash() + "\n")
    return OUT


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
