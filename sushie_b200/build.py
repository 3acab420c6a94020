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
CAPI = os.path.join(HERE, "csrc", "capi.cu")            # host side of the C ABI + small kernels
KERNELS = os.path.join(HERE, "csrc", "ibss_kernels.cu")  # IBSS kernel instantiations, compiled once per ancestry count
DEPS = [CAPI, KERNELS, os.path.join(HERE, "csrc", "ibss_kernel.cuh"), os.path.join(ROOT, "include", "sushie_b200.h")]
# development: SUSHIE_B200_OUT builds a differently configured engine next to the shipped one (loaded through
# SUSHIE_B200_LIB), SUSHIE_B200_DEFS adds -D switches (e.g. "-DSB_I8_ROWTEAMS=0")
OUT = os.environ.get("SUSHIE_B200_OUT", os.path.join(HERE, "_lib", "libsushie_b200.so"))
OBJ = os.path.join(HERE, "_lib", "obj" + ("" if "SUSHIE_B200_OUT" not in os.environ else "_" + os.path.basename(OUT)))
EXTRA_DEFS = os.environ.get("SUSHIE_B200_DEFS", "").split()

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-lpthread"]


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
    h.update(" ".join(NVCC_FLAGS + EXTRA_DEFS).encode())
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
    """Compile capi.cu and the six per-ancestry-count kernel units in parallel, then link the shared library."""
    if not force and up_to_date():
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    nvcc = find_nvcc()
    dev = os.environ.get("SUSHIE_B200_ONLY_K")  # development: compile one ancestry count only
    ks = [int(dev)] if dev else [1, 2, 3, 4, 5, 6]
    vflags = ["-Xptxas", "-v"] if verbose else []
    jobs = [([nvcc] + NVCC_FLAGS + EXTRA_DEFS + ([f"-DSB_ONLY_K={int(dev)}"] if dev else []) + ["-c", "-o", os.path.join(OBJ, "capi.o"), CAPI],
             os.path.join(OBJ, "capi.o"))]
    for k in ks:
        obj = os.path.join(OBJ, f"ibss_k{k}.o")
        jobs.append(([nvcc] + NVCC_FLAGS + EXTRA_DEFS + vflags + [f"-DSB_K={k}", "-c", "-o", obj, KERNELS], obj))
    procs = [(cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)) for cmd, _ in jobs]
    logs = []
    for cmd, proc in procs:
        out, _ = proc.communicate()
        logs.append(out)
        if proc.returncode != 0:
            for _, other in procs:
                if other.poll() is None:
                    other.kill()
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + out)
    link = [nvcc] + LINK_FLAGS + ["-o", OUT] + [obj for _, obj in jobs]
    proc = subprocess.run(link, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + " ".join(link) + "\n" + proc.stdout + proc.stderr)
    if verbose:
        sys.stderr.write("".join(logs))
    with open(STAMP, "w") as fh:
        fh.write(_source_hash() + "\n")
    return OUT


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
