"""Build recipe for libvpc.so (hand-written sm_100a CUDA, C ABI in include/vpc.h).

The library is built IN-TREE (vtkcloudpoint_b200/libvpc.so) so that it travels to the
GPU box with the repo snapshot.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "libvpc.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",            # reference arithmetic is IEEE binary64 without FMA contraction
    "-Xcompiler", "-fPIC", "-shared",
]


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + [ROOT / "include" / "vpc.h"]


def _fingerprint() -> str:
    """Hash of every source and of the flags.  File times do not survive a snapshot copy (the GPU box would rebuild a fresh
    library, once per rank and at the same time), contents do."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS + os.environ.get("VPC_NVCC_EXTRA", "").split()).encode())
    for src in _sources() + sorted((CSRC / "host").glob("*.hpp")):
        h.update(src.name.encode()); h.update(src.read_bytes())
    return h.hexdigest()


STAMP = PKG / "libvpc.so.stamp"


def stale() -> bool:
    if not LIB.exists() or not STAMP.exists():
        return True
    return STAMP.read_text().strip() != _fingerprint()


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libvpc.so cannot be built (there is no CPU fallback)")


def build_lib(force: bool = False, verbose: bool = False) -> Path:
    if not force and not stale():
        return LIB
    extra = os.environ.get("VPC_NVCC_EXTRA", "").split()      # developer knob for A/B builds (e.g. -DVPC_ICP_ITER_BLOCK=128)
    tmp = f"{LIB}.{os.getpid()}.tmp"                            # several ranks may build at once: no shared temporary
    cmd = [nvcc_path(), *NVCC_FLAGS, *extra, "-o", tmp, str(CSRC / "vpc_api.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    os.replace(tmp, LIB)
    STAMP.write_text(_fingerprint())
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build_lib(force=True, verbose=True))
