"""In-tree build of libzzflate_b200.so (kernels + C-ABI + host driver) for sm_100a.

nvcc cross-compiles without a GPU; the resulting .so stays next to this file (git-ignored) so that it
travels with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB = PKG_DIR / "libzzflate_b200.so"
SOURCES = ["zz_kernels.cu", "zz_cabi.cu", "zz_host.cpp"]
DEPS = SOURCES + ["zz_kernels.cuh", "../../include/zzgpu.h", "../../include/zzflate.h",
                  "../../include/encoder.h", "../../include/crc.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "static"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any((CSRC / d).resolve().stat().st_mtime > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    cmd = [nvcc_path(), *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []),
           *[str(CSRC / s) for s in SOURCES], "-o", str(LIB)]
    subprocess.check_call(cmd, cwd=str(CSRC))
    return LIB


SYNTH_LIB = PKG_DIR / "libzzsynth.so"


def build_synth(force: bool = False) -> Path:
    src = CSRC / "zz_synth.c"
    if force or not SYNTH_LIB.exists() or src.stat().st_mtime > SYNTH_LIB.stat().st_mtime:
        cc = shutil.which("gcc") or "cc"
        subprocess.check_call([cc, "-O2", "-std=c11", "-fPIC", "-shared", "-o", str(SYNTH_LIB), str(src), "-lpthread"])
    return SYNTH_LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_synth(force="--force" in sys.argv))
