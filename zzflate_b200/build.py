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
SOURCES = ["zz_kernels.cu", "zz_cabi.cu", "zz_host.cpp", "zz_decoder.cpp"]
DEPS = SOURCES + ["zz_kernels.cuh", "../../include/zzgpu.h", "../../include/zzflate.h",
                  "../../include/encoder.h", "../../include/crc.h", "../../include/outputbitstream.h", "../../include/decoder.h"]
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


BOUNDARY_SRC = PKG_DIR.parent / "tests" / "cpp" / "boundary_test.cpp"
BOUNDARY_BIN = PKG_DIR.parent / "tests" / "cpp" / "_build" / "boundary_test"


def build_boundary_test(force: bool = False) -> Path:
    """tests/cpp/boundary_test.cpp: a C++14 caller compiled against include/*.h and linked to the library (the
    reference's own tests re-hosted); run by the GPU tests."""
    build()
    deps = [BOUNDARY_SRC, LIB] + [PKG_DIR.parent / "include" / h for h in ("zzflate.h", "encoder.h", "crc.h", "huffman.h", "outputbitstream.h")]
    if force or not BOUNDARY_BIN.exists() or any(d.stat().st_mtime > BOUNDARY_BIN.stat().st_mtime for d in deps):
        BOUNDARY_BIN.parent.mkdir(parents=True, exist_ok=True)
        cxx = shutil.which("g++") or "c++"
        subprocess.check_call([cxx, "-std=c++14", "-O1", "-Wall", "-o", str(BOUNDARY_BIN), str(BOUNDARY_SRC),
                               "-L" + str(PKG_DIR), "-lzzflate_b200", "-lz", "-Wl,-rpath," + str(PKG_DIR)])
    return BOUNDARY_BIN


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_synth(force="--force" in sys.argv))
    print(build_boundary_test(force="--force" in sys.argv))
