"""In-tree build of libflashjoin_b200.so (nvcc, sm_100a) and the pybind11 module flash_join.

Used by setup.py (``python setup.py build_ext --inplace``, the reference's build command,
/root/reference/README.md:107) and by __graft_entry__.build().  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
ROOT = PKG.parent
OBJ = PKG / "_obj"
LIB = PKG / "libflashjoin_b200.so"
CU_SOURCES = ["fj_scalar.cu", "fj_radix.cu", "fj_part.cu", "fj_engine.cu", "fj_dist.cu"]
HEADERS = ["fj_common.cuh", "fj_kernels.h", "fj_dist.h", "../../include/flashjoin_b200.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.sep not in c or os.path.exists(c)):
            return c
    raise RuntimeError("nvcc not found")


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def ext_path() -> Path:
    return PKG / f"flash_join{sysconfig.get_config_var('EXT_SUFFIX')}"


def build(force: bool = False, verbose: bool = False) -> None:
    OBJ.mkdir(exist_ok=True)
    hdrs = [CSRC / h for h in HEADERS]
    nvcc = _nvcc()

    def compile_one(src: str) -> Path:
        o = OBJ / (src + ".o")
        if force or _stale(o, [CSRC / src, *hdrs]):
            cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(o)]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.check_call(cmd)
        return o

    with ThreadPoolExecutor(max_workers=4) as ex:
        objs = list(ex.map(compile_one, CU_SOURCES))
    if force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *map(str, objs), "-ldl", "-lpthread"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    ext = ext_path()
    src = CSRC / "flash_join_py.cpp"
    if force or _stale(ext, [src, LIB, ROOT / "include" / "flashjoin_b200.h"]):
        import pybind11

        inc = sysconfig.get_paths()["include"]
        cmd = [
            "g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden",
            f"-I{pybind11.get_include()}", f"-I{inc}", str(src), "-o", str(ext),
            f"-L{PKG}", "-lflashjoin_b200", "-Wl,-rpath,$ORIGIN",
        ]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built", LIB, ext_path())
