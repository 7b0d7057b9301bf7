"""In-tree build of the product libraries (nvcc cross-compiles sm_100a without a GPU).

    nano_b200/lib/libnano_b200.so        engine + C-ABI (include/nano_b200.h)     <- csrc/engine.cu, kernels.cuh
    nano_b200/lib/libnano_infer_b200.so  reference-API shim (infer.h symbols)     <- csrc/infer_b200.c  (if present)

The .so files are git-ignored but travel to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "nano_b200", "csrc")
LIB = os.path.join(ROOT, "nano_b200", "lib")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

ENGINE_SO = os.path.join(LIB, "libnano_b200.so")
SHIM_SO = os.path.join(LIB, "libnano_infer_b200.so")


def _newer(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources if os.path.exists(s))


def build_engine(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB, exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    srcs.append(os.path.join(ROOT, "include", "nano_b200.h"))
    if force or _newer(ENGINE_SO, srcs):
        cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
               "-Xcompiler", "-fPIC", "-shared", os.path.join(CSRC, "engine.cu"), "-o", ENGINE_SO]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.run(cmd, check=True)
    return ENGINE_SO


def build_shim(force: bool = False) -> str | None:
    src = os.path.join(CSRC, "infer_b200.c")
    if not os.path.exists(src):
        return None
    hdrs = [os.path.join(ROOT, "include", "nano_infer_abi.h"), os.path.join(ROOT, "include", "nano_b200.h")]
    if force or _newer(SHIM_SO, [src] + hdrs):
        cmd = ["gcc", "-O2", "-Wall", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), src, "-o", SHIM_SO,
               "-L" + LIB, "-lnano_b200", "-Wl,-rpath,$ORIGIN", "-lm"]
        subprocess.run(cmd, check=True)
    return SHIM_SO


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_engine(force, verbose)
    build_shim(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built:", ENGINE_SO)
