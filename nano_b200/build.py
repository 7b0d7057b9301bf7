"""In-tree build of the product libraries (nvcc cross-compiles sm_100a without a GPU).

    nano_b200/lib/libnano_b200.so        engine + C-ABI (include/nano_b200.h)        <- csrc/engine.cu, kernels.cuh
    nano_b200/lib/libnano_infer_b200.so  reference host API (infer.h symbols)        <- csrc/infer_b200.c
    nano_b200/lib/libnano_refhost.so     the REFERENCE's own host objects that stay in the link line
                                         (tokenizer.c utils.c hal_{ram,fs,os,misc}_linux.c), compiled unmodified
                                         where they lie under /root/reference -- only when that tree exists
    nano_b200/lib/nano_cli               infer/main_cli.c linked against our libraries ("nano_cli links unchanged")

The outputs are git-ignored but travel to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "nano_b200", "csrc")
LIB = os.path.join(ROOT, "nano_b200", "lib")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
REF = os.environ.get("NANO_REFERENCE", "/root/reference/infer")

ENGINE_SO = os.path.join(LIB, "libnano_b200.so")
SHIM_SO = os.path.join(LIB, "libnano_infer_b200.so")
REFHOST_SO = os.path.join(LIB, "libnano_refhost.so")
NANO_CLI = os.path.join(LIB, "nano_cli")
REFHOST_SRCS = ["tokenizer.c", "utils.c", "hal_ram_linux.c", "hal_fs_linux.c", "hal_os_linux.c", "hal_misc_linux.c"]


def _newer(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources if os.path.exists(s))


# translation units of libnano_b200.so and what each depends on (besides itself)
UNITS = {
    "engine.cu": ["kernels.cuh", "expf_ref.cuh", "stream_args.h", "stream_host.h", "sample_host.h", "../../include/nano_b200.h"],
    "stream.cu": ["kernels.cuh", "expf_ref.cuh", "stream_args.h", "stream_host.h", "stream.cuh"],
    "sample.cu": ["kernels.cuh", "expf_ref.cuh", "sample_host.h"],
}


def build_engine(force: bool = False, verbose: bool = False) -> str:
    """One object per translation unit (compiled in parallel, only when its sources changed), then one link."""
    os.makedirs(LIB, exist_ok=True)
    base = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
    if verbose:
        base.insert(1, "-Xptxas=-v")
    objs, procs = [], []
    for unit, deps in UNITS.items():
        obj = os.path.join(LIB, unit[:-3] + ".o")
        objs.append(obj)
        if force or _newer(obj, [os.path.join(CSRC, unit)] + [os.path.join(CSRC, d) for d in deps]):
            procs.append((unit, subprocess.Popen(base + ["-c", os.path.join(CSRC, unit), "-o", obj])))
    failed = [u for u, p in procs if p.wait() != 0]
    if failed:
        raise RuntimeError("nvcc failed for " + ", ".join(failed))
    if procs or _newer(ENGINE_SO, objs):
        subprocess.run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", ENGINE_SO] + objs, check=True)
    return ENGINE_SO


TRACE_SO = os.path.join(LIB, "libnano_b200_trace.so")


def build_trace_variant() -> str:
    """Developer build of the streaming kernel with its fine cycle stamps compiled in (-DNB200_FINE_TRACE; they cost ~200
    cycles each, so the product library never carries them).  Loaded by tools/gpu_trace.py through NB200_ENGINE_SO."""
    build_engine()
    obj = os.path.join(LIB, "stream_trace.o")
    subprocess.run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
                    "-DNB200_FINE_TRACE", "-c", os.path.join(CSRC, "stream.cu"), "-o", obj], check=True)
    objs = [os.path.join(LIB, u[:-3] + ".o") for u in UNITS if u != "stream.cu"] + [obj]
    subprocess.run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", TRACE_SO] + objs, check=True)
    return TRACE_SO


def build_shim(force: bool = False) -> str:
    src = os.path.join(CSRC, "infer_b200.c")
    hdrs = [os.path.join(ROOT, "include", "nano_infer_abi.h"), os.path.join(ROOT, "include", "nano_b200.h")]
    if force or _newer(SHIM_SO, [src] + hdrs + [ENGINE_SO]):
        cmd = ["gcc", "-O2", "-Wall", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), src, "-o", SHIM_SO,
               "-L" + LIB, "-lnano_b200", "-Wl,-rpath,$ORIGIN", "-lm"]
        subprocess.run(cmd, check=True)
    return SHIM_SO


def reference_present() -> bool:
    return os.path.exists(os.path.join(REF, "main_cli.c"))


def build_refhost(force: bool = False):
    """The reference's unchanged host-side objects as a shared library (so ctypes can load the shim on the GPU box)."""
    if not reference_present():
        return REFHOST_SO if os.path.exists(REFHOST_SO) else None
    srcs = [os.path.join(REF, s) for s in REFHOST_SRCS]
    if force or _newer(REFHOST_SO, srcs):
        subprocess.run(["gcc", "-DNANO_CLI", "-O2", "-w", "-fPIC", "-shared", "-pthread", "-I" + REF] + srcs +
                       ["-o", REFHOST_SO, "-lm"], check=True)
    return REFHOST_SO


def build_nano_cli(force: bool = False):
    """infer/Makefile:145-147 with `tensor.c infer.c` replaced by our two libraries."""
    if not reference_present():
        return NANO_CLI if os.path.exists(NANO_CLI) else None
    srcs = [os.path.join(REF, s) for s in ["main_cli.c"] + REFHOST_SRCS]
    if force or _newer(NANO_CLI, srcs + [SHIM_SO, ENGINE_SO]):
        subprocess.run(["gcc", "-DNANO_CLI", "-O3", "-march=native", "-ffast-math", "-w", "-fopenmp", "-pthread", "-I" + REF] + srcs +
                       ["-o", NANO_CLI, "-L" + LIB, "-lnano_infer_b200", "-lnano_b200", "-Wl,-rpath,$ORIGIN", "-lm"], check=True)
    return NANO_CLI


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_engine(force, verbose)
    build_shim(force)
    build_refhost(force)
    build_nano_cli(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built:", ENGINE_SO, SHIM_SO, REFHOST_SO if os.path.exists(REFHOST_SO) else "(no refhost)",
          NANO_CLI if os.path.exists(NANO_CLI) else "(no nano_cli)")
