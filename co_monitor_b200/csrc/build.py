"""Builds libcomonitor_sm100.so in-tree with nvcc for sm_100a (no torch involved).

    python -m co_monitor_b200.csrc.build [--force]
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["geometry_host.cu", "visibility.cu", "emissivity.cu", "reff.cu", "io_host.cu", "capi.cu", "volume.cu"]
LIB = os.path.join(HERE, "libcomonitor_sm100.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
    "-shared", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cu", ".h", ".cuh"))]
    deps.append(os.path.join(HERE, "..", "..", "include", "comonitor_sm100.h"))
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    srcs = [os.path.join(HERE, s) for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    tmp = LIB + f".tmp{os.getpid()}"  # linked beside the target, then renamed over it: a process that has the old
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", tmp, *srcs]  # library mapped keeps its (unlinked) file intact
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError(f"nvcc failed ({proc.returncode}): {' '.join(cmd)}")
    os.replace(tmp, LIB)
    with open(os.path.join(HERE, "ptxas_info.txt"), "w") as fh:
        fh.write(proc.stderr)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
