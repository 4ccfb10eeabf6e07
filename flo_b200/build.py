"""Builds flo_b200/libflo_b200.so (the C-ABI library of include/flo_b200.h) with nvcc for sm_100a.

In-tree build so the .so travels to the GPU box with the repo snapshot.  nvcc cross-compiles
without a GPU.  Usage: python -m flo_b200.build [--force]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libflo_b200.so")
SOURCES = ["flo_encode_nt512.cu", "flo_encode_nt256.cu", "flo_encode_nt128.cu", "flo_kernels.cu", "flo_decode.cu", "flo_api.cu"]
HEADERS = [os.path.join(CSRC, "flo_internal.h"), os.path.join(CSRC, "encode_v3_body.cuh"), os.path.join(os.path.dirname(HERE), "include", "flo_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2", "--fmad=false",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; flo_b200 cannot be built (there is no CPU fallback)")
    return exe


def stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines: tuple = (), out: str = SO) -> str:
    """defines/out are for experiments (e.g. defines=("FLO_NT=1024",), out=".../libflo_b200_x.so")."""
    if not force and not stale() and out == SO:
        return SO
    from concurrent.futures import ThreadPoolExecutor
    tag = "" if out == SO else "_" + os.path.basename(out).replace(".so", "")

    def compile_one(src: str) -> str:
        obj = os.path.join(CSRC, src.replace(".cu", tag + ".o"))
        cmd = [nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(len(SOURCES)) as ex:          # the three kernel variants dominate: build them side by side
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc(), "-shared", "-o", out, *objs, "-cudart", "static"]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    defs = tuple(a[2:] for a in sys.argv[1:] if a.startswith("-D"))
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs,
                out=os.path.join(HERE, outs[0]) if outs else SO))
