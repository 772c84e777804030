"""Build recipe for libimm3gpu.so (nvcc, sm_100a only).

The shared library is built IN-TREE (immutable3_b200/libimm3gpu.so) so that it travels to the GPU
box with the repo snapshot; it is git-ignored.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libimm3gpu.so")

SOURCES = ["meta.cpp", "writer.cpp", "plan.cpp", "sql.cpp", "store.cpp", "kernels.cu", "engine.cu", "encode.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".hpp", ".cuh", ".h"))) + ["../../include/imm3.h"]  # every header is a dependency

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libimm3gpu.so cannot be built (there is no CPU fallback)")


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_lib(force: bool = False, verbose: bool = False, extra_flags=(), out_path: str = LIB, objdir_name: str = "build") -> str:
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    if not force and not _stale(out_path, deps):
        return out_path
    objdir = os.path.join(HERE, objdir_name)
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s) + ".o")
        objs.append(o)
        cmd = [_nvcc(), *NVCC_FLAGS, *extra_flags, "-x", "cu", "-c", s, "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for cmd, p in procs:
        out, _ = p.communicate()
        log.append("$ " + " ".join(cmd) + "\n" + out)
        if p.returncode:
            raise RuntimeError("nvcc failed:\n" + log[-1])
    link = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out_path, *objs]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log.append("$ " + " ".join(link) + "\n" + r.stdout)
    if r.returncode:
        raise RuntimeError("link failed:\n" + log[-1])
    with open(os.path.join(objdir, "build.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return out_path


if __name__ == "__main__":
    build_lib(force="--force" in sys.argv, verbose=True)
    print(LIB)
