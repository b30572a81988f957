"""Build liboip_b200.so (hand-written CUDA for sm_100a + the C ABI) in-tree with nvcc.

    python -m opticalimageprocessor_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liboip_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",                      # parity: the reference arithmetic has no FMA contraction
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-fvisibility=default",
    "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.hpp")) + [
        os.path.join(HERE, "host", "oip_cli.cpp"),
        os.path.join(HERE, "..", "include", "oip_b200.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    objs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(HERE, "build", os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [NVCC] + FLAGS + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(out)
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f"nvcc failed on {src}")
    with open(os.path.join(HERE, "build", "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-lcufft"]
    subprocess.check_call(cmd)
    build_cli()
    return LIB


CLI_BIN = os.path.join(HERE, "OpticalImageProcessor")


def build_cli() -> str:
    """the reference-grammar command line (C++17 host code over the C ABI)"""
    src = os.path.join(HERE, "host", "oip_cli.cpp")
    cmd = [os.environ.get("CXX", "g++"), "-std=c++17", "-O2", "-Wall", "-o", CLI_BIN, src, "-L" + HERE, "-loip_b200",
           "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + HERE]
    subprocess.check_call(cmd)
    return CLI_BIN


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
