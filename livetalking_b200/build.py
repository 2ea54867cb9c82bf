"""Build libltb200.so (sm_100a only) in-tree with nvcc.

    python -m livetalking_b200.build            # incremental
    python -m livetalking_b200.build --force

The library is written to livetalking_b200/lib/libltb200.so so that it travels with the
source tree (git-ignored, not gpurun-ignored).  There is exactly one target architecture:
-gencode arch=compute_100a,code=sm_100a.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libltb200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-std=c++17", "-O3", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libltb200 cannot be built")


def sources(diag: bool = False):
    """Product sources = csrc/*.cu; csrc/diag/*.cu (hardware probes) go only into the diagnostic library."""
    src = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    if diag:
        d = os.path.join(CSRC, "diag")
        src += sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cu"))
    return src


def _deps_mtime() -> float:
    mt = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".h", ".cuh")):
                mt = max(mt, os.path.getmtime(os.path.join(root, f)))
    return mt


def build(force: bool = False, verbose: bool = False, defines=(), tag: str = "") -> str:
    """defines/tag: diagnostic variants only (e.g. tools/diag_halo.py builds lib/libltb200_diag.so with -DLTB_HALO_DIAG)."""
    LIB = os.path.join(LIBDIR, f"libltb200{tag}.so")
    OBJDIR = os.path.join(HERE, "build" + tag)
    NVCC_FLAGS = [*globals()["NVCC_FLAGS"], *[f"-D{d}" for d in defines]]
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()
    hdr_mt = _deps_mtime()
    jobs = []
    objs = []
    for src in sources(diag=bool(tag)):
        obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_mt):
            cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for log in ex.map(run, jobs):
                if verbose and log:
                    print(log)
    if jobs or not os.path.exists(LIB):
        run([nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcuda"])
    return LIB


if __name__ == "__main__":
    if "--diag" in sys.argv:     # diagnostic library: product kernels + LTB_HALO_DIAG knock-outs + the csrc/diag probes
        print(build(force="--force" in sys.argv, defines=("LTB_HALO_DIAG",), tag="_diag"))
        sys.exit(0)
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
