"""Builds neon_b200/lib/libneon_lbm.so — the C-ABI library of hand-written sm_100a kernels (include/neon_lbm.h).

In-tree, explicit nvcc (no JIT cache): the .so travels with the repository snapshot to the GPU box.

    python -m neon_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libneon_lbm.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC,-O2,-Wall", "-I", os.path.join(ROOT, "include")]
# translation unit -> extra flags.  *_ref / setup units reproduce the reference's CPU rounding: no FMA contraction.
UNITS = {
    "lbm_api.cu": [],
    "lbm_step_fast.cu": [],
    "lbm_step_ref.cu": ["-fmad=false"],
    "lbm_setup.cu": ["-fmad=false"],
    "lbm_halo.cu": [],
    "lbm_block_fast.cu": [],
    "lbm_block_ref.cu": ["-fmad=false"],
}


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the sm_100a kernels cannot be built (there is no CPU fallback)")
    return exe


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + [
        os.path.join(ROOT, "include", "neon_lbm.h"), os.path.abspath(__file__)]


def _stale(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    deps = _deps()
    units = {u: f for u, f in UNITS.items() if os.path.exists(os.path.join(CSRC, u))}
    jobs = []
    for unit, extra in units.items():
        src = os.path.join(CSRC, unit)
        obj = os.path.join(OBJ, unit.replace(".cu", ".o"))
        if force or _stale(obj, [src] + deps):
            cmd = [nvcc()] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for out in ex.map(run, jobs):
                if verbose and out:
                    print(out, file=sys.stderr)
    objs = [os.path.join(OBJ, u.replace(".cu", ".o")) for u in units]
    if force or jobs or _stale(LIB, objs):
        run([nvcc()] + ARCH + ["-shared", "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
