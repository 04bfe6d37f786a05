"""Build libmauv_b200.so (sm_100a only) in-tree with nvcc.

The shared library exports the plain C-ABI declared in include/mauv_b200.h; it has no
torch / Python dependency. Rebuilds only when a source is newer than the library.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB_DIR = HERE / "mauv" / "lib"
LIB = LIB_DIR / "libmauv_b200.so"

SOURCES = ["runtime.cu", "gemm_tc.cu", "stem_pool.cu", "sample_weights.cu", "elementwise.cu", "head.cu", "reduce.cu", "backward.cu", "x3.cu", "train_bwd.cu", "optim.cu", "membench.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: mauv_b200 has no prebuilt or fallback path")


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + list(CSRC.glob("*.cuh")) + [Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    LIB_DIR.mkdir(parents=True, exist_ok=True)
    obj_dir = HERE / "build"
    obj_dir.mkdir(exist_ok=True)
    nvcc = nvcc_path()
    objs = []
    procs = []
    for s in SOURCES:
        o = obj_dir / (s + ".o")
        objs.append(str(o))
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / s), "-o", str(o)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {s} ---\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libmauv_b200.so")
    cmd = [nvcc, "-shared", "-o", str(LIB), *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
