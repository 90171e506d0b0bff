"""Build libgeomb200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m geomjax_b200.build [--force] [--verbose]

One translation unit per sampler, compiled in parallel; objects and the .so live under
geomjax_b200/_lib/ (git-ignored, but shipped to the GPU box by gpurun).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB_DIR = HERE / "_lib"
LIB = LIB_DIR / "libgeomb200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC"]
# experiment builds, e.g. GEOMB200_EXTRA_NVCC_FLAGS=-DGB_FT_TIMING for tools/ft_stamps.py (part of the build digest:
# unsetting it rebuilds the shipped library)
FLAGS += os.environ.get("GEOMB200_EXTRA_NVCC_FLAGS", "").split()
LPC_GROUPS = (1, 2, 4, 8, 32)


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _digest():
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
                    + [HERE.parent / "include" / "geomb200.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    LIB_DIR.mkdir(exist_ok=True)
    stamp = LIB_DIR / "build.sha256"
    dig = _digest()
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == dig:
        return LIB
    if not Path(NVCC).exists():
        raise RuntimeError(f"nvcc not found at {NVCC}; cannot build {LIB}")

    jobs = []
    for src in _sources():
        if src.name.endswith("_launch.cu"):  # one object per lanes-per-chain group
            jobs += [(src, lpc) for lpc in LPC_GROUPS]
        else:
            jobs.append((src, None))

    def compile_one(job):
        src, lpc = job
        obj = LIB_DIR / (src.stem + (f"_lpc{lpc}" if lpc else "") + ".o")
        cmd = [NVCC, *FLAGS, "-c", str(src), "-o", str(obj)]
        if lpc:
            cmd.append(f"-DGB_LPC={lpc}")
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, jobs))
    cmd = [NVCC, "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fPIC"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
