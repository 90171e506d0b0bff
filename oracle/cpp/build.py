"""Build the C++/OpenMP CPU restatement (oracle/cpp/*.cpp) into oracle/_cpp/liboraclecpu.so.

    python -m oracle.cpp.build [--force]

TEST INFRASTRUCTURE (see oracle/__init__.py): the library is loaded only by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  oracle/_cpp/ is git-ignored but travels to the GPU box.
prng.cpp is compiled with -ffp-contract=off (bit-exact normal transform); the rest may contract to FMA.
"""
from __future__ import annotations

import hashlib
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE.parent / "_cpp"
LIB = OUT / "liboraclecpu.so"
CXX = "g++"
# -march=x86-64-v3 (AVX2 + FMA), not -march=native: the library is built in the CPU container and travels to the
# GPU box, whose host CPU may be a different model
COMMON = ["-O3", "-march=x86-64-v3", "-fopenmp", "-fPIC", "-std=c++17", "-fno-math-errno"]


def _digest():
    h = hashlib.sha256()
    for p in sorted(HERE.glob("*.cpp")) + sorted(HERE.glob("*.h")):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(COMMON).encode())
    return h.hexdigest()


def build(force: bool = False) -> Path:
    OUT.mkdir(exist_ok=True)
    stamp = OUT / "build.sha256"
    dig = _digest()
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == dig:
        return LIB
    objs = []
    for src, extra in (("prng.cpp", ["-ffp-contract=off"]), ("geom_cpu.cpp", [])):
        obj = OUT / (src[:-4] + ".o")
        r = subprocess.run([CXX, *COMMON, *extra, "-c", str(HERE / src), "-o", str(obj)], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"{CXX} failed for {src}:\n{r.stderr}")
        objs.append(str(obj))
    r = subprocess.run([CXX, "-shared", "-fopenmp", "-o", str(LIB), *objs], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stderr}")
    stamp.write_text(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
