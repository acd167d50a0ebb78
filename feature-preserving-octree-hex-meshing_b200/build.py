"""Build recipe for libfpohm.so (hand-written sm_100a CUDA + C-ABI), in-tree, no JIT cache.

`python build.py` or `__graft_entry__.build()`.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OBJ = HERE / "build"
LIB = HERE / "libfpohm.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# -fmad=false: the parity contract is "same IEEE-754 operations as the reference's SSE2 build" 
NVCC_FLAGS = ARCH + ["-O3", "-std=c++17", "-lineinfo", "-fmad=false", "-Xcompiler", "-fPIC,-O2,-fno-fast-math,-ffp-contract=off",
                     "-ccbin", "/usr/bin/g++", "-Xptxas", "-v", "-I", str(HERE.parent / "include")]
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-pthread", "-fno-fast-math", "-ffp-contract=off", "-I", "/usr/local/cuda/include",
             "-I", str(HERE.parent / "include")]


def _newer(src: Path, dst: Path, deps) -> bool:
    if not dst.exists():
        return True
    t = dst.stat().st_mtime
    return any(p.stat().st_mtime > t for p in [src, *deps])


def build(verbose: bool = False, force: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    headers = list(CSRC.glob("*.h")) + [HERE.parent / "include" / "fpohm.h", Path(__file__)]
    cu = sorted(CSRC.glob("*.cu"))
    cpp = sorted(CSRC.glob("*.cpp"))
    jobs = []
    for src in cu:
        o = OBJ / (src.stem + ".o")
        if force or _newer(src, o, headers):
            jobs.append(([NVCC, *NVCC_FLAGS, "-c", str(src), "-o", str(o)], src))
    for src in cpp:
        o = OBJ / (src.stem + ".o")
        if force or _newer(src, o, headers):
            jobs.append((["/usr/bin/g++", *CXX_FLAGS, "-c", str(src), "-o", str(o)], src))

    def run(job):
        cmd, src = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        (OBJ / (src.stem + ".log")).write_text(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"compile failed: {src}\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stdout + r.stderr)
        return src

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    objs = [str(OBJ / (s.stem + ".o")) for s in cu + cpp]
    if jobs or not LIB.exists():
        r = subprocess.run([NVCC, *ARCH, "-shared", "-o", str(LIB), *objs, "-ccbin", "/usr/bin/g++"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
