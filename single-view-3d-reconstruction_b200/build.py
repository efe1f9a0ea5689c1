"""Builds csrc/*.cu into the in-tree C-ABI shared library libsvr_b200.so (sm_100a only).

nvcc cross-compiles without a GPU.  The kernels are torch-free translation units behind
include/svr_b200.h, so a full rebuild takes well under a minute."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libsvr_b200.so"
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    sources = sorted(CSRC.glob("*.cu"))
    headers = sorted(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "svr_b200.h"]
    objdir = CSRC / "build"
    objdir.mkdir(exist_ok=True)
    logs = []

    def compile_one(src: Path):
        obj = objdir / (src.stem + ".o")
        if force or _stale(obj, [src] + headers):
            r = subprocess.run([NVCC, *FLAGS, "-c", str(src), "-o", str(obj)], capture_output=True, text=True)
            logs.append((src.name, r.stderr))
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        objs = list(ex.map(compile_one, sources))
    if force or _stale(LIB, objs):
        r = subprocess.run([NVCC, "-shared", "-o", str(LIB), *map(str, objs)], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stderr}")
    if verbose:
        for name, log in logs:
            print(f"== {name}\n{log}")
        (objdir / "ptxas.log").write_text("\n".join(f"== {n}\n{l}" for n, l in logs))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
