"""Builds libb200ann.so (the C-ABI library, include/b200ann.h) for sm_100a with nvcc, in-tree.

    python the-algorithm_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels with gpurun snapshots.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB_DIR = HERE / "lib"
LIB = LIB_DIR / "libb200ann.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]
OBJ_DIR = LIB_DIR / "obj"


def sources():
    return sorted(CSRC.glob("*.cu"))


def headers():
    return list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list(CSRC.glob("*.hpp")) + [HERE.parent / "include" / "b200ann.h"]


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(p.stat().st_mtime > t for p in sources() + headers())


def build(force: bool = False, verbose: bool = False) -> Path:
    """One object per .cu (compiled in parallel, rebuilt only when the file or a header is newer), then one link."""
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        if LIB.exists():  # GPU box image always has nvcc, but do not fail on a prebuilt library
            return LIB
        raise RuntimeError("nvcc not found and no prebuilt libb200ann.so")
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    hdr_t = max(p.stat().st_mtime for p in headers())
    jobs = []
    for src in sources():
        obj = OBJ_DIR / (src.stem + ".o")
        if force or not obj.exists() or obj.stat().st_mtime < max(src.stat().st_mtime, hdr_t):
            jobs.append((src, obj))
    log = []

    def compile_one(job):
        src, obj = job
        cmd = [nvcc, *NVCC_FLAGS, "-c", "-o", str(obj), str(src)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, res

    from concurrent.futures import ThreadPoolExecutor
    failed = False
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for cmd, res in ex.map(compile_one, jobs):
            log.append(" ".join(cmd) + "\n" + res.stdout + res.stderr)
            if verbose or res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
            failed = failed or res.returncode != 0
    if not failed:
        tmp = LIB_DIR / "libb200ann.so.tmp"
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(tmp),
               *[str(OBJ_DIR / (s.stem + ".o")) for s in sources()], "-lcudart", "-lcuda", "-lpthread"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log.append(" ".join(cmd) + "\n" + res.stdout + res.stderr)
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
        failed = res.returncode != 0
        if not failed:
            os.replace(tmp, LIB)
    (LIB_DIR / "build.log").write_text("\n".join(log))
    if failed:
        raise RuntimeError("nvcc failed; see the-algorithm_b200/lib/build.log")
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
