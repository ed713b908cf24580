"""Build libxbitinfo_b200.so in-tree with nvcc for sm_100a (no torch extension machinery:
the library is a plain C-ABI shared object, loaded with ctypes).

    python -m xbitinfo_b200.csrc.build [--force] [--verbose]
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libxbitinfo_b200.so")
SOURCES = ["xbi_api.cu", "xbi_count_generic.cu", "xbi_count_tma.cu", "xbi_count_cols.cu", "xbi_count_xy.cu", "xbi_epilogue.cu", "xbi_finish.cu", "xbi_bitround.cu", "xbi_shuffle.cu", "xbi_chunked.cu", "xbi_chunked_walk.cu", "xbi_hostpool.cu"]
HEADERS = ["xbi_common.cuh", "xbi_elem.cuh", "xbi_mask.cuh", "xbi_epilogue.cuh", "xbi_tma.cuh", "xbi_tmap.h", "xbi_internal.h", "../../include/xbitinfo_b200.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unknown-pragmas,-pthread",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or add /usr/local/cuda/bin to PATH)")


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    hdrs = [os.path.join(HERE, h) for h in HEADERS] + [os.path.abspath(__file__)]
    objs, jobs = [], []
    for src in SOURCES:
        s = os.path.join(HERE, src)
        o = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            jobs.append((s, o))
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)

    def compile_one(job):
        s, o = job
        cmd = [nvcc, *NVCC_FLAGS, "-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return s, r

    with cf.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for s, r in ex.map(compile_one, jobs):
            log = os.path.join(HERE, "build", os.path.basename(s) + ".ptxas.log")
            with open(log, "w") as f:
                f.write(r.stdout + r.stderr)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError(f"nvcc failed on {s}")
            if verbose:
                sys.stderr.write(r.stderr)
    if jobs or force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB, *objs, "-cudart", "static", "-Xlinker", "--no-undefined", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
