"""In-tree nvcc build of the sm_100a kernel library (csrc/*.cu -> csrc/build/libunetk.so).

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels with the source
snapshot to the GPU box.  Rebuilds only the objects whose sources (or headers) changed.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(CSRC, "build")
LIB = os.path.join(BUILD, "libunetk.so")
SOURCES = ["runtime.cu", "igemm.cu", "wgrad.cu", "elementwise.cu", "head_loss.cu", "head_fused.cu", "first_conv.cu", "seg_loss.cu",
           "eval_metrics.cu", "preprocess.cu", "optim.cu", "api.cu"]
HEADERS = ["common.cuh", "kernels.h", os.path.join("..", "..", "include", "unetk.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _digest(paths):
    h = hashlib.sha256()
    h.update(" ".join(FLAGS).encode())
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    headers = [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    jobs = []
    objs = []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        stamp = obj + ".sha"
        dig = _digest([sp] + headers)
        objs.append(obj)
        old = open(stamp).read() if os.path.exists(stamp) else ""
        if force or not os.path.exists(obj) or old != dig:
            jobs.append((sp, obj, stamp, dig))

    def compile_one(job):
        sp, obj, stamp, dig = job
        cmd = [NVCC] + FLAGS + ["-c", sp, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(obj + ".log", "w") as f:
            f.write(log)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {sp}:\n{log}")
        with open(stamp, "w") as f:
            f.write(dig)
        return log

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for log in ex.map(compile_one, jobs):
                if verbose:
                    print(log)
    if jobs or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                     "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
