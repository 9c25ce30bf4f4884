"""In-tree build of libagdiff_b200.so with nvcc for sm_100a (no torch dependency in the library).

    python -m agdiff_b200.build [--force] [--verbose]

The shared object lands next to this file so it travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "csrc", "_obj")
LIB_PATH = os.path.join(HERE, "libagdiff_b200.so")
SOURCES = ["api.cu", "edges.cu", "encoder.cu", "schnet.cu", "tc_filter.cu", "tc_filter16.cu", "tc_cfconv.cu", "tc_mlp16.cu", "tc_mlp.cu", "tc_node.cu", "tc_node16.cu", "gin.cu", "step.cu", "ops.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the CUDA extension cannot be built")


def _digest() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for name in sorted(os.listdir(root)):
            p = os.path.join(root, name)
            if os.path.isfile(p) and name.endswith((".cu", ".cuh", ".h")):
                h.update(name.encode())
                h.update(open(p, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(os.environ.get("AGD_BUILD_DEFS", "").encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    stamp = os.path.join(OBJ_DIR, "stamp")
    dig = _digest()
    if (not force and os.path.exists(LIB_PATH) and os.path.exists(stamp)
            and open(stamp).read().strip() == dig):
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    extra = ["-Xptxas", "-v"] if verbose else []
    extra += os.environ.get("AGD_BUILD_DEFS", "").split()     # e.g. AGD_BUILD_DEFS=-DAGD_F16_TIMING (diagnostic builds)

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    objs = []
    with cf.ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        for src, obj, r in ex.map(compile_one, SOURCES):
            if verbose or r.returncode != 0:
                sys.stderr.write("== %s\n%s%s\n" % (src, r.stdout, r.stderr))
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for %s" % src)
            objs.append(obj)
    cmd = [nvcc, "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s%s" % (r.stdout, r.stderr))
    with open(stamp, "w") as f:
        f.write(dig)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
