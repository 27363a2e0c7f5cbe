"""Build libdic_b200.so in-tree with nvcc for sm_100a.

    python -m deep_interpolation_clustering_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  Objects are cached under csrc/build/ keyed on the
source/header mtimes; the shared library lands next to this file so it travels with
the tree (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(REPO, "include")
OUT = os.path.join(HERE, "libdic_b200.so")
BUILD = os.path.join(CSRC, "build")

SOURCES = ["api.cu", "upload.cu", "interp_sci.cu", "interp_cci.cu", "interp_rbf.cu", "dec.cu", "kmeans.cu", "kmeans_tc.cu", "pairwise_tc.cu", "cluster_eval.cu", "lstm.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    return hdrs


# the sources that define the kernels of the headline (c2) step: what profiles/dram_traffic.json is tied to
STEP_SOURCES = ["common.cuh", "interp_stage.cuh", "interp_sci.cu", "interp_cci.cu", "interp_rbf.cu", "dec.cu"]


def sources_sha256():
    """SHA-256 over the sources of the headline step's kernels (sorted by name): what a measured ncu capture of that
    step is tied to (changes to the k-means / pairwise / LSTM sources do not invalidate it)."""
    import hashlib
    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC, f) for f in STEP_SOURCES)
    for f in files:
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=False):
    nvcc = _nvcc()
    os.makedirs(BUILD, exist_ok=True)
    deps = _deps()
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]

    def compile_one(src):
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        path = os.path.join(CSRC, src)
        if force or _stale(obj, [path] + deps):
            # NVCC_EXTRA: extra compile flags for a BENCHMARK build (e.g. -DDIC_TC_PROFILE, benchmarks/README.md);
            # a build-time switch of this script, not something the library reads
            cmd = [nvcc] + NVCC_FLAGS + os.environ.get("NVCC_EXTRA", "").split() + \
                ["-I", INCLUDE, "-I", CSRC, "-c", path, "-o", obj]
            if verbose:
                cmd[1:1] = ["-Xptxas", "-v"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or r.returncode:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError(f"nvcc failed on {src}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    if force or _stale(OUT, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return OUT


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
