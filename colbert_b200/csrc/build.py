"""In-tree build of libcolbert_b200.so with nvcc for sm_100a (cross-compiles without a GPU).

    python -m colbert_b200.csrc.build [--force] [--verbose]

The shared object is written next to the sources (git-ignored, but it travels to the GPU box with
the repo snapshot).  Only the CUDA runtime is linked (statically); the one driver symbol needed for
TMA descriptors is resolved at run time through cudaGetDriverEntryPoint.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SOURCES = ["api.cu", "rerank.cu", "topk.cu", "gather.cu", "partition.cu", "exhaustive.cu", "rerank_generic.cu", "rerank_umma.cu", "rerank_wide.cu", "score_allpairs.cu", "rerank_wide_stream.cu"]
# self-test / issue-rate probes of the tcgen05 building blocks: a separate library (tests and benchmarks only), linked against
# the product library for error reporting and launch accounting
PROBE_SOURCES = ["umma_probe.cu"]
HEADERS = ["cbk_common.cuh", "umma.cuh", os.path.join(ROOT, "include", "colbert_b200.h"),
           os.path.join(ROOT, "include", "colbert_b200_probe.h")]
LIB = os.path.join(HERE, "libcolbert_b200.so")
PROBE_LIB = os.path.join(HERE, "libcolbert_b200_probe.so")
STAMP = os.path.join(HERE, ".build_stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _digest() -> str:
    h = hashlib.sha256()
    for f in SOURCES + PROBE_SOURCES + HEADERS + [os.path.abspath(__file__)]:
        path = f if os.path.isabs(f) else os.path.join(HERE, f)
        with open(path, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    digest = _digest()
    if (not force and os.path.exists(LIB) and os.path.exists(PROBE_LIB) and os.path.exists(STAMP)
            and open(STAMP).read().strip() == digest):
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src: str) -> str:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(HERE, src), "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log = res.stdout + res.stderr
        with open(obj + ".log", "w") as fh:
            fh.write(" ".join(cmd) + "\n" + log)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{log}")
        if verbose:
            print(log, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES) + len(PROBE_SOURCES))) as ex:
        all_objs = list(ex.map(compile_one, SOURCES + PROBE_SOURCES))
    objs, probe_objs = all_objs[:len(SOURCES)], all_objs[len(SOURCES):]
    link = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    link = [nvcc, "-shared", "-o", PROBE_LIB, *probe_objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
            "-L" + HERE, "-lcolbert_b200", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN"]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link of the probe library failed:\n" + res.stdout + res.stderr)
    with open(STAMP, "w") as fh:
        fh.write(digest)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
