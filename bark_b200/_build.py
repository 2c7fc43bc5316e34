"""Build recipe for libbark_b200.so (hand-written sm_100a CUDA + the C ABI of include/bark_b200.h).

`nvcc` cross-compiles without a GPU.  The library is built IN-TREE (bark_b200/libbark_b200.so) so that it
travels to the GPU box with the repo snapshot."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libbark_b200.so")
SOURCES = ["forest.cu", "gram.cu", "gram_umma.cu", "mll.cu", "mcmc.cu", "predict.cu", "predict_umma.cu", "kinv.cu", "prior.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "bark_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA sources into bark_b200/libbark_b200.so (no-op when up to date)."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libbark_b200.so")
    extra = os.environ.get("BARK_NVCC_EXTRA", "").split()  # e.g. -DBARK_PHASE_TIMING for the instrumented build
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-o", LIB_PATH + ".tmp", *[os.path.join(CSRC, s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv))
