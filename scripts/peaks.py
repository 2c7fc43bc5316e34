#!/usr/bin/env python
"""Build and run scripts/peaks.cu on the GPU box; writes profiles/peaks_r2.json (or the path given).

    python scripts/peaks.py [out.json]     # under gpurun; `--build-only` cross-compiles here without a GPU
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "scripts", "peaks.cu")
BIN = os.path.join(ROOT, "scripts", "peaks_bin")


def build():
    if os.path.exists(BIN) and os.path.getmtime(BIN) >= os.path.getmtime(SRC):
        return BIN
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-o", BIN, SRC],
                   check=True)
    return BIN


if __name__ == "__main__":
    build()
    if "--build-only" in sys.argv:
        sys.exit(0)
    out = next((a for a in sys.argv[1:] if not a.startswith("--")), os.path.join(ROOT, "profiles", "peaks_r2.json"))
    res = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    sys.stderr.write(res.stderr)
    if res.returncode != 0:
        sys.exit(res.returncode)
    with open(out, "w") as f:
        f.write(res.stdout)
    print(res.stdout)
