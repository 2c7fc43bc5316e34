"""Stage the UNMODIFIED reference hot-path modules under `baseline/_ref/src/` (git-ignored, shipped by gpurun).

TEST / BENCH INFRASTRUCTURE ONLY.  `/root/reference` is mounted read-only in the build container and does not
exist on the GPU box; the reference is pure Python (numba), so "building" it is copying the ten modules of the
hot path byte for byte to a place that travels with the repo snapshot.  Nothing is committed (see `.gitignore`).
`bench.py --impl reference` and the `cpu_baseline` leg import them through `oracle/ref_shim.py` (bofire / gpytorch
stubs, `gammaln` overload) and drive the reference's own `_run_bark_sampler_multichain`
(src/bark/fitting/bark_sampler.py:120-213) and `forest_predict` (src/bark/tree_kernels/tree_gps.py:80-113).

    python -m oracle.build_ref            # called by __graft_entry__.build() when /root/reference is present
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = "/root/reference/src"
DST = os.path.join(ROOT, "baseline", "_ref", "src")

FILES = [
    "bark/forest.py",
    "bark/fitting/bark_sampler.py",
    "bark/fitting/tree_proposals.py",
    "bark/fitting/tree_traversal.py",
    "bark/fitting/quick_inverse.py",
    "bark/fitting/noise_scale_proposals.py",
    "bark/utils/bit_operations.py",
    "bark/tree_kernels/tree_gps.py",
    "bark/tree_kernels/tree_model_kernel.py",
    "bofire_mixed/domain.py",
]
PACKAGES = ["bark", "bark/fitting", "bark/utils", "bark/tree_kernels", "bofire_mixed"]  # empty __init__.py upstream


def build() -> str | None:
    """Copy the files; returns the destination, or None when the reference is not mounted (GPU box)."""
    if not os.path.isdir(os.path.join(SRC, "bark")):
        return DST if os.path.isdir(os.path.join(DST, "bark")) else None
    manifest = {}
    for pkg in PACKAGES:
        os.makedirs(os.path.join(DST, pkg), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, pkg, "__init__.py"), os.path.join(DST, pkg, "__init__.py"))
    for rel in FILES:
        shutil.copyfile(os.path.join(SRC, rel), os.path.join(DST, rel))
        with open(os.path.join(DST, rel), "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(os.path.dirname(DST), "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "sha256": manifest}, f, indent=1)
    return DST


if __name__ == "__main__":
    print(build())
