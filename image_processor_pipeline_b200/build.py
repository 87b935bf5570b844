"""Build libippgpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m image_processor_pipeline_b200.build [--force]
"""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT = HERE / "libippgpu.so"
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-ffp-contract=off",
              "-Xcompiler", "-fPIC", "-shared"]


def sources():
    return sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "ippgpu.h"])


def up_to_date() -> bool:
    if not OUT.is_file():
        return False
    t = OUT.stat().st_mtime
    return all(p.stat().st_mtime <= t for p in sources())


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and up_to_date():
        return OUT
    cmd = ["nvcc", *NVCC_FLAGS, "-o", str(OUT), str(CSRC / "ippgpu_all.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
