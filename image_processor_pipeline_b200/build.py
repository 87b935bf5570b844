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


def build(force: bool = False, verbose: bool = False, out: Path = OUT, defines=()) -> Path:
    """``out`` / ``defines``: tuning variants (e.g. -DIPP_SEG_THREADS=512) built next to the product library
    and selected at run time with IPPGPU_LIB=<path>; the product build takes neither."""
    if not force and out == OUT and up_to_date():
        return OUT
    out.parent.mkdir(parents=True, exist_ok=True)
    cmd = ["nvcc", *NVCC_FLAGS, *defines, "-o", str(out), str(CSRC / "ippgpu_all.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    return out


JPEG_OUT = HERE / "libippjpeg.so"
JPEG_SRC = HERE / "csrc_jpeg" / "ippjpeg.cpp"


def build_jpeg(force: bool = False) -> Path:
    """libippjpeg.so: the nvJPEG codec layer (include/ippjpeg.h), its own library so that libippgpu does not depend on libnvjpeg."""
    srcs = [JPEG_SRC, HERE.parent / "include" / "ippjpeg.h"]
    if not force and JPEG_OUT.is_file() and all(p.stat().st_mtime <= JPEG_OUT.stat().st_mtime for p in srcs):
        return JPEG_OUT
    cmd = ["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-fPIC", "-shared", "-o", str(JPEG_OUT), str(JPEG_SRC), "-lnvjpeg"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
    return JPEG_OUT


VIDEO_OUT = HERE / "libippvideo.so"
VIDEO_SRC = HERE / "csrc_video" / "ippvideo.cu"


def build_video(force: bool = False) -> Path:
    """libippvideo.so: the NVDEC frame source (include/ippvideo.h).  libnvcuvid belongs to the driver and is opened at run time
    (dlopen): the library builds and loads on machines without it."""
    srcs = [VIDEO_SRC, HERE.parent / "include" / "ippvideo.h"]
    if not force and VIDEO_OUT.is_file() and all(p.stat().st_mtime <= VIDEO_OUT.stat().st_mtime for p in srcs):
        return VIDEO_OUT
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
           "-o", str(VIDEO_OUT), str(VIDEO_SRC), "-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
    return VIDEO_OUT


if __name__ == "__main__":
    if "--variant" in sys.argv:                       # python -m ...build --variant NAME -DX=1 -DY=2
        name = sys.argv[sys.argv.index("--variant") + 1]
        print(build(force=True, verbose="-v" in sys.argv, out=HERE / "_variants" / f"libippgpu_{name}.so",
                    defines=[a for a in sys.argv if a.startswith("-D")]))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
        print(build_jpeg(force="--force" in sys.argv))
        print(build_video(force="--force" in sys.argv))
