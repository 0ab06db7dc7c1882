"""Compile libphamclust_b200.so IN-TREE with nvcc for sm_100a (and only that).

Usage: python -m phamclust_b200.build_native [--force]
The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import pathlib
import shlex
import shutil
import subprocess
import sys

HERE = pathlib.Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT = HERE / "libphamclust_b200.so"
SOURCES = ["pc_api.cu", "pc_contract_popc.cu", "pc_contract_i8.cu", "pc_align.cu", "pc_gather.cu",
           "pc_io.cu", "pc_dedup.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xptxas", "-v",
              "-lcuda", "--threads", "8"]


def _newest_source_mtime():
    files = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
    files.append(HERE.parent / "include" / "phamclust_b200.h")
    return max(f.stat().st_mtime for f in files)


def build(force=False, verbose=False, out=None):
    """``out``: build a variant (PC_NVCC_EXTRA defines) next to the product library."""
    if out is not None:
        return _compile(pathlib.Path(out), verbose, log=False)
    if not force and OUT.exists() and OUT.stat().st_mtime >= _newest_source_mtime():
        return OUT
    return _compile(OUT, verbose, log=True)


def _compile(OUT, verbose, log):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    extra = shlex.split(os.environ.get("PC_NVCC_EXTRA", ""))      # e.g. -DPC_NW_XL_FMA (experiments)
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-o", str(OUT), *[str(CSRC / s) for s in SOURCES]]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode:
        raise RuntimeError("nvcc failed building libphamclust_b200.so")
    if log:
        (HERE / "build.log").write_text(proc.stdout + proc.stderr)
    return OUT


if __name__ == "__main__":
    variant = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose=True, out=variant[0] if variant else None))
