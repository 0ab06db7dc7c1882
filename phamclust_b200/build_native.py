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
SOURCES = ["pc_api.cu", "pc_contract_popc.cu", "pc_contract_i8.cu", "pc_contract_i8p.cu", "pc_align.cu", "pc_gather.cu", "pc_resident.cu",
           "pc_io.cu", "pc_dedup.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xptxas", "-v",
              "-lcuda", "--threads", "8"]


def _newest_source_mtime():
    files = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
    files.append(HERE.parent / "include" / "phamclust_b200.h")
    return max(f.stat().st_mtime for f in files)


PYPACK_OUT = HERE / "libphamclust_pypack.so"
BUILD_RECORD = HERE / "build_record.json"


def build_pypack(force=False):
    """gcc: the CPython-API packer (csrc/pc_pypack.c, list[Genome] -> CSR + arena in one walk);
    resolved against the running interpreter when ctypes.PyDLL loads it."""
    import sysconfig
    src = CSRC / "pc_pypack.c"
    if not force and PYPACK_OUT.exists() and PYPACK_OUT.stat().st_mtime >= src.stat().st_mtime:
        return PYPACK_OUT, False
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-Wall", "-I", sysconfig.get_paths()["include"],
           "-o", str(PYPACK_OUT), str(src)]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("gcc failed building libphamclust_pypack.so")
    return PYPACK_OUT, True


def build(force=False, verbose=False, out=None):
    """``out``: build a variant (PC_NVCC_EXTRA defines) next to the product library.
    Writes build_record.json (what was compiled in this call and what was reused, with the
    sha256 of the product library) so a run can say which build its numbers come from."""
    if out is not None:
        return _compile(pathlib.Path(out), verbose, log=False)
    _, packed = build_pypack(force)
    if not force and OUT.exists() and OUT.stat().st_mtime >= _newest_source_mtime():
        _record("reused", packed)
        return OUT
    so = _compile(OUT, verbose, log=True)
    _record("compiled", packed)
    return so


def library_sha256(path=None):
    import hashlib
    h = hashlib.sha256()
    with open(path or OUT, "rb") as fh:
        for chunk in iter(lambda: fh.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def _record(mode, pypack_compiled):
    import json
    import time
    try:
        BUILD_RECORD.write_text(json.dumps({
            "build_mode": mode, "pypack": "compiled" if pypack_compiled else "reused",
            "library": OUT.name, "sha256": library_sha256(), "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
            "nvcc_flags": NVCC_FLAGS}, indent=1) + "\n")
    except OSError:
        pass


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
