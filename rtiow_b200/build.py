"""In-tree build of librtiow_cuda.so (nvcc, sm_100a only) and of the C++ host mirror's example binary.

`python -m rtiow_b200.build` or `rtiow_b200.build.build_all()`.  The .so lands in rtiow_b200/lib/
(git-ignored, NOT gpurun-ignored, so it travels to the GPU box with the snapshot).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB_DIR = PKG / "lib"
LIB = LIB_DIR / "librtiow_cuda.so"
HOST_EXAMPLE = LIB_DIR / "rtiow_host_example"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-Wall",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found: librtiow_cuda.so cannot be built (there is no CPU fallback)")
    return cand


def _stale(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def sources():
    return sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cpp")))


def headers():
    return sorted(list(CSRC.glob("*.cuh")) + [ROOT / "include" / "rtiow_cuda.h"])


def build_lib(force: bool = False, verbose: bool = False) -> Path:
    LIB_DIR.mkdir(exist_ok=True)
    srcs = sources()
    if not force and not _stale(LIB, srcs + headers() + [Path(__file__)]):
        return LIB
    env = dict(os.environ)
    env.pop("CC", None), env.pop("CXX", None)      # the image exports a wrapper gcc that lacks specs
    extra = os.environ.get("RTIOW_NVCC_EXTRA", "").split()        # experiments only (tools/): e.g. -DRT_FLUSH_PER_PATH=1
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-shared", "-ccbin", "/usr/bin/g++", "-o", str(LIB), *map(str, srcs), "-ldl"]      # NCCL is dlopen()ed on first multi-GPU use
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB


def build_host_example(force: bool = False) -> Path:
    """C++ host mirror of the reference API (rtiow_b200/host/rtiow.hpp) + its example main."""
    src = PKG / "host" / "example_main.cpp"
    hdr = PKG / "host" / "rtiow.hpp"
    if not src.exists():
        return HOST_EXAMPLE
    build_lib()
    if not force and not _stale(HOST_EXAMPLE, [src, hdr, LIB]):
        return HOST_EXAMPLE
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-I", str(ROOT / "include"), "-I", str(PKG / "host"), str(src),
           "-o", str(HOST_EXAMPLE), "-L", str(LIB_DIR), "-lrtiow_cuda", f"-Wl,-rpath,{LIB_DIR}", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return HOST_EXAMPLE


def build_all(force: bool = False, verbose: bool = False):
    lib = build_lib(force, verbose)
    build_host_example(force)
    return lib


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
