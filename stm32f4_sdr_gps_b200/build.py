"""In-tree native build of the engine.

Two shared libraries are produced under ``stm32f4_sdr_gps_b200/lib/`` (git-ignored, but they travel
to the GPU box with the repo snapshot):

* ``libgpsb_cuda.so``  - hand-written sm_100a kernels + the C ABI of ``include/gpsb.h`` (nvcc)
* ``libgpsb_host.so``  - host-side C mirror of the reference's acquisition / tracking / gps_master
  state machines (``include/gpsb_host.h``), linked against the CUDA library (gcc)

Nothing here depends on torch; nvcc cross-compiles for sm_100a without a GPU present.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO_DIR = PKG_DIR.parent
LIB_DIR = PKG_DIR / "lib"
CSRC_DIR = PKG_DIR / "csrc"
HOST_DIR = PKG_DIR / "host"
CORE_DIR = PKG_DIR / "core"      # sources compiled into BOTH libraries (the per-millisecond loop step)
INCLUDE_DIR = REPO_DIR / "include"

CUDA_LIB = LIB_DIR / "libgpsb_cuda.so"
HOST_LIB = LIB_DIR / "libgpsb_host.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",            # the only float ops are the detector's add/sqrt; never contract
    "-Xcompiler", "-fPIC", "-shared",
]
# host C: IEEE semantics exactly like the reference build recipe (no fast-math, no FMA contraction)
GCC_FLAGS = ["-std=gnu11", "-O2", "-fPIC", "-shared", "-fno-strict-aliasing", "-ffp-contract=off",
             "-fno-fast-math", "-pthread", "-D_GNU_SOURCE", "-Wall", "-Wextra", "-Wno-unused-parameter"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build the sm_100a kernels")


def _newer(target: Path, sources) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(Path(s).stat().st_mtime <= t for s in sources)


def _run(cmd) -> None:
    proc = subprocess.run([str(c) for c in cmd], capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("build failed: %s\n%s\n%s" % (" ".join(map(str, cmd)), proc.stdout, proc.stderr))


def build_cuda(force: bool = False) -> Path:
    srcs = sorted(CSRC_DIR.glob("*.cu"))
    deps = srcs + sorted(CSRC_DIR.glob("*.cuh")) + sorted(INCLUDE_DIR.glob("*.h")) + sorted(CORE_DIR.glob("*.h"))
    if not force and _newer(CUDA_LIB, deps):
        return CUDA_LIB
    LIB_DIR.mkdir(exist_ok=True)
    extra = os.environ.get("GPSB_NVCC_EXTRA", "").split()      # experiments only (e.g. -DGPSB_LOOP_EXPERIMENTS)
    _run([_nvcc(), *NVCC_FLAGS, *extra, "-I", INCLUDE_DIR, "-o", CUDA_LIB, *srcs])
    return CUDA_LIB


def build_host(force: bool = False) -> Path:
    srcs = sorted(HOST_DIR.glob("*.c"))
    if not srcs:
        return HOST_LIB
    deps = srcs + sorted(HOST_DIR.glob("*.h")) + sorted(INCLUDE_DIR.glob("*.h")) + sorted(CORE_DIR.glob("*.h")) + [CUDA_LIB]
    if not force and _newer(HOST_LIB, deps):
        return HOST_LIB
    LIB_DIR.mkdir(exist_ok=True)
    cc = os.environ.get("CC") or shutil.which("gcc") or "gcc"
    _run([cc, *GCC_FLAGS, "-I", INCLUDE_DIR, "-I", HOST_DIR, "-o", HOST_LIB, *srcs,
          "-L", LIB_DIR, "-lgpsb_cuda", "-Wl,-rpath,$ORIGIN", "-lm"])
    return HOST_LIB


def build_oracle() -> None:
    """Compile the checker (oracle/liboracle.so and, where /root/reference exists, oracle/_ref).

    Building the checker is not using it: the product libraries never link against it."""
    _run(["make", "-C", REPO_DIR / "oracle", "all"])


def build_all(force: bool = False) -> None:
    build_cuda(force)
    build_host(force)


if __name__ == "__main__":
    build_all(force=True)
    build_oracle()
    print("built:", CUDA_LIB, HOST_LIB)
