"""Build the in-tree native libraries.

* ``holo_b200/lib/libholo_spf.so`` — the product: CUDA kernels (sm_100a) + C ABI
  (include/holo_spf.h) + host-side LSDB flatteners.  Built with nvcc; it
  cross-compiles without a GPU.
* ``oracle/_build/liboracle.so`` — TEST INFRASTRUCTURE: the CPU restatement of the
  reference algorithm (plain g++).  Building it is not using it.
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "holo_b200" / "csrc"
LIBDIR = ROOT / "holo_b200" / "lib"
PRODUCT_LIB = LIBDIR / "libholo_spf.so"
ORACLE_DIR = ROOT / "oracle"
ORACLE_LIB = ORACLE_DIR / "_build" / "liboracle.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-Wall", "-shared",
    "-I", str(ROOT / "include"),
]


def _newer(target: Path, sources) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(Path(s).stat().st_mtime <= t for s in sources)


def _run(cmd):
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("build failed: " + " ".join(map(str, cmd)))
    return proc.stdout + proc.stderr


def product_sources():
    return sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cc")))


def build_product(force: bool = False, verbose: bool = False) -> Path:
    srcs = product_sources()
    deps = srcs + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list((ROOT / "include").glob("*.h"))
    if not force and _newer(PRODUCT_LIB, deps):
        return PRODUCT_LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    LIBDIR.mkdir(parents=True, exist_ok=True)
    cmd = [nvcc, *NVCC_FLAGS]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", str(PRODUCT_LIB), *map(str, srcs)]
    out = _run(cmd)
    if verbose:
        print(out)
    return PRODUCT_LIB


def oracle_sources():
    return sorted(ORACLE_DIR.glob("*.cc"))


def build_oracle(force: bool = False) -> Path:
    srcs = oracle_sources()
    deps = srcs + list(ORACLE_DIR.glob("*.h")) + list((ROOT / "include").glob("*.h"))
    if not force and _newer(ORACLE_LIB, deps):
        return ORACLE_LIB
    ORACLE_LIB.parent.mkdir(parents=True, exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-pthread",
           "-I", str(ROOT / "include"), "-o", str(ORACLE_LIB), *map(str, srcs)]
    _run(cmd)
    return ORACLE_LIB


def build_all(force: bool = False, verbose: bool = False):
    return build_product(force, verbose), build_oracle(force)


if __name__ == "__main__":
    p, o = build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
    print(o)
