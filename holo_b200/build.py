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


def _fingerprint(tool: str, flags, deps) -> str:
    """What the library was built from: compiler identity, flags and the bytes of every source and header.
    (Not mtimes: a prebuilt .so that travels to another box, or sources restored by a checkout, must neither
    be trusted when something differs nor rebuilt when nothing does.)"""
    import hashlib
    h = hashlib.sha256()
    try:
        ver = subprocess.run([tool, "--version"], capture_output=True, text=True).stdout
    except OSError:
        ver = "missing"
    h.update(ver.encode())
    # paths relative to the checkout: the same tree under another root is the same build
    h.update("\0".join(str(f).replace(str(ROOT), ".") for f in flags).encode())
    for d in sorted(Path(x).resolve() for x in deps):
        h.update(str(d.relative_to(ROOT)).encode())
        h.update(d.read_bytes())
    return h.hexdigest()


def _up_to_date(target: Path, fp: str) -> bool:
    stamp = target.with_suffix(target.suffix + ".stamp")
    return target.exists() and stamp.exists() and stamp.read_text().strip() == fp


def _write_stamp(target: Path, fp: str):
    target.with_suffix(target.suffix + ".stamp").write_text(fp + "\n")


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
    nvcc = os.environ.get("NVCC", "nvcc")
    fp = _fingerprint(nvcc, NVCC_FLAGS, deps)
    if not force and _up_to_date(PRODUCT_LIB, fp):
        return PRODUCT_LIB
    LIBDIR.mkdir(parents=True, exist_ok=True)
    cmd = [nvcc, *NVCC_FLAGS]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", str(PRODUCT_LIB), *map(str, srcs)]
    out = _run(cmd)
    _write_stamp(PRODUCT_LIB, fp)
    if verbose:
        print(out)
    return PRODUCT_LIB


def oracle_sources():
    return sorted(ORACLE_DIR.glob("*.cc"))


def build_oracle(force: bool = False) -> Path:
    srcs = oracle_sources()
    deps = srcs + list(ORACLE_DIR.glob("*.h")) + list((ROOT / "include").glob("*.h"))
    flags = ["-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-pthread", "-I", str(ROOT / "include")]
    fp = _fingerprint("g++", flags, deps)
    if not force and _up_to_date(ORACLE_LIB, fp):
        return ORACLE_LIB
    ORACLE_LIB.parent.mkdir(parents=True, exist_ok=True)
    _run(["g++", *flags, "-o", str(ORACLE_LIB), *map(str, srcs)])
    _write_stamp(ORACLE_LIB, fp)
    return ORACLE_LIB


def build_all(force: bool = False, verbose: bool = False):
    return build_product(force, verbose), build_oracle(force)


if __name__ == "__main__":
    p, o = build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
    print(o)
