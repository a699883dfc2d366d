"""CPU: the C-ABI library loads, exports every declared symbol, and the Python
struct layouts match the C ones.  No compute call is made (no GPU here)."""
import ctypes
import re
from pathlib import Path

import pytest

from holo_b200 import capi, ospfv2

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    syms = set()
    for h in ("holo_spf.h", "holo_spf_lsdb.h"):
        text = (ROOT / "include" / h).read_text()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        for m in re.finditer(r"\b(hspf_[a-z0-9_]+)\s*\(", text):
            syms.add(m.group(1))
    return sorted(syms)


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(str(built[0]))
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/*.h but not exported"
    for s in capi.EXPORTS:
        assert s in syms


def test_struct_layouts_match(built):
    assert ospfv2.abi_sizes_from_library() == ospfv2.abi_sizes_expected()


def test_ctx_create_fails_loudly_without_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.HspfError):
        capi.Context(0)


def test_atom_decode_host_only(built):
    from holo_b200 import synth
    t = synth.random_topology(50, 220, synth.SEED_BASE + 20, lan_fraction=0.2)
    csr = synth.topology_csr(t)
    L = len(t.lans)
    root = t.lans[0][0][0] + L
    n = capi.atom_count(csr, root)
    deg = int(csr.row_ptr[root + 1] - csr.row_ptr[root])
    assert n > deg
    seen = set()
    for a in range(n):
        tail, e = capi.atom_decode(csr, root, a)
        assert csr.row_ptr[tail] <= e < csr.row_ptr[tail + 1]
        assert (tail == root) == (a < deg)
        seen.add((tail, e))
    assert len(seen) == n
    with pytest.raises(capi.HspfError):
        capi.atom_decode(csr, root, n)


def test_build_fingerprint_covers_flags_sources_and_not_the_checkout_root(tmp_path, built):
    """A prebuilt library is reused only when compiler, flags and every source byte are the ones it was built
    from (holo_b200/build.py); the same tree under another root is the same build."""
    from holo_b200 import build
    src = sorted(build.CSRC.glob("*.h"))[:2]
    a = build._fingerprint("g++", ["-O2", "-I", str(build.ROOT / "include")], src)
    assert a == build._fingerprint("g++", ["-O2", "-I", str(build.ROOT / "include")], src)
    assert a != build._fingerprint("g++", ["-O3", "-I", str(build.ROOT / "include")], src)
    assert a != build._fingerprint("g++", ["-O2", "-I", str(build.ROOT / "include")], src[:1])
    assert a != build._fingerprint("nvcc", ["-O2", "-I", str(build.ROOT / "include")], src)
    stamp = built[0].with_suffix(".so.stamp")
    assert stamp.exists() and build._up_to_date(built[0], stamp.read_text().strip())
    assert not build._up_to_date(built[0], "0" * 64)


def test_integration_doc_names_only_declared_entry_points():
    """Every `hspf_*` function INTEGRATION.md binds or calls is declared in include/*.h (and so exported)."""
    text = (ROOT / "INTEGRATION.md").read_text()
    named = set(re.findall(r"\b(hspf_[a-z0-9_]+)\s*\(", text))
    declared = set(declared_symbols())
    # shorthand the text uses for families of calls
    named = {n for n in named if not n.endswith("_")}
    missing = sorted(n for n in named if n not in declared)
    assert not missing, missing
    assert len(named) > 30
