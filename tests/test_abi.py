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
