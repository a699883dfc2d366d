"""holo_b200 — B200-native batched SPF engine for holo's OSPF/IS-IS hot path.

Only what the path needs: `csrc/` (CUDA kernels + C ABI + LSDB flatteners),
`capi` (ctypes twin of the C ABI), `synth` (seeded synthetic LSDBs of the
BASELINE.json shapes).  See DESIGN.md.
"""
__version__ = "0.1.0"
