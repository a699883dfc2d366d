#!/bin/bash
# Round-2 development check on one B200 (run under gpurun): micro-benchmark, parity of the
# engine-level suite, then bench digests of the old kernel and the quad kernel at several CTA sizes.
mkdir -p gpurun_out
tag=${1:-r2a}
{
timeout 120 scripts/ubench/smem_ops_bench
echo "== parity"
timeout 1500 python -m pytest tests/test_engine_gpu.py -x -q -m gpu 2>&1 | tail -8
echo "== bench digests"
HSPF_NO_QUAD=1 timeout 300 python scripts/bench_brief.py --steps 10 --warmup 3 --no-cpu-baseline
for T in 384 256 512 128; do
  HSPF_QUAD_T=$T timeout 300 python scripts/bench_brief.py --steps 10 --warmup 3 --no-cpu-baseline
done
echo "== phase profile (T=384)"
timeout 300 python scripts/quad_profile.py
HSPF_QUAD_T=256 timeout 300 python scripts/quad_profile.py
} > gpurun_out/${tag}.log 2>&1
tail -60 gpurun_out/${tag}.log
