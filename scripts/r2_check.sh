#!/bin/bash
# Round-2 development check on one B200 (run under gpurun): parity of the engine-level
# suite, bench digests of the quad kernel at several CTA sizes / bucket widths, phase
# counters, and (with NCU=1) one `ncu --set full` capture of the kernel.
mkdir -p gpurun_out
tag=${1:-r2b}
{
echo "== parity"
timeout 1500 python -m pytest tests/test_engine_gpu.py -x -q -m gpu 2>&1 | tail -8
echo "== bench digests"
for T in ${TS:-384 512 256}; do
  echo "T=$T"; HSPF_QUAD_T=$T timeout 300 python scripts/bench_brief.py --steps 10 --warmup 3 --no-cpu-baseline
done
for D in ${DELTAS:-128 512}; do
  echo "T=${TBEST:-512} delta=$D"; HSPF_QUAD_T=${TBEST:-512} timeout 300 python scripts/bench_brief.py --steps 10 --warmup 3 --no-cpu-baseline --delta $D
done
echo "== phase profile"
for T in ${TS:-384 512}; do echo "T=$T"; HSPF_QUAD_T=$T timeout 300 python scripts/quad_profile.py; done
if [ -n "$NCU" ]; then
  echo "== ncu"
  HSPF_QUAD_T=${TBEST:-512} timeout 900 ncu --set full --clock-control none --import-source on -k regex:spf_quad_kernel -s 3 -c 1 \
      -f -o gpurun_out/${tag}_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1
  tail -3 gpurun_out/${tag}_ncu_full.log
fi
} > gpurun_out/${tag}.log 2>&1
tail -70 gpurun_out/${tag}.log
