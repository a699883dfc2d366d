#!/bin/bash
# Round-2 development check on one B200 (run under gpurun): parity of the engine-level
# suite, bench digests of the quad kernel at several CTA sizes / bucket widths, phase
# counters, and (with NCU=1) one `ncu --set full` capture of the kernel.
# Every step has a tight timeout and the script stops at the first failure: a hung kernel
# must not eat the GPU budget.
mkdir -p gpurun_out
tag=${1:-r2b}
log=gpurun_out/${tag}.log
: > $log
step() { echo "== $1" >> $log; shift; timeout "$@" >> $log 2>&1; rc=$?; if [ $rc -ne 0 ]; then echo "STEP FAILED rc=$rc" >> $log; tail -40 $log; exit 1; fi; }
step parity 200 python -m pytest tests/test_engine_gpu.py -x -q -m gpu
for T in ${TS:-384 512 256}; do
  HSPF_QUAD_T=$T step "bench T=$T" 120 python scripts/bench_brief.py --steps 10 --warmup 3 --no-cpu-baseline
done
for D in ${DELTAS:-128 512}; do
  HSPF_QUAD_T=${TBEST:-512} step "bench T=${TBEST:-512} delta=$D" 120 python scripts/bench_brief.py --steps 10 --warmup 3 --no-cpu-baseline --delta $D
done
for SUB in ${SUBS:-}; do
  HSPF_QUAD_SUB=$SUB HSPF_QUAD_T=${TBEST:-512} step "bench T=${TBEST:-512} sub_rounds=$SUB delta=${DBEST:-0}" 120 python scripts/bench_brief.py --steps 10 --warmup 3 --no-cpu-baseline --delta ${DBEST:-0}
done
for J in ${JOBSWEEP:-}; do
  HSPF_QUAD_T=${TBEST:-512} step "bench T=${TBEST:-512} jobs=$J" 120 python scripts/bench_brief.py --steps 10 --warmup 3 --no-cpu-baseline --jobs $J
done
for T in ${TS:-384 512}; do HSPF_QUAD_T=$T step "phase profile T=$T" 100 python scripts/quad_profile.py; done
if [ -n "$NCU" ]; then
  HSPF_QUAD_T=${TBEST:-512} step ncu 400 ncu --set full --clock-control none --import-source on -k regex:spf_quad_kernel -s 3 -c 1 \
      -f -o gpurun_out/${tag}_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline
fi
tail -70 $log
