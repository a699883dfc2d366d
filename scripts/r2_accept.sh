#!/bin/bash
# Lean acceptance on one B200 (under gpurun): whole -m gpu suite, smoke(), the default bench line, the ncu
# launch list of the same command.  Most important first: the GPU budget may cut the tail.
mkdir -p gpurun_out
tag=${1:-r2accept}
log=gpurun_out/${tag}.log
: > $log
step() { echo "== $1" >> $log; shift; timeout "$@" >> $log 2>&1; rc=$?; if [ $rc -ne 0 ]; then echo "STEP FAILED rc=$rc" >> $log; fi; }
step "pytest -m gpu" 400 python -m pytest tests -q -m gpu
step "smoke" 120 python __graft_entry__.py smoke
echo "== default bench" >> $log
timeout 300 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 300 gpurun_out/${tag}_bench.json >> $log
step "ncu launch list" 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e-variants
grep -v "^==PROF\|^==WARN\|^{" $log | tail -30
