#!/bin/bash
# Device route stage on one B200 (under gpurun): its parity tests and the C5 bench line with `route_stage.device_batch`.
mkdir -p gpurun_out
tag=${1:-r2routes}
log=gpurun_out/${tag}.log
: > $log
step() { echo "== $1" >> $log; shift; timeout "$@" >> $log 2>&1; rc=$?; if [ $rc -ne 0 ]; then echo "STEP FAILED rc=$rc" >> $log; fi; }
step "pytest route stage" 300 python -m pytest tests/test_ospfv2_gpu.py -q -m gpu -k "batch"
step "pytest ospfv2 + engine" 300 python -m pytest tests/test_ospfv2_gpu.py tests/test_engine_gpu.py -q -m gpu -x
echo "== C5 bench" >> $log
timeout 300 python bench.py --config C5 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e-variants > gpurun_out/${tag}_C5.json 2> gpurun_out/${tag}_C5.err
tail -c 1500 gpurun_out/${tag}_C5.json >> $log; tail -5 gpurun_out/${tag}_C5.err >> $log
grep -v "^==PROF\|^==WARN" $log | tail -40
