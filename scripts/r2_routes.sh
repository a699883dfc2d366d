#!/bin/bash
# Device route stage + in-place cost update on one B200 (under gpurun): parity tests, the C5 bench line with
# `route_stage.device_batch`, one ncu capture of route_cells_kernel.
mkdir -p gpurun_out
tag=${1:-r2routes}
log=gpurun_out/${tag}.log
: > $log
step() { echo "== $1" >> $log; shift; timeout "$@" >> $log 2>&1; rc=$?; if [ $rc -ne 0 ]; then echo "STEP FAILED rc=$rc" >> $log; fi; }
step "pytest update_costs + ospfv2" 300 python -m pytest tests/test_engine_gpu.py tests/test_ospfv2_gpu.py -q -m gpu -k "update_costs or batch or cost_change or run_area"
step "route stage timing" 120 python scripts/route_stage_profile.py 1000
echo "== C5 bench" >> $log
timeout 300 python bench.py --config C5 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e-variants > gpurun_out/${tag}_C5.json 2> gpurun_out/${tag}_C5.err
tail -c 1300 gpurun_out/${tag}_C5.json >> $log; tail -5 gpurun_out/${tag}_C5.err >> $log
step "ncu route kernel" 200 ncu --set full --clock-control none --import-source on -k regex:route_cells_kernel -s 1 -c 1 -f -o gpurun_out/${tag}_cells python scripts/route_stage_profile.py 1000
grep -v "^==PROF\|^==WARN" $log | tail -40
