#!/bin/bash
# GPU-side check used during development: parity tests, then a short delta sweep.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_ospfv2_gpu.py -x -q -m gpu 2>&1 | tail -6
for d in "$@"; do
  timeout 200 python scripts/bench_brief.py --steps 10 --warmup 3 --no-cpu-baseline --delta $d
done
