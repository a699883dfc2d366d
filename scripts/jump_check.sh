#!/bin/bash
# GPU-side check: full parity suite, phase profile, short bench (+ optional delta sweep).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 100 python scripts/phase_profile.py 0
timeout 200 python scripts/bench_brief.py --steps 10 --warmup 3 --no-cpu-baseline
for d in "$@"; do
  timeout 200 python scripts/bench_brief.py --steps 10 --warmup 3 --no-cpu-baseline --delta $d
done
