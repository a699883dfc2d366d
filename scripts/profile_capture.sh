#!/bin/bash
# Round profile capture on one B200 (run under gpurun).  $1 = tag (e.g. r1).
#   1. default bench line                          -> gpurun_out/<tag>_bench.json
#   2. ncu launch list of the same command         -> gpurun_out/<tag>_launches.csv
#   3. one `ncu --set full` capture of the kernel  -> gpurun_out/<tag>_full.ncu-rep
tag=${1:-r1}
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -c 600 gpurun_out/${tag}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${tag}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spf_batch_kernel -s 3 -c 1 \
    -f -o gpurun_out/${tag}_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${tag}_ncu_full.log 2>&1
ls -la gpurun_out | tail -8
