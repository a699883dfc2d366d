#!/bin/bash
# one N=8 validation: bench with and without SM reservation for the overlapped all-gather
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --reserve-sms $1 2>gpurun_out/n8_$1.err | grep "^{" | tee gpurun_out/n8_$1.json | python -c "import sys,json; l=json.loads(sys.stdin.read()); print('reserve','$1','value',round(l['value']),'ms/step',round(l['ms_per_step'],3),'kernel',round(l['roofline']['kernel_ms'],3),'e2e',round(l['e2e']['value']))"; }
mkdir -p gpurun_out
run 40
run 0
tail -3 gpurun_out/n8_0.err
