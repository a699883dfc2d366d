#!/bin/bash
# 2-GPU check of the peer exchange: self-test, then (unless $1 = selftest) the bench with both modes.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29533 scripts/xchg_selftest.py --steps 40 --slot-mb 8 > gpurun_out/n2_selftest.log 2>&1
grep -v "OMP_NUM_THREADS\|^\*\*\*" gpurun_out/n2_selftest.log | tail -25
[ "$1" = "selftest" ] && exit 0
timeout 300 $TR --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/n2_p2p.json 2> gpurun_out/n2_p2p.err
timeout 300 $TR --master-port 29535 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --exchange nccl > gpurun_out/n2_nccl.json 2> gpurun_out/n2_nccl.err
python - <<'PY'
import json
for f in ("n2_p2p", "n2_nccl"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "kernel", round(d["roofline"]["kernel_ms"], 3), "|", d["config"]["exchange"][:40])
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
