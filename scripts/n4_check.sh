#!/bin/bash
# N-GPU check of the peer exchange ($1 = N): self-test, then the bench (p2p).
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29533 scripts/xchg_selftest.py --steps 30 --slot-mb 8 > gpurun_out/n${N}_selftest.log 2>&1
grep "xchg_selftest\|Error\|error" gpurun_out/n${N}_selftest.log | tail -5
timeout 300 $TR --master-port 29534 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/n${N}_p2p.json 2> gpurun_out/n${N}_p2p.err
python - <<PY
import json
f = "n${N}_p2p"
try:
    d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
    print(f, "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "kernel", round(d["roofline"]["kernel_ms"], 3), "|", d["config"]["exchange"][:40])
except Exception as e:
    print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
