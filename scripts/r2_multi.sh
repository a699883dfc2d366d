#!/bin/bash
# Round-2 multi-GPU check (run under `gpurun --gpus N`): exchange self-test, then the bench with the
# peer exchange (vertex planes and all planes) and with NCCL.  $1 = N.  Tight timeouts everywhere.
N=${1:-2}
mkdir -p gpurun_out
log=gpurun_out/r2_n${N}.log
: > $log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== pytest exchange" >> $log
[ -z "$NOTEST" ] && timeout 300 python -m pytest tests/test_xchg_gpu.py -x -q -m gpu 2>&1 | tail -4 >> $log
echo "== selftest" >> $log
timeout 150 $TR --master-port 29533 scripts/xchg_selftest.py --steps 40 --slot-mb 8 2>&1 | grep -v "OMP_NUM_THREADS\|^\*\*\*" | tail -5 >> $log
run() {  # name, extra args
  name=$1; shift
  echo "== bench $name" >> $log
  timeout 200 $TR --master-port 295$((40 + RANDOM % 50)) bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-e2e-variants "$@" \
      > gpurun_out/r2_n${N}_$name.json 2> gpurun_out/r2_n${N}_$name.err || { echo "FAILED rc=$?" >> $log; tail -8 gpurun_out/r2_n${N}_$name.err >> $log; }
}
run fused
[ -z "$QUICK" ] && run p2p_vertex --exchange p2p
[ -n "$FULL" ] && run p2p_all --exchange p2p --xchg-planes all
[ -n "$FULL" ] && run nccl_vertex --exchange nccl
python - >> $log <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_n${N}_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "kernel", round(d["roofline"]["kernel_ms"], 3),
              "e2e", round(d["e2e"]["value"]), "|", d["config"].get("exchange_bytes", "")[:60])
    except Exception as e:
        print(f, "ERR", e)
PY
cat $log
