run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --reserve-sms $2 2>/dev/null | grep "^{" | python -c "import sys,json; l=json.loads(sys.stdin.read()); print('channels','$1','reserve','$2','value',round(l['value']),'ms/step',round(l['ms_per_step'],3),'kernel',round(l['roofline']['kernel_ms'],3))"; }
NCCL_MAX_NCHANNELS=8 run 8 8
NCCL_MAX_NCHANNELS=8 run 8 12
NCCL_MAX_NCHANNELS=16 run 16 16
NCCL_MAX_NCHANNELS=4 run 4 6
run default 40
