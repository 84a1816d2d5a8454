#!/bin/bash
# BASELINE configs[3] on N GPUs of one box (run under `gpurun --gpus N`): weak scaling at 4096 and 8192 windows per GPU, strong
# scaling at a fixed global batch of 65536.  Lines go to gpurun_out/.
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
O=gpurun_out/r02e_scale_${N}gpu
$TR --master-port 29601 bench.py --gpus $N --steps 20 --warmup 5 > ${O}_weak4096.json 2> ${O}_weak4096.err
$TR --master-port 29602 bench.py --gpus $N --steps 20 --warmup 5 --batch 8192 > ${O}_weak8192.json 2> ${O}_weak8192.err
$TR --master-port 29603 bench.py --gpus $N --steps 20 --warmup 5 --global-batch 65536 > ${O}_strong65536.json 2> ${O}_strong65536.err
for f in ${O}_*.json; do echo "== $f"; python - "$f" <<'PY'
import json, sys
try:
    l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: l.get(k) for k in ("value", "ms_per_step", "n_gpus", "scaling")}, l.get("config", {}).get("global_batch"))
except Exception as e:
    print("unreadable:", e)
PY
done
