#!/bin/bash
# usage: scale_modes.sh N "name ENV=... ENV=..." ...   (developer probe: bench.py --light under different exchange settings)
N=$1; shift
for spec in "$@"; do
  set -- $spec; name=$1; shift
  env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --light > gpurun_out/s${N}_$name.json 2> gpurun_out/s${N}_$name.err
  python - <<PY
import json
try:
    s=open("gpurun_out/s${N}_$name.json").read(); d=json.loads(s[s.index("{"):]); print("N=$N $name", d["value"], d["ms_per_step"])
except Exception as e: print("$name failed", e)
PY
done
