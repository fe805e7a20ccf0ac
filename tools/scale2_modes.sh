run() { # name, env...
  name=$1; shift
  env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --light > gpurun_out/s2_$name.json 2> gpurun_out/s2_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/s2_$name.json")); print("$name", d["value"], d["ms_per_step"], d["config"]["launch"][:20])
except Exception as e: print("$name failed", e)
PY
}
run arena PN_BENCH_ALLREDUCE=arena
run ov0 PN_BENCH_ALLREDUCE=overlap
run ov16 PN_BENCH_ALLREDUCE=overlap PN_BENCH_COMM_SMS=16
run ov32 PN_BENCH_ALLREDUCE=overlap PN_BENCH_COMM_SMS=32
