#!/bin/bash
# bench.py on N GPUs under every gradient-exchange mode (weak and strong scaling); one JSON line per run in gpurun_out/scale_N.jsonl
# usage: tools/scale_modes.sh N "mode1 mode2 ..." "weak strong"
N=${1:-2}; MODES=${2:-"sparse sparse_p2p overlap"}; SCALINGS=${3:-"weak strong"}
mkdir -p gpurun_out
for sc in $SCALINGS; do for m in $MODES; do
  echo "== N=$N exchange=$m scaling=$sc" >&2
  PN_BENCH_EXCHANGE=$m timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus $N --steps 20 --warmup 5 --light --scaling $sc 2> gpurun_out/scale_err.log | tail -1 | \
    python -c "import sys,json; l=sys.stdin.read().strip(); d=json.loads(l) if l else {}; print(json.dumps({'n':$N,'exchange':'$m','scaling':'$sc','ms_per_step':d.get('ms_per_step'),'value':d.get('value'),'bytes_per_rank':d.get('config',{}).get('exchange_bytes_per_rank'),'touched':d.get('config',{}).get('touched_rows_this_rank'),'kernel_ms':(d.get('roofline') or {}).get('kernel_ms')}))" | tee -a gpurun_out/scale_$N.jsonl
  tail -3 gpurun_out/scale_err.log | grep -i "error\|Traceback" >&2
done; done
