# N-GPU check of the overlapped sparse exchange: correctness (2 tests) + weak-scaling step time with and without it
N=${1:-2}
timeout 400 python -m pytest tests/test_gpu_multirank.py -q -x -k "sparse_overlap" 2>&1 | tail -6
for ov in 0 1; do for rs in 16; do
  PN_OVERLAP_EXCHANGE=$ov PN_OVERLAP_RESERVE_SMS=$rs PN_BENCH_EXCHANGE=sparse timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus $N --steps 20 --warmup 5 --light 2> gpurun_out/ov_err.log | tail -1 | python -c "import sys,json; l=sys.stdin.read().strip(); d=json.loads(l) if l else {}; print('overlap=$ov reserve=$rs N=$N', d.get('ms_per_step'), d.get('value'))"
  tail -3 gpurun_out/ov_err.log | grep -i "error\|Traceback"
done; done
