#!/bin/bash
# k-NN (config 4) on one GPU: tests, bench line, ncu launch list and full capture of the two kernels -> gpurun_out/knn/
O=gpurun_out/knn; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_knn.py -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log; tail -40 $O/tests.log
timeout 300 python bench.py --config knn > $O/bench_knn.json 2> $O/bench_knn.err; echo "bench rc=$?"; tail -5 $O/bench_knn.err; cut -c1-1500 $O/bench_knn.json
if [[ -z "${SKIP_NCU:-}" ]]; then
timeout 200 python bench.py --config knn --light --no-graph --steps 2 --warmup 3 > $O/light.json 2> $O/light.err && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_knn.csv \
  python bench.py --config knn --light --no-graph --steps 2 --warmup 3 > $O/ncu_launches.log 2>&1
timeout 600 ncu --set full --metrics lts__t_sectors_op_red.sum,lts__t_sectors.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum \
  --clock-control none --import-source on -k regex:'k_knn_fwd|k_knn_bwd' --launch-skip 6 -c 2 -o $O/knn_full -f \
  python bench.py --config knn --light --no-graph --steps 2 --warmup 3 > $O/ncu_full.log 2>&1
fi
ls -la $O
