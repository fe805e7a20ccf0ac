#!/bin/bash
# One-GPU evidence run for profiles/r2_*: GPU test suite, bench lines of every config, reference arm, ncu launch list and one
# `ncu --set full` capture of the decoder kernels (each ncu pass only after the same command exited 0 without ncu).
# usage (on the GPU box): tools/r2_evidence.sh [tag]     -> gpurun_out/<tag>/*
TAG=${1:-r2}; O=gpurun_out/$TAG; mkdir -p $O
NCU_EXTRA="lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,lts__t_sectors.sum"
set -x
if [[ -z "${SKIP_TESTS:-}" ]]; then timeout 900 python -m pytest tests -m gpu -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log; tail -3 $O/tests.log; fi
timeout 300 python bench.py > $O/bench_mapping.json 2> $O/bench_mapping.err; echo "rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "rc=$?"
for c in tracking dense mesh256 imap knn; do
  timeout 300 python bench.py --config $c > $O/bench_$c.json 2> $O/bench_$c.err; echo "$c rc=$?"
done
if [[ -z "${SKIP_NCU:-}" ]]; then
timeout 200 python bench.py --light --no-graph --steps 2 --warmup 3 > $O/light.json 2> $O/light.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_mapping.csv \
  python bench.py --light --no-graph --steps 2 --warmup 3 > $O/ncu_launches.log 2>&1
timeout 900 ncu --set full --metrics $NCU_EXTRA --clock-control none --import-source on -k regex:'k_grid_mlp_|k_wgrad_tc' --launch-skip 21 -c 7 \
  -o $O/mapping_full -f python bench.py --light --no-graph --steps 2 --warmup 3 > $O/ncu_full.log 2>&1
timeout 200 python bench.py --config imap --light --no-graph --steps 1 --warmup 3 > $O/light_imap.json 2> $O/light_imap.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_imap.csv \
  python bench.py --config imap --light --no-graph --steps 1 --warmup 3 > $O/ncu_launches_imap.log 2>&1
fi
ls -la $O
