#!/bin/bash
# Round-2 checkpoint on the GPU box: parity tests, bench lines (headline kf500, reference arm, configs A and B), ncu launch list of
# the bench command, per-launch metrics of the steady-state kf500 frames, one full capture.  usage: tools/checkpoint_r2.sh <tag>
tag=${1:-r2_x}
out=gpurun_out
set -x
timeout 400 python -m pytest tests -m gpu -x -q > $out/${tag}_gpu_tests.log 2>&1; echo rc=$? >> $out/${tag}_gpu_tests.log
timeout 400 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo rc=$?
timeout 400 python bench.py --impl reference > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err; echo rc=$?
timeout 300 python bench.py --config A --map live --batch 1 --streams 1 > $out/${tag}_bench_A.json 2> $out/${tag}_bench_A.err; echo rc=$?
timeout 300 python bench.py --config B --map live > $out/${tag}_bench_B.json 2> $out/${tag}_bench_B.err; echo rc=$?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $out/${tag}_launches_ncu.csv \
  python bench.py --steps 2 --warmup 3 --skip-cpu-baseline --skip-latency > $out/${tag}_ncu_bench.log 2>&1; echo rc=$?
timeout 400 tools/step_metrics.sh 1600 ${tag}; echo rc=$?
python tools/step_metrics_summary.py $out/${tag}_step_metrics_ncu.csv 16 $out/${tag}_traffic.json > $out/${tag}_step_metrics_summary.txt 2>&1
timeout 300 tools/prof_multi.sh ${tag}_full "k_feature_ring|k_seg_emit|k_grid_scan|k_map_knn" 30 6; echo rc=$?
