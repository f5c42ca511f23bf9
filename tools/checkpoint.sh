#!/bin/bash
# Round checkpoint on the GPU box: parity tests, bench lines (both arms, both map modes), ncu launch list, per-kernel
# metrics and one full capture of the kernels changed last.  usage: tools/checkpoint.sh <tag>   (outputs: gpurun_out/<tag>_*)
tag=${1:-r1_x}
out=gpurun_out
set -x
timeout 300 python -m pytest tests -m gpu -x -q > $out/${tag}_gpu_tests.log 2>&1; echo rc=$? >> $out/${tag}_gpu_tests.log
timeout 300 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo rc=$?
timeout 300 python bench.py --map live > $out/${tag}_bench_live.json 2> $out/${tag}_bench_live.err; echo rc=$?
timeout 300 python bench.py --impl reference > $out/${tag}_bench_reference_arm.json 2>> $out/${tag}_bench.err; echo rc=$?
timeout 300 python bench.py --impl reference --map live > $out/${tag}_bench_reference_arm_live.json 2>> $out/${tag}_bench_live.err; echo rc=$?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $out/${tag}_launches_ncu.csv \
  python bench.py --steps 4 --warmup 3 --skip-cpu-baseline > $out/ncu_bench.log 2>&1; echo rc=$?
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,smsp__thread_inst_executed_per_inst_executed.ratio,lts__t_sector_hit_rate.pct \
  --clock-control none --csv --log-file $out/${tag}_step_metrics_ncu.csv python tools/prof_run.py 16 8 > $out/prof_ncu.log 2>&1; echo rc=$?
timeout 240 ncu --set full --clock-control none --import-source on -k regex:'k_map_knn|k_ccl_merge|k_seg_emit' -c 8 -f -o $out/${tag}_full \
  python tools/prof_run.py 16 8 > $out/prof_full.log 2>&1; echo rc=$?
