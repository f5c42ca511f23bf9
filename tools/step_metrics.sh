#!/bin/bash
# per-launch metrics (duration, warp instructions, DRAM bytes, occupancy) of the steady-state frames of the kf500 probe, B = 16
COUNT=${1:-1600}; TAG=${2:-r2}
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,lts__t_sector_hit_rate.pct \
  --clock-control none --profile-from-start off -c $COUNT --csv --log-file gpurun_out/${TAG}_step_metrics_ncu.csv \
  python tools/kf500_gpu_probe.py 16 500 33 C --ncu-range > gpurun_out/${TAG}_step_metrics.log 2>&1
echo rc=$?
