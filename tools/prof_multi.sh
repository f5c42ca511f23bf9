#!/bin/bash
# ncu --set full captures (with source) of several kernels of the kf500 probe (B=16): tools/prof_multi.sh <tag> <regex> [skip] [count]
TAG=$1; K=$2; SKIP=${3:-40}; CNT=${4:-8}
ncu --set full --clock-control none --import-source on -k regex:"$K" -s $SKIP -c $CNT -o gpurun_out/prof_$TAG -f \
  python tools/kf500_gpu_probe.py ${PROBE_ARGS:-16 40 9 C} > gpurun_out/prof_$TAG.log 2>&1
echo rc=$?
