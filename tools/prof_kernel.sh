#!/bin/bash
# ncu --set full capture of one kernel of the kf500 probe (B=16): tools/prof_kernel.sh <kernel regex> <tag> [skip]
K=$1; TAG=$2; SKIP=${3:-20}
ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 2 -o gpurun_out/prof_$TAG -f \
  python tools/kf500_gpu_probe.py 16 40 9 C > gpurun_out/prof_$TAG.log 2>&1
echo rc=$?
