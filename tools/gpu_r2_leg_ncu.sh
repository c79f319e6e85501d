#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none --csv --log-file gpurun_out/r2_leg_launches.csv python tools/leg_once.py 64 > gpurun_out/r2_leg_ncu.log 2>&1
echo "ncu exit $?"; grep -c k_leg gpurun_out/r2_leg_launches.csv
