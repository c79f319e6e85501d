#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m overlapnet_b200.build > gpurun_out/r2_build.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${KERNEL:-k_conv3_pair_tc}" -s ${SKIP:-2} -c 1 \
  -o gpurun_out/r2_prof_one -f python bench.py --steps 1 --warmup 1 --no-extras > gpurun_out/r2_ncu_one.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/r2_ncu_one.log
