#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m overlapnet_b200.build > gpurun_out/r2_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_network.py -m gpu -q -s -k "leg" > gpurun_out/r2_pytest_leg.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/r2_pytest_leg.log; grep "\[parity\]" gpurun_out/r2_pytest_leg.log
timeout 600 python tools/time_leg.py 2 2>&1 | tee gpurun_out/r2_time_leg.log
