#!/bin/bash
# GPU session: leg tests first (falls back to OVN_LEG_WIDE=0 if they fail), leg timing, every GPU test, the full bench
# line, and the ncu launch list of the bench command.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m overlapnet_b200.build > gpurun_out/r2_build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_network.py -m gpu -q -s -k "leg" > gpurun_out/r2_pytest_leg.log 2>&1
rc=$?
echo "leg pytest exit $rc"; grep "\[parity\]" gpurun_out/r2_pytest_leg.log
if [ $rc -ne 0 ]; then tail -30 gpurun_out/r2_pytest_leg.log; export OVN_LEG_WIDE=0; echo "FALLBACK OVN_LEG_WIDE=0"; fi
if [ $rc -eq 0 ]; then timeout 300 python tools/time_leg.py 2 2>&1 | tee gpurun_out/r2_time_leg.log; fi
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_pytest.log
tail -6 gpurun_out/r2_pytest.log
grep "\[parity\]" gpurun_out/r2_pytest.log | sort | uniq
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
echo "bench exit $?"; tail -3 gpurun_out/r2_bench2.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_bench2.json'))
print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e']['value'], d['parity_check'], d['roofline']['share_of_step'], d['roofline']['frac_burst'])
print({k: v for k, v in d.get('extras', {}).items()} if 'extras' in d else [k for k in d])
PY
