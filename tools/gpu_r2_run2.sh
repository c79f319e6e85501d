#!/bin/bash
# GPU session: all GPU tests (no -x) with parity prints + headline bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m overlapnet_b200.build > gpurun_out/r2_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_pytest.log
tail -15 gpurun_out/r2_pytest.log
grep "\[parity\]" gpurun_out/r2_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 ${BENCH_ARGS} > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
echo "bench exit $?"; tail -3 gpurun_out/r2_bench2.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_bench2.json'))
print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e']['value'], d['parity_check'], d['roofline']['share_of_step'], d['roofline']['frac_burst'])
PY
