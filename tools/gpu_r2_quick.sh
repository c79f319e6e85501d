#!/bin/bash
# quick iteration: rebuild, a few parity tests, headline bench without extras
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m overlapnet_b200.build > gpurun_out/r2_build.log 2>&1 || cat gpurun_out/r2_build.log | tail -20
timeout 600 python -m pytest tests/test_gpu_network.py tests/test_gpu_errors.py -m gpu -q -x -k "${TESTS:-pair or heads_match or full_size or timeout}" > gpurun_out/r2_pytest_quick.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/r2_pytest_quick.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/r2_bench_quick.json 2> gpurun_out/r2_bench_quick.err
echo "bench exit $?"; tail -3 gpurun_out/r2_bench_quick.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_bench_quick.json'))
sh = d['roofline']['share_of_step']; ms = d['ms_per_step']
print({k: d[k] for k in ('value', 'ms_per_step')}, 'e2e', round(d['e2e']['value']), 'frac_burst', round(d['roofline']['frac_burst'], 4))
print({k: round(v * ms, 4) for k, v in sh.items()})
PY
