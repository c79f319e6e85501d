#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 240 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
echo "bench exit $?"; tail -3 gpurun_out/r2_bench2.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_bench2.json'))
print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e']['value'], d['parity_check']['max_abs_overlap_err'], d['roofline']['frac_burst'])
print(d['leg_batch256']); print(d['all_pairs']['measured']); print(d['latency_1pair'])
PY
