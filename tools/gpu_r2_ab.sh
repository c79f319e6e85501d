#!/bin/bash
# A/B on ONE box: default build (tests + bench), then a build with $ALT_FLAGS (bench only)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m overlapnet_b200.build --force > gpurun_out/r2_build.log 2>&1 || tail -20 gpurun_out/r2_build.log
timeout 600 python -m pytest tests/test_gpu_network.py tests/test_gpu_errors.py -m gpu -q -x -k "${TESTS:-pair or heads_match or full_size or timeout}" > gpurun_out/r2_pytest_quick.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/r2_pytest_quick.log
report() { python - "$1" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
sh = d['roofline']['share_of_step']; ms = d['ms_per_step']
print(sys.argv[1], {k: d[k] for k in ('value', 'ms_per_step')}, 'frac_burst', round(d['roofline']['frac_burst'], 4), 'parity', d['parity_check']['max_abs_overlap_err'])
print('   ', {k: round(v * ms, 4) for k, v in sh.items()})
PY
}
for rep in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/r2_bench_A$rep.json 2> gpurun_out/r2_bench_A.err; report gpurun_out/r2_bench_A$rep.json
done
OVN_NVCC_EXTRA="$ALT_FLAGS" python -m overlapnet_b200.build --force > gpurun_out/r2_build_alt.log 2>&1 || tail -20 gpurun_out/r2_build_alt.log
for rep in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/r2_bench_B$rep.json 2> gpurun_out/r2_bench_B.err; report gpurun_out/r2_bench_B$rep.json
done
