#!/bin/bash
# 2-GPU session: sharded LCD test, non-current-device test, bench at N=2 (symm transport and collective fallback)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m overlapnet_b200.build > gpurun_out/r2_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_lcd.py tests/test_gpu_errors.py -m gpu -q -s -k "two_gpus or non_current" > gpurun_out/r2_pytest_2gpu.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/r2_pytest_2gpu.log
for tr in ${TRANSPORTS:-auto}; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 20 --warmup 3 --transport $tr > gpurun_out/r2_bench_2gpu_$tr.json 2> gpurun_out/r2_bench_2gpu_$tr.err
echo "bench[$tr] exit $?"; tail -5 gpurun_out/r2_bench_2gpu_$tr.err
python - <<PY
import json
try:
  d = json.load(open('gpurun_out/r2_bench_2gpu_$tr.json'))
  print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}, d['e2e'], d['config']['parallelism'], d.get('bank4541'), d.get('all_pairs', {}).get('measured'))
except Exception as e:
  print('no json', e)
PY
done
