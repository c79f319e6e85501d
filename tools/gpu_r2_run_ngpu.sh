#!/bin/bash
# N-GPU bench (N from $NGPU), symm transport (auto) and optionally the collective fallback
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${NGPU:-8}
python -m overlapnet_b200.build > gpurun_out/r2_build.log 2>&1
for tr in ${TRANSPORTS:-auto}; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps 20 --warmup 3 --transport $tr > gpurun_out/r2_bench_${N}gpu_$tr.json 2> gpurun_out/r2_bench_${N}gpu_$tr.err
echo "bench[$N,$tr] exit $?"; tail -3 gpurun_out/r2_bench_${N}gpu_$tr.err
python - <<PY
import json
try:
  txt = open('gpurun_out/r2_bench_${N}gpu_$tr.json').read()
  d = json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
  print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}, d['e2e']['value'], d['config']['parallelism'])
  print(d.get('bank4541')); print(d.get('all_pairs', {}).get('measured')); print(d.get('all_pairs', {}).get('projected_full_matrix_s'))
except Exception as e:
  print('no json', e)
PY
done
