#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m overlapnet_b200.build > gpurun_out/r2_build.log 2>&1
cat > /tmp/one_leg.py <<'PY'
import sys; sys.path.insert(0, '.')
import torch, bench
from overlapnet_b200 import synth
from overlapnet_b200.engine import Engine
eng = Engine(model=bench.MODEL, precision='f16_tc', max_batch_scans=64, max_batch_pairs=1)
eng.load_weights(bench.make_weights(4))
x = torch.from_numpy(synth.range_like_images(5, 8, 4)).to(eng.device).repeat(8, 1, 1, 1)
for _ in range(3): eng.leg(x)
torch.cuda.synchronize()
PY
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,dram__bytes.sum --clock-control none -s 24 -c 12 --csv --log-file gpurun_out/r2_leg_launches.csv python /tmp/one_leg.py > gpurun_out/r2_leg_ncu.log 2>&1
echo "ncu exit $?"
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/r2_leg_launches.csv')))
hdr = None
for r in rows:
  if r and r[0] == 'ID': hdr = r; continue
  if hdr and len(r) == len(hdr):
    d = dict(zip(hdr, r))
    print(d['ID'], d['Kernel Name'][:44], d['Grid Size'], d['Metric Name'][:40], d['Metric Value'])
PY
