#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/leg1.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from overlapnet_b200 import synth
from overlapnet_b200.engine import Engine
from oracle import network as N
MODEL = {'additional_unsymmetric_layer3a': True, 'strides_layer1': [2, 2]}
eng = Engine(model=MODEL, precision='f16_tc', max_batch_scans=4, max_batch_pairs=16)
eng.load_weights(N.glorot_weights(4, MODEL, seed=0))
x = torch.from_numpy(synth.range_like_images(1, 1, 4)).cuda()
for _ in range(3): eng.leg(x)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_stream_tc -s 25 -c 1 -o gpurun_out/prof_leg python /tmp/leg1.py > gpurun_out/ncu_leg.log 2>&1
echo done
