#!/usr/bin/env python3
"""One batched leg call per input width (for an ncu launch list): python tools/leg_once.py [batch]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from overlapnet_b200 import synth
from overlapnet_b200.engine import Engine
import bench

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for channels, use in ((4, {}), (25, {'use_intensity': True, 'use_class_probabilities': True})):
  eng = Engine(use=use, model=bench.MODEL, precision='f16_tc', max_batch_scans=batch, max_batch_pairs=1)
  eng.load_weights(bench.make_weights(channels))
  x = torch.from_numpy(synth.range_like_images(5, 8, channels)).to(eng.device).repeat(batch // 8, 1, 1, 1)
  eng.leg(x)
  torch.cuda.synchronize()
  eng.close()
