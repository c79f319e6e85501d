#!/usr/bin/env python3
"""A second, independent float64 restatement of both heads for ONE full-size pair (360 x 128 volumes),
written with NumPy einsum / explicit index arithmetic only (no torch, no conv primitive), straight from
the reference's layer definitions:

  DeltaLayer            generateNet.py:45-59    d[i, j, c] = |L[i, c] - R[j, c]|
  c_conv1 (1x15, s 15)  generateNet.py:96-100   linear
  c_conv2 (15x1, s 15)  generateNet.py:102-106  ReLU
  c_conv3 (3x3)         generateNet.py:108-110  ReLU
  Flatten + Dense(1)    generateNet.py:112-114  sigmoid
  RangePadding2D + NormalizedCorrelation2D(normalize='none')   RangePadding2D.py:31-38,
                        NormalizedCorrelation2D.py:96-109      corr[k] = sum_j <L[(k + j + 180) % 360], R[j]>

The network oracle (oracle/network.py) stays "parity unpinned" (no TensorFlow / weights offline); this
fixture only pins it against regressions with an independently written second opinion.
Writes tests/golden/heads_pair_einsum.npz (inputs are re-generated from seeds by the test).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from overlapnet_b200 import synth, weights as W  # noqa: E402

MODEL = {'additional_unsymmetric_layer3a': True, 'strides_layer1': [2, 2]}
SEED_W, SEED_FV = 21, 33


def inputs():
  w = W.glorot_init(4, MODEL, seed=SEED_W)
  rng = np.random.default_rng(SEED_W + 1)
  w = {k: (kern, rng.uniform(-0.05, 0.05, b.shape).astype(np.float32)) for k, (kern, b) in w.items()}
  fv = synth.feature_volumes(SEED_FV, 2)[:, 0]
  return w, fv[0], fv[1]


def heads_einsum(L, R, w):
  L = L.astype(np.float64)
  R = R.astype(np.float64)
  k1, b1 = [a.astype(np.float64) for a in w['c_conv1']]      # (1, 15, 128, 64)
  k2, b2 = [a.astype(np.float64) for a in w['c_conv2']]      # (15, 1, 64, 128)
  k3, b3 = [a.astype(np.float64) for a in w['c_conv3']]      # (3, 3, 128, 256)
  kd, bd = [a.astype(np.float64) for a in w['overlap_output']]
  d = np.abs(L[:, None, :] - R[None, :, :])                   # (360 i, 360 j, 128 c): rows = LEFT pixels
  # c_conv1: kernel (1, 15) slides over j with stride 15
  o1 = np.einsum('ibdc,dco->ibo', d.reshape(360, 24, 15, 128), k1[0]) + b1          # (360, 24, 64)
  # c_conv2: kernel (15, 1) slides over i with stride 15
  o2 = np.maximum(np.einsum('adbo,don->abn', o1.reshape(24, 15, 24, 64), k2[:, 0]) + b2, 0)   # (24, 24, 128)
  # c_conv3: 3x3 valid
  o3 = np.zeros((22, 22, 256))
  for dy in range(3):
    for dx in range(3):
      o3 += np.einsum('yxc,cn->yxn', o2[dy:dy + 22, dx:dx + 22, :], k3[dy, dx])
  o3 = np.maximum(o3 + b3, 0)
  z = float(o3.reshape(-1) @ kd[:, 0] + bd[0])                 # Flatten is row-major over (H, W, C)
  overlap = 1.0 / (1.0 + np.exp(-z))
  G = L @ R.T                                                   # (360 rows of L, 360 rows of R)
  j = np.arange(360)
  corr = np.array([G[(k + j + 180) % 360, j].sum() for k in range(360)])
  return {'o1_sum': o1.sum(), 'o1_abs': np.abs(o1).sum(), 'o2_sum': o2.sum(), 'o3_sum': o3.sum(),
          'o1_probe': o1[[0, 17, 359], [0, 5, 23], [0, 31, 63]], 'o2_probe': o2[[0, 11, 23], [0, 7, 23], [0, 64, 127]],
          'logit': z, 'overlap': overlap, 'corr': corr, 'yaw': 180 - int(np.argmax(corr))}


def main():
  w, L, R = inputs()
  a = heads_einsum(L, R, w)            # LEFT = volume 0, RIGHT = volume 1
  b = heads_einsum(R, L, w)            # the head is not symmetric: the swapped pair is a second vector
  out = os.path.join(ROOT, 'tests', 'golden', 'heads_pair_einsum.npz')
  np.savez(out, **{'lr_' + k: v for k, v in a.items()}, **{'rl_' + k: v for k, v in b.items()})
  print('wrote', out, 'logits', a['logit'], b['logit'], 'yaw', a['yaw'], b['yaw'])


if __name__ == '__main__':
  main()
