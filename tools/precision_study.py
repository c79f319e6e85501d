#!/usr/bin/env python3
"""Precision budget of the tensor-core heads path, emulated on the CPU in float64 with explicit
fp16 roundings at the places the kernels round (tools/, not product; uses the oracle).

  python tools/precision_study.py [--case heads|bank] [--spread 1.5]

For every variant: max |overlap - overlap_ref| over the test pairs with the Dense layer rescaled
to the given logit spread (tests/test_gpu_network.py), and the logit error as a fraction of the
logit standard deviation.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from oracle import network as N  # noqa: E402
from overlapnet_b200 import synth  # noqa: E402

MODEL = {'additional_unsymmetric_layer3a': True, 'strides_layer1': [2, 2]}


def h(x):
  return x.to(torch.float16).to(torch.float64)


def hilo(x):
  hi = x.to(torch.float16).to(torch.float64)
  lo = (x - hi).to(torch.float16).to(torch.float64)
  return hi, lo


def emu_logit(L, R, w, v):
  """L, R: (360,128) float64 tensors.  v: dict of variant switches."""
  W1, b1 = [torch.as_tensor(a, dtype=torch.float64) for a in w['c_conv1']]
  W2, b2 = [torch.as_tensor(a, dtype=torch.float64) for a in w['c_conv2']]
  W3, b3 = [torch.as_tensor(a, dtype=torch.float64) for a in w['c_conv3']]
  Wd, bd = [torch.as_tensor(a, dtype=torch.float64) for a in w['overlap_output']]
  fv = v.get('fv', 'f16')
  if fv == 'center':
    mu = h(0.5 * (L.mean(0) + R.mean(0)))
    L, R = L - mu, R - mu
  if fv == 'exact':
    d = (L[:, None, :] - R[None, :, :])
  elif fv in ('f16', 'center'):
    d = h(h(L)[:, None, :] - h(R)[None, :, :])            # HSUB2 rounds the difference
  elif fv == 'hilo':
    Lh, Ll = hilo(L)
    Rh, Rl = hilo(R)
    d = h(h(Lh[:, None, :] - Rh[None, :, :]) + h(Ll[:, None, :] - Rl[None, :, :]))
  elif fv == 'hilo_fma':          # t = Lh - Rh ; d = (t + Ll) - Rl with two more roundings
    Lh, Ll = hilo(L)
    Rh, Rl = hilo(R)
    d = h(h(h(Lh[:, None, :] - Rh[None, :, :]) + Ll[:, None, :]) - Rl[None, :, :])
  elif fv == 'r_exact':
    Rh, Rl = hilo(R)
    d = h(h(h(L)[:, None, :] - Rh[None, :, :]) - Rl[None, :, :])
  elif fv == 'fp32sub':
    d = (L.to(torch.float32)[:, None, :] - R.to(torch.float32)[None, :, :]).to(torch.float64)
  else:
    raise ValueError(fv)
  d = d.abs()
  if v.get('d16', True):
    d = h(d)
  w1 = h(W1) if v.get('w1_16', True) else W1
  # o1[i, jb, o] = sum_{dj, c} d[i, 15 jb + dj, c] W1[0, dj, c, o]   (bias folded into b2eff)
  dd = d.reshape(360, 24, 15, 128)
  o1 = torch.einsum('ijdc,dco->ijo', dd, w1[0])
  if not v.get('fold_b1', True):
    o1 = o1 + b1
  if v.get('o1_center', False):               # fp16(o1 - mean per channel): c_conv2 is linear, the mean folds into its bias
    mo = h(o1.mean((0, 1)))
    o1 = h(o1 - mo) + mo
  elif v.get('o1_16', True):
    o1 = h(o1)
  w2 = h(W2) if v.get('w2_16', True) else W2
  # x3[ib, jb, n] = relu(b2eff + sum_{di, o} o1[15 ib + di, jb, o] W2[di, 0, o, n])
  oo = o1.reshape(24, 15, 24, 64)
  b2eff = b2 + (torch.einsum('o,don->n', b1, W2[:, 0]) if v.get('fold_b1', True) else 0)
  b2eff = b2eff.to(torch.float32).to(torch.float64)
  x3 = torch.relu(torch.einsum('idjo,don->ijn', oo, w2[:, 0]) + b2eff)
  x3_mean = None
  if v.get('x3_center', False):               # fp16(x3 - mean per channel): folds into the c_conv3 bias
    mx = h(x3.mean((0, 1)))
    if v.get('w3_fold_exact', False):         # the mean's image under c_conv3 is folded with the fp32 weights
      x3_mean = mx
      x3 = (x3 - mx) if v.get('x3_keep_exact', False) else h(x3 - mx)
    else:
      x3 = h(x3 - mx) + mx
  elif v.get('x3_16', True):
    x3 = h(x3)
  elif v.get('x3_hilo', False):
    a, b = hilo(x3)
    x3 = a + b
  w3 = h(W3) if v.get('w3_16', True) else W3
  y = F.conv2d(x3.permute(2, 0, 1)[None], w3.permute(3, 2, 0, 1).contiguous(), b3)
  if x3_mean is not None:
    y = y + torch.einsum('n,yxnm->m', x3_mean, W3)[None, :, None, None]
  y = torch.relu(y)[0].permute(1, 2, 0).reshape(-1)          # Flatten (H, W, C)
  return float(y @ Wd[:, 0] + bd[0])


EX = dict(d16=False, w1_16=False, o1_16=False, w2_16=False, x3_16=False, w3_16=False)
VARIANTS = [
    ('all exact (sanity)', dict(fv='exact', **EX)),
    ('round-1 f16_tc (all fp16)', dict()),
    ('PRODUCTION r2: centres + W2 hi/lo', dict(fv='center', w2_16=False, o1_center=True, x3_center=True, w3_fold_exact=True)),
    ('--- single terms (everything else exact) ---', None),
    ('only fv fp16 (round 1)', dict(fv='f16', **EX)),
    ('only fv fp16 CENTRED', dict(fv='center', **EX)),
    ('only fv hi/lo subtraction', dict(fv='hilo', **EX)),
    ('only |d| fp16', dict(fv='exact', **{**EX, 'd16': True})),
    ('only W1 fp16', dict(fv='exact', **{**EX, 'w1_16': True})),
    ('only o1 fp16', dict(fv='exact', **{**EX, 'o1_16': True})),
    ('only o1 fp16 CENTRED', dict(fv='exact', **{**EX, 'o1_center': True})),
    ('only W2 fp16 (round 1)', dict(fv='exact', **{**EX, 'w2_16': True})),
    ('only x3 fp16', dict(fv='exact', **{**EX, 'x3_16': True})),
    ('only x3 fp16 CENTRED + exact fold', dict(fv='exact', **{**EX, 'x3_center': True, 'w3_fold_exact': True})),
    ('only W3 fp16 (x3 raw)', dict(fv='exact', **{**EX, 'w3_16': True})),
    ('only W3 fp16 (x3 centred, exact fold)', dict(fv='exact', **{**EX, 'w3_16': True, 'x3_center': True, 'w3_fold_exact': True,
                                                            'x3_keep_exact': True})),
    ('--- what more would buy ---', None),
    ('production + fv hi/lo', dict(fv='hilo', w2_16=False, o1_center=True, x3_center=True, w3_fold_exact=True)),
    ('production + x3, o1 exact', dict(fv='center', w2_16=False, o1_16=False, x3_16=False)),
    ('production + W3 exact', dict(fv='center', w2_16=False, o1_center=True, x3_center=True, w3_fold_exact=True, w3_16=False)),
]


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--case', default='heads')
  ap.add_argument('--spread', type=float, default=1.5)
  ap.add_argument('--pairs', type=int, default=8)
  ap.add_argument('--only', default='')
  args = ap.parse_args()
  w = N.glorot_weights(4, MODEL, seed=0)
  if args.case == 'heads':
    x = synth.range_like_images(1234, 6, 4)
    bank = N.leg_forward(x, w, MODEL)[:, 0]
    rng = np.random.default_rng(5)
    bank[1] = np.roll(bank[5], 37, axis=0) + np.abs(rng.normal(0, 0.01, bank[5].shape)).astype(np.float32)
    bank[2] = np.roll(bank[5], -120, axis=0)
    left = np.array([0, 1, 2, 3, 4, 5, 5, 2])[:args.pairs]
    right = np.array([5, 5, 5, 5, 5, 5, 0, 1])[:args.pairs]
  else:
    n = 1101
    bank = synth.feature_volumes(11, n)[:, 0] * np.float32(0.2)
    rs = np.random.default_rng(0)
    left = rs.choice(n, args.pairs, replace=False)
    right = np.full(args.pairs, 17)
  Lb = torch.as_tensor(bank, dtype=torch.float64)
  exact = dict(fv='exact', d16=False, w1_16=False, o1_16=False, w2_16=False, x3_16=False, w3_16=False)
  z_ref = np.array([emu_logit(Lb[a], Lb[b], w, exact) for a, b in zip(left, right)])
  # cross-check the emulation against the oracle proper
  _, _, _, z0 = N.heads_forward(bank[left][:, None], bank[right][:, None], w, MODEL, batch=2, return_logit=True)
  print('emulation vs oracle logits: max abs diff %.3e (logit std %.4f)' % (np.abs(z_ref - z0).max(), z0.std()))
  kd, bd = w['overlap_output']
  raw = z0 - float(bd[0])
  g = args.spread / raw.std()
  med = np.median(raw)
  ov_ref = 1 / (1 + np.exp(-g * (raw - med)))
  print('overlaps (ref) at spread %.2f: %s' % (args.spread, np.round(ov_ref, 3)))
  print('%-28s %12s %12s %12s' % ('variant', 'max|dz|/std', 'rms|dz|/std', 'max|d ov|'))
  for name, v in VARIANTS:
    if v is None:
      print(name)
      continue
    if args.only and args.only not in name:
      continue
    z = np.array([emu_logit(Lb[a], Lb[b], w, v) for a, b in zip(left, right)])
    dz = (z - z_ref)
    ov = 1 / (1 + np.exp(-g * (z - float(bd[0]) - med)))
    print('%-28s %12.3e %12.3e %12.3e' % (name, np.abs(dz).max() / raw.std(), np.sqrt((dz ** 2).mean()) / raw.std(),
                                          np.abs(ov - ov_ref).max()))


if __name__ == '__main__':
  main()
