"""The network oracle is parity-UNPINNED by the reference (no TF/Keras, no weights).  These tests
pin it to what the reference does state: the circular-padding KAT (RangePadding2D.py:5), the
shift property of the correlation head, the layer table / parameter counts of generateNet.py, and
independent loop restatements of every op."""
import numpy as np
import torch

from oracle import network as N

MODEL = {'additional_unsymmetric_layer3a': True, 'strides_layer1': [2, 2]}


def test_range_padding_kat():
  # RangePadding2D.py:5  pad([1 2 3 4], 2) -> [3, 4, 1, 2, 3, 4, 1]
  x = torch.tensor([1., 2., 3., 4.]).reshape(1, 1, 4, 1)
  assert N.range_padding(x, 2).reshape(-1).tolist() == [3, 4, 1, 2, 3, 4, 1]
  # self-demo of RangePadding2D.py:44-92: padding=3 on [1..6] -> width 2W-1 = 11
  y = N.range_padding(torch.arange(1., 7.).reshape(1, 1, 6, 1), 3).reshape(-1).tolist()
  assert y == [4, 5, 6, 1, 2, 3, 4, 5, 6, 1, 2]


def test_leg_shapes_and_parameter_counts():
  """SURVEY 8a layer table: 64x900xC -> 1x360x128; params 1 104 112 (C=4), head 665 025."""
  w = N.glorot_weights(4, MODEL, seed=0)
  leg = sum(w[n][0].size + w[n][1].size for n, _, _, _ in N.leg_layers(MODEL))
  head = sum(w[n][0].size + w[n][1].size for n in ('c_conv1', 'c_conv2', 'c_conv3', 'overlap_output'))
  assert leg == 1104112 and head == 665025
  w25 = N.glorot_weights(25, MODEL, seed=0)
  assert sum(w25[n][0].size + w25[n][1].size for n, _, _, _ in N.leg_layers(MODEL)) == 1129312
  x = np.random.default_rng(0).standard_normal((1, 64, 900, 4)).astype(np.float32)
  acts = N.leg_forward(x, w, MODEL, return_all=True)
  shapes = [a.shape[1:] for a in acts]
  assert shapes == [(30, 443, 16), (14, 429, 32), (6, 415, 64), (2, 404, 64), (1, 396, 128), (1, 388, 128),
                    (1, 380, 128), (1, 372, 128), (1, 366, 128), (1, 362, 128), (1, 360, 128)]


def test_conv_against_naive_loops():
  rng = np.random.default_rng(1)
  x = rng.standard_normal((9, 40, 3))
  k = rng.standard_normal((3, 15, 3, 5)).astype(np.float32)
  b = rng.standard_normal(5).astype(np.float32)
  ref = N.conv2d_valid_naive(x, k, b, (2, 1), True)
  got = N._conv(torch.tensor(x).permute(2, 0, 1)[None], k, b, (2, 1), True, torch.float64)[0].permute(1, 2, 0).numpy()
  assert np.allclose(ref, got, atol=1e-12)


def test_correlation_head_against_naive_and_shift_kat():
  rng = np.random.default_rng(2)
  W, C = 36, 8
  L = np.abs(rng.standard_normal((1, 1, W, C))).astype(np.float32)
  R = np.abs(rng.standard_normal((1, 1, W, C))).astype(np.float32)
  c = N.correlation_head(L, R)[0]
  assert np.allclose(c, N.correlation_naive(L[0, 0], R[0, 0]), atol=1e-10)
  # shift property: R = roll(L, s) along the width  =>  argmax = (W//2 - s) mod W, i.e. yaw = s
  for s in (0, 1, 5, -7, W // 2):
    Rs = np.roll(L, s, axis=2)
    k = int(np.argmax(N.correlation_head(L, Rs)[0]))
    assert k == (W // 2 - s) % W
  # the demo of NormalizedCorrelation2D.py:112-144: ramp vs ramp rolled by +1 (resolution 6)
  img1 = np.arange(6, dtype=np.float32).reshape(1, 1, 6, 1)
  img2 = np.roll(img1, 1, axis=2)
  assert int(np.argmax(N.correlation_head(img1, img2)[0])) == (3 - 1) % 6


def test_delta_head_against_naive():
  rng = np.random.default_rng(3)
  W, C, s = 45, 128, 15
  w = N.glorot_weights(4, MODEL, seed=4)
  # shrink the dense layer to the small geometry: 45 -> (45,3,64) -> (3,3,128) -> (1,1,256)
  w['overlap_output'] = (rng.standard_normal((256, 1)).astype(np.float32) * 0.1, np.array([0.05], np.float32))
  L = np.abs(rng.standard_normal((2, 1, W, C))).astype(np.float32)
  R = np.abs(rng.standard_normal((2, 1, W, C))).astype(np.float32)
  acts, z, o = N.delta_head(L, R, w, MODEL, return_all=True)
  for b in range(2):
    o1 = N.delta_conv1_naive(L[b, 0], R[b, 0], w['c_conv1'][0], w['c_conv1'][1], s)
    assert np.allclose(acts[0][b], o1, atol=1e-10)
    o2 = N.conv2d_valid_naive(o1, w['c_conv2'][0], w['c_conv2'][1], (s, 1), True)
    assert np.allclose(acts[1][b], o2, atol=1e-10)
    o3 = N.conv2d_valid_naive(o2, w['c_conv3'][0], w['c_conv3'][1], (1, 1), True)
    zz = o3.reshape(-1) @ w['overlap_output'][0].astype(np.float64)[:, 0] + w['overlap_output'][1][0]
    assert np.allclose(z[b, 0], zz, atol=1e-10)
    assert np.allclose(o[b, 0], 1 / (1 + np.exp(-zz)), atol=1e-12)
  # the delta layer is NOT symmetric in (L, R): c_conv1 slides over R's columns, c_conv2 over L's
  o_swapped = N.delta_head(R, L, w, MODEL)
  assert not np.allclose(o_swapped, o.astype(np.float32))


def test_readout_first_max_and_range():
  corr = np.zeros((3, 360))
  corr[0, 0] = 1          # yaw 180
  corr[1, 359] = 1        # yaw -179
  corr[2, [10, 200]] = 1  # tie -> first maximum (np.argmax)
  _, yaw = N.readout(np.zeros(3), corr)
  assert yaw.tolist() == [180, -179, 170]


def test_heads_match_independent_einsum_restatement():
  """tests/golden/heads_pair_einsum.npz: one full-size pair (both orderings) through an independent
  NumPy-einsum restatement of the two heads (tools/make_golden_heads.py).  The network oracle stays
  'parity unpinned' by the reference; this catches regressions of oracle/network.py."""
  import os
  import sys
  from conftest import GOLDEN, ROOT
  sys.path.insert(0, os.path.join(ROOT, 'tools'))
  import make_golden_heads as G
  g = np.load(os.path.join(GOLDEN, 'heads_pair_einsum.npz'))
  w, L, R = G.inputs()
  for tag, (l, r) in (('lr_', (L, R)), ('rl_', (R, L))):
    acts, z, ov = N.delta_head(l[None, None], r[None, None], w, G.MODEL, return_all=True)
    o1, o2, o3 = acts[0][0], acts[1][0], acts[2][0]
    assert abs(float(z[0, 0]) - float(g[tag + 'logit'])) <= 1e-10 * max(1.0, abs(float(g[tag + 'logit'])))
    assert abs(float(ov[0, 0]) - float(g[tag + 'overlap'])) <= 1e-12
    for name, val in (('o1_sum', o1.sum()), ('o1_abs', np.abs(o1).sum()), ('o2_sum', o2.sum()), ('o3_sum', o3.sum())):
      assert abs(val - float(g[tag + name])) <= 1e-9 * abs(float(g[tag + name]))
    assert np.allclose(o1[[0, 17, 359], [0, 5, 23], [0, 31, 63]], g[tag + 'o1_probe'], rtol=1e-10, atol=1e-12)
    assert np.allclose(o2[[0, 11, 23], [0, 7, 23], [0, 64, 127]], g[tag + 'o2_probe'], rtol=1e-10, atol=1e-12)
    corr = N.correlation_head(l[None, None], r[None, None])[0]
    assert np.allclose(corr, g[tag + 'corr'], rtol=1e-11, atol=1e-9)
    assert 180 - int(np.argmax(corr)) == int(g[tag + 'yaw'])
