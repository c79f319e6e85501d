"""torch-CPU oracle for the network stages (TEST INFRASTRUCTURE, see oracle/__init__.py).

PARITY UNPINNED by the reference: TensorFlow / Keras / h5py cannot be installed offline and the
pretrained ``data/model_geo.weight`` is not shipped, so no reference activation exists.  This file
restates the Keras graph line by line with library ops (torch conv2d in float64 or float32) and is
pinned only by (a) the reference's own known-answer statements -- the circular padding KAT in
``RangePadding2D.py:5`` and the shift property of the correlation head -- and (b) the independent
pure-loop restatements at the bottom of this file (``*_naive``), compared in
``tests/test_oracle_network.py``.

Keras semantics honoured throughout: ``padding='valid'``, ``use_bias=True``, kernels stored HWIO
``(kh, kw, cin, cout)``, ``channels_last``, ``Flatten`` row-major over (H, W, C), Dense kernel
``(in, out)``, and TF's conv = cross-correlation (no kernel flip).
"""
import numpy as np
import torch
import torch.nn.functional as F


def leg_layers(model_cfg=None):
  """Layer table of ``generate360OutputkLegs`` (generateNet.py:119-219; the ``...Fixed`` variant
  :222-324 is inference-identical): (name, (kh, kw), (sh, sw), cout), all ReLU, all valid."""
  cfg = dict(model_cfg or {})
  s1 = tuple(cfg.get('strides_layer1', (2, 2)))                 # generateNet.py:143-144
  layers = [('s_conv1', (5, 15), s1, 16),                       # :161-164
            ('s_conv2', (3, 15), (2, 1), 32),                   # :167-170
            ('s_conv3', (3, 15), (2, 1), 64)]                   # :173-176
  if cfg.get('additional_unsymmetric_layer3a', False):          # :145-146, :178-182
    layers.append(('s_conv3a', (3, 12), (2, 1), 64))
  layers += [('s_conv4', (2, 9), (2, 1), 128),                  # :184-187
             ('s_conv5', (1, 9), (1, 1), 128),                  # :189-192
             ('s_conv6', (1, 9), (1, 1), 128),                  # :194-197
             ('s_conv7', (1, 9), (1, 1), 128),                  # :199-202
             ('s_conv8', (1, 7), (1, 1), 128),                  # :204-207
             ('s_conv9', (1, 5), (1, 1), 128),                  # :209-212
             ('s_conv10', (1, 3), (1, 1), 128)]                 # :214-217
  return layers


def head_layers(model_cfg=None):
  """Conv layers of ``generateDeltaLayerConv1NetworkHead`` (generateNet.py:64-116):
  (name, (kh,kw), (sh,sw), cout, activation)."""
  cfg = dict(model_cfg or {})
  s = int(cfg.get('conv1NetworkHead_conv1size', 15))            # :88-89
  return [('c_conv1', (1, s), (1, s), 64, 'linear'),            # :96-100
          ('c_conv2', (s, 1), (s, 1), 128, 'relu'),             # :102-106
          ('c_conv3', (3, 3), (1, 1), 256, 'relu')]             # :108-110


def glorot_weights(in_channels, model_cfg=None, seed=0, bias_range=0.05, leg_out_width=360):
  """Seeded synthetic weights keyed by the Keras layer names: Glorot-uniform kernels
  (Keras' default initialiser) in HWIO layout, biases U(-bias_range, bias_range) so that the bias
  path is exercised (Keras' default is zeros).  Returns {name: (kernel f32, bias f32)}."""
  rng = np.random.default_rng(seed)
  w = {}
  cin = in_channels
  for name, (kh, kw), _, cout in leg_layers(model_cfg):
    lim = np.sqrt(6.0 / (kh * kw * cin + kh * kw * cout))
    w[name] = (rng.uniform(-lim, lim, (kh, kw, cin, cout)).astype(np.float32),
               rng.uniform(-bias_range, bias_range, cout).astype(np.float32))
    cin = cout
  hw = leg_out_width
  hh = leg_out_width
  for name, (kh, kw), (sh, sw), cout, _ in head_layers(model_cfg):
    lim = np.sqrt(6.0 / (kh * kw * cin + kh * kw * cout))
    w[name] = (rng.uniform(-lim, lim, (kh, kw, cin, cout)).astype(np.float32),
               rng.uniform(-bias_range, bias_range, cout).astype(np.float32))
    hh = (hh - kh) // sh + 1
    hw = (hw - kw) // sw + 1
    cin = cout
  n_in = hh * hw * cin
  lim = np.sqrt(6.0 / (n_in + 1))
  w['overlap_output'] = (rng.uniform(-lim, lim, (n_in, 1)).astype(np.float32),
                         rng.uniform(-bias_range, bias_range, 1).astype(np.float32))
  return w


def _conv(x_nchw, kernel_hwio, bias, stride, relu, dtype):
  k = torch.as_tensor(np.ascontiguousarray(np.transpose(kernel_hwio, (3, 2, 0, 1))), dtype=dtype)
  b = torch.as_tensor(bias, dtype=dtype)
  y = F.conv2d(x_nchw, k, b, stride=stride)
  return torch.relu(y) if relu else y


def leg_forward(x_nhwc, weights, model_cfg=None, dtype=torch.float64, return_all=False):
  """Leg encoder (generateNet.py:161-217) applied to ``x_nhwc`` (B, 64, 900, C).  Returns the
  feature volumes (B, 1, 360, 128) as float32 numpy (what ``Infer.create_feature_volumes``
  returns, infer.py:262-265), computed in ``dtype``."""
  x = torch.as_tensor(np.asarray(x_nhwc), dtype=dtype).permute(0, 3, 1, 2).contiguous()
  acts = []
  for name, _, stride, _ in leg_layers(model_cfg):
    k, b = weights[name]
    x = _conv(x, k, b, stride, True, dtype)
    if return_all:
      acts.append(x.permute(0, 2, 3, 1).contiguous().numpy())
  out = x.permute(0, 2, 3, 1).contiguous()
  if return_all:
    return acts
  return out.numpy().astype(np.float32)


def delta_layer(l, r):
  """``DeltaLayer`` (generateNet.py:15-61): l, r (B, w, h, C) -> abs(l[i] - r[j]) of shape
  (B, w*h, w*h, C); index 1 runs over the LEFT volume's pixels, index 2 over the RIGHT's."""
  B, w, h, C = l.shape
  rl = l.reshape(B, w * h, 1, C)                                # :48-49
  rr = r.reshape(B, 1, w * h, C)                                # :50-51
  return (rl - rr).abs()                                        # :53-59 (tile is a broadcast)


def delta_head(l_fv, r_fv, weights, model_cfg=None, dtype=torch.float64, return_all=False):
  """Overlap head (generateNet.py:64-116): l_fv, r_fv (B, 1, 360, 128) -> overlap (B, 1).
  The delta tensor is materialised like Keras does (66 MB per pair in float32)."""
  l = torch.as_tensor(np.asarray(l_fv), dtype=dtype)
  r = torch.as_tensor(np.asarray(r_fv), dtype=dtype)
  x = delta_layer(l, r).permute(0, 3, 1, 2).contiguous()        # NCHW: (B, 128, 360, 360)
  acts = []
  for name, _, stride, _, act in head_layers(model_cfg):
    k, b = weights[name]
    x = _conv(x, k, b, stride, act == 'relu', dtype)
    acts.append(x.permute(0, 2, 3, 1).contiguous())
  flat = acts[-1].reshape(acts[-1].shape[0], -1)                # Flatten over (H, W, C), :112
  kd, bd = weights['overlap_output']
  z = flat @ torch.as_tensor(kd, dtype=dtype) + torch.as_tensor(bd, dtype=dtype)
  out = torch.sigmoid(z)                                        # :114
  if return_all:
    return [a.numpy() for a in acts], z.numpy(), out.numpy()
  return out.numpy().astype(np.float32)


def range_padding(x, padding):
  """``RangePadding2D.call`` (RangePadding2D.py:31-38): width-axis wrap padding
  [x[padding:], x, x[:padding-1]] -> width 2W-1.  KAT (RangePadding2D.py:5):
  pad([1 2 3 4], 2) -> [3, 4, 1, 2, 3, 4, 1]."""
  return torch.cat([x[:, :, padding:, :], x, x[:, :, :padding - 1, :]], dim=2)


def correlation_head(l_fv, r_fv, dtype=torch.float64):
  """Yaw head (generateNet.py:327-354 with normalize='none'; NormalizedCorrelation2D.py:43-109):
  per sample, valid cross-correlation of the circularly padded LEFT volume (1, 719, 128) with the
  RIGHT volume as the kernel (1, 360, 128) -> 360 scores:
  corr[k] = sum_{j,c} L[(k + j + 180) mod 360, c] * R[j, c]."""
  l = torch.as_tensor(np.asarray(l_fv), dtype=dtype)
  r = torch.as_tensor(np.asarray(r_fv), dtype=dtype)
  B, H, W, C = l.shape
  pad = range_padding(l, W // 2)                                # NormalizedCorrelation2D.py:77
  out = []
  for b in range(B):                                            # tf.scan over samples, :79-82
    disp = pad[b:b + 1].permute(0, 3, 1, 2)                     # (1, C, H, 2W-1)
    ker = r[b].permute(2, 0, 1).unsqueeze(0)                    # (1, C, H, W): HWIO with O=1
    out.append(F.conv2d(disp, ker).reshape(-1))                 # :96-109
  return torch.stack(out).numpy()                               # Flatten, generateNet.py:352


def readout(overlap, corr):
  """``Infer`` readout (infer.py:157-158, 197-198, 232-233): yaw = 180 - argmax (first max)."""
  return np.asarray(overlap, np.float32), 180 - np.argmax(corr, axis=1)


def heads_forward(l_fv, r_fv, weights, model_cfg=None, dtype=torch.float64, batch=4, return_logit=False):
  """Both heads over n pairs, batched to bound the size of the materialised delta tensor.
  Returns (overlap (n,) f32, yaw (n,) int64, corr (n, 360) float64/32 numpy[, logit (n,)])."""
  n = l_fv.shape[0]
  ov, cr, zs = [], [], []
  for s in range(0, n, batch):
    _, z, o = delta_head(l_fv[s:s + batch], r_fv[s:s + batch], weights, model_cfg, dtype, return_all=True)
    ov.append(o[:, 0].astype(np.float32))
    zs.append(z[:, 0])
    cr.append(correlation_head(l_fv[s:s + batch], r_fv[s:s + batch], dtype))
  overlap = np.concatenate(ov) if ov else np.zeros((0,), np.float32)
  corr = np.concatenate(cr) if cr else np.zeros((0, l_fv.shape[2]))
  o, yaw = readout(overlap, corr)
  if return_logit:
    return o, yaw, corr, (np.concatenate(zs) if zs else np.zeros((0,)))
  return o, yaw, corr


def spread_dense(weights, logits, target_std=1.5):
  """Test helper: rescale / recentre the Dense(1) layer so that the given raw logits map to
  overlaps spread over (0,1) -- a much stricter check of an absolute 1e-3 gate than Glorot's
  near-constant 0.5.  Dense is linear: new logit = g * (raw - median(raw))."""
  k, b = weights['overlap_output']
  raw = np.asarray(logits, np.float64) - float(b[0])
  g = target_std / max(float(np.std(raw)), 1e-12)
  med = float(np.median(raw))
  w2 = dict(weights)
  w2['overlap_output'] = ((k.astype(np.float64) * g).astype(np.float32), np.array([-g * med], np.float32))
  return w2


# ----------------------------------------------------------------------------------------------
# Independent loop restatements (small shapes only) used to cross-check the library-op versions.
# ----------------------------------------------------------------------------------------------

def conv2d_valid_naive(x_hwc, kernel_hwio, bias, stride, relu):
  """Direct valid cross-correlation of one (H, W, C) image in float64."""
  x = np.asarray(x_hwc, np.float64)
  k = np.asarray(kernel_hwio, np.float64)
  kh, kw, cin, cout = k.shape
  sh, sw = stride
  H, W, _ = x.shape
  Ho, Wo = (H - kh) // sh + 1, (W - kw) // sw + 1
  y = np.empty((Ho, Wo, cout))
  for i in range(Ho):
    for j in range(Wo):
      patch = x[i * sh:i * sh + kh, j * sw:j * sw + kw, :]
      y[i, j] = np.tensordot(patch, k, axes=([0, 1, 2], [0, 1, 2])) + bias
  return np.maximum(y, 0) if relu else y


def correlation_naive(l_wc, r_wc):
  """corr[k] = sum_j <L[(k + j + W//2) mod W], R[j]> in float64 for one pair of (W, C) volumes."""
  L = np.asarray(l_wc, np.float64)
  R = np.asarray(r_wc, np.float64)
  W = L.shape[0]
  out = np.zeros(W)
  for k in range(W):
    idx = (k + np.arange(W) + W // 2) % W
    out[k] = np.sum(L[idx] * R)
  return out


def delta_conv1_naive(l_wc, r_wc, kernel, bias, size):
  """c_conv1 on the delta tensor without materialising it:
  o1[i, jb, o] = b[o] + sum_{dj<size, c} |L[i,c] - R[size*jb+dj, c]| * K[0, dj, c, o]."""
  L = np.asarray(l_wc, np.float64)
  R = np.asarray(r_wc, np.float64)
  K = np.asarray(kernel, np.float64)[0]                         # (size, C, cout)
  W = L.shape[0]
  nb = (W - size) // size + 1
  out = np.empty((W, nb, K.shape[-1]))
  for i in range(W):
    for jb in range(nb):
      d = np.abs(L[i][None, :] - R[size * jb:size * jb + size, :])   # (size, C)
      out[i, jb] = np.tensordot(d, K, axes=([0, 1], [0, 1])) + bias
  return out
