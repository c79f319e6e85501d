"""Weight container keyed by the reference's Keras layer names.

The reference loads a Keras HDF5 file with ``load_weights(file, by_name=True)`` (infer.py:117-120):
matching is by layer name -- ``s_conv1..s_conv10`` (+ ``s_conv3a``), ``c_conv1..c_conv3``,
``overlap_output`` (generateNet.py:99-114,162-214) -- with Keras layouts: conv kernel
``(kh, kw, cin, cout)``, dense kernel ``(in, out)``, bias ``(cout,)``.  h5py is not installable
offline, so the native container here is an ``.npz`` with entries ``<layer>/kernel`` and
``<layer>/bias``; ``load_keras_h5`` imports the reference's HDF5 files when h5py is available.
"""
import os

import numpy as np

LEG_TABLE = [  # (name, kh, kw, sh, sw, cout, optional) -- generateNet.py:161-217
    ('s_conv1', 5, 15, None, None, 16, False), ('s_conv2', 3, 15, 2, 1, 32, False),
    ('s_conv3', 3, 15, 2, 1, 64, False), ('s_conv3a', 3, 12, 2, 1, 64, True),
    ('s_conv4', 2, 9, 2, 1, 128, False), ('s_conv5', 1, 9, 1, 1, 128, False),
    ('s_conv6', 1, 9, 1, 1, 128, False), ('s_conv7', 1, 9, 1, 1, 128, False),
    ('s_conv8', 1, 7, 1, 1, 128, False), ('s_conv9', 1, 5, 1, 1, 128, False),
    ('s_conv10', 1, 3, 1, 1, 128, False)]


def layer_shapes(in_channels, model_cfg, H=64, W=900):
  """{name: (kernel_shape, bias_shape)} for the configured model, in Keras layouts."""
  s1 = tuple(model_cfg.get('strides_layer1', (2, 2)))
  use3a = bool(model_cfg.get('additional_unsymmetric_layer3a', False))
  size = int(model_cfg.get('conv1NetworkHead_conv1size', 15))
  shapes = {}
  cin, h, w = in_channels, H, W
  for name, kh, kw, sh, sw, cout, opt in LEG_TABLE:
    if opt and not use3a:
      continue
    if sh is None:
      sh, sw = s1
    shapes[name] = ((kh, kw, cin, cout), (cout,))
    h, w, cin = (h - kh) // sh + 1, (w - kw) // sw + 1, cout
  wf = w
  shapes['c_conv1'] = ((1, size, cin, 64), (64,))
  shapes['c_conv2'] = ((size, 1, 64, 128), (128,))
  shapes['c_conv3'] = ((3, 3, 128, 256), (256,))
  n = wf // size
  shapes['overlap_output'] = (((n - 2) * (n - 2) * 256, 1), (1,))
  return shapes


def glorot_init(in_channels, model_cfg, seed=None, H=64, W=900):
  """Keras' default initialisation (glorot_uniform kernels, zero biases): what the reference runs
  with when ``pretrained_weightsfilename`` is empty (infer.py:117-122)."""
  rng = np.random.default_rng(seed)
  out = {}
  for name, (ks, bs) in layer_shapes(in_channels, model_cfg, H, W).items():
    if len(ks) == 4:
      fan_in, fan_out = ks[0] * ks[1] * ks[2], ks[0] * ks[1] * ks[3]
    else:
      fan_in, fan_out = ks
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    out[name] = (rng.uniform(-lim, lim, ks).astype(np.float32), np.zeros(bs, np.float32))
  return out


def save_npz(path, weights):
  flat = {}
  for name, (k, b) in weights.items():
    flat[name + '/kernel'] = np.asarray(k, np.float32)
    flat[name + '/bias'] = np.asarray(b, np.float32)
  np.savez(path, **flat)


def load_npz(path):
  z = np.load(path)
  names = sorted({k.split('/')[0] for k in z.files})
  return {n: (z[n + '/kernel'].astype(np.float32), z[n + '/bias'].astype(np.float32)) for n in names}


def load_keras_h5(path):
  """Import a Keras full-model / weights HDF5 file as written by the reference's training
  (training.py:211-212,347-349).  Needs h5py, which this image does not ship."""
  try:
    import h5py
  except ImportError as e:
    raise Exception('Reading Keras HDF5 weights needs h5py, which is not installed; convert the '
                    'file to the .npz container (overlapnet_b200.weights.save_npz) instead') from e
  out = {}
  with h5py.File(path, 'r') as f:
    g = f['model_weights'] if 'model_weights' in f else f
    def visit(name, obj):
      if isinstance(obj, h5py.Dataset):
        parts = name.split('/')
        layer = parts[0]
        kind = 'kernel' if 'kernel' in parts[-1] else ('bias' if 'bias' in parts[-1] else None)
        if kind:
          out.setdefault(layer, {})[kind] = np.asarray(obj, np.float32)
    g.visititems(visit)
  return {n: (d['kernel'], d['bias']) for n, d in out.items() if 'kernel' in d and 'bias' in d}


def load(path):
  """Load weights from ``.npz`` (native) or Keras HDF5 (anything else, like the reference's
  ``data/model_geo.weight``)."""
  if not os.path.exists(path):
    raise Exception('Weights file not found: %s' % path)
  if path.endswith('.npz'):
    return load_npz(path)
  with open(path, 'rb') as f:
    magic = f.read(8)
  if magic[:6] == b'\x93NUMPY' or magic[:2] == b'PK':
    return load_npz(path)
  return load_keras_h5(path)
