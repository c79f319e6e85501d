"""Oracle restatement of the ``Infer`` class call semantics (src/two_heads/infer.py:22-265) on top
of oracle/network.py: which file is LEFT / RIGHT, bank indexing, return shapes and dtypes.
TEST INFRASTRUCTURE (see oracle/__init__.py); parity unpinned by the reference (no Keras)."""
import os

import numpy as np

from . import network as N


class InferRef:
  def __init__(self, config, weights):
    self.config = config
    self.w = weights
    self.model = config['model']
    self.datasetpath = config['data_root_folder']
    self.seq = config['infer_seqs']
    self.feature_volumes = []

  def _load(self, name):
    """prepareOneInput (ImagePairOverlapOrientationSequence.py:130-207): depth, normal,
    probabilities, intensity in that channel order, raw values."""
    chans = []
    base = os.path.join(self.datasetpath, self.seq)
    if self.config.get('use_depth', True):
      chans.append(np.load(os.path.join(base, 'depth', name + '.npy'))[..., None])
    if self.config.get('use_normals', True):
      chans.append(np.load(os.path.join(base, 'normal', name + '.npy')))
    if self.config.get('use_class_probabilities', False):
      sub = 'probability_pca' if self.config.get('use_class_probabilities_pca', False) else 'probability'
      chans.append(np.load(os.path.join(base, sub, name + '.npy')))
    if self.config.get('use_intensity', False):
      chans.append(np.load(os.path.join(base, 'intensity', name + '.npy'))[..., None])
    return np.concatenate(chans, axis=-1).astype(np.float32)

  def create_feature_volumes(self, filenames):          # infer.py:240-265
    x = np.stack([self._load(f) for f in filenames])
    return N.leg_forward(x, self.w, self.model)

  def _heads(self, bank, pairs):
    # x1 = bank[pairs[:,0]] (LEFT), x2 = bank[pairs[:,1]] (RIGHT): FeatureVolume.py:44-45
    ov, yaw, corr = N.heads_forward(bank[pairs[:, 0]], bank[pairs[:, 1]], self.w, self.model)
    return ov[:, None], yaw, corr

  def infer_one(self, filepath1, filepath2):             # infer.py:124-160
    f1 = os.path.basename(filepath1).replace('.bin', '')
    f2 = os.path.basename(filepath2).replace('.bin', '')
    fv = self.create_feature_volumes([f2, f1])
    ov, yaw, corr = self._heads(fv, np.array([[0, 1]]))
    return ov[0], yaw, corr

  def infer_multiple(self, current_frame_id, reference_frame_id):     # infer.py:162-203
    self.feature_volumes.append(self.create_feature_volumes([str(current_frame_id).zfill(6)])[0])
    if len(reference_frame_id) == 0:
      return None
    pairs = np.zeros((len(reference_frame_id), 2), dtype=int)
    pairs[:, 1] = current_frame_id
    pairs[:, 0] = reference_frame_id
    ov, yaw, corr = self._heads(np.array(self.feature_volumes), pairs)
    return ov.squeeze(), yaw, corr

  def infer_multiple_vs_multiple(self, file_names, first_idxs, second_idxs):   # infer.py:205-238
    file_names = [os.path.basename(v).replace('.bin', '') for v in file_names]
    self.feature_volumes = self.create_feature_volumes(file_names)
    if len(second_idxs) == 0:
      return None
    pairs = np.zeros((len(second_idxs), 2), dtype=int)
    pairs[:, 1] = first_idxs
    pairs[:, 0] = second_idxs
    ov, yaw, corr = self._heads(self.feature_volumes, pairs)
    return ov.squeeze(), yaw, corr
