"""``Infer`` with the growing loop-closure bank sharded over the GPUs of a process group.

The reference's online flow (demo/demo3_lcd.py:85-123 -> ``Infer.infer_multiple``, infer.py:162-203)
appends one feature volume per frame to a host list and re-stacks the whole list for every query
(infer.py:193).  Here every rank of the group runs the same driver loop; frame ``i`` is stored on
rank ``i % world`` (its GPU-resident bank grows by one row), rank ``src`` reads the preprocessed cue
files and encodes the current frame, ONE broadcast ships the 184 KB volume, every rank scores the
reference frames it owns and ONE gather returns the 8-byte (overlap, yaw) records to ``src``.
With world size 1 (or no process group) it behaves exactly like ``Infer``.
"""
import numpy as np

from .engine import FEAT_C
from .infer import Infer
from .search import ShardedBank


class ShardedInfer(Infer):
  """Same constructor as ``Infer`` plus ``group`` / ``src``.  ``infer_multiple`` returns the reference's
  result on rank ``src`` and None on every other rank."""

  def __init__(self, config, group=None, src=0, **kwargs):
    super().__init__(config, **kwargs)
    self._sb = ShardedBank(self._encode_frame, self._append_one, self._heads_local,
                           (self.network_output_size, FEAT_C), self._engine.device, group=group, src=src)

  def _encode_frame(self, frame_id):
    return self._create_feature_volumes_device([str(frame_id).zfill(6)])[0]

  def _append_one(self, fv):
    self._append_bank(fv.reshape(1, self.network_output_size, FEAT_C))
    self._fv_as_array = False

  def _heads_local(self, local_rows, query):
    ov, yaw, _ = self._engine.heads_1vsN(self._bank[:self._bank_n], query, cand_idx=local_rows)
    return ov, yaw

  def infer_multiple(self, current_frame_id, reference_frame_id):
    """infer.py:162-203 on the sharded bank.  LEFT = reference frames, RIGHT = the current frame."""
    res = self._sb.step(current_frame_id, reference_frame_id, on_query=self._on_query)
    self._engine.check()
    if res is None:
      return None
    ov, yaw = res
    overlap = ov.cpu().numpy()[:, None]
    return overlap.squeeze(), yaw.cpu().numpy().astype(np.int64)

  def _on_query(self, frame_id, q):
    # every rank derives the numeric centres of the tensor-core heads from frame 0 (ovn_calibrate), like a
    # single-GPU Infer does implicitly with its first bank row: sharded and unsharded results are bit-identical
    if frame_id == 0:
      self._engine.calibrate(q)

  @property
  def local_frames(self):
    """Frame ids whose volumes live on this rank, in local row order."""
    return list(range(self._sb.rank, self._sb.n_frames, self._sb.world))
