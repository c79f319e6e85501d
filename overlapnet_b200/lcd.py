"""Loop-closure detection driver: the per-frame logic of the reference's demo 3
(demo/demo3_lcd.py:85-176) without the matplotlib animation.

For frame idx: frames younger than ``inactive_time_thres`` (100) only get encoded; otherwise the
candidates are the frames that are at least 100 frames old AND more than ``inactive_dist_thres``
(50 m) of travelled distance back AND inside the 3-sigma covariance ellipse of the current pose;
``Infer.infer_multiple(idx, candidates)`` scores them and a loop closure is reported when the best
overlap exceeds ``overlap_thres`` (0.3).  The gating is host-side NumPy (it is O(frames)); the
scoring runs on the GPU-resident bank held by ``Infer``.
"""
import numpy as np


def get_cov_ellipse(cov, center, nstd):
  """demo3_lcd.py:125-140 without the matplotlib patch: returns (center, width, height, angle_deg)."""
  eigvals, eigvecs = np.linalg.eigh(cov)
  order = eigvals.argsort()[::-1]
  eigvals, eigvecs = eigvals[order], eigvecs[:, order]
  vx, vy = eigvecs[:, 0][0], eigvecs[:, 0][1]
  theta = np.arctan2(vy, vx)
  width, height = 2 * nstd * np.sqrt(eigvals[:2])
  return np.asarray(center), float(width), float(height), float(np.degrees(theta))


def gate_candidates(idx, traj, traj_length, ellipse, inactive_time_thres=100, inactive_dist_thres=50):
  """Candidate gating of ``get_predictions`` (demo3_lcd.py:92-115).  ``traj``: (>=idx+1, 2) xy
  positions, ``traj_length``: travelled distance per frame, ``ellipse``: result of get_cov_ellipse.
  Returns the int array of reference frame ids (may be empty)."""
  indices = np.arange(idx - inactive_time_thres)
  if indices.size == 0:
    return indices
  dist_delta = traj_length[idx] - np.array(traj_length)[indices]
  indices = indices[dist_delta > inactive_dist_thres]
  _, width, height, angle = ellipse
  cos_angle = np.cos(np.radians(180. - angle))
  sin_angle = np.sin(np.radians(180. - angle))
  xc = traj[idx, 0] - traj[indices, 0]
  yc = traj[idx, 1] - traj[indices, 1]
  xct = xc * cos_angle - yc * sin_angle
  yct = xc * sin_angle + yc * cos_angle
  rad_cc = (xct ** 2 / (width / 2.) ** 2) + (yct ** 2 / (height / 2.) ** 2)
  return indices[rad_cc < 1]


class LoopClosureDetector:
  """Stateful driver: call ``step(idx, pose_xy, cov6x6)`` for idx = 0, 1, 2, ... (the bank index is
  the frame id, infer.py:166-170).  ``infer`` is an ``overlapnet_b200.Infer`` (or anything with its
  ``infer_multiple``)."""

  def __init__(self, infer, inactive_time_thres=100, inactive_dist_thres=50, overlap_thres=0.3, nstd=3):
    self.infer = infer
    self.inactive_time_thres = inactive_time_thres
    self.inactive_dist_thres = inactive_dist_thres
    self.overlap_thres = overlap_thres
    self.nstd = nstd
    self.traj = []
    self.traj_length = []

  def step(self, idx, pose_xy, cov):
    """Returns the frame id of the detected loop closure or None (demo3_lcd.py:85-123,149-176)."""
    assert idx == len(self.traj), 'frames must arrive in order 0,1,2,...'
    self.traj.append(np.asarray(pose_xy, dtype=float))
    traj = np.asarray(self.traj)
    if idx > 0:
      self.traj_length.append(self.traj_length[-1] + np.linalg.norm(traj[idx] - traj[idx - 1]))
    else:
      self.traj_length.append(0)
    if idx < self.inactive_time_thres:
      self.infer.infer_multiple(idx, [])
      return None
    cov = np.asarray(cov, dtype=float).reshape(6, 6)
    ellipse = get_cov_ellipse(cov[:2, :2], traj[idx], self.nstd)
    reference_idx = gate_candidates(idx, traj, self.traj_length, ellipse, self.inactive_time_thres,
                                    self.inactive_dist_thres)
    if len(reference_idx) > 0:
      res = self.infer.infer_multiple(idx, reference_idx)
      if res is None:                      # ShardedInfer: ranks other than the source only follow along
        return None
      overlaps, _ = res
      if np.max(overlaps) > self.overlap_thres:
        return int(reference_idx[np.argmax(overlaps)])
      return None
    self.infer.infer_multiple(idx, [])
    return None
