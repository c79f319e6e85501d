#!/usr/bin/env python3
"""Generate tests/golden/gt_overlap_yaw.npz from the REFERENCE's com_overlap_yaw (build container only).

The reference function reads KITTI .bin files, so the five test clouds are written to a temporary
folder first: the two real fixture scans, rigidly moved copies of fixtures 0 and 1 (known relative
poses -> large, pose-dependent overlaps) and one seeded synthetic cloud.  Stored: the poses, the
[frame, reference, overlap, yaw bin] rows for frame_idx 0 and 3, and the clouds that cannot be
regenerated from tests/golden/kitti_*.npz + overlapnet_b200.synth alone (the two moved copies are
rebuilt by ``gt_test_clouds`` below, which the tests import).

  python tools/make_golden_gt.py
"""
import io
import os
import sys
import tempfile
from contextlib import redirect_stdout

import numpy as np

REF = '/root/reference'
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
from overlapnet_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden', 'gt_overlap_yaw.npz')


def se3(yaw_deg, pitch_deg, roll_deg, t):
  y, p, r = np.deg2rad([yaw_deg, pitch_deg, roll_deg])
  Rz = np.array([[np.cos(y), -np.sin(y), 0], [np.sin(y), np.cos(y), 0], [0, 0, 1]])
  Ry = np.array([[np.cos(p), 0, np.sin(p)], [0, 1, 0], [-np.sin(p), 0, np.cos(p)]])
  Rx = np.array([[1, 0, 0], [0, np.cos(r), -np.sin(r)], [0, np.sin(r), np.cos(r)]])
  T = np.eye(4)
  T[:3, :3] = Rz @ Ry @ Rx
  T[:3, 3] = t
  return T


def gt_test_poses():
  return np.stack([np.eye(4), se3(1.5, 0.2, -0.1, [0.9, 0.05, 0.01]), se3(10.0, 0.5, 0.3, [2.0, -0.5, 0.05]),
                   se3(93.0, -1.0, 0.8, [9.0, 4.0, -0.2]), se3(-140.0, 0.0, 0.0, [-20.0, 7.0, 0.3])])


def gt_test_clouds(golden_dir):
  """The five clouds (float32 (N,4)) the fixture was generated on."""
  k0 = np.load(os.path.join(golden_dir, 'kitti_000000.npz'))['points']
  k1 = np.load(os.path.join(golden_dir, 'kitti_000001.npz'))['points']
  poses = gt_test_poses()
  clouds = [k0, k1]
  # cloud 2: fixture 0 (at pose 0) seen from pose 2; cloud 3: fixture 1 (at pose 1) seen from pose 3;
  # p_j = inv(T_j) T_src p, stored as float32
  for j, (src, src_pose) in ((2, (k0, poses[0])), (3, (k1, poses[1]))):
    h = np.ones((src.shape[0], 4))
    h[:, :3] = src[:, :3]
    moved = np.linalg.inv(poses[j]).dot(src_pose.dot(h.T)).T
    c = src.copy()
    c[:, :3] = moved[:, :3].astype(np.float32)
    clouds.append(c)
  clouds.append(synth.kitti_like_cloud(77, n_points=60000))
  return clouds, poses


def main():
  sys.path.insert(0, os.path.join(REF, 'src/utils'))
  from com_overlap_yaw import com_overlap_yaw  # the reference's function, unmodified
  golden_dir = os.path.join(ROOT, 'tests', 'golden')
  clouds, poses = gt_test_clouds(golden_dir)
  out = {'poses': poses}
  with tempfile.TemporaryDirectory() as tmp:
    paths = []
    for i, c in enumerate(clouds):
      p = os.path.join(tmp, '%06d.bin' % i)
      np.ascontiguousarray(c, np.float32).tofile(p)
      paths.append(p)
    for frame in (0, 3):
      with redirect_stdout(io.StringIO()):
        out['mapping_frame%d' % frame] = com_overlap_yaw(paths, poses, frame_idx=frame)
  np.savez_compressed(OUT, **out)
  for k, v in out.items():
    print(k, v.shape)
    if k.startswith('mapping'):
      print(v)


if __name__ == '__main__':
  main()
