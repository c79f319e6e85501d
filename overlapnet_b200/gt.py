"""GPU drop-in for the reference's ground-truth generator ``com_overlap_yaw``
(``src/utils/com_overlap_yaw.py:10-68``, used by ``demo/demo4_gen_gt_files.py:77``): same name,
arguments and return value.  The O(N) pose-transformed float64 range projections and the pixel
comparisons run in csrc/gt_overlap.cu through the C ABI (``ovn_gt_range_batch``,
``ovn_gt_overlap_count``); the yaw bin is 4x4 host arithmetic exactly as the reference writes it.
There is no CPU fallback."""
import math

import numpy as np

from .preprocess import _engine, _read_scan


def load_poses(pose_path):
  """``load_poses`` (utils.py:10-36): KITTI ``.txt`` (12 numbers per line) or ``.npz['arr_0']`` -> (n,4,4)."""
  poses = []
  try:
    if '.txt' in pose_path:
      with open(pose_path, 'r') as f:
        for line in f.readlines():
          T = np.array(line.split(), dtype=float).reshape(3, 4)
          poses.append(np.vstack((T, [0, 0, 0, 1])))
    else:
      poses = np.load(pose_path)['arr_0']
  except FileNotFoundError:
    print('Ground truth poses are not avaialble.')
  return np.array(poses)


def load_calib(calib_path):
  """``load_calib`` (utils.py:39-56): the ``Tr:`` line of a KITTI calib file -> T_cam_velo (4,4)."""
  T_cam_velo = []
  try:
    with open(calib_path, 'r') as f:
      for line in f.readlines():
        if 'Tr:' in line:
          T = np.array(line.replace('Tr:', '').split(), dtype=float).reshape(3, 4)
          T_cam_velo = np.vstack((T, [0, 0, 0, 1]))
  except FileNotFoundError:
    print('Calibrations are not avaialble.')
  return np.array(T_cam_velo)


def euler_angles_from_rotation_matrix(R):
  """(psi, theta, phi) = roll, pitch, yaw after Slabaugh, as ``utils.py:178-216``."""
  def isclose(x, y, rtol=1.e-5, atol=1.e-8):
    return abs(x - y) <= atol + rtol * abs(y)
  phi = 0.0
  if isclose(R[2, 0], -1.0):
    theta = math.pi / 2.0
    psi = math.atan2(R[0, 1], R[0, 2])
  elif isclose(R[2, 0], 1.0):
    theta = -math.pi / 2.0
    psi = math.atan2(-R[0, 1], -R[0, 2])
  else:
    theta = -math.asin(R[2, 0])
    cos_theta = math.cos(theta)
    psi = math.atan2(R[2, 1] / cos_theta, R[2, 2] / cos_theta)
    phi = math.atan2(R[1, 0] / cos_theta, R[0, 0] / cos_theta)
  return psi, theta, phi


def overlap_yaw_from_clouds(clouds, poses, frame_idx, leg_output_width=360, scans_per_launch=16):
  """The body of ``com_overlap_yaw``: rows [frame_idx, reference_idx, overlap, yaw bin] as float64.
  ``clouds``: list of (N,4) float32 arrays, or of zero-argument callables returning one (read lazily,
  ``scans_per_launch`` at a time, so a whole KITTI sequence never sits in host memory)."""
  poses = np.asarray(poses, dtype=np.float64)
  n = len(clouds)
  eng = _engine(3.0, -25.0, 64, 900, 50)

  def get(i):
    c = clouds[i]
    return np.ascontiguousarray(c() if callable(c) else c, np.float32)

  cur = eng.gt_range(eng.upload_clouds([get(frame_idx)]))[0]
  current_pose = poses[frame_idx]
  cur_inv = np.linalg.inv(current_pose)
  counts = np.zeros(n, np.int64)
  valid_num = 0
  for s0 in range(0, n, scans_per_launch):
    s1 = min(n, s0 + scans_per_launch)
    batch = eng.upload_clouds([get(i) for i in range(s0, s1)])
    ref = eng.gt_range(batch, pose_ref=poses[s0:s1], pose_cur_inv=cur_inv)
    c = eng.gt_overlap_count(ref, cur).cpu().numpy()
    counts[s0:s1] = c[:-1]
    valid_num = int(c[-1])
  mapping = np.zeros((n, 4))
  mapping[:, 0] = np.ones(n) * frame_idx
  mapping[:, 1] = np.arange(n)
  yaw_resolution = leg_output_width
  for r in range(n):
    mapping[r, 2] = int(counts[r]) / valid_num                                   # com_overlap_yaw.py:44-46
    relative_transform = cur_inv.dot(poses[r])                                   # :49
    _, _, yaw = euler_angles_from_rotation_matrix(relative_transform[:3, :3])    # :50-51
    mapping[r, 3] = int(- (yaw / np.pi) * yaw_resolution//2 + yaw_resolution//2)  # :54, same expression
  return mapping


def com_overlap_yaw(scan_paths, poses, frame_idx, leg_output_width=360):
  """Drop-in for ``com_overlap_yaw`` (com_overlap_yaw.py:10-68): ground-truth overlap and yaw bin of
  every scan in ``scan_paths`` against scan ``frame_idx``, from the ground-truth ``poses`` (n,4,4).
  Returns the (n, 4) float64 array [current_frame_idx, reference_frame_idx, overlap, yaw]."""
  print('Start to compute ground truth overlap and yaw ...')
  clouds = [(lambda p=p: _read_scan(p)) for p in scan_paths]      # streamed like the reference (one scan at a time there)
  mapping = overlap_yaw_from_clouds(clouds, poses, frame_idx, leg_output_width)
  print('Finish generating ground_truth_mapping!')
  return mapping
