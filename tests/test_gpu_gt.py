"""GPU parity of the ground-truth overlap / yaw generator (csrc/gt_overlap.cu through the C ABI)
against the golden vectors produced by the reference's ``com_overlap_yaw`` and against the oracle.

Tolerance: yaw bins and frame indices exact.  The range images are float64 arithmetic followed by a
float32 store; the only freedom is the rounding of the two 4x4 pose products and of atan2 / asin
(<= 1-2 ulp of a float64), which can move a point across a bin edge or the |dr| < 1 threshold with
probability ~1e-12 per point: at most MAX_PIXELS pixels per image may differ, so overlaps agree to
MAX_PIXELS / valid_num (about 7e-5)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
from make_golden_gt import gt_test_clouds  # noqa: E402
from oracle import gt as G  # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
MAX_PIXELS = 3


@pytest.fixture(scope='module')
def fixture():
  clouds, poses = gt_test_clouds(GOLDEN)
  return clouds, poses, np.load(os.path.join(GOLDEN, 'gt_overlap_yaw.npz'))


def test_gt_range_images_match_oracle(engine_fp32, fixture):
  clouds, poses, _ = fixture
  eng = engine_fp32
  cur_inv = np.linalg.inv(poses[3])
  got = eng.gt_range(eng.upload_clouds(clouds), pose_ref=poses, pose_cur_inv=cur_inv).cpu().numpy()
  for r, c in enumerate(clouds):
    v = cur_inv.dot(poses[r].dot(G.homogeneous_points(c).T)).T
    want = G.range_image_f64(v)
    assert got[r].dtype == np.float32
    assert np.count_nonzero(got[r].view(np.uint32) != want.view(np.uint32)) <= MAX_PIXELS
  # no transform: the current scan's own image (com_overlap_yaw.py:29-30)
  own = eng.gt_range(eng.upload_clouds([clouds[0]])).cpu().numpy()[0]
  assert np.count_nonzero(own.view(np.uint32) != G.range_image_f64(G.homogeneous_points(clouds[0])).view(np.uint32)) <= MAX_PIXELS


def test_gt_overlap_count_matches_numpy(engine_fp32):
  import torch
  rng = np.random.default_rng(5)
  ref = np.where(rng.random((4, 64, 900)) < 0.2, -1.0, rng.uniform(0, 50, (4, 64, 900))).astype(np.float32)
  cur = np.where(rng.random((64, 900)) < 0.2, -1.0, rng.uniform(0, 50, (64, 900))).astype(np.float32)
  ref[1] = cur                                     # identical image: every valid pixel counts
  got = engine_fp32.gt_overlap_count(torch.from_numpy(ref).cuda(), torch.from_numpy(cur).cuda()).cpu().numpy()
  want = [G.overlap_counts(cur, r) for r in ref] + [int(np.count_nonzero(cur > 0))]
  assert got.tolist() == want
  assert got[1] == got[4]


@pytest.mark.parametrize('frame', [0, 3])
def test_mapping_matches_reference_golden(fixture, frame):
  from overlapnet_b200.gt import overlap_yaw_from_clouds
  clouds, poses, gold = fixture
  rows = overlap_yaw_from_clouds(clouds, poses, frame, scans_per_launch=2)       # exercises the chunking
  want = gold['mapping_frame%d' % frame]
  assert rows.dtype == np.float64 and rows.shape == want.shape
  assert np.array_equal(rows[:, [0, 1, 3]], want[:, [0, 1, 3]])
  valid_num = np.count_nonzero(G.range_image_f64(G.homogeneous_points(clouds[frame])) > 0)
  assert np.max(np.abs(rows[:, 2] - want[:, 2])) <= MAX_PIXELS / valid_num


def test_com_overlap_yaw_drop_in_reads_bin_files(fixture, tmp_path, capsys):
  from overlapnet_b200 import com_overlap_yaw
  clouds, poses, gold = fixture
  paths = []
  for i, c in enumerate(clouds):
    p = tmp_path / ('%06d.bin' % i)
    np.ascontiguousarray(c, np.float32).tofile(p)
    paths.append(str(p))
  rows = com_overlap_yaw(paths, poses, frame_idx=0)
  assert 'Finish generating ground_truth_mapping!' in capsys.readouterr().out       # the reference prints this
  want = gold['mapping_frame0']
  assert np.array_equal(rows[:, [0, 1, 3]], want[:, [0, 1, 3]])
  assert np.max(np.abs(rows[:, 2] - want[:, 2])) < 1e-4
