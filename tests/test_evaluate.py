"""Host logic of the evaluation flow (testing.py:207-352): npz formats, scan dedup, statistics."""
import numpy as np
import pytest

from overlapnet_b200 import evaluate as E


def _gt_table():
  return np.array([[0, 3, 0.91, 170.0], [0, 1, 0.72, 178.0], [2, 1, 0.10, 20.0], [3, 2, 0.95, 359.0]])


def test_load_overlap_npz_both_formats(tmp_path):
  t = _gt_table()
  seq = np.empty((4, 2), dtype=object); seq[:] = '07'
  np.savez_compressed(str(tmp_path / 'new.npz'), overlaps=t, seq=seq)       # demo4_gen_gt_files.py:111
  np.savez(str(tmp_path / 'old.npz'), t)                                    # single-array format
  f1, f2, d1, d2, ov, ori = E.load_overlap_npz([str(tmp_path / 'new.npz')], shuffle=False)
  assert f1 == ['000000', '000000', '000002', '000003'] and f2 == ['000003', '000001', '000001', '000002']
  assert d1 == ['07'] * 4 and d2 == ['07'] * 4
  assert np.array_equal(ov, t[:, 2]) and np.array_equal(ori, t[:, 3])
  f1o, _, d1o, _, ovo, _ = E.load_overlap_npz([str(tmp_path / 'old.npz')], shuffle=False)
  assert f1o == f1 and d1o == [''] * 4 and np.array_equal(ovo, ov)
  # shuffle keeps rows together; two files concatenate
  np.random.seed(0)
  f1s, f2s, _, _, ovs, oris = E.load_overlap_npz([str(tmp_path / 'new.npz'), str(tmp_path / 'old.npz')], shuffle=True)
  assert len(f1s) == 8 and sorted(zip(f1s, f2s, ovs.tolist(), oris.tolist())) == sorted(
      2 * list(zip(f1, f2, ov.tolist(), ori.tolist())))


def test_testdata_file_selection_rules():
  assert E.testdata_files({'data_root_folder': '/d', 'testing_seqs': '07', 'training_seqs': '03 05'}) == \
      ['/d/07/ground_truth/ground_truth_overlap_yaw.npz']
  assert E.testdata_files({'data_root_folder': '/d', 'training_seqs': '03 05'}) == \
      ['/d/03/ground_truth/validation_set.npz', '/d/05/ground_truth/validation_set.npz']
  assert E.testdata_files({'testdata_npzfile': 'x.npz'}) == ['x.npz']


def test_pair_indices_dedup():
  f1, f2 = ['000000', '000000', '000002', '000003'], ['000003', '000001', '000001', '000002']
  imgs, idx = E.pair_indices(f1, f2)
  assert imgs == ['000000', '000001', '000002', '000003']
  assert [imgs[i] for i in idx[:, 0]] == f1 and [imgs[i] for i in idx[:, 1]] == f2


def test_error_statistics_circular_yaw():
  t = _gt_table()
  model_ov = np.array([0.81, 0.82, 0.10, 0.95])
  model_arg = np.array([172, 170, 200, 1])               # last: |1 - 359| = 358 -> circular 2
  s = E.error_statistics(model_ov, model_arg, t[:, 2], t[:, 3])
  assert s['overlap_mean'] == pytest.approx(0.05) and s['overlap_max'] == pytest.approx(0.10)
  assert s['overlap_rms'] == pytest.approx(np.sqrt((0.01 + 0.01) / 4))
  assert s['yaw_pairs'] == 3                             # ground-truth overlap > 0.7 only
  assert s['yaw_mean'] == pytest.approx((2 + 8 + 2) / 3) and s['yaw_max'] == 8
  none = E.error_statistics(model_ov, model_arg, np.full(4, 0.5), t[:, 3])
  assert none['yaw_pairs'] == 0 and np.isnan(none['yaw_mean'])
