"""Host logic of the loop-closure driver (demo/demo3_lcd.py:85-140): candidate gating and the
decision rule, with a stub in place of the GPU ``Infer``."""
import numpy as np

from overlapnet_b200 import lcd


def reference_gate(idx, traj, traj_length, cov, inactive_time_thres=100, inactive_dist_thres=50):
  """Line-by-line restatement of demo3_lcd.py:92-115 + get_cov_ellipse :125-140."""
  eigvals, eigvecs = np.linalg.eigh(cov)
  order = eigvals.argsort()[::-1]
  eigvals, eigvecs = eigvals[order], eigvecs[:, order]
  theta = np.arctan2(eigvecs[:, 0][1], eigvecs[:, 0][0])
  width, height = 2 * 3 * np.sqrt(eigvals[:2])
  angle = np.degrees(theta)
  indices = np.arange(idx - inactive_time_thres)
  dist_delta = traj_length[idx] - np.array(traj_length)[indices]
  indices = indices[dist_delta > inactive_dist_thres]
  ca, sa = np.cos(np.radians(180. - angle)), np.sin(np.radians(180. - angle))
  xc, yc = traj[idx, 0] - traj[indices, 0], traj[idx, 1] - traj[indices, 1]
  xct, yct = xc * ca - yc * sa, xc * sa + yc * ca
  rad = (xct ** 2 / (width / 2.) ** 2) + (yct ** 2 / (height / 2.) ** 2)
  return indices[rad < 1]


class StubInfer:
  def __init__(self, revisit_of):
    self.calls = []
    self.revisit_of = revisit_of

  def infer_multiple(self, idx, refs):
    self.calls.append((idx, list(refs)))
    if len(refs) == 0:
      return None
    ov = np.array([0.9 if r == self.revisit_of.get(idx, -1) else 0.05 for r in refs], np.float32)
    return ov, np.zeros(len(refs), np.int64)


def loop_trajectory(n=260):
  """A square loop revisited: frames 200.. retrace frames 0.. (0.6 m per frame)."""
  t = np.arange(n) * 0.6
  side = 30.0
  s = t % (4 * side)
  x = np.where(s < side, s, np.where(s < 2 * side, side, np.where(s < 3 * side, 3 * side - s, 0.0)))
  y = np.where(s < side, 0.0, np.where(s < 2 * side, s - side, np.where(s < 3 * side, side, 4 * side - s)))
  return np.stack([x, y], 1)


def test_gating_matches_reference_restatement():
  traj = loop_trajectory()
  tl = np.concatenate([[0], np.cumsum(np.linalg.norm(np.diff(traj, axis=0), axis=1))])
  rng = np.random.default_rng(0)
  for idx in (100, 150, 205, 230, 259):
    a = rng.uniform(0.5, 4.0, (2, 2))
    cov = a @ a.T
    got = lcd.gate_candidates(idx, traj, tl, lcd.get_cov_ellipse(cov, traj[idx], 3))
    want = reference_gate(idx, traj, tl, cov)
    assert np.array_equal(got, want)
  # nothing is old enough before frame 100 + the 50 m rule
  assert lcd.gate_candidates(100, traj, tl, lcd.get_cov_ellipse(np.eye(2), traj[100], 3)).size == 0


def test_driver_decisions_and_bank_protocol():
  traj = loop_trajectory()
  revisit = {i: i - 200 for i in range(200, 260)}              # frame i looks at the same place as frame i-200
  stub = StubInfer(revisit)
  det = lcd.LoopClosureDetector(stub)
  cov = np.zeros((6, 6))
  cov[:2, :2] = np.diag([4.0, 4.0])                              # 3 sigma = 6 m search radius
  found = {}
  for i in range(len(traj)):
    r = det.step(i, traj[i], cov)
    if r is not None:
      found[i] = r
  # infer_multiple is called exactly once per frame, in order (the bank index is the frame id)
  assert [c[0] for c in stub.calls] == list(range(len(traj)))
  assert all(len(c[1]) == 0 for c in stub.calls[:100])
  assert found and all(found[i] == i - 200 for i in found)
  assert min(found) >= 200


def _golden_overlap(idx, ref):
  """tools/make_golden_lcd.py:overlap_field (the seeded overlap field the golden run used)."""
  if ref == idx - 200:
    return 0.9
  h = (idx * 7919 + ref * 104729) % 1000
  return 0.45 if h < 12 else 0.001 * (h % 250)


class FieldInfer:
  def __init__(self):
    self.calls = []

  def infer_multiple(self, idx, refs):
    refs = [int(r) for r in refs]
    self.calls.append((int(idx), refs))
    if not refs:
      return None
    return np.array([_golden_overlap(idx, r) for r in refs], np.float32), np.zeros(len(refs), np.int64)


def test_driver_matches_the_reference_run():
  """Pinned to the reference itself: tests/golden/lcd_demo3.npz holds what demo3_lcd.py's own
  get_predictions / get_cov_ellipse (executed by tools/make_golden_lcd.py) asked of Infer for 300
  frames and which loop closures they reported."""
  import os
  from conftest import GOLDEN
  g = np.load(os.path.join(GOLDEN, 'lcd_demo3.npz'))
  inf = FieldInfer()
  det = lcd.LoopClosureDetector(inf)
  dec = []
  for i in range(len(g['traj'])):
    r = det.step(i, g['traj'][i], g['covs'][i])
    dec.append(-1 if r is None else int(r))
  assert np.array_equal(np.array(dec), g['decisions'])
  assert [c[0] for c in inf.calls] == g['call_idx'].tolist()
  offs = g['call_offsets']
  for k, (_, refs) in enumerate(inf.calls):
    assert refs == g['call_refs'][offs[k]:offs[k + 1]].tolist()
  assert (g['decisions'] >= 0).sum() > 50 and len(g['call_refs']) > 500      # the golden run is not trivial
