#!/usr/bin/env python3
"""Golden vectors for the loop-closure driver, produced by EXECUTING the reference's own
``AnimatedLCD.get_predictions`` / ``get_cov_ellipse`` / the trajectory bookkeeping of ``update``
(demo/demo3_lcd.py:85-176) in this container.

demo3_lcd.py cannot be imported (matplotlib / keras are not installed), so the two methods are cut
out of its source by AST, compiled unchanged, and bound to a stand-in object whose ``infer`` records
the calls and returns a seeded overlap field; ``Ellipse`` is a four-attribute stand-in.
Writes tests/golden/lcd_demo3.npz.  Only runs where /root/reference is mounted.
"""
import ast
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference/demo/demo3_lcd.py'


class Ellipse:                                   # matplotlib.patches.Ellipse stand-in (attributes only)
  def __init__(self, xy, width, height, angle=0.0, **kwargs):
    self.center, self.width, self.height, self.angle = xy, width, height, angle


def overlap_field(idx, ref, revisit_lag=200):
  """Deterministic synthetic overlaps: the true revisit scores 0.9, a few distractors exceed 0.3."""
  if ref == idx - revisit_lag:
    return 0.9
  h = (idx * 7919 + ref * 104729) % 1000
  return 0.45 if h < 12 else 0.001 * (h % 250)


class RecordingInfer:
  def __init__(self):
    self.calls = []

  def infer_multiple(self, idx, refs):
    refs = [int(r) for r in refs]
    self.calls.append((int(idx), refs))
    if len(refs) == 0:
      return None
    return np.array([overlap_field(idx, r) for r in refs], np.float32), np.zeros(len(refs), np.int64)


def loop_trajectory(n=300):
  t = np.arange(n) * 0.6
  side = 30.0
  s = t % (4 * side)
  x = np.where(s < side, s, np.where(s < 2 * side, side, np.where(s < 3 * side, 3 * side - s, 0.0)))
  y = np.where(s < side, 0.0, np.where(s < 2 * side, s - side, np.where(s < 3 * side, side, 4 * side - s)))
  rng = np.random.default_rng(7)
  return np.stack([x, y], 1) + rng.normal(0, 0.05, (n, 2))


def main():
  src = open(REF).read()
  tree = ast.parse(src)
  cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == 'AnimatedLCD')
  keep = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in ('get_predictions', 'get_cov_ellipse')]
  assert len(keep) == 2
  mod = ast.Module(body=[ast.ClassDef(name='RefLCD', bases=[], keywords=[], body=keep, decorator_list=[])],
                   type_ignores=[])
  ast.fix_missing_locations(mod)
  ns = {'np': np, 'Ellipse': Ellipse}
  exec(compile(mod, REF, 'exec'), ns)
  obj = ns['RefLCD']()
  obj.infer = RecordingInfer()
  obj.traj_length = []
  obj.inactive_time_thres, obj.inactive_dist_thres, obj.overlap_thres = 100, 50, 0.3   # demo3_lcd.py:53-55

  traj_all = loop_trajectory()
  rng = np.random.default_rng(11)
  covs = np.zeros((len(traj_all), 36))
  for i in range(len(traj_all)):
    a = rng.uniform(0.3, 2.5, (2, 2))
    c = np.zeros((6, 6))
    c[:2, :2] = a @ a.T
    covs[i] = c.reshape(-1)
  decisions = np.full(len(traj_all), -1, np.int64)
  for idx in range(len(traj_all)):
    traj = traj_all[:idx + 1]
    cov = covs[idx].reshape(6, 6)
    # the bookkeeping of AnimatedLCD.update (demo3_lcd.py:155-163), restated because update() also draws
    if idx > 0:
      obj.traj_length.append(obj.traj_length[-1] + np.linalg.norm(traj[idx] - traj[idx - 1]))
    else:
      obj.traj_length.append(0)
    ellipse = obj.get_cov_ellipse(cov[:2, :2], traj[idx], 3, linewidth=1, edgecolor='r', facecolor='none')
    r = obj.get_predictions(idx, traj, ellipse)
    if r is not None:
      decisions[idx] = int(r)
  calls = obj.infer.calls
  flat = np.concatenate([np.asarray(c[1], np.int64) for c in calls]) if calls else np.zeros(0, np.int64)
  offs = np.cumsum([0] + [len(c[1]) for c in calls])
  out = os.path.join(ROOT, 'tests', 'golden', 'lcd_demo3.npz')
  np.savez_compressed(out, traj=traj_all, covs=covs, decisions=decisions, call_idx=np.array([c[0] for c in calls]),
                      call_refs=flat, call_offsets=offs)
  print('wrote', out, 'frames', len(traj_all), 'loop closures', int((decisions >= 0).sum()),
        'scored candidates', len(flat))


if __name__ == '__main__':
  sys.exit(main())
