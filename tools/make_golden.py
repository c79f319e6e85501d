#!/usr/bin/env python3
"""Generate tests/golden/ from the REFERENCE implementation (run in the build container only).

Imports /root/reference/src/utils/utils.py unmodified (NumPy only), runs it on the reference's two
fixture scans and on seeded synthetic clouds, checks that the results equal the reference's own
shipped fixtures (data/preprocess_data_demo/**), and stores inputs + outputs as compressed .npz
so that the parity tests can run on the GPU box, where /root/reference does not exist.

  python tools/make_golden.py            # writes tests/golden/*.npz and MANIFEST.json
"""
import hashlib
import json
import os
import sys

import numpy as np

REF = '/root/reference'
sys.path.insert(0, os.path.join(REF, 'src/utils'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import utils as ref_utils  # noqa: E402  (the reference's module)
from overlapnet_b200 import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')


def sha(a):
  return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_reference(pts, probs=None):
  rng, vert, inten, idx = ref_utils.range_projection(pts)
  nrm = ref_utils.gen_normal_map(rng, vert)
  _, _, _, idx_inf = ref_utils.range_projection(pts, max_range=np.inf)
  out = dict(range=rng, intensity=inten, idx=idx, normal=nrm, idx_inf=idx_inf, vertex_sha=sha(vert))
  if probs is not None:
    sem = np.full((64, 900, probs.shape[1]), -1, dtype=np.float32)       # gen_semantic_data.py:42-46
    sem[idx_inf >= 0] = probs[idx_inf[idx_inf >= 0]]
    out['semantic_sha'] = sha(sem)
    out['semantic_rows'] = sem[24:28].copy()                            # a 4-row crop for a direct compare
  return out


def main():
  os.makedirs(OUT, exist_ok=True)
  manifest = {}
  # ---- the reference's two real KITTI fixture scans
  for name in ('000000', '000001'):
    pts = np.fromfile(os.path.join(REF, 'data/scans', name + '.bin'), dtype=np.float32).reshape(-1, 4)
    labels = np.fromfile(os.path.join(REF, 'data/semantic_probs', name + '.label'), dtype=np.float32).reshape(-1, 20)
    g = run_reference(pts, labels)
    demo = os.path.join(REF, 'data/preprocess_data_demo')
    assert np.array_equal(g['range'], np.load(os.path.join(demo, 'depth', name + '.npy')))
    assert np.array_equal(g['normal'], np.load(os.path.join(demo, 'normal', name + '.npy')))
    assert np.array_equal(g['intensity'], np.load(os.path.join(demo, 'intensity', name + '.npy')))
    shipped_sem = np.load(os.path.join(demo, 'semantic', name + '.npy'))
    assert sha(shipped_sem) == g['semantic_sha'], 'reference semantic fixture mismatch'
    # the real .label file is 10 MB; tests use seeded synthetic probabilities instead and the hash of
    # the shipped semantic fixture is only checked when /root/reference is present
    sp = synth.random_probs(int(name) + 100, pts.shape[0])
    gs = run_reference(pts, sp)
    np.savez_compressed(os.path.join(OUT, 'kitti_%s.npz' % name), points=pts, range=g['range'],
                        intensity=g['intensity'], idx=g['idx'], idx_inf=g['idx_inf'], normal=g['normal'],
                        semantic_rows=gs['semantic_rows'])
    manifest['kitti_' + name] = dict(n_points=int(pts.shape[0]), vertex_sha=g['vertex_sha'],
                                     semantic_sha_synthetic_probs=gs['semantic_sha'],
                                     semantic_sha_shipped=g['semantic_sha'], probs_seed=int(name) + 100)
  # ---- seeded synthetic clouds (zero-depth points exercise the filtered-index quirk)
  for seed, n, zeros in ((3, 40000, 7), (5, 9000, 3), (11, 257, 2)):
    pts = synth.kitti_like_cloud(seed, n_points=n, zero_points=zeros)
    probs = synth.random_probs(seed + 100, n)
    g = run_reference(pts, probs)
    np.savez_compressed(os.path.join(OUT, 'synth_%d.npz' % seed), points=pts, range=g['range'],
                        intensity=g['intensity'], idx=g['idx'], idx_inf=g['idx_inf'], normal=g['normal'],
                        semantic_rows=g['semantic_rows'])
    manifest['synth_%d' % seed] = dict(n_points=n, zero_points=zeros, vertex_sha=g['vertex_sha'],
                                       semantic_sha_synthetic_probs=g['semantic_sha'], probs_seed=seed + 100)
  manifest['_generator'] = 'tools/make_golden.py on numpy %s, reference commit 4188c3a' % np.__version__
  with open(os.path.join(OUT, 'MANIFEST.json'), 'w') as f:
    json.dump(manifest, f, indent=1, sort_keys=True)
  print(json.dumps(manifest, indent=1))


if __name__ == '__main__':
  main()
