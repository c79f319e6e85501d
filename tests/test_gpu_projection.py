"""GPU parity of stage 1 (projection / normals / semantic / packing) through the C ABI:
bit-exact against the reference-generated golden vectors and against the oracle."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, load_golden
from oracle import projection as P
from overlapnet_b200 import synth

pytestmark = pytest.mark.gpu


def bits(a):
  a = np.ascontiguousarray(a)
  return a.view(np.uint32) if a.dtype == np.float32 else a


def sha(a):
  return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize('case', GOLDEN_CASES)
def test_projection_golden_bit_exact(engine_fp32, case, manifest):
  g = load_golden(case)
  eng = engine_fp32
  batch = eng.upload_clouds([g['points']])
  out = eng.project(batch)
  assert np.array_equal(bits(out['range'][0].cpu().numpy()), bits(g['range']))
  assert np.array_equal(bits(out['intensity'][0].cpu().numpy()), bits(g['intensity']))
  assert np.array_equal(out['idx'][0].cpu().numpy(), g['idx'])
  assert sha(out['vertex'][0].cpu().numpy()) == manifest[case]['vertex_sha']
  nrm = eng.normals(out['range'], out['vertex'])
  assert np.array_equal(bits(nrm[0].cpu().numpy()), bits(g['normal']))
  # max_range = inf (gen_semantic_data.py:39) + the filtered-index gather
  idx_inf = eng.project(batch, max_range=float('inf'), want=('idx',))['idx']
  assert np.array_equal(idx_inf[0].cpu().numpy(), g['idx_inf'])
  probs = synth.random_probs(manifest[case]['probs_seed'], g['points'].shape[0])
  sem = eng.semantic(idx_inf, torch.from_numpy(probs).to(eng.device), batch.offsets)[0].cpu().numpy()
  assert sha(sem) == manifest[case]['semantic_sha_synthetic_probs']
  assert np.array_equal(sem[24:28], g['semantic_rows'])


def test_fused_preprocess_matches_separate_stages_and_golden(engine_fp32):
  eng = engine_fp32
  cases = [load_golden(c) for c in GOLDEN_CASES]
  batch = eng.upload_clouds([g['points'] for g in cases])        # ragged batch: 124668 ... 257 points
  x = eng.preprocess(batch).cpu().numpy()
  assert x.shape == (len(cases), 64, 900, 4)
  for i, g in enumerate(cases):
    assert np.array_equal(bits(x[i, :, :, 0]), bits(g['range']))
    assert np.array_equal(bits(x[i, :, :, 1:4]), bits(g['normal']))


def test_oracle_parity_synthetic_batch(engine_fp32):
  """Seeded KITTI-shaped clouds incl. zero-depth points, one empty cloud and a 1-point cloud."""
  eng = engine_fp32
  clouds = [synth.kitti_like_cloud(20 + s, n_points=n, zero_points=z)
            for s, (n, z) in enumerate([(124668, 0), (60000, 11), (33, 1), (5000, 0)])]
  clouds.insert(2, np.zeros((0, 4), np.float32))
  clouds.append(np.array([[2.0, 1.0, -0.2, 0.7]], np.float32))
  batch = eng.upload_clouds(clouds)
  out = eng.project(batch)
  nrm = eng.normals(out['range'], out['vertex']).cpu().numpy()
  out = {k: v.cpu().numpy() for k, v in out.items()}
  for i, c in enumerate(clouds):
    rng, vert, inten, idx = P.range_projection(c)
    assert np.array_equal(bits(out['range'][i]), bits(rng)), i
    assert np.array_equal(bits(out['vertex'][i]), bits(vert)), i
    assert np.array_equal(bits(out['intensity'][i]), bits(inten)), i
    assert np.array_equal(out['idx'][i], idx), i
    assert np.array_equal(bits(nrm[i]), bits(P.gen_normal_map(rng, vert))), i


def test_more_scans_than_workspace(engine_fp32):
  """n_scans > max_batch_scans (8 here) is chunked by the wrapper; results stay per-scan exact."""
  eng = engine_fp32
  clouds = [synth.kitti_like_cloud(100 + s, n_points=3000 + 97 * s) for s in range(19)]
  x = eng.preprocess(eng.upload_clouds(clouds)).cpu().numpy()
  for i in (0, 7, 8, 18):
    rng, vert, _, _ = P.range_projection(clouds[i])
    ref = P.pack_input(rng, P.gen_normal_map(rng, vert))
    assert np.array_equal(bits(x[i]), bits(ref)), i


def test_pack_input_and_5_channel_fused():
  from overlapnet_b200.engine import Engine
  eng = Engine(use={'use_intensity': True}, precision='fp32', model={'additional_unsymmetric_layer3a': True},
               max_batch_scans=2, max_batch_pairs=1)
  assert eng.C == 5
  g = load_golden('synth_5')
  batch = eng.upload_clouds([g['points']])
  x = eng.preprocess(batch).cpu().numpy()[0]
  ref = P.pack_input(g['range'], g['normal'], None, g['intensity'])
  assert np.array_equal(bits(x), bits(ref))
  dev = eng.device
  y = eng.pack_input(torch.from_numpy(g['range'])[None].to(dev), torch.from_numpy(g['normal'])[None].to(dev), None,
                     torch.from_numpy(g['intensity'])[None].to(dev)).cpu().numpy()[0]
  assert np.array_equal(bits(y), bits(ref))
  eng.close()


def test_projection_idempotent_and_order_independent(engine_fp32):
  """Size-independent properties at full scan size: shuffling the points only changes proj_idx
  (min-depth winner is order independent except exact-depth ties), and re-running is identical."""
  eng = engine_fp32
  pts = synth.kitti_like_cloud(77)
  a = eng.project(eng.upload_clouds([pts]))
  b = eng.project(eng.upload_clouds([pts]))
  for k in a:
    assert torch.equal(a[k], b[k])
  perm = np.random.default_rng(0).permutation(pts.shape[0])
  c = eng.project(eng.upload_clouds([pts[perm]]))
  assert torch.equal(a['range'], c['range'])


def test_reference_style_functions(tmp_path):
  """The drop-in module functions keep the reference's signatures, file names and formats."""
  from overlapnet_b200 import preprocess as pp
  g = load_golden('synth_5')
  rng, vert, inten, idx = pp.range_projection(g['points'])
  assert rng.dtype == np.float32 and idx.dtype == np.int32 and vert.shape == (64, 900, 4)
  assert np.array_equal(bits(rng), bits(g['range'])) and np.array_equal(idx, g['idx'])
  assert np.array_equal(bits(pp.gen_normal_map(rng, vert)), bits(g['normal']))
  scans = tmp_path / 'scans'
  scans.mkdir()
  sem_dir = tmp_path / 'sem'
  sem_dir.mkdir()
  g['points'].tofile(str(scans / '000042.bin'))
  probs = synth.random_probs(105, g['points'].shape[0])
  probs.tofile(str(sem_dir / '000042.label'))
  dst = tmp_path / 'out'
  dst.mkdir()
  d = pp.gen_depth_data(str(scans), str(dst))
  n = pp.gen_normal_data(str(scans), str(dst))
  it = pp.gen_intensity_data(str(scans), str(dst))
  s = pp.gen_semantic_data(str(sem_dir), str(scans), str(dst))
  assert np.array_equal(np.load(str(dst / 'depth' / '000000.npy')), g['range']) and np.array_equal(d[0], g['range'])
  assert np.array_equal(np.load(str(dst / 'normal' / '000000.npy')), g['normal']) and len(n) == 1
  assert np.array_equal(np.load(str(dst / 'intensity' / '000000.npy')), g['intensity']) and len(it) == 1
  sem = np.load(str(dst / 'semantic' / '000042.npy'))            # named by scan basename
  assert sem.shape == (64, 900, 20) and np.array_equal(sem[24:28], g['semantic_rows']) and len(s) == 1
