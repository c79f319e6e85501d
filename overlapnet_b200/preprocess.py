"""GPU drop-ins for the reference's preprocessing functions (same names, arguments and return
values as ``src/utils/utils.py`` and ``src/utils/gen_*_data.py``), executed by the CUDA kernels
in csrc/projection.cu through the C ABI.  There is no CPU fallback."""
import os

import numpy as np
import torch

from .engine import Engine

_engines = {}


def _engine(fov_up, fov_down, proj_H, proj_W, max_range=50):
  """One projection-only handle per (fov, image size, device); max_range is a per-call argument."""
  key = (float(fov_up), float(fov_down), int(proj_H), int(proj_W), torch.cuda.current_device())
  if key not in _engines:
    _engines[key] = Engine(precision='fp32', proj_H=proj_H, proj_W=proj_W, fov_up=fov_up, fov_down=fov_down,
                           max_range=50.0, max_batch_scans=16, max_batch_pairs=1,
                           model={'additional_unsymmetric_layer3a': True})
  return _engines[key]


def range_projection(current_vertex, fov_up=3.0, fov_down=-25.0, proj_H=64, proj_W=900, max_range=50):
  """Drop-in for ``range_projection`` (utils.py:59-134): (N,4) float32 cloud ->
  (proj_range (H,W) f32, proj_vertex (H,W,4) f32, proj_intensity (H,W) f32, proj_idx (H,W) i32)."""
  pts = np.ascontiguousarray(current_vertex, dtype=np.float32)
  if pts.ndim != 2 or pts.shape[1] != 4:
    raise ValueError('range_projection expects an (N, 4) array [x, y, z, intensity]')
  eng = _engine(fov_up, fov_down, proj_H, proj_W, max_range)
  batch = eng.upload_clouds([pts])
  out = eng.project(batch, max_range=float(max_range))
  return (out['range'][0].cpu().numpy(), out['vertex'][0].cpu().numpy(), out['intensity'][0].cpu().numpy(),
          out['idx'][0].cpu().numpy())


def gen_normal_map(current_range, current_vertex, proj_H=64, proj_W=900):
  """Drop-in for ``gen_normal_map`` (utils.py:137-175)."""
  eng = _engine(3.0, -25.0, proj_H, proj_W, 50)
  r = torch.from_numpy(np.ascontiguousarray(current_range, np.float32)).to(eng.device)[None]
  v = torch.from_numpy(np.ascontiguousarray(current_vertex, np.float32)).to(eng.device)[None]
  return eng.normals(r, v)[0].cpu().numpy()


def load_files(folder):
  """utils.py:233-239."""
  file_paths = [os.path.join(dp, f) for dp, dn, fn in os.walk(os.path.expanduser(folder)) for f in fn]
  file_paths.sort()
  return file_paths


def _read_scan(path):
  return np.fromfile(path, dtype=np.float32).reshape((-1, 4))   # gen_depth_data.py:31-32


def _dst(dst_folder, sub):
  d = os.path.join(dst_folder, sub)
  try:
    os.stat(d)
    print('generating %s data in: ' % sub, d)
  except OSError:
    print('creating new %s folder: ' % sub, d)
    os.mkdir(d)
  return d


def _batched(scan_paths, eng):
  for s0 in range(0, len(scan_paths), eng.max_batch_scans):
    paths = scan_paths[s0:s0 + eng.max_batch_scans]
    yield s0, eng.upload_clouds([_read_scan(p) for p in paths])


def gen_depth_data(scan_folder, dst_folder, normalize=False):
  """Drop-in for gen_depth_data.py:10-48: writes ``<dst>/depth/%06d.npy`` (64,900) f32."""
  dst = _dst(dst_folder, 'depth')
  scan_paths = load_files(scan_folder)
  eng = _engine(3.0, -25.0, 64, 900, 50)
  depths = []
  for s0, batch in _batched(scan_paths, eng):
    rng = eng.project(batch, want=('range',))['range'].cpu().numpy()
    for i in range(rng.shape[0]):
      proj_range = rng[i]
      if normalize:
        proj_range = proj_range / np.max(proj_range)              # gen_depth_data.py:37-38
      dst_path = os.path.join(dst, str(s0 + i).zfill(6))
      np.save(dst_path, proj_range)
      depths.append(proj_range)
      print('finished generating depth data at: ', dst_path)
  return depths


def gen_normal_data(scan_folder, dst_folder):
  """Drop-in for gen_normal_data.py:10-46: writes ``<dst>/normal/%06d.npy`` (64,900,3) f32."""
  dst = _dst(dst_folder, 'normal')
  scan_paths = load_files(scan_folder)
  eng = _engine(3.0, -25.0, 64, 900, 50)
  normals = []
  for s0, batch in _batched(scan_paths, eng):
    out = eng.project(batch, want=('range', 'vertex'))
    nrm = eng.normals(out['range'], out['vertex']).cpu().numpy()
    for i in range(nrm.shape[0]):
      dst_path = os.path.join(dst, str(s0 + i).zfill(6))
      np.save(dst_path, nrm[i])
      normals.append(nrm[i])
      print('finished generating intensity data at: ', dst_path)   # sic, gen_normal_data.py:43
  return normals


def gen_intensity_data(scan_folder, dst_folder):
  """Drop-in for gen_intensity_data.py:10-43: writes ``<dst>/intensity/%06d.npy`` (64,900) f32."""
  dst = _dst(dst_folder, 'intensity')
  scan_paths = load_files(scan_folder)
  eng = _engine(3.0, -25.0, 64, 900, 50)
  intensities = []
  for s0, batch in _batched(scan_paths, eng):
    inten = eng.project(batch, want=('intensity',))['intensity'].cpu().numpy()
    for i in range(inten.shape[0]):
      dst_path = os.path.join(dst, str(s0 + i).zfill(6))
      np.save(dst_path, inten[i])
      intensities.append(inten[i])
      print('finished generating intensity data at: ', dst_path)
  return intensities


def gen_semantic_data(semantic_folder, scan_folder, dst_folder, proj_H=64, proj_W=900):
  """Drop-in for gen_semantic_data.py:11-57: writes ``<dst>/semantic/<scan basename>.npy``
  (64,900,20) f32; projection with max_range=inf and the filtered-index gather of the reference."""
  dst = _dst(dst_folder, 'semantic')
  prob_paths = load_files(semantic_folder)
  scan_paths = load_files(scan_folder)
  eng = _engine(3.0, -25.0, proj_H, proj_W, 50)
  semantics = []
  for idx in range(len(prob_paths)):
    probs = np.fromfile(prob_paths[idx], dtype=np.float32).reshape((-1, 20))
    batch = eng.upload_clouds([_read_scan(scan_paths[idx])])
    pidx = eng.project(batch, max_range=float('inf'), want=('idx',))['idx']
    d_probs = torch.from_numpy(probs).to(eng.device)
    proj_prob = eng.semantic(pidx, d_probs, batch.offsets)[0].cpu().numpy()
    base_name = os.path.basename(scan_paths[idx]).replace('.bin', '')
    dst_path = os.path.join(dst, base_name)
    np.save(dst_path, proj_prob)
    semantics.append(proj_prob)
    print('finished generating semantic data at: ', dst_path)
  return semantics
