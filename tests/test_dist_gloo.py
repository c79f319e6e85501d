"""world_size-2 gloo test of the sharded search logic (broadcast query, per-shard scoring, gather in
global order) with the CPU oracle standing in for the CUDA engine."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT  # noqa: F401


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _cheap_heads(bank, query):
  """A cheap, order-revealing stand-in with the same signature as the engine's 1-vs-N call:
  'overlap' = sigmoid of a bank/query inner product, 'yaw' = 180 - argmax of the circular
  correlation computed with the oracle's definition (small C keeps it fast)."""
  from oracle import network as N
  b = bank.numpy()
  q = query.numpy()
  ov = 1.0 / (1.0 + np.exp(-(b * q[None]).sum(axis=(1, 2)) / b[0].size))
  corr = np.stack([N.correlation_naive(b[i], q) for i in range(b.shape[0])]) if b.shape[0] else np.zeros((0, b.shape[1]))
  yaw = b.shape[1] // 2 - corr.argmax(axis=1) if b.shape[0] else np.zeros((0,), np.int64)
  return torch.from_numpy(ov.astype(np.float32)), torch.from_numpy(yaw.astype(np.int32))


def _make_bank(n, W=12, C=4):
  rng = np.random.default_rng(0)
  return np.abs(rng.standard_normal((n, W, C))).astype(np.float32)


def _worker(rank, world, port, n_total, out_path):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  from overlapnet_b200.search import ShardedSearch, shard_range
  bank = _make_bank(n_total)
  lo, hi = shard_range(n_total, rank, world)
  ss = ShardedSearch(_cheap_heads, torch.from_numpy(bank[lo:hi].copy()), n_total)
  q = torch.from_numpy(bank[3].copy()) if rank == 0 else torch.zeros(bank.shape[1:])   # only rank 0 knows it
  res = ss.query(q)
  ap = ss.all_pairs()
  if rank == 0:
    np.savez(out_path, ov=res[0].numpy(), yaw=res[1].numpy(), ap_ov=ap[0].numpy(), ap_yaw=ap[1].numpy())
  else:
    assert res is None and ap is None
  dist.barrier()
  dist.destroy_process_group()


def test_shard_range_partitions():
  from overlapnet_b200.search import shard_range
  for n in (0, 1, 7, 1101, 4541):
    for world in (1, 2, 4, 8):
      r = [shard_range(n, k, world) for k in range(world)]
      assert r[0][0] == 0 and r[-1][1] == n
      assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
      assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
  assert shard_range(4541, 0, 8) == (0, 568) and shard_range(4541, 7, 8) == (3974, 4541)


def test_sharded_search_world2_matches_single_process(tmp_path):
  n_total = 7                          # odd: shards of 4 and 3
  out = str(tmp_path / 'res.npz')
  port = _free_port()
  mp.spawn(_worker, args=(2, port, n_total, out), nprocs=2, join=True)
  got = np.load(out)
  bank = _make_bank(n_total)
  ov, yaw = _cheap_heads(torch.from_numpy(bank), torch.from_numpy(bank[3]))
  assert np.array_equal(got['ov'], ov.numpy()) and np.array_equal(got['yaw'], yaw.numpy())
  assert got['yaw'][3] == 0            # the query against itself
  for i in range(n_total):
    o, y = _cheap_heads(torch.from_numpy(bank), torch.from_numpy(bank[i]))
    assert np.array_equal(got['ap_ov'][i], o.numpy()) and np.array_equal(got['ap_yaw'][i], y.numpy())


# ---- the growing, sharded loop-closure bank (search.ShardedBank) under gloo ----------------------
def _bank_worker(rank, world, port, out_path):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  from overlapnet_b200.search import ShardedBank
  from overlapnet_b200 import lcd
  from conftest import GOLDEN
  g = np.load(os.path.join(GOLDEN, 'lcd_demo3.npz'))
  W, C = 12, 4
  rng = np.random.default_rng(0)
  vols = np.abs(rng.standard_normal((len(g['traj']), W, C))).astype(np.float32)     # "encoded" frame i
  local = []                                                                       # this rank's rows

  def heads(local_rows, q):
    bank = torch.from_numpy(np.stack([local[int(r)] for r in local_rows.tolist()]))
    return _cheap_heads(bank, q)

  sb = ShardedBank(lambda fid: torch.from_numpy(vols[fid]), lambda fv: local.append(fv.numpy().copy()), heads,
                   (W, C), torch.device('cpu'))

  class Adapter:                       # what lcd.LoopClosureDetector needs of an Infer
    def __init__(self):
      self.results = []

    def infer_multiple(self, idx, refs):
      res = sb.step(idx, refs)
      if res is None:
        if rank != 0 and len(refs) > 0:
          return np.zeros(len(refs), np.float32), np.zeros(len(refs), np.int64)    # non-source ranks only follow
        return None
      ov, yaw = res
      self.results.append((idx, list(map(int, refs)), ov.numpy().copy(), yaw.numpy().copy()))
      return ov.numpy(), yaw.numpy()

  ad = Adapter()
  det = lcd.LoopClosureDetector(ad, overlap_thres=2.0)       # never "finds" one: every frame is scored and appended
  for i in range(260):
    det.step(i, g['traj'][i], g['covs'][i])
  assert len(local) == len(range(rank, 260, world))          # frame i lives on rank i % world
  for k, v in enumerate(local):
    assert np.array_equal(v, vols[rank + k * world])
  if rank == 0:
    np.savez(out_path, n=len(ad.results), idx=np.array([r[0] for r in ad.results]),
             refs=np.concatenate([np.array(r[1]) for r in ad.results]),
             ov=np.concatenate([r[2] for r in ad.results]), yaw=np.concatenate([r[3] for r in ad.results]),
             offs=np.cumsum([0] + [len(r[1]) for r in ad.results]))
  dist.barrier()
  dist.destroy_process_group()


def test_sharded_growing_bank_world2_matches_single_process(tmp_path):
  """Frames 0..259 of the golden LCD run through ShardedBank on 2 ranks: every scored candidate gets
  the value a single process computes from the full bank, in the caller's candidate order."""
  from conftest import GOLDEN
  out = str(tmp_path / 'bank.npz')
  mp.spawn(_bank_worker, args=(2, _free_port(), out), nprocs=2, join=True)
  got = np.load(out)
  g = np.load(os.path.join(GOLDEN, 'lcd_demo3.npz'))
  rng = np.random.default_rng(0)
  vols = np.abs(rng.standard_normal((len(g['traj']), 12, 4))).astype(np.float32)
  assert got['n'] > 20
  for k in range(int(got['n'])):
    idx = int(got['idx'][k])
    refs = got['refs'][got['offs'][k]:got['offs'][k + 1]]
    ov, yaw = _cheap_heads(torch.from_numpy(vols[refs]), torch.from_numpy(vols[idx]))
    assert np.array_equal(got['ov'][got['offs'][k]:got['offs'][k + 1]], ov.numpy())
    assert np.array_equal(got['yaw'][got['offs'][k]:got['offs'][k + 1]], yaw.numpy())
