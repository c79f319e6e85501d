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
