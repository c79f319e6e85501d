"""Multi-GPU loop-closure search: the candidate feature bank sharded across ranks
(one process per GPU, torch.distributed; NCCL on GPUs, gloo in the CPU tests).

The reference is single-process; its 1-vs-N entry point is ``Infer.infer_multiple``
(infer.py:162-203) and its many-vs-many entry point ``infer_multiple_vs_multiple`` (:205-238).
All units are independent, so the path shards with no data-path collective except:
  1 x N  : ONE broadcast of the query volume (360x128 fp32 = 184 320 B) and ONE gather of
           (overlap f32, yaw i32) per candidate;
  N x N  : ONE all_gather of the encoded bank, then rank r scores rows [lo_r, hi_r) of the ordered
           pair matrix against the full bank; one gather of the result rows.
The compute is injected (``heads_1vsN_fn(bank_local, query) -> (overlap, yaw)``) so that the
sharding logic is testable on CPU with the oracle standing in for the CUDA engine.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
  """Contiguous block [lo, hi) of rank ``rank`` when n items are split over ``world`` ranks."""
  base, rem = divmod(n, world)
  lo = rank * base + min(rank, rem)
  return lo, lo + base + (1 if rank < rem else 0)


class ShardedSearch:
  """heads_1vsN_fn(bank_local [m,360,128], query [360,128]) -> (overlap [m] f32, yaw [m] i32) on
  the same device as its inputs.  ``bank_local`` is this rank's contiguous block of the bank."""

  def __init__(self, heads_1vsN_fn, bank_local, n_total, group=None):
    self.fn = heads_1vsN_fn
    self.bank = bank_local
    self.n_total = int(n_total)
    self.group = group
    self.rank = dist.get_rank(group) if dist.is_initialized() else 0
    self.world = dist.get_world_size(group) if dist.is_initialized() else 1
    self.lo, self.hi = shard_range(self.n_total, self.rank, self.world)
    if bank_local.shape[0] != self.hi - self.lo:
      raise Exception('rank %d holds %d volumes, its shard is [%d, %d)' % (self.rank, bank_local.shape[0],
                                                                          self.lo, self.hi))
    self.max_shard = max(shard_range(self.n_total, r, self.world)[1] - shard_range(self.n_total, r, self.world)[0]
                         for r in range(self.world))

  def query(self, query_fv, src=0):
    """query_fv: [360,128] tensor (contents only matter on ``src``).  Returns on rank ``src``
    (overlap [n_total] f32, yaw [n_total] i32) in global candidate order; None elsewhere."""
    q = query_fv.contiguous()
    if self.world > 1:
      dist.broadcast(q, src, group=self.group)
    ov, yaw = self.fn(self.bank, q)
    if self.world == 1:
      return ov, yaw
    dev = q.device
    pad_ov = torch.zeros(self.max_shard, dtype=torch.float32, device=dev)
    pad_yaw = torch.zeros(self.max_shard, dtype=torch.int32, device=dev)
    pad_ov[:ov.numel()] = ov
    pad_yaw[:yaw.numel()] = yaw
    if self.rank == src:
      g_ov = [torch.empty_like(pad_ov) for _ in range(self.world)]
      g_yaw = [torch.empty_like(pad_yaw) for _ in range(self.world)]
    else:
      g_ov = g_yaw = None
    dist.gather(pad_ov, g_ov, dst=src, group=self.group)
    dist.gather(pad_yaw, g_yaw, dst=src, group=self.group)
    if self.rank != src:
      return None
    sizes = [shard_range(self.n_total, r, self.world) for r in range(self.world)]
    return (torch.cat([g_ov[r][:hi - lo] for r, (lo, hi) in enumerate(sizes)]),
            torch.cat([g_yaw[r][:hi - lo] for r, (lo, hi) in enumerate(sizes)]))

  def all_pairs(self, dst=0):
    """Ordered all-pairs matrix (the delta head is not symmetric): entry [i, j] is
    LEFT = bank[j], RIGHT = bank[i] (i.e. row i = query i against every candidate j).
    One all_gather of the bank; rank r fills rows [lo_r, hi_r).  Returns on ``dst``
    (overlap [n,n] f32, yaw [n,n] i32); None elsewhere."""
    dev = self.bank.device
    if self.world > 1:
      pad = torch.zeros((self.max_shard,) + tuple(self.bank.shape[1:]), dtype=self.bank.dtype, device=dev)
      pad[:self.bank.shape[0]] = self.bank
      parts = [torch.empty_like(pad) for _ in range(self.world)]
      dist.all_gather(parts, pad, group=self.group)
      sizes = [shard_range(self.n_total, r, self.world) for r in range(self.world)]
      full = torch.cat([parts[r][:hi - lo] for r, (lo, hi) in enumerate(sizes)])
    else:
      full = self.bank
    rows_ov = torch.zeros((self.max_shard, self.n_total), dtype=torch.float32, device=dev)
    rows_yaw = torch.zeros((self.max_shard, self.n_total), dtype=torch.int32, device=dev)
    for k in range(self.hi - self.lo):
      ov, yaw = self.fn(full, full[self.lo + k])
      rows_ov[k], rows_yaw[k] = ov, yaw
    if self.world == 1:
      return rows_ov[:self.n_total], rows_yaw[:self.n_total]
    if self.rank == dst:
      g_ov = [torch.empty_like(rows_ov) for _ in range(self.world)]
      g_yaw = [torch.empty_like(rows_yaw) for _ in range(self.world)]
    else:
      g_ov = g_yaw = None
    dist.gather(rows_ov, g_ov, dst=dst, group=self.group)
    dist.gather(rows_yaw, g_yaw, dst=dst, group=self.group)
    if self.rank != dst:
      return None
    sizes = [shard_range(self.n_total, r, self.world) for r in range(self.world)]
    return (torch.cat([g_ov[r][:hi - lo] for r, (lo, hi) in enumerate(sizes)]),
            torch.cat([g_yaw[r][:hi - lo] for r, (lo, hi) in enumerate(sizes)]))


def engine_heads_fn(engine):
  """Adapter: the CUDA engine as the compute of a ShardedSearch."""
  def fn(bank, query):
    ov, yaw, _ = engine.heads_1vsN(bank, query, n_cand=int(bank.shape[0]))
    return ov, yaw
  return fn
