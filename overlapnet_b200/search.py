"""Multi-GPU loop-closure search: the candidate feature bank sharded across ranks
(one process per GPU, torch.distributed; NCCL on GPUs, gloo in the CPU tests).

The reference is single-process; its 1-vs-N entry point is ``Infer.infer_multiple``
(infer.py:162-203) and its many-vs-many entry point ``infer_multiple_vs_multiple`` (:205-238).
All units are independent, so the path shards with no data-path collective except moving the
query volume (360x128 fp32 = 184 320 B) to every rank and the (overlap f32, yaw i32) records back:

  transport 'collective' : ONE broadcast + ONE gather of packed 8-byte records (works on gloo / NCCL);
  transport 'symm'       : no collective at all.  The query volume and the result table live in
      symmetric (peer-mapped) memory: every rank's kernels READ the query straight from the source
      rank's buffer over NVLink and the kernels that finish a pair (k_dense_finalize /
      k_corr_finalize) STORE overlap / yaw straight into the source rank's result table; the only
      synchronisation is one device-side signal per peer each way (torch symmetric-memory signal
      pads, stream-ordered, no host involvement).

  N x N  : ONE all_gather of the encoded bank, then rank r scores rows [lo_r, hi_r) of the ordered
           pair matrix against the full bank; one gather of the result rows.

The compute is injected (``heads_1vsN_fn(bank_local, query[, out=(ov, yaw)]) -> (overlap, yaw)``)
so that the sharding logic is testable on CPU with the oracle standing in for the CUDA engine.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
  """Contiguous block [lo, hi) of rank ``rank`` when n items are split over ``world`` ranks."""
  base, rem = divmod(n, world)
  lo = rank * base + min(rank, rem)
  return lo, lo + base + (1 if rank < rem else 0)


def pack_records(ov, yaw):
  """(overlap f32 [m], yaw i32 [m]) -> int32 [m, 2]: one 8-byte record per candidate."""
  return torch.stack([ov.contiguous().view(torch.int32), yaw.to(torch.int32)], dim=1).contiguous()


def unpack_records(rec):
  return rec[:, 0].contiguous().view(torch.float32), rec[:, 1].contiguous()


def balanced_sizes(n, world, src, src_discount):
  """Shard sizes when rank ``src`` also encodes the query: it gets ``src_discount`` fewer candidates
  (= encode time / time per candidate), the rest is spread evenly.  Returns a list of ``world`` sizes."""
  d = int(max(0, min(src_discount, n // max(world, 1))))
  if world == 1:
    return [n]
  base, rem = divmod(n + d, world)
  sizes = [base + (1 if r < rem else 0) for r in range(world)]
  sizes[src] -= d
  if sizes[src] < 0:
    return [hi - lo for lo, hi in (shard_range(n, r, world) for r in range(world))]
  assert sum(sizes) == n
  return sizes


class _SymmTransport:
  """Peer-mapped query buffer + result table + step flags (torch.distributed._symmetric_memory)."""

  def __init__(self, n_slots, fv_shape, device, group):
    import torch.distributed._symmetric_memory as symm_mem
    self.group = group if group is not None else dist.group.WORLD
    self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
    self.fv_shape = tuple(fv_shape)
    self.n_slots = int(n_slots)
    qn = int(np.prod(self.fv_shape))
    self.q_words = qn
    # one allocation: [query fp32 | overlap fp32 x world x n_slots | yaw i32 x world x n_slots | flags i32 x 2 x world]
    self.flag_off = qn + 2 * self.world * self.n_slots
    total = self.flag_off + 2 * self.world
    self.buf = symm_mem.empty(total, dtype=torch.float32, device=device)
    self.hdl = symm_mem.rendezvous(self.buf, self.group)
    self.buf.zero_()
    self.hdl.barrier(channel=2)
    self.step = 0
    # flags[c][r] on rank d = "rank r finished step <value> of channel c" (c 0: query in place, c 1: results stored)
    self.my_flags = [self.hdl.get_buffer(self.rank, (self.world,), torch.int32, self.flag_off + c * self.world)
                     for c in range(2)]
    self.peer_flag_addr = [[self.hdl.get_buffer(d, (1,), torch.int32, self.flag_off + c * self.world + self.rank).data_ptr()
                            for d in range(self.world)] for c in range(2)]

  def query_view(self, rank):
    return self.hdl.get_buffer(rank, self.fv_shape, torch.float32, 0)

  def result_views(self, rank, slot_rank):
    """Views into ``rank``'s result table for the slots of ``slot_rank``."""
    off = self.q_words + slot_rank * self.n_slots
    ov = self.hdl.get_buffer(rank, (self.n_slots,), torch.float32, off)
    yaw = self.hdl.get_buffer(rank, (self.n_slots,), torch.int32, off + self.world * self.n_slots)
    return ov, yaw


class ShardedSearch:
  """heads_1vsN_fn(bank_local [m,360,128], query [360,128]) -> (overlap [m] f32, yaw [m] i32) on
  the same device as its inputs.  ``bank_local`` is this rank's contiguous block of the bank.
  With ``transport='symm'`` the function must also accept ``out=(overlap, yaw)`` tensors to write into."""

  def __init__(self, heads_1vsN_fn, bank_local, n_total, group=None, transport='collective', sizes=None):
    self.fn = heads_1vsN_fn
    self.bank = bank_local
    self.n_total = int(n_total)
    self.group = group
    self.rank = dist.get_rank(group) if dist.is_initialized() else 0
    self.world = dist.get_world_size(group) if dist.is_initialized() else 1
    if sizes is None:                      # contiguous blocks, equal up to one (``sizes``: explicit, e.g. balanced_sizes)
      self.sizes = [shard_range(self.n_total, r, self.world) for r in range(self.world)]
    else:
      offs = np.concatenate([[0], np.cumsum(sizes)])
      assert len(sizes) == self.world and int(offs[-1]) == self.n_total
      self.sizes = [(int(offs[r]), int(offs[r + 1])) for r in range(self.world)]
    self.lo, self.hi = self.sizes[self.rank]
    if bank_local.shape[0] != self.hi - self.lo:
      raise Exception('rank %d holds %d volumes, its shard is [%d, %d)' % (self.rank, bank_local.shape[0],
                                                                          self.lo, self.hi))
    self.max_shard = max(hi - lo for lo, hi in self.sizes)
    # own signalling kernels (ovn_peer_signal / ovn_peer_wait: one launch each) when the compute is the engine
    self.engine = getattr(heads_1vsN_fn, 'engine', None)
    self.transport = 'collective'
    self.symm = None
    if transport in ('symm', 'auto') and self.world > 1 and bank_local.is_cuda:
      ok = 1
      try:
        self.symm = _SymmTransport(self.max_shard, bank_local.shape[1:], bank_local.device, group)
      except Exception as e:                      # no peer access / old torch: fall back to the collectives
        ok = 0
        self.symm_error = repr(e)
      flag = torch.tensor([ok], dtype=torch.int32, device=bank_local.device)
      dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)       # every rank must take the same path
      if int(flag.item()) == 1:
        self.transport = 'symm'
      else:
        self.symm = None
        if transport == 'symm':
          raise Exception('symmetric-memory transport unavailable: %s' % getattr(self, 'symm_error', 'on another rank'))
    if self.world > 1 and self.transport == 'collective':
      dev = bank_local.device
      self._pad = torch.zeros((self.max_shard, 2), dtype=torch.int32, device=dev)
      self._gathered = [torch.empty_like(self._pad) for _ in range(self.world)] if self.rank == 0 else None

  # ---- 1 x N ---------------------------------------------------------------------------------
  def query(self, query_fv, src=0):
    """query_fv: [360,128] tensor (contents only matter on ``src``).  Returns on rank ``src``
    (overlap [n_total] f32, yaw [n_total] i32) in global candidate order; None elsewhere."""
    if self.world == 1:
      return self.fn(self.bank, query_fv.contiguous())
    if self.transport == 'symm':
      return self._query_symm(query_fv, src)
    q = query_fv.contiguous()
    dist.broadcast(q, src, group=self.group)                     # collective 1 of 2
    ov, yaw = self.fn(self.bank, q)
    m = ov.numel()
    self._pad[:m] = pack_records(ov, yaw)
    if self.rank == src and self._gathered is None:
      self._gathered = [torch.empty_like(self._pad) for _ in range(self.world)]
    dist.gather(self._pad, self._gathered if self.rank == src else None, dst=src, group=self.group)   # 2 of 2
    if self.rank != src:
      return None
    rec = torch.cat([self._gathered[r][:hi - lo] for r, (lo, hi) in enumerate(self.sizes)])
    return unpack_records(rec)

  def _query_symm(self, query_fv, src):
    s = self.symm
    m = self.hi - self.lo
    s.step += 1
    eng = self.engine
    if self.rank == src:
      s.query_view(src).copy_(query_fv)                            # local store into the symmetric buffer
      if eng is not None:                                          # "query k is in place": ONE launch for all peers
        eng.peer_signal([a for d, a in enumerate(s.peer_flag_addr[0]) if d != src], s.step)
      else:
        for r in range(self.world):
          if r != src:
            s.hdl.put_signal(r, channel=0)
    elif eng is not None:
      eng.peer_wait(s.my_flags[0][src:src + 1], 1, -1, s.step)
    else:
      s.hdl.wait_signal(src, channel=0)
    q = s.query_view(src)                                          # peers read it over NVLink inside their kernels
    ov_out, yaw_out = s.result_views(src, self.rank)               # rows of the SOURCE rank's table
    self.fn(self.bank, q, out=(ov_out[:m], yaw_out[:m]))
    if self.rank != src:
      if eng is not None:                                          # "my results are in your table"
        eng.peer_signal([s.peer_flag_addr[1][src]], s.step)
      else:
        s.hdl.put_signal(src, channel=1)
      return None
    if eng is not None:
      eng.peer_wait(s.my_flags[1], self.world, src, s.step)       # ONE launch waits for every peer
    else:
      for r in range(self.world):
        if r != src:
          s.hdl.wait_signal(r, channel=1)
    ov_all, yaw_all = [], []
    for r, (lo, hi) in enumerate(self.sizes):
      o, y = s.result_views(src, r)
      ov_all.append(o[:hi - lo])
      yaw_all.append(y[:hi - lo])
    return torch.cat(ov_all), torch.cat(yaw_all)

  # ---- N x N ---------------------------------------------------------------------------------
  def gather_bank(self):
    """The whole bank on every rank (ONE all_gather)."""
    if self.world == 1:
      return self.bank
    dev = self.bank.device
    pad = torch.zeros((self.max_shard,) + tuple(self.bank.shape[1:]), dtype=self.bank.dtype, device=dev)
    pad[:self.bank.shape[0]] = self.bank
    parts = [torch.empty_like(pad) for _ in range(self.world)]
    dist.all_gather(parts, pad, group=self.group)
    return torch.cat([parts[r][:hi - lo] for r, (lo, hi) in enumerate(self.sizes)])

  def all_pairs(self, dst=0, rows_fn=None, gather=True):
    """Ordered all-pairs matrix (the delta head is not symmetric): entry [i, j] is
    LEFT = bank[j], RIGHT = bank[i] (i.e. row i = query i against every candidate j).
    One all_gather of the bank; rank r fills rows [lo_r, hi_r).  ``rows_fn(full_bank, lo, hi) ->
    (overlap [hi-lo, n], yaw [hi-lo, n])`` scores a block of rows in one call (the engine loops over
    the rows inside the C ABI); without it the rows are scored one ``fn`` call at a time.
    Returns on ``dst`` (overlap [n,n] f32, yaw [n,n] i32); None elsewhere (``gather=False``: every
    rank returns its own row block)."""
    dev = self.bank.device
    full = self.gather_bank()
    if rows_fn is not None:
      blk_ov, blk_yaw = rows_fn(full, self.lo, self.hi)
    else:
      blk_ov = torch.zeros((self.hi - self.lo, self.n_total), dtype=torch.float32, device=dev)
      blk_yaw = torch.zeros((self.hi - self.lo, self.n_total), dtype=torch.int32, device=dev)
      for k in range(self.hi - self.lo):
        ov, yaw = self.fn(full, full[self.lo + k])
        blk_ov[k], blk_yaw[k] = ov, yaw
    if self.world == 1 or not gather:
      return blk_ov, blk_yaw
    rows = torch.zeros((self.max_shard, self.n_total, 2), dtype=torch.int32, device=dev)
    rows[:self.hi - self.lo, :, 0] = blk_ov.contiguous().view(torch.int32)
    rows[:self.hi - self.lo, :, 1] = blk_yaw
    g = [torch.empty_like(rows) for _ in range(self.world)] if self.rank == dst else None
    dist.gather(rows, g, dst=dst, group=self.group)
    if self.rank != dst:
      return None
    rec = torch.cat([g[r][:hi - lo] for r, (lo, hi) in enumerate(self.sizes)])
    return rec[..., 0].contiguous().view(torch.float32), rec[..., 1].contiguous()


class ShardedBank:
  """A GROWING bank sharded over the ranks for the online loop-closure flow (demo3_lcd.py:85-123 on
  top of Infer.infer_multiple, infer.py:162-203): frame ``i`` lives on rank ``i % world`` at local row
  ``i // world``.  Every rank runs the same driver loop (SPMD):

    encode_fn(frame_id) -> volume [360,128]   (called on rank ``src`` only; the volume is broadcast)
    append_fn(volume)                          (called on the owning rank: stores it as the next local row)
    heads_fn(local_rows int32 tensor, query volume) -> (overlap, yaw)   scores this rank's candidates

  ``step(frame_id, reference_ids)`` returns on ``src`` (overlap, yaw) in the order of ``reference_ids``
  (None when the list is empty); other ranks get None.  Two collectives per scored frame: one
  broadcast (query volume) and one gather (8-byte records)."""

  def __init__(self, encode_fn, append_fn, heads_fn, fv_shape, device, group=None, src=0):
    self.encode_fn, self.append_fn, self.heads_fn = encode_fn, append_fn, heads_fn
    self.group, self.src = group, src
    self.device = device
    self.fv_shape = tuple(fv_shape)
    self.rank = dist.get_rank(group) if dist.is_initialized() else 0
    self.world = dist.get_world_size(group) if dist.is_initialized() else 1
    self.n_frames = 0

  def owner(self, frame_id):
    return int(frame_id) % self.world

  def step(self, frame_id, reference_ids, on_query=None):
    """``on_query(frame_id, volume)`` is called on every rank once the current frame's volume is there."""
    if int(frame_id) != self.n_frames:
      raise Exception('frames must arrive in order 0,1,2,... (got %d, expected %d)' % (frame_id, self.n_frames))
    if self.rank == self.src:
      q = self.encode_fn(frame_id).reshape(self.fv_shape).contiguous()
    else:
      q = torch.empty(self.fv_shape, dtype=torch.float32, device=self.device)
    if self.world > 1:
      dist.broadcast(q, self.src, group=self.group)
    if on_query is not None:
      on_query(int(frame_id), q)
    if self.owner(frame_id) == self.rank:
      self.append_fn(q)
    self.n_frames += 1
    refs = np.asarray(reference_ids, dtype=np.int64).reshape(-1)
    if refs.size == 0:
      return None
    if refs.min() < 0 or refs.max() >= self.n_frames:
      raise IndexError('reference frame id out of range')
    mine = np.nonzero(refs % self.world == self.rank)[0]                 # positions in the caller's list
    local_rows = torch.from_numpy((refs[mine] // self.world).astype(np.int32)).to(self.device)
    if mine.size:
      ov, yaw = self.heads_fn(local_rows, q)
    else:
      ov = torch.empty((0,), dtype=torch.float32, device=self.device)
      yaw = torch.empty((0,), dtype=torch.int32, device=self.device)
    if self.world == 1:
      return ov, yaw
    counts = [int(np.count_nonzero(refs % self.world == r)) for r in range(self.world)]
    pad = torch.zeros((max(counts), 2), dtype=torch.int32, device=self.device)
    pad[:mine.size] = pack_records(ov, yaw)
    g = [torch.empty_like(pad) for _ in range(self.world)] if self.rank == self.src else None
    dist.gather(pad, g, dst=self.src, group=self.group)
    if self.rank != self.src:
      return None
    out = torch.zeros((refs.size, 2), dtype=torch.int32, device=self.device)
    for r in range(self.world):
      pos = np.nonzero(refs % self.world == r)[0]
      if pos.size:
        out[torch.from_numpy(pos).to(self.device)] = g[r][:pos.size]
    return unpack_records(out)


def engine_heads_fn(engine):
  """Adapter: the CUDA engine as the compute of a ShardedSearch."""
  def fn(bank, query, out=None):
    ov, yaw, _ = engine.heads_1vsN(bank, query, n_cand=int(bank.shape[0]), out=out)
    return ov, yaw
  fn.engine = engine
  return fn


def engine_rows_fn(engine):
  """Adapter: a block of all-pairs rows in one C-ABI call (ovn_heads_rows_vs_bank)."""
  def fn(full_bank, lo, hi):
    return engine.heads_rows_vs_bank(full_bank, lo, hi)
  return fn
