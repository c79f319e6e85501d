#!/usr/bin/env python3
"""bench.py -- BASELINE.json's metric: scan-pairs/s (overlap + yaw) of the 1-query-vs-N-candidate
loop-closure search at 64x900, on synthetic KITTI-shaped data.

  python bench.py --gpus N --steps K --warmup W            # this repo's B200 path
  python bench.py --impl reference --steps K --warmup W    # the CPU port of the reference, host cores
  (N > 1: python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...)

A step = one query scan through the whole hot path:
  raw cloud (124 668 pts) -> projection + normals (64x900x4) -> leg -> 1 x 1101 delta + correlation
  heads over the rank's candidate bank -> (overlap, yaw) per candidate.
N > 1 (weak scaling, BASELINE config "bank sharded across GPUs"): every rank holds its own
1101-candidate shard; rank 0 encodes the query, ONE NCCL broadcast ships the 184 KB query volume,
every rank scores its shard, ONE gather returns (overlap, yaw).  value = all ranks' pairs / time.

Printed line: value = pairs/s with the query cloud already in HBM; e2e = the same through the
host-buffer C-ABI call (H2D of the cloud and D2H of the results inside the timed region);
roofline = k_delta_conv1_tc (the dominant kernel) against the measured bf16 tensor peak;
cpu_baseline = oracle (torch-CPU fp32 port of the reference graph) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

MODEL = {'additional_unsymmetric_layer3a': True, 'strides_layer1': [2, 2]}
N_CAND = 1101                     # KITTI-07 length (BASELINE config 2)
N_SRC_SCANS = 32                  # distinct synthetic scans behind the bank (yaw-rolled to 1101 volumes)
FLOP_DELTA_CONV1 = 2 * 1061683200 # per pair, c_conv1
FLOP_PAIR = 2 * (1275323392 + 16588800)
METRIC = 'scan-pairs/sec (overlap+yaw) at 64x900'


def parse():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=20)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  ap.add_argument('--precision', default='f16_tc', choices=['f16_tc', 'fp32'])
  ap.add_argument('--cpu-pairs', type=int, default=24, help='pairs in the cpu_baseline sample')
  return ap.parse_args()


def peaks():
  p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(p):
    with open(p) as f:
      d = json.load(f)
    return d.get('bf16_tflops_sustained', d.get('bf16_tflops')), d.get('hbm_gbs'), 'measured (MEASURED_PEAKS.json, sustained bf16)'
  return 1400.0, 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler(threading.Thread):
  """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

  def __init__(self, index=0):
    super().__init__(daemon=True)
    self.index, self.rows, self.proc = index, [], None

  def wait_first_sample(self, timeout=5.0):
    t0 = time.time()
    while not self.rows and time.time() - t0 < timeout:
      time.sleep(0.02)

  def run(self):
    q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
    try:
      self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
                                    '--format=csv,noheader,nounits', '-lms', '20'],
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      for line in self.proc.stdout:
        self.rows.append([x.strip() for x in line.split(',')])
    except Exception:
      pass

  def stop(self):
    if self.proc:
      self.proc.terminate()
    sm, mx, reasons = [], 0, set()
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    for r in self.rows:
      try:
        sm.append(float(r[0]))
        mx = max(mx, float(r[1]))
        for i, n in enumerate(names):
          if r[2 + i].lower().startswith('active'):
            reasons.add(n)
      except Exception:
        continue
    busy = sorted(s for s in sm if s > 0.5 * mx) or sorted(sm)
    return {'sm_mhz': busy[len(busy) // 2] if busy else None, 'sm_max_mhz': mx or None,
            'reasons': sorted(reasons), 'samples': len(sm)}


def make_weights():
  from oracle import network as onet      # seeded Glorot weights keyed by the Keras layer names
  return onet.glorot_weights(4, MODEL, seed=0)


def cpu_pairs_per_s(w, n_pairs, threads=None):
  """The oracle port of the reference graph on the host cores: 1 projection + 1 leg + n_pairs heads
  (delta tensor materialised like Keras, batch 16 like config/network.yml:41), torch fp32."""
  from oracle import network as onet
  from oracle import projection as oproj
  from overlapnet_b200 import synth
  if threads:
    torch.set_num_threads(threads)
  cloud = synth.kitti_like_cloud(4242)
  bank = synth.feature_volumes(7, n_pairs)
  t0 = time.perf_counter()
  rng, vert, _, _ = oproj.range_projection(cloud)
  x = oproj.pack_input(rng, oproj.gen_normal_map(rng, vert))[None]
  q = onet.leg_forward(x, w, MODEL, dtype=torch.float32)
  onet.heads_forward(bank, np.repeat(q, n_pairs, 0), w, MODEL, dtype=torch.float32, batch=16)
  dt = time.perf_counter() - t0
  return n_pairs / dt, dt


def run_reference(args):
  """--impl reference: the reference's own implementation is Python/TF and cannot be installed
  offline, so this arm times the oracle port on all host cores (kind = "port")."""
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  w = make_weights()
  cores = os.cpu_count()
  torch.set_num_threads(cores)
  n = args.cpu_pairs
  for _ in range(min(args.warmup, 1)):
    cpu_pairs_per_s(w, 4)
  t = []
  steps = max(1, min(args.steps, 5))
  for _ in range(steps):
    t.append(cpu_pairs_per_s(w, n)[1])
  dt = float(np.mean(t))
  val = n / dt
  sample = '%d of the 1101 pairs per step (+1 projection, +1 leg), torch-CPU fp32, delta tensor materialised' % n
  print(json.dumps({
      'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'pairs/s', 'n_gpus': args.gpus, 'steps': steps,
      'warmup': min(args.warmup, 1), 'ms_per_step': dt * 1e3 * N_CAND / n, 'higher_is_better': True,
      'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': '1 query x 1101 candidates, geo-only 64x900 (BASELINE config 2), bounded sample'},
      'cpu_baseline': {'value': val, 'unit': 'pairs/s', 'cores': cores, 'kind': 'port', 'sample': sample},
      'e2e': {'value': val, 'unit': 'pairs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
  }))


def main():
  args = parse()
  if args.impl == 'reference':
    run_reference(args)
    return
  import torch.distributed as dist
  from overlapnet_b200 import synth
  from overlapnet_b200.engine import Engine

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  if args.gpus != world:
    if world == 1 and args.gpus > 1:
      raise SystemExit('launch with torch.distributed.run --nproc-per-node %d for --gpus %d' % (args.gpus, args.gpus))
  torch.cuda.set_device(local)
  dev = torch.device('cuda', local)
  if world > 1:
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('NCCL_DEBUG', 'WARN')            # keep rank 0's stdout to the one JSON line
    dist.init_process_group('nccl', device_id=dev)

  w = make_weights()
  eng = Engine(model=MODEL, precision=args.precision, device=local, max_batch_scans=N_SRC_SCANS,
               max_batch_pairs=N_CAND)
  eng.load_weights(w)

  # ---- candidate bank of this rank: 32 synthetic scans encoded by the product path, yaw-rolled to 1101
  clouds = [synth.kitti_like_cloud(1000 * rank + s) for s in range(N_SRC_SCANS)]
  fv_src = eng.leg(eng.preprocess(eng.upload_clouds(clouds)))
  src = torch.arange(N_CAND, device=dev) % N_SRC_SCANS
  roll = (torch.arange(N_CAND, device=dev) // N_SRC_SCANS) * 10
  rows = (torch.arange(eng.Wf, device=dev)[None, :] - roll[:, None]) % eng.Wf          # one gather, no per-row kernels
  bank = fv_src[src[:, None], rows].contiguous()
  del fv_src
  eng.bank_prepare(bank)          # the candidate bank is static: keep its tensor-core operand copies resident

  # ---- query clouds: a fresh scan per step (pinned host copies for the e2e leg)
  n_q = 4
  q_host = [torch.from_numpy(synth.kitti_like_cloud(50000 + s)).pin_memory() for s in range(n_q)]
  q_dev = [eng.upload_clouds([q.numpy()]) for q in q_host]
  ov_all = torch.empty((world, N_CAND), dtype=torch.float32, device=dev) if rank == 0 else None
  yaw_all = torch.empty((world, N_CAND), dtype=torch.int32, device=dev) if rank == 0 else None
  qfv = torch.empty((eng.Wf, 128), dtype=torch.float32, device=dev)

  def step(i):
    """Device-resident step: query cloud already in HBM."""
    if rank == 0:
      qfv.copy_(eng.leg(eng.preprocess(q_dev[i % n_q]))[0])
    if world > 1:
      dist.broadcast(qfv, 0)                                   # 184 320 B, the only data-path collective in
    ov, yaw, _ = eng.heads_1vsN(bank, qfv, n_cand=N_CAND)
    if world > 1:
      dist.gather(ov, list(ov_all.unbind(0)) if rank == 0 else None, 0)     # 8 B per candidate back
      dist.gather(yaw, list(yaw_all.unbind(0)) if rank == 0 else None, 0)
    return ov, yaw

  ov_h = np.empty((N_CAND,), np.float32)
  yaw_h = np.empty((N_CAND,), np.int32)

  def step_e2e(i):
    """Public host-buffer entry point: H2D of the query cloud + D2H of the results every step."""
    if world == 1:
      eng.query_cloud_vs_bank_host(q_host[i % n_q], bank, n_cand=N_CAND, out_overlap=ov_h, out_yaw=yaw_h)
    else:
      # rank 0 encodes from the host buffer, the volume is broadcast, every rank copies its results to the host
      if rank == 0:
        qd = q_host[i % n_q].to(dev, non_blocking=True)
        off = torch.tensor([0, qd.shape[0]], dtype=torch.int64, device=dev)
        from overlapnet_b200.engine import CloudBatch
        qfv.copy_(eng.leg(eng.preprocess(CloudBatch(qd, off, [0, qd.shape[0]])))[0])
      dist.broadcast(qfv, 0)
      ov, yaw, _ = eng.heads_1vsN(bank, qfv, n_cand=N_CAND)
      dist.gather(ov, list(ov_all.unbind(0)) if rank == 0 else None, 0)
      dist.gather(yaw, list(yaw_all.unbind(0)) if rank == 0 else None, 0)
      if rank == 0:
        ov_all.cpu(), yaw_all.cpu()
      torch.cuda.synchronize()

  def timed(fn, steps, warmup):
    for i in range(warmup):
      fn(i)
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
      fn(warmup + i)
    b.record()
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b)], device=dev)
    if world > 1:
      dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())

  # ---- timed region 1: device-resident
  sampler = ClockSampler(local) if rank == 0 else None
  if sampler:
    sampler.start()
    sampler.wait_first_sample()
  eng.profile_enable(True)
  for k in ('delta_conv1', 'conv2', 'conv3', 'corr', 'leg', 'project_scatter', 'project_gather'):
    eng.profile_read(k)
  l0 = eng.launch_count()
  # warm-up is inside timed(); reset the profile after it by timing warmup separately
  timed(step, 0, args.warmup)
  for k in ('delta_conv1', 'conv2', 'conv3', 'corr', 'leg', 'project_scatter', 'project_gather'):
    eng.profile_read(k)
  l0 = eng.launch_count()
  ms = timed(step, args.steps, 0)
  launches = eng.launch_count() - l0
  prof = {k: eng.profile_read(k) for k in ('delta_conv1', 'conv2', 'conv3', 'corr', 'leg', 'project_scatter',
                                           'project_gather')}
  eng.profile_enable(False)
  # ---- timed region 2: end to end through the host-buffer entry point
  ms_e2e = timed(step_e2e, args.steps, args.warmup)
  clocks = sampler.stop() if sampler else None

  if rank == 0:
    tflops_peak, hbm_peak, peak_src = peaks()
    pairs = world * N_CAND * args.steps
    value = pairs / (ms * 1e-3)
    e2e = pairs / (ms_e2e * 1e-3)
    k_ms, k_n = prof['delta_conv1']
    ach = (N_CAND * FLOP_DELTA_CONV1 / 1e12) / (k_ms / max(k_n, 1) * 1e-3) if k_n else None
    shares = {k: round(v[0] / ms, 4) for k, v in prof.items()}
    proj_ms = (prof['project_scatter'][0] + prof['project_gather'][0]) / max(prof['project_scatter'][1], 1)
    npts = int(q_host[0].shape[0])
    cpu_val, cpu_dt = cpu_pairs_per_s(w, args.cpu_pairs, threads=os.cpu_count())
    traffic = None
    tp = os.path.join(ROOT, 'profiles', 'r1_delta_traffic.json')
    if os.path.exists(tp) and args.precision == 'f16_tc':
      with open(tp) as f:
        tj = json.load(f)
      traffic = (tj['dram_bytes_read'] + tj['dram_bytes_write']) * N_CAND / tj['pairs_per_launch']
    line = {
        'metric': METRIC, 'value': value, 'unit': 'pairs/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f16' if args.precision == 'f16_tc' else 'f32', 'data': 'synthetic',
        'config': {'workload': '1 query x 1101 candidates per GPU, geo-only 64x900 (BASELINE config 2)',
                   'candidates_per_gpu': N_CAND, 'query_points': int(q_host[0].shape[0]),
                   'l2': 'inputs larger than L2: the fp32 candidate bank is 203 MB per step',
                   'parallelism': 'bank sharded x%d, NCCL broadcast(query 184 KB) + gather(8 B/candidate)' % world,
                   'precision': args.precision, 'weights': 'seeded Glorot (no pretrained weights offline)'},
        'gpu_launches': int(launches),
        'e2e': {'value': e2e, 'unit': 'pairs/s', 'ms_per_step': ms_e2e / args.steps,
                'h2d_bytes_per_step': int(q_host[0].numel() * 4),
                'd2h_bytes_per_step': int(N_CAND * 8 * world)},
        'roofline': {'kernel': 'k_delta_conv1_tc' if args.precision == 'f16_tc' else 'k_simt_gemm<DeltaOperand>',
                     'bound': 'tensor', 'achieved': ach, 'peak': tflops_peak, 'unit': 'TFLOP/s',
                     'frac': (ach / tflops_peak) if ach else None, 'traffic': traffic,
                     'traffic_note': 'DRAM bytes per launch from the committed ncu capture (profiles/r1_ncu_summary_v2.txt); algorithmic bytes 1.32e9 (108 MB of LEFT volumes + 1.21 GB of o1)',
                     'peak_source': peak_src,
                     'flop_per_launch': N_CAND * FLOP_DELTA_CONV1, 'avg_launch_ms': k_ms / max(k_n, 1),
                     'share_of_step': shares},
        'range_proj': {'mpts_per_s': npts / (proj_ms * 1e-3) / 1e6 if proj_ms else None, 'scans_per_launch': 1,
                       'ms_per_scan': proj_ms, 'algorithmic_bytes': npts * 16 + 64 * 900 * 16,
                       'gb_per_s': (npts * 16 + 64 * 900 * 16) / (proj_ms * 1e-3) / 1e9 if proj_ms else None,
                       'hbm_peak_gb_per_s': hbm_peak,
                       'note': 'one 124 668-point scan per launch is launch-latency bound; see profiles/ for the batched figure'},
        'whole_path_tflops': pairs * FLOP_PAIR / 1e12 / (ms * 1e-3),
        'cpu_baseline': {'value': cpu_val, 'unit': 'pairs/s', 'cores': os.cpu_count(), 'kind': 'port',
                         'sample': '%d of the 1101 pairs (+1 projection, +1 leg) in %.1f s, torch-CPU fp32'
                                   % (args.cpu_pairs, cpu_dt)},
        'clocks': clocks,
    }
    print(json.dumps(line))
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
