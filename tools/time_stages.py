"""Quick CUDA-event timings of each stage (development aid, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from overlapnet_b200 import synth
from overlapnet_b200.engine import Engine
from oracle import network as N

MODEL = {'additional_unsymmetric_layer3a': True, 'strides_layer1': [2, 2]}

def timeit(fn, iters=10, warm=3):
  for _ in range(warm): fn()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(True), torch.cuda.Event(True)
  a.record()
  for _ in range(iters): fn()
  b.record(); torch.cuda.synchronize()
  return a.elapsed_time(b) / iters

prec = sys.argv[1] if len(sys.argv) > 1 else 'fp32'
npairs = int(sys.argv[2]) if len(sys.argv) > 2 else 128
eng = Engine(model=MODEL, precision=prec, max_batch_scans=64, max_batch_pairs=max(npairs, 16))
eng.load_weights(N.glorot_weights(4, MODEL, seed=0))
clouds = [synth.kitti_like_cloud(s) for s in range(64)]
batch = eng.upload_clouds(clouds)
npts = sum(c.shape[0] for c in clouds)
eng.profile_enable(True)
t = timeit(lambda: eng.preprocess(batch))
for k in ('project_scatter','project_gather'):
  ms, n = eng.profile_read(k); print('  %s: %.1f us per launch (%d launches)' % (k, ms / max(n,1) * 1e3, n))
eng.profile_enable(False)
print('preprocess 64 scans: %.3f ms  -> %.1f Mpts/s, %.1f GB/s algorithmic' % (t, npts / t / 1e3, (npts * 16 + 64 * 57600 * 16) / t / 1e6))
t = timeit(lambda: eng.project(batch))
print('project(all outputs incl idx) 64 scans: %.3f ms -> %.1f Mpts/s' % (t, npts / t / 1e3))
x = eng.preprocess(batch)
t = timeit(lambda: eng.leg(x), iters=3, warm=1)
print('leg 64 scans: %.3f ms -> %.3f ms/scan, %.2f TFLOP/s' % (t, t / 64, 64 * 1.7332e-3 / t * 1e3 / 1e3 * 1e3))
fv = eng.leg(x)
bank = fv.repeat((npairs + 63) // 64, 1, 1)[:npairs].contiguous()
t = timeit(lambda: eng.heads_1vsN(bank, fv[0], n_cand=npairs), iters=3, warm=1)
print('heads 1x%d: %.3f ms -> %.1f pairs/s, %.2f TFLOP/s' % (npairs, t, npairs / t * 1e3, npairs * 2.5838e-3 / t))
print('launches', eng.launch_count())
