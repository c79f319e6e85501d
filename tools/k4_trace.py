"""Development aid: per-role clock64() timeline of CTA 0 of k_delta_conv1_tc (one jb).
Needs a trace build:  nvcc ... -DOVN_K4_TRACE -o overlapnet_b200/libovn_b200_trace.so  (see tools/gpu_trace.sh)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from overlapnet_b200 import _cabi
_cabi.LIB_PATH = os.path.join(os.path.dirname(_cabi.LIB_PATH), os.environ.get('OVN_TRACE_LIB', 'libovn_b200_trace.so'))
from overlapnet_b200.engine import Engine
from oracle import network as N

MODEL = {'additional_unsymmetric_layer3a': True, 'strides_layer1': [2, 2]}
npairs = 1101
eng = Engine(model=MODEL, precision='f16_tc', max_batch_scans=16, max_batch_pairs=npairs)
eng.load_weights(N.glorot_weights(4, MODEL, seed=0))
g = torch.Generator(device='cuda').manual_seed(0)
bank = torch.rand((npairs, 360, 128), device='cuda', generator=g)
for _ in range(3): eng.heads_1vsN(bank, bank[0], n_cand=npairs)
torch.cuda.synchronize()
eng.profile_enable(True)
for _ in range(5): eng.heads_1vsN(bank, bank[0], n_cand=npairs)
torch.cuda.synchronize()
ms, n = eng.profile_read('delta_conv1'); print('delta_conv1: %.4f ms per launch (%d)' % (ms / max(n, 1), n))
eng.profile_enable(False)
tr = np.zeros((8, 64, 4), np.int64)
rc = _cabi.lib().ovn_debug_k4_trace(tr.ctypes.data_as(C.c_void_p), C.c_longlong(tr.nbytes))
assert rc == 0, rc
t0 = tr[0, 0, 0]
m = tr[0, :60]
print('MMA issuer: step  top  ready  waited  issued   (clk rel. to step 0; delta to prev step top)')
for s in range(60):
  print('  %2d  %6d  %d  %6d  %6d   d=%d' % (s, m[s, 0] - t0, m[s, 3], m[s, 1] - t0, m[s, 2] - t0, m[s, 0] - m[s - 1, 0] if s else 0))
print('jb span (60 steps): %d clk -> %.1f clk/step' % (m[59, 2] - m[0, 0], (m[59, 2] - m[0, 0]) / 60.0))
for grp in range(3):
  pr = tr[1 + grp]
  print('producer group %d: step  begin  slot_ok  st_issued  arrived' % grp)
  for s in range(60):
    if pr[s, 0]:
      ex = tr[5 + grp]
      print('  %2d  %6d  %6d  %6d  %6d   | half0 stores issued +%d, all issued +%d, wait::st +%d, arrive +%d' % (
          s, pr[s, 0] - t0, pr[s, 1] - t0, pr[s, 2] - t0, pr[s, 3] - t0, ex[s, 0] - pr[s, 1], pr[s, 2] - pr[s, 1],
          ex[s, 1] - pr[s, 2], pr[s, 3] - ex[s, 1]))
e = tr[4]
print('epilogue (jb 1 -> rel. to MMA step 0 of jb 2): wait_begin %d  d_full %d  tile (ld issued, released, staged, stored) %s' %
      (e[0, 0] - t0, e[0, 1] - t0, [(int(e[1 + t, 1] - t0), int(e[1 + t, 0] - t0), int(e[1 + t, 3] - t0), int(e[1 + t, 2] - t0)) for t in range(3)]))
