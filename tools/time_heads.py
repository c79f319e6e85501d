"""Per-kernel CUDA-event times of the 1 x 1101 heads (development aid for parameter sweeps)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from overlapnet_b200.engine import Engine
from oracle import network as N

MODEL = {'additional_unsymmetric_layer3a': True, 'strides_layer1': [2, 2]}
n = 1101
eng = Engine(model=MODEL, precision='f16_tc', max_batch_scans=1, max_batch_pairs=n)
eng.load_weights(N.glorot_weights(4, MODEL, seed=0))
g = torch.Generator(device='cuda').manual_seed(0)
bank = torch.rand((n, 360, 128), device='cuda', generator=g)
eng.bank_prepare(bank, 0, n)
for _ in range(3): eng.heads_1vsN(bank, bank[0], n_cand=n)
torch.cuda.synchronize()
eng.profile_enable(True)
for _ in range(10): eng.heads_1vsN(bank, bank[0], n_cand=n)
torch.cuda.synchronize()
out = []
for k in ('delta_conv1', 'conv2', 'conv3', 'corr'):
  ms, c = eng.profile_read(k)
  out.append('%s %.4f' % (k, ms / max(c, 1)))
print(os.environ.get('TAG', ''), ' '.join(out))
