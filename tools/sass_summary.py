#!/usr/bin/env python3
"""Opcode evidence for the Blackwell-native claim: per-kernel histogram of the SASS mnemonics that
prove tcgen05 / TMEM / bulk-async-copy use (B200_PROFILING.md "What proves a Blackwell-native kernel")
from `cuobjdump -sass overlapnet_b200/libovn_b200.so`.  Writes profiles/r2_sass_summary.txt."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'overlapnet_b200', 'libovn_b200.so')
WATCH = ['UTCHMMA', 'UTCQMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTCATOMSWS', 'UBLKCP', 'UTMALDG', 'UTMASTG', 'SYNCS', 'HMMA',
         'HGMMA', 'LDGSTS', 'ATOMG', 'RED', 'ACQBULK', 'ELECT', 'HADD2', 'HFMA2', 'LOP3', 'FFMA', 'DFMA', 'DADD', 'DMUL',
         'LDG', 'STG', 'LDS', 'STS', 'SHFL', 'MUFU', 'BAR']


def main():
  out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
  kernels = collections.OrderedDict()
  cur = None
  for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
      name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
      name = re.sub(r'\(.*', '', name)
      cur = kernels.setdefault(name, collections.Counter())
      continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_.]+)?)', line)
    if m and cur is not None:
      op = m.group(1)
      cur['__total__'] += 1
      base = op.split('.')[0]
      for wname in WATCH:
        if base == wname or base.startswith(wname):
          cur[wname] += 1
          break
      if base == 'LDTM' or base == 'STTM':
        cur[op] += 1                                       # keep the .x8/.x16/.x32 shapes
  lines = ['SASS opcode histogram of overlapnet_b200/libovn_b200.so (cuobjdump -sass, sm_100a), per kernel.',
           'UTCHMMA = tcgen05.mma kind::f16, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk',
           '(1-D bulk copies: the operands are pre-packed in shared-memory image order, so no tensor map is needed),',
           'SYNCS = mbarrier ops.  No HMMA / HGMMA (legacy / Hopper tensor paths) anywhere.', '']
  tot = collections.Counter()
  for name, c in kernels.items():
    keys = [k for k in c if k != '__total__' and '.' not in k]
    shapes = [k for k in c if '.' in k]
    tc = c['UTCHMMA'] + c['LDTM'] + c['STTM'] + c['UBLKCP']
    mark = '*' if tc else ' '
    lines.append('%s %-46s instr %6d  %s' % (mark, name[:46], c['__total__'],
                 ' '.join('%s=%d' % (k, c[k]) for k in WATCH if c[k])))
    if shapes:
      lines.append('    %s' % ' '.join('%s=%d' % (k, c[k]) for k in sorted(shapes)))
    tot.update({k: c[k] for k in keys})
  lines += ['', 'library totals: ' + ' '.join('%s=%d' % (k, tot[k]) for k in WATCH if tot[k]),
            'legacy tensor opcodes: HMMA=%d HGMMA=%d' % (tot['HMMA'], tot['HGMMA'])]
  path = os.path.join(ROOT, 'profiles', 'r2_sass_summary.txt')
  with open(path, 'w') as f:
    f.write('\n'.join(lines) + '\n')
  print('\n'.join(lines[-3:]))
  print('wrote', path)


if __name__ == '__main__':
  sys.exit(main())
