#!/usr/bin/env python3
"""Summarise ncu artefacts into text for profiles/:  launch list (csv) -> per-kernel shares of one
bench step; .ncu-rep (--set full) -> the raw metrics the roofline numbers come from."""
import collections
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_bytes.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed.sum', 'launch__shared_mem_per_block_dynamic', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']


def launches(path):
  rows = [r for r in csv.reader(open(path)) if len(r) > 5]
  hdr = next(r for r in rows if 'Kernel Name' in r)
  start = rows.index(hdr)
  ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
  seq = []
  for r in rows[start + 1:]:
    try:
      seq.append((r[ki], float(r[vi].replace(',', ''))))
    except ValueError:
      pass
  marks = [i for i, (n, _) in enumerate(seq) if 'k_project_scatter' in n]
  if len(marks) >= 2:
    seq = seq[marks[-2]:marks[-1]]          # one complete step
  tot = sum(v for _, v in seq)
  print('one bench step, ncu per-launch durations (cold-cache, serialised): total %.1f us' % (tot / 1e3))
  agg = collections.OrderedDict()
  for n, v in seq:
    k = n.split('(')[0][:60]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
  for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print('  %-62s n=%3d %10.1f us %5.1f%%' % (k, n, v / 1e3, 100 * v / tot))


def report(path):
  out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
  rows = list(csv.reader(out.splitlines()))
  hdr, units = rows[0], rows[1]
  for vals in rows[2:]:
    name = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
    print('kernel:', name[:100])
    for i, h in enumerate(hdr):
      if any(h == k or h.startswith(k + ' ') for k in KEYS):
        print('  %-70s %14s %s' % (h, vals[i], units[i]))


if __name__ == '__main__':
  for a in sys.argv[1:]:
    print('==', a)
    launches(a) if a.endswith('.csv') else report(a)
