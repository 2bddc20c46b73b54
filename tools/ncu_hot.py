"""Aggregates the per-instruction stall samples of an ncu source-page CSV.

  ncu -i rep.ncu-rep --page source --csv --print-source sass > src.csv
  python tools/ncu_hot.py src.csv [top]

Prints sample totals per stall reason, per opcode class, and the hottest SASS
instructions with their dominant stall reason.
"""
import csv
import sys
from collections import Counter


def main():
  path = sys.argv[1]
  top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
  rows = list(csv.reader(open(path)))
  hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
  hdr = rows[hdr_i]
  col = {h: i for i, h in enumerate(hdr)}
  stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
  data = rows[hdr_i + 1:]
  tot = Counter()
  by_op = Counter()
  exec_by_op = Counter()
  lines = []
  for n, r in enumerate(data):
    if len(r) < len(hdr):
      continue
    samples = int(r[col['# Samples']] or 0)
    execd = int(r[col['Instructions Executed']] or 0)
    src = r[col['Source']].strip()
    op = src.split()[0] if src else '?'
    if op.startswith('@'):
      op = src.split()[1]
    op = op.split('.')[0]
    by_op[op] += samples
    exec_by_op[op] += execd
    st = {h: int(r[col[h]] or 0) for h in stall_cols}
    for h, v in st.items():
      tot[h] += v
    lines.append((samples, n, src, execd, st))
  total = sum(s for s, *_ in lines)
  print('total samples', total)
  print('--- by stall reason')
  for h, v in tot.most_common():
    if v:
      print('  %-24s %7d %5.1f%%' % (h, v, 100.0 * v / total))
  print('--- by opcode (samples, executed warp-instr)')
  for op, v in by_op.most_common(25):
    print('  %-10s %7d %5.1f%%  exec %d' % (op, v, 100.0 * v / total, exec_by_op[op]))
  print('--- hottest instructions')
  for samples, n, src, execd, st in sorted(lines, reverse=True)[:top]:
    dom = max(st.items(), key=lambda kv: kv[1])
    print('  #%5d %6d  %-60s exec %9d  %s=%d' % (n, samples, src[:60], execd, dom[0], dom[1]))


if __name__ == '__main__':
  main()
