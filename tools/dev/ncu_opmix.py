"""Dynamic instruction mix of a kernel from an `ncu --page source --csv` dump:
executed warp instructions per opcode, as a share and per unit of the most executed
instruction (one hot-loop iteration)."""
import csv
import sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]
i_s, i_e = hdr.index('Source'), hdr.index('Instructions Executed')
data = [(int(r[i_e]), r[i_s].strip()) for r in rows[hi + 1:] if len(r) > i_e and r[i_e].isdigit()]
tot = sum(e for e, _ in data)
mx = max(e for e, _ in data)
c = Counter()
for e, t in data:
  op = t.split()[1] if t.startswith('@') else t.split()[0]
  c['.'.join(op.split('.')[:2])] += e
print('executed warp instructions: %d; most executed instruction: %d times' % (tot, mx))
for k, v in c.most_common(28):
  print('  %-22s %6.2f%%   %8.2f per hot-loop iteration' % (k, 100.0 * v / tot, v / mx))
