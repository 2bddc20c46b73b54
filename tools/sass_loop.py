"""Instruction mix of the hottest loop of a kernel in a cuobjdump -sass listing.

  cuobjdump -sass -fun <mangled> file.o | python tools/sass_loop.py

The hot loop is taken to be the backward branch (span <= 1500 instructions,
or argv[1]) whose span holds the most FP64 instructions (DFMA/DMUL/DADD/DSETP) -- or FFMA when there are none.
"""
import re
import sys
from collections import Counter

ins = []
for line in sys.stdin:
  m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
  if m:
    ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_index = {a: i for i, (a, _) in enumerate(ins)}


def opcode(text):
  t = text.split()
  op = t[1] if t[0].startswith('@') else t[0]
  return op.split('.')[0]


best = None
limit = int(sys.argv[1]) if len(sys.argv) > 1 else 1500   # max loop length considered
for i, (a, text) in enumerate(ins):
  m = re.search(r'BRA\S*\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)', text)
  if not m:
    continue
  tgt = int(m.group(1), 16)
  if tgt <= a and tgt in addr_index:
    j = addr_index[tgt]
    body = ins[j:i + 1]
    if len(body) > limit:
      continue
    score = sum(opcode(t) in ('DFMA', 'DMUL', 'DADD', 'DSETP') for _, t in body)
    if score == 0:
      score = sum(opcode(t) == 'FFMA' for _, t in body) / 1000.0
    if best is None or score > best[0]:
      best = (score, j, i)
score, j, i = best
body = ins[j:i + 1]
c = Counter(opcode(t) for _, t in body)
print('loop 0x%x..0x%x: %d instructions' % (ins[j][0], ins[i][0], len(body)))
dp = sum(c[k] for k in ('DFMA', 'DMUL', 'DADD', 'DSETP'))
print('FP64-pipe %d' % dp)
print(c.most_common())
