"""Per-instruction shared-memory wavefronts / bank conflicts of an ncu source-page CSV
(`ncu -i rep --page source --csv --print-source sass`): which instruction pays them."""
import csv
import sys


def main():
  rows = list(csv.reader(open(sys.argv[1])))
  top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
  hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
  hdr = rows[hdr_i]
  data = [r for r in rows[hdr_i + 1:] if len(r) >= len(hdr)]
  src = hdr.index('Source')
  ex = hdr.index('Instructions Executed')
  want = [i for i, h in enumerate(hdr)
          if ('onflict' in h or 'Wavefronts' in h) and ('Shared' in h or 'shared' in h)]
  print('columns:', [hdr[i] for i in want])
  for i in want:
    def val(r):
      try:
        return float(r[i] or 0)
      except ValueError:
        return 0.0
    tot = sum(val(r) for r in data)
    print('--- %s (total %.0f)' % (hdr[i], tot))
    for r in sorted(data, key=val, reverse=True)[:top]:
      if val(r) > 0:
        print('  %12.0f  exec %10s  %s' % (val(r), r[ex], r[src].strip()[:70]))


if __name__ == '__main__':
  main()
