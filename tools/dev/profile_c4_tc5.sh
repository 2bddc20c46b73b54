#!/bin/bash
# ncu --set full of the C4 tcgen05 kernel (2 M paths) + digest; run on the GPU box.
O=gpurun_out
TAG=${1:-r2v}
export TQF_MVGBM_TC5=1
ncu --set full --clock-control none --import-source on -k regex:mvgbm_tc5 -s 3 -c 1 -o /tmp/c4tc5 -f \
  python bench.py --only --steps 1 --warmup 1 --workload c4 --paths 2000000 > $O/${TAG}_ncu.log 2>&1
{
  echo "# ncu --set full --clock-control none -k regex:mvgbm_tc5 -s 3 -c 1 python bench.py --only --steps 1 --warmup 1 --workload c4 --paths 2000000 (TQF_MVGBM_TC5=1)"
  python tools/ncu_summary.py /tmp/c4tc5.ncu-rep
  ncu -i /tmp/c4tc5.ncu-rep --page raw --csv > /tmp/c4tc5.raw.csv
  python tools/dev/ncu_extra.py < /tmp/c4tc5.raw.csv
  ncu -i /tmp/c4tc5.ncu-rep --page source --csv --print-source sass > /tmp/c4tc5.src.csv
  python tools/ncu_hot.py /tmp/c4tc5.src.csv 40
  python tools/dev/ncu_smem.py /tmp/c4tc5.src.csv 14
} > $O/${TAG}_c4_tc5.txt 2>&1
