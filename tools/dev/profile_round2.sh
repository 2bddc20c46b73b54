# Round-2 evidence run (on the GPU box): the full bench line, one `ncu --set full`
# capture per dominant kernel digested to text right here (the .ncu-rep files are
# 10-25 MB each; only the text summaries travel back), and the launch list.
set -x
OUT=gpurun_out
python bench.py > $OUT/r2f_bench.json 2> $OUT/r2f_bench.err
NCU="ncu --set full --clock-control none --import-source on"
digest() {   # name kernel-regex bench-args...
  name=$1; regex=$2; shift 2
  timeout 400 $NCU -k regex:$regex -s 3 -c 1 -o /tmp/$name python bench.py --only --steps 1 --warmup 1 "$@" > /dev/null 2>&1
  {
    echo "# ncu --set full --clock-control none -k regex:$regex -s 3 -c 1 python bench.py --only --steps 1 --warmup 1 $*"
    python tools/ncu_summary.py /tmp/$name.ncu-rep
    ncu -i /tmp/$name.ncu-rep --page raw --csv | python tools/dev/ncu_extra.py
    ncu -i /tmp/$name.ncu-rep --page source --csv --print-source sass > /tmp/$name.src.csv
    python tools/ncu_hot.py /tmp/$name.src.csv 25
    python tools/dev/ncu_opmix.py /tmp/$name.src.csv
  } > $OUT/r2f_$name.txt 2>&1
  rm -f /tmp/$name.ncu-rep /tmp/$name.src.csv
}
digest c2 path_kernel --paths 2000000
digest c3 path_kernel --workload c3 --paths 4000000
digest c2qe path_kernel --workload c2_qe --paths 2000000
digest c5gen path_kernel --workload c5
digest c5lsm lsm_persistent --workload c5
digest c1 path_kernel --workload c1
digest c4_tc5 mvgbm_tc5 --workload c4 --paths 2000000
TQF_MVGBM_TC5=0 digest c4_mma mvgbm_mma --workload c4 --paths 2000000
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/r2f_launches.csv python bench.py --steps 2 --warmup 1 > /dev/null 2>&1
ls -la $OUT/r2f_*
