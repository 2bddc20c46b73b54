#!/bin/bash
# One GPU-box pass: the GPU test suite, the default bench line, and compute-sanitizer
# (memcheck, then racecheck) over the small parity tests of every kernel family.
# Usage: tools/gpu_check.sh TAG   -> gpurun_out/TAG_*.log
TAG=${1:-check}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${TAG}_tests.log 2>&1
echo "tests rc=$?" >> $O/${TAG}_tests.log
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
SMALL='tests/test_gpu_lsm.py tests/test_gpu_hull_white.py tests/test_gpu_mvgbm.py tests/test_gpu_mvgbm_tc5.py tests/test_brownian_bridge.py tests/test_tangents.py tests/test_halton.py tests/test_qmc.py tests/test_milstein.py'
KEXPR='reference_kats or tabulated or two_dimensional or column_sums or swaption_reference_kat or discount_curve_paths or basket_price or fused_continuous or fused_delta or halton_matches or 1000-5 or 777-3 or 128-1 or 1-2-7 or randomisation or digital_net_argument or nd_state or heston_tangent_paths'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest $SMALL tests/test_gpu_parity.py -m gpu -q -x \
  -k "$KEXPR or fused_price or c1_notebook or ragged or per_path_initial" \
  > $O/${TAG}_memcheck.log 2>&1
echo "memcheck rc=$?" >> $O/${TAG}_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_lsm.py tests/test_gpu_parity.py tests/test_gpu_mvgbm.py tests/test_gpu_mvgbm_tc5.py -m gpu -q -x \
  -k "reference_kats or fused_price or basket_price or c1_notebook or 777-3 or 1-2-7" \
  > $O/${TAG}_racecheck.log 2>&1
echo "racecheck rc=$?" >> $O/${TAG}_racecheck.log
tail -3 $O/${TAG}_tests.log; tail -4 $O/${TAG}_memcheck.log; tail -4 $O/${TAG}_racecheck.log
