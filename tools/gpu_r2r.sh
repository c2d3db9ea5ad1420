#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out; : > $O/r2r_diag.log
for so in libpairing_b200 exp_gcdlane; do
  echo "== $so inf" | tee -a $O/r2r_diag.log
  PAIRING_B200_LIB=$PWD/pairing_b200/lib/$so.so timeout 60 python tools/diag_pairing.py 4096 inf 2>&1 | tail -2 | tee -a $O/r2r_diag.log
done
timeout 400 python -m pytest tests -m gpu -x -q --durations=4 > $O/r2r_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2r_pytest.log
tail -9 $O/r2r_pytest.log
timeout 300 bash tools/bench_variants.sh pairing 2>&1 | tee $O/r2r_pair_variants.log
BENCH_PATHS_ARGS="--log2 16" timeout 300 bash tools/bench_variants.sh mm,g1,g2 2>&1 | tee $O/r2r_paths.log
