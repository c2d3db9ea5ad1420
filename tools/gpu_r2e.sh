#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
python -m pytest tests/test_gpu_wide.py tests/test_gpu_parity.py -m gpu -x -q > $O/r2e_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2e_pytest.log
tail -4 $O/r2e_pytest.log
for so in libpairing_b200 exp_p2noinline; do
  echo "== $so" | tee -a $O/r2e_latency.log
  PAIRING_B200_LIB=$PWD/pairing_b200/lib/$so.so python tools/bench_latency.py 2>&1 | tee -a $O/r2e_latency.log | tail -24
done
bash tools/bench_variants.sh pairing 2>&1 | tee $O/r2e_pair_variants.log
