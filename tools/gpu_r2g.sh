#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
python -m pytest tests/test_gpu_wide.py tests/test_gpu_parity.py -m gpu -x -q > $O/r2g_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2g_pytest.log
tail -4 $O/r2g_pytest.log
python tools/bench_latency.py --mm-log2 17 2>&1 | tee $O/r2g_latency.log | head -20
for so in libpairing_b200 exp_g2k3 exp_g2b2 exp_g2b4 exp_g2k1; do
  echo "== $so" | tee -a $O/r2g_paths.log
  PAIRING_B200_LIB=$PWD/pairing_b200/lib/$so.so python tools/bench_paths.py --log2 20 --skip pairing,mm,g1 2>&1 | grep -E "config|norm|mismatch|Error|exact" | tee -a $O/r2g_paths.log
done
