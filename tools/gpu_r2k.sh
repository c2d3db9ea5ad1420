#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
for so in libpairing_b200 exp_g1k4 exp_g1k4b3 exp_g1k4b5 exp_g1k5; do
  echo "== $so" | tee -a $O/r2k_paths.log
  PAIRING_B200_LIB=$PWD/pairing_b200/lib/$so.so python tools/bench_paths.py --log2 22 --skip pairing,mm,g2 2>&1 | grep -E "config|norm|mismatch|Error|exact" | tee -a $O/r2k_paths.log
done
