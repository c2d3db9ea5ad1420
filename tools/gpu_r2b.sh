#!/bin/bash
# round-2 GPU call B: parity of the warp-cooperative engine / product tail / multi-device entry points, latency sweep,
# G1 wNAF occupancy variants
mkdir -p gpurun_out; O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r2b_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2b_pytest.log
tail -5 $O/r2b_pytest.log
python tools/bench_latency.py > $O/r2b_latency.log 2>&1; tail -40 $O/r2b_latency.log
for so in libpairing_b200 exp_g1b3 exp_g1b4 exp_g1b4k2 exp_g1b6; do
  [ -f pairing_b200/lib/$so.so ] || continue
  echo "== $so" >> $O/r2b_paths.log
  PAIRING_B200_LIB=$PWD/pairing_b200/lib/$so.so python tools/bench_paths.py --log2 22 --skip pairing,mm,g2 2>&1 | grep -E "config|norm|mismatch|Error|exact" >> $O/r2b_paths.log
done
cat $O/r2b_paths.log
