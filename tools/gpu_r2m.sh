#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
for rep in 1 2 3; do for so in libpairing_b200 exp_mmnever; do
  echo "== $so run $rep" | tee -a $O/r2m_mm.log
  PAIRING_B200_LIB=$PWD/pairing_b200/lib/$so.so python tools/bench_latency.py --only-mm --mm-sweep 2>&1 | grep "trips\|2^20" | tee -a $O/r2m_mm.log
done; done
bash tools/bench_variants.sh pairing 2>&1 | tee $O/r2m_pair_variants.log
python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q -k "miller" 2>&1 | tail -3
