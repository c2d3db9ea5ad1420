#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
for so in libpairing_b200 exp_mmsingle; do
  echo "== $so" | tee -a $O/r2l_mm.log
  PAIRING_B200_LIB=$PWD/pairing_b200/lib/$so.so python tools/bench_latency.py --only-mm --mm-log2 17 --mm-sweep 2>&1 | tee -a $O/r2l_mm.log | tail -14
done
bash tools/bench_variants.sh pairing 2>&1 | tee $O/r2l_pair_variants.log
PAIRING_B200_LIB=$PWD/pairing_b200/lib/exp_mmsingle.so python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q -k "miller" 2>&1 | tail -3
